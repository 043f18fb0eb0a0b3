// Score-grid Gram kernel with the ENROL operand in tensor memory (tcgen05.mma "TS" form).
//
//     out[e, t] = (L[e,:] . R[t,:] + row[e] + col[t] - zmean[e]) * zinv[e]            (K6, src/pldamodule.cpp:258-277)
//
// Why a second kernel.  At d <= 256 the shared-memory port bounds gemm_bf16x3_kernel (gemm_tc.cu): per 256 x 256 of
// output a CTA moves ~780 KB through it (A + B fill, the operand reads of three MMAs per k-step, the epilogue's
// st.shared + TMA-store reads) against 4 992 cycles of MMA issue.  The enrol tile of a row sweep is at most 256
// columns of K wide, i.e. it FITS IN TMEM next to two 128-column accumulators:
//
//   TMEM (512 columns per CTA, lane = enrol row):  [0,128) acc 0 | [128,256) acc 1 | [256,384) A_hi | [384,512) A_lo
//   (A packed two bf16 per 32-bit column, K-major: column 256 + k/2 holds A[m][k], A[m][k+1])
//
// so the A operand never touches shared memory during the sweep: no A fill, and each MMA reads only its B half from
// shared memory (~520 KB per 256 x 256 of output, below the MMA time).
//
// Anatomy (CTA pair, cta_group::2, M = 256 = 128 rows per CTA, N = 128, K = 16; 512 threads):
//   warp 0      TMA producer: B_hi / B_lo tiles of 64 test rows x 64 k per CTA, 128B swizzle, 6-stage ring (16 KB/stage)
//   warp 1      TMEM allocator; in the leader CTA one thread issues the MMAs: per k-step
//               A_hi[tmem] x B_hi, A_hi[tmem] x B_lo, A_lo[tmem] x B_hi  -> one fp32 accumulator
//   warp 2      relay (peer CTA): "my A is in TMEM" -> leader, cluster-scope release
//   warps 4-7   A loader: thread = TMEM lane = enrol row; 16-byte global loads of the row's hi / lo planes ->
//               tcgen05.st.  Runs once per row block of the pair's tile range (tiles are handed out as contiguous
//               ranges in row-block-major order, so a pair reloads A once or twice per launch)
//   warps 8-15  epilogue (the FAST epilogue of gemm_tc.cu): tcgen05.ld -> row / column / z-norm terms -> swizzled
//               staging -> TMA store
// Sharded B (GemmShard) is supported like in gemm_tc.cu (flag wait before the first tile touching an owner's rows).
#include <algorithm>

#include "runtime.h"

namespace pb {
namespace {

constexpr int TS_BM = 128;                 // rows per CTA
constexpr int TS_BN = 128;                 // tile width == UMMA N
constexpr int TS_BK = 64;
constexpr int TS_STAGES = 6;
constexpr int TS_B_TILE = (TS_BN / 2) * TS_BK * 2;          // 8 KB: one plane, this CTA's half of the B tile
constexpr int TS_STAGE_BYTES = 2 * TS_B_TILE;               // hi + lo
constexpr int TS_EPI_WARPS = 8;
constexpr int TS_EPI_BUF = 32 * 32 * 4;
constexpr int TS_EPI_BYTES = TS_EPI_WARPS * TS_EPI_BUF;
constexpr int TS_COLC_BYTES = 2 * TS_BN * 4;
constexpr int TS_BAR_BYTES = 256;
constexpr int TS_SMEM = TS_STAGES * TS_STAGE_BYTES + TS_EPI_BYTES + TS_COLC_BYTES + TS_BAR_BYTES;
constexpr int TS_THREADS = 512;
constexpr int TS_CTRL_REGS = 40, TS_EPI_REGS = 152;     // the loader warpgroup keeps its 128
constexpr int TS_ACOL_HI = 256, TS_ACOL_LO = 384;

struct TsParams {
  int m, n;
  int m_units, n_tiles;
  int nkb, last_ksteps, k16;
  int n_rot;
  const __nv_bfloat16* a_hi;
  const __nv_bfloat16* a_lo;
  long long lda;
  GemmEpilogue epi;
  GemmShard shard;
  long long* dbg;     // optional stall counters (env PLDA_B200_DBG=1)
  int skip_store;     // debug (PLDA_B200_EPI=skip): no epilogue stores -> isolates the TMA / MMA loop
};

__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem, 128 lanes per CTA x 16 k] * B[smem desc]^T, cta_group::2
__device__ __forceinline__ void umma_bf16_ts_2cta(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                                  uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mbar_wait_acq_cluster(uint64_t* bar, uint32_t parity) {
  uint64_t t0 = 0;
  uint32_t spins = 0;
  while (true) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return;
    if (t0 == 0) t0 = global_timer_ns();
    if ((++spins & 0x3ff) == 0 && global_timer_ns() - t0 > 4000000000ull) {
      printf("plda_b200: gemm_ts mbarrier wait timeout block %d thread %d\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_arrive_rel_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __noinline__ void ts_shard_wait(const unsigned* f, unsigned epoch, unsigned* err) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
  if (static_cast<int>(v - epoch) >= 0) return;
  if (*reinterpret_cast<volatile unsigned*>(err) != 0u) return;
  const uint64_t t0 = global_timer_ns();
  uint32_t spins = 0;
  while (true) {
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
    if (static_cast<int>(v - epoch) >= 0) return;
    if ((++spins & 0xff) == 0 &&
        (global_timer_ns() - t0 > 2000000000ull || *reinterpret_cast<volatile unsigned*>(err) != 0u)) {
      atomicAdd(err, 1u);
      return;
    }
  }
}

__global__ void __launch_bounds__(TS_THREADS, 1)
gemm_ts_kernel(const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
               const __grid_constant__ CUtensorMap tm_out, const TsParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) {
    if (threadIdx.x == 0) printf("plda_b200: dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  uint8_t* epi_base = smem + TS_STAGES * TS_STAGE_BYTES;
  float* colc = reinterpret_cast<float*>(epi_base + TS_EPI_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(epi_base + TS_EPI_BYTES + TS_COLC_BYTES);
  uint64_t* empty = full + TS_STAGES;
  uint64_t* tfull = empty + TS_STAGES;
  uint64_t* tempty = tfull + 2;
  uint64_t* a_ready = tempty + 2;      // this CTA's loader has written its rows of A
  uint64_t* a_ready_x = a_ready + 1;   // leader only: the peer's A is in its TMEM
  uint64_t* a_free = a_ready_x + 1;    // every MMA that read the current A has retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_free + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const long long tiles = static_cast<long long>(p.m_units) * p.n_tiles;
  const int pairs = gridDim.x >> 1, pair = blockIdx.x >> 1;
  const int t0 = static_cast<int>(tiles * pair / pairs), t1 = static_cast<int>(tiles * (pair + 1) / pairs);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_b_hi);
    tma_prefetch_desc(&tm_b_lo);
    tma_prefetch_desc(&tm_out);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < TS_STAGES; ++s) {
        mbar_init(&full[s], 1);
        mbar_init(&empty[s], 1);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(&tfull[a], 1);
        mbar_init(&tempty[a], 2 * TS_EPI_WARPS);
      }
      mbar_init(a_ready, 4);
      mbar_init(a_ready_x, 1);
      mbar_init(a_free, 1);
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc_2cta(tmem_slot, 512);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(TS_CTRL_REGS));
    if (warp == 0) {
      // ===================== TMA producer: B tiles (both CTAs) ===================== //
      if (lane == 0) {
        int stage = 0;
        uint32_t phase = 0;
        const uint32_t tx_cta = TS_STAGE_BYTES;
        uint32_t shard_ready = 0;
        uint32_t lead_full[TS_STAGES];
#pragma unroll
        for (int i = 0; i < TS_STAGES; ++i) lead_full[i] = mapa_shared(smem_u32(&full[i]), 0);
        for (int t = t0; t < t1; ++t) {
          int nb = t % p.n_tiles + p.n_rot;
          if (nb >= p.n_tiles) nb -= p.n_tiles;
          const int brow = nb * TS_BN + static_cast<int>(rank) * (TS_BN / 2);
          if (p.shard.flags != nullptr) {
            const int b_lo = brow, b_hi = min(brow + TS_BN / 2, p.n);
            bool waited = false;
            for (int r = 0; r < p.shard.world; ++r) {
              if ((shard_ready >> r) & 1u) continue;
              if (p.shard.bounds[r] >= b_hi || p.shard.bounds[r + 1] <= b_lo) continue;
              ts_shard_wait(p.shard.flags + r, p.shard.epoch, p.shard.err);
              shard_ready |= 1u << r;
              waited = true;
            }
            if (waited) asm volatile("fence.proxy.async;" ::: "memory");
          }
          for (int kb = 0; kb < p.nkb; ++kb) {
            mbar_wait(&empty[stage], phase ^ 1);
            uint8_t* s = smem + stage * TS_STAGE_BYTES;
            if (rank == 0) mbar_arrive_expect_tx(&full[stage], 2 * tx_cta);
            uint32_t lf = lead_full[0];
#pragma unroll
            for (int i = 1; i < TS_STAGES; ++i) lf = stage == i ? lead_full[i] : lf;
            tma_load_2d_2sm(s, &tm_b_hi, lf, kb * TS_BK, brow);
            tma_load_2d_2sm(s + TS_B_TILE, &tm_b_lo, lf, kb * TS_BK, brow);
            if (++stage == TS_STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer (leader CTA) ===================== //
      if (lane == 0 && rank == 0) {
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0, a_phase = 0;
        int cur_rb = -1;
        const uint32_t idesc = umma_idesc_bf16_f32(2 * TS_BM, TS_BN);
        long long dbg_full = 0, dbg_tempty = 0, dbg_a = 0;
        const long long t_start = clock64();
        for (int t = t0; t < t1; ++t) {
          const int rb = t / p.n_tiles;
          if (rb != cur_rb) {
            // the loaders may overwrite A once every MMA issued so far has retired; then wait for the new rows
            if (cur_rb >= 0) umma_commit_2cta(a_free, 3);
            const long long ta = clock64();
            mbar_wait(a_ready, a_phase);
            mbar_wait_acq_cluster(a_ready_x, a_phase);
            dbg_a += clock64() - ta;
            a_phase ^= 1;
            tc_fence_after();
            cur_rb = rb;
          }
          long long tw = clock64();
          mbar_wait(&tempty[acc], acc_phase ^ 1);
          dbg_tempty += clock64() - tw;
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * TS_BN;
          for (int kb = 0; kb < p.nkb; ++kb) {
            tw = clock64();
            mbar_wait(&full[stage], phase);
            dbg_full += clock64() - tw;
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + stage * TS_STAGE_BYTES);
            const uint64_t b_hi = umma_desc_kmajor(sa, 128);
            const uint64_t b_lo = umma_desc_kmajor(sa + TS_B_TILE, 128);
            const int nks = kb == p.nkb - 1 ? p.last_ksteps : (TS_BK / 16);
            for (int ks = 0; ks < nks; ++ks) {
              const uint64_t off = static_cast<uint64_t>(ks * 2);         // +32 B inside the swizzle atom
              const uint32_t acol = kb * (TS_BK / 2) + ks * 8;            // 16 bf16 = 8 TMEM columns
              const uint32_t first = (kb == 0 && ks == 0) ? 0u : 1u;
              umma_bf16_ts_2cta(d_tmem, tmem_base + TS_ACOL_HI + acol, b_hi + off, idesc, first);
              umma_bf16_ts_2cta(d_tmem, tmem_base + TS_ACOL_HI + acol, b_lo + off, idesc, 1u);
              umma_bf16_ts_2cta(d_tmem, tmem_base + TS_ACOL_LO + acol, b_hi + off, idesc, 1u);
            }
            umma_commit_2cta(&empty[stage], 3);
            if (++stage == TS_STAGES) { stage = 0; phase ^= 1; }
          }
          umma_commit_2cta(&tfull[acc], 3);
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (p.dbg != nullptr && blockIdx.x < 2) {
          p.dbg[blockIdx.x * 16 + 2] = dbg_full;
          p.dbg[blockIdx.x * 16 + 3] = dbg_tempty;
          p.dbg[blockIdx.x * 16 + 4] = clock64() - t_start;
          p.dbg[blockIdx.x * 16 + 5] = t1 - t0;
          p.dbg[blockIdx.x * 16 + 13] = dbg_a;
        }
      }
    } else if (warp == 2) {
      // ===================== relay (peer CTA): its A is in TMEM -> the leader's MMA thread ===================== //
      if (lane == 0 && rank == 1) {
        uint32_t a_phase = 0;
        int cur_rb = -1;
        for (int t = t0; t < t1; ++t) {
          const int rb = t / p.n_tiles;
          if (rb == cur_rb) continue;
          cur_rb = rb;
          mbar_wait(a_ready, a_phase);
          a_phase ^= 1;
          tc_fence_after();
          mbar_arrive_rel_cluster(mapa_shared(smem_u32(a_ready_x), 0));
        }
      }
    }
  } else if (warp < 8) {
    // ===================== A loader: thread = TMEM lane = enrol row ===================== //
    const int q = warp & 3;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const int nvec = p.k16 >> 3;                       // 16-byte vectors per plane row (even: k16 % 16 == 0)
    uint32_t free_phase = 0;
    int cur_rb = -1;
    for (int t = t0; t < t1; ++t) {
      const int rb = t / p.n_tiles;
      if (rb == cur_rb) continue;
      if (cur_rb >= 0) { mbar_wait(a_free, free_phase); free_phase ^= 1; tc_fence_after(); }
      cur_rb = rb;
      const long long m = (static_cast<long long>(rb) * 2 + rank) * TS_BM + q * 32 + lane;
      const bool valid = m < p.m;
#pragma unroll 1
      for (int plane = 0; plane < 2; ++plane) {
        const uint4* src = reinterpret_cast<const uint4*>((plane ? p.a_lo : p.a_hi) + m * p.lda);
        const uint32_t col0 = lane_base + (plane ? TS_ACOL_LO : TS_ACOL_HI);
#pragma unroll 1
        for (int v0 = 0; v0 < nvec; v0 += 8) {
          uint4 buf[8];
#pragma unroll
          for (int j = 0; j < 8; ++j)
            buf[j] = (valid && v0 + j < nvec) ? __ldg(src + v0 + j) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
          for (int j = 0; j < 8; j += 2) {
            if (v0 + j < nvec) {                       // warp uniform
              const uint32_t r[8] = {buf[j].x, buf[j].y, buf[j].z, buf[j].w, buf[j + 1].x, buf[j + 1].y, buf[j + 1].z,
                                     buf[j + 1].w};
              tmem_st_32x32b_x8(col0 + (v0 + j) * 4, r);
            }
          }
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_ready);
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(TS_EPI_REGS));
    // ===================== epilogue (warps 8..15) ===================== //
    const int ew = warp - 8;
    const int q = warp & 3;
    const int h = ew >> 2;
    uint8_t* sb = epi_base + ew * TS_EPI_BUF;
    const uint32_t sb_row = smem_u32(sb) + lane * 128;
    const GemmEpilogue& e = p.epi;
    int acc = 0;
    uint32_t acc_phase = 0;
    const bool col_cached = e.col_add != nullptr;
    const bool col_late = p.shard.flags != nullptr;
    long long dbg_tfull = 0;
    const long long t_start = clock64();
    for (int t = t0; t < t1; ++t) {
      const int rb = t / p.n_tiles;
      int nb = t - rb * p.n_tiles + p.n_rot;
      if (nb >= p.n_tiles) nb -= p.n_tiles;
      const int m0 = (rb * 2 + static_cast<int>(rank)) * TS_BM;
      const int n0 = nb * TS_BN;
      const int m = m0 + q * 32 + lane;
      const bool mvalid = m < p.m;
      const bool warp_rows_valid = (m0 + q * 32) < p.m;
      float ra = 0.f, zm = 0.f, zi = 1.f;
      if (mvalid) {
        if (e.row_add) ra = __ldg(e.row_add + m);
        if (e.zmean) { zm = __ldg(e.zmean + m); zi = __ldg(e.zinv + m); }
      }
      const float radd = ra - zm;
      const int ncols = min(TS_BN, p.n - n0);
      const int nchunks = (ncols + 31) >> 5;
      float cpre[2] = {0.f, 0.f};
      if (col_cached && !col_late) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int cc = h + 2 * i;
          if (cc < nchunks) cpre[i] = __ldg(e.col_add + n0 + cc * 32 + lane);
        }
      }
      float* colslot = colc + acc * TS_BN;
      // the four warps sharing a slot half meet once per tile: everyone has finished tile i-1 (hence its reads of
      // tile i-2's slot) before the slot is rewritten (see gemm_tc.cu)
      if (col_cached) asm volatile("bar.sync %0, 128;" ::"r"(1 + h) : "memory");
      const uint32_t tbase = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * TS_BN;
      const long long tw = clock64();
      mbar_wait(&tfull[acc], acc_phase);
      dbg_tfull += clock64() - tw;
      tc_fence_after();
      if (col_cached && col_late) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int cc = h + 2 * i;
          if (cc < nchunks) cpre[i] = __ldcg(e.col_add + n0 + cc * 32 + lane);
        }
      }
      if (col_cached) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int cc = h + 2 * i;
          if (cc < nchunks) colslot[cc * 32 + lane] = cpre[i];
        }
        __syncwarp();
      }
      uint32_t r[2][32];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int cc = h + 2 * i;
        if (cc < nchunks) tmem_ld_32x32b_x32(tbase + cc * 32, r[i]);
      }
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&tempty[acc]), 0));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int cc = h + 2 * i;
        if (cc >= nchunks) continue;
        const int nbase = n0 + cc * 32;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[i][j]);
        if (col_cached) {
          const float4* cp = reinterpret_cast<const float4*>(colslot + cc * 32);
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 tq = cp[j4];
            v[4 * j4 + 0] += tq.x; v[4 * j4 + 1] += tq.y; v[4 * j4 + 2] += tq.z; v[4 * j4 + 3] += tq.w;
          }
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = (v[j] + radd) * zi;
        if (warp_rows_valid && !p.skip_store) {
          if (lane == 0) tma_store_wait_read<0>();
          __syncwarp();
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const uint32_t addr = sb_row + ((j4 ^ (lane & 7)) << 4);
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v[4 * j4 + 0]), "f"(v[4 * j4 + 1]),
                         "f"(v[4 * j4 + 2]), "f"(v[4 * j4 + 3])
                         : "memory");
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tm_out, sb, nbase, m0 + q * 32);
            tma_store_commit();
          }
        }
      }
    }
    if (lane == 0) tma_store_wait<0>();
    if (p.dbg != nullptr && blockIdx.x < 2 && lane == 0 && (ew == 0 || ew == 7)) {
      const int o = blockIdx.x * 16 + (ew == 0 ? 6 : 9);
      p.dbg[o + 0] = dbg_tfull;
      p.dbg[o + 2] = clock64() - t_start;
    }
  }

  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc_2cta(tmem_base, 512);
  }
}

}  // namespace

// Returns false when the problem does not qualify (the caller then uses gemm_bf16x3_kernel).
bool gemm_ts_score(Context& ctx, const SplitOperand& a, const SplitOperand& b, int64_t m, int64_t n, int64_t k,
                   const GemmEpilogue& epi, const GemmShard* shard) {
  const int64_t k16 = round_up(k, 16);
  if (!ctx.gemm_ts || k16 > 256 || m < 512 || n < 256) return false;
  if (epi.out == nullptr || epi.grp != nullptr || epi.rsum != nullptr || epi.lse_max != nullptr || epi.mom != nullptr ||
      epi.hist_n != nullptr)
    return false;
  if (epi.ldo % 4 != 0 || (reinterpret_cast<uintptr_t>(epi.out) & 15) != 0) return false;
  if (a.ld % 8 != 0 || b.ld % 8 != 0 || a.ld < k16 || b.ld < k16) return false;
  if (((reinterpret_cast<uintptr_t>(a.hi) | reinterpret_cast<uintptr_t>(a.lo) | reinterpret_cast<uintptr_t>(b.hi) |
        reinterpret_cast<uintptr_t>(b.lo)) & 15) != 0)
    return false;
  if ((ctx.epi_mode != 0 && ctx.epi_mode != 2) || ctx.epi_sector || !ctx.gemm_two_cta) return false;
  TsParams p{};
  p.m = static_cast<int>(m);
  p.n = static_cast<int>(n);
  p.m_units = static_cast<int>(ceil_div(m, 2 * TS_BM));
  p.n_tiles = static_cast<int>(ceil_div(n, TS_BN));
  p.k16 = static_cast<int>(k16);
  p.nkb = static_cast<int>(ceil_div(k16, TS_BK));
  p.last_ksteps = static_cast<int>((k16 - static_cast<int64_t>(TS_BK) * (p.nkb - 1)) / 16);
  p.a_hi = a.hi;
  p.a_lo = a.lo;
  p.lda = a.ld;
  p.epi = epi;
  p.dbg = ctx.gemm_dbg.size() >= 32 ? ctx.gemm_dbg.get() : nullptr;
  p.skip_store = ctx.epi_mode == 2 ? 1 : 0;
  if (shard != nullptr && shard->flags != nullptr) {
    p.shard = *shard;
    const int first = shard->bounds[shard->rank] < n ? shard->bounds[shard->rank] : 0;
    p.n_rot = first / TS_BN;
  }
  CUtensorMap tb_hi, tb_lo, tout;
  encode_tmap_2d(&tb_hi, TmaType::BF16, b.hi, k16, b.rows, b.ld * 2, TS_BK, TS_BN / 2, 128);
  encode_tmap_2d(&tb_lo, TmaType::BF16, b.lo, k16, b.rows, b.ld * 2, TS_BK, TS_BN / 2, 128);
  encode_tmap_2d(&tout, TmaType::F32, epi.out, n, m, epi.ldo * 4, 32, 32, 128);
  const long long tiles = static_cast<long long>(p.m_units) * p.n_tiles;
  const int pairs = static_cast<int>(std::min<long long>(tiles, ctx.num_sms / 2));
  static std::once_flag once;
  std::call_once(once, [] {
    cudaFuncSetAttribute(gemm_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SMEM);
  });
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (ctx.profile_gemm) {
    PB_CUDA(cudaEventCreate(&e0));
    PB_CUDA(cudaEventCreate(&e1));
    PB_CUDA(cudaEventRecord(e0, ctx.stream));
  }
  const bool pdl = ctx.pdl_pending && !ctx.profile_gemm;
  ctx.pdl_pending = false;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(TS_THREADS);
  cfg.dynamicSmemBytes = TS_SMEM;
  cfg.stream = ctx.stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 2 : 1;
  PB_CUDA(cudaLaunchKernelEx(&cfg, gemm_ts_kernel, tb_hi, tb_lo, tout, p));
  if (ctx.profile_gemm) {
    PB_CUDA(cudaEventRecord(e1, ctx.stream));
    ctx.gemm_events.emplace_back(e0, e1);
  }
  ctx.count_launch();
  return true;
}

}  // namespace pb
