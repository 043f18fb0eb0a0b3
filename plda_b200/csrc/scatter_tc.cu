// K3 fused: within-class scatter of the stats pass without a materialised operand.
//
//     S = sum_p w_p (x_p - m_c(p)) (x_p - m_c(p))^T ,   w_p = 1 / n_c(p)  (PLDA)  or 1  (LDA)
//
// Replaces PldaStats::AddSamples -> SpMatrix::AddMat2 per speaker (src/pldamodule.cpp:94-98) and the class-centred
// Gram of python/liblda/lda.py:184-191.  Round 1 read the rows for the class means, read them again to write a
// centred / scaled / transposed split-bf16 operand and read that twice more in a split-K SYRK (~4x the algorithmic
// HBM traffic).  Here ONE kernel reads every row once:
//
//   * the rows are centred on an ANCHOR of their class that needs no prior pass -- the class's first row g_c --
//     so the values fed to the tensor cores have within-class magnitude (no cancellation in bf16):
//         S = sum_p w_p (x_p - g_c)(x_p - g_c)^T  -  sum_c n_c w_c delta_c delta_c^T ,   delta_c = m_c - g_c
//     the class sums sum_p (x_p - g_c) fall out of the same registers (vector fp32 reductions at class changes),
//     giving delta_c, the class means, and the rank-K correction (a second, tiny launch of the same kernel on the
//     delta rows);
//   * producer warps load the rows with 16-byte loads (through the `order` gather of the label sort), centre, scale by
//     sqrt(w_p), split into bf16 hi / lo and store straight into the shared-memory layout tcgen05 reads for an
//     MN-MAJOR operand (rows of x are the reduction axis, columns the M / N axis: no transpose anywhere);
//   * a CTA pair (cta_group::2) owns a contiguous range of rows and accumulates a 256 x DP block row of S in TMEM
//     (fp32) over its whole range: 3 MMAs per k-step (hi*hi + hi*lo + lo*hi), M = 256, N = 256; for DP = 512 two
//     pair types (block rows 0 / 1) walk the same ranges, so the second read of a row hits L2;
//   * partial block rows are reduced (and symmetrised) in fp64.
//
// Shared-memory operand layout (per 128-column block, per plane): canonical MN-major, 128-byte swizzle --
//   [column group of 64][k row 0..63][64 columns = 128 B], 16-byte chunk index ^= (k row & 7);
//   descriptor: leading byte offset (between column groups) 8192, stride byte offset (between 8-row groups) 1024.
#include <algorithm>

#include "kernels.h"

namespace pb {
namespace {

constexpr int SC_KROWS = 64;                                  // reduction rows per k-block
constexpr int SC_TILE_BYTES = SC_KROWS * 128 * 2;             // one 128-column block, one plane: 16 KB
constexpr int SC_BLOCK_BYTES = 2 * SC_TILE_BYTES;             // hi + lo
constexpr int SC_STAGES = 3;
constexpr int SC_PROD_WARPS = 16;
constexpr int SC_THREADS = 64 + 32 * SC_PROD_WARPS;           // warp 0 relay, warp 1 MMA issuer, warps 2..17 producers
constexpr int SC_BAR_BYTES = 256;

struct ScatterParams {
  const void* x;
  long long ld;
  int d;
  const int4* meta;      // [nkb * 64] {source row, anchor row (-1: none), sqrt(weight) bits, class}
  void* csum;            // [classes][csum_ld] class sums of (x - anchor) in the rows' own type T, or null
  int csum_ld;
  float* partial;        // [ranges][dp][dp]
  int dp;
  int nkb;
  int ranges;
};

__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
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
      printf("plda_b200: scatter mbarrier wait timeout block %d thread %d\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
// remote arrive that PUBLISHES: everything the arriving thread has observed (the local full barrier, i.e. the
// producers' shared-memory stores) happens-before a cluster-scope acquire wait on the target barrier
__device__ __forceinline__ void mbar_arrive_release_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}
// MN-major operand tile, 128-byte swizzle (see the file header)
__device__ __forceinline__ uint64_t umma_desc_mnmajor_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(8192 >> 4) << 16;     // leading byte offset: next group of 64 columns
  d |= static_cast<uint64_t>(1024 >> 4) << 32;     // stride byte offset: next group of 8 k rows
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// One CTA pair = (block row I of S, row range).  T: float / double rows.  NBLK: 128-column blocks staged per CTA
// (1: d <= 256, 2: d <= 512) = number of 256-column accumulators = number of pair types.
template <typename T, int NBLK>
__global__ void __launch_bounds__(SC_THREADS, 1)
scatter_syrk_kernel(const ScatterParams p) {
  constexpr int CPL = 4 * NBLK;                          // columns per producer lane
  // a producer warp owns 4 rows of every k-block and converts them in NB batches of RB rows; the next batch is in
  // flight (registers, ping-pong) while one is converted
  constexpr int RB = (sizeof(T) == 8 && NBLK == 2) ? 1 : 2;
  constexpr int NB = 4 / RB;
  constexpr int STAGE_BYTES = NBLK * SC_BLOCK_BYTES;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) {
    if (threadIdx.x == 0) printf("plda_b200: dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + SC_STAGES * STAGE_BYTES);
  uint64_t* xfull = full + SC_STAGES;      // leader only: the peer's stage is written
  uint64_t* empty = xfull + SC_STAGES;
  uint64_t* tfull = empty + SC_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1;
  const int blk_row = pair % NBLK;                       // block row I of S this pair accumulates
  const int range = pair / NBLK;
  const int base = p.nkb / p.ranges, rem = p.nkb % p.ranges;
  const int kb0 = range * base + min(range, rem);
  const int cnt = base + (range < rem ? 1 : 0);

  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < SC_STAGES; ++s) {
        mbar_init(&full[s], SC_PROD_WARPS);
        mbar_init(&xfull[s], 1);
        mbar_init(&empty[s], 1);
      }
      mbar_init(tfull, 1);
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc_2cta(tmem_slot, NBLK * 256);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== relay (peer CTA): its producers' stage is complete -> tell the leader's MMA thread, cluster scope =====
    if (lane == 0 && rank == 1) {
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < cnt; ++i) {
        mbar_wait(&full[stage], phase);
        mbar_arrive_release_cluster(mapa_shared(smem_u32(&xfull[stage]), 0));
        if (++stage == SC_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA) =====
    if (lane == 0 && rank == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t idesc = umma_idesc_bf16_f32(256, 256) | (1u << 15) | (1u << 16);   // A and B MN-major
      for (int i = 0; i < cnt; ++i) {
        mbar_wait(&full[stage], phase);
        mbar_wait_cluster(&xfull[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
        const uint32_t a_hi = sa + blk_row * SC_BLOCK_BYTES, a_lo = a_hi + SC_TILE_BYTES;
#pragma unroll
        for (int j = 0; j < NBLK; ++j) {
          const uint32_t b_hi = sa + j * SC_BLOCK_BYTES, b_lo = b_hi + SC_TILE_BYTES;
          const uint32_t d_tmem = tmem_base + j * 256;
#pragma unroll
          for (int ks = 0; ks < SC_KROWS / 16; ++ks) {
            const uint32_t off = ks * 16 * 128;            // 16 k rows of 128 B
            const uint32_t acc = (i == 0 && ks == 0) ? 0u : 1u;
            umma_bf16_ss_2cta(d_tmem, umma_desc_mnmajor_sw128(a_hi + off), umma_desc_mnmajor_sw128(b_hi + off), idesc, acc);
            umma_bf16_ss_2cta(d_tmem, umma_desc_mnmajor_sw128(a_hi + off), umma_desc_mnmajor_sw128(b_lo + off), idesc, 1u);
            umma_bf16_ss_2cta(d_tmem, umma_desc_mnmajor_sw128(a_lo + off), umma_desc_mnmajor_sw128(b_hi + off), idesc, 1u);
          }
        }
        umma_commit_2cta(&empty[stage], 3);
        if (++stage == SC_STAGES) { stage = 0; phase ^= 1; }
      }
      umma_commit_2cta(tfull, 3);
    }
  } else {
    // ===== producers (both CTAs): rows -> centred, scaled, split operand tiles =====
    // 16 warps; warp pw converts the k rows pw, pw + 16, pw + 32, pw + 48 of every k-block (one BATCH) -- its next
    // batch is in flight (registers, ping-pong) while the current one is converted.  Everything that does not change
    // from row to row is hoisted: the column state of the lane, the swizzled store offset (k row & 7 == pw & 7).
    const int pw = warp - 2;
    const int cl = lane * CPL;                                         // column inside the CTA's staged span
    const int lb = cl >> 7;                                            // local block (0 .. NBLK-1)
    const int cb = cl & 127;                                           // column inside the block
    const int gcol = 128 * (2 * lb + static_cast<int>(rank)) + cb;     // column of x
    const bool aligned = ((reinterpret_cast<uintptr_t>(p.x) | (static_cast<uintptr_t>(p.ld) * sizeof(T))) & 15) == 0;
    // 1: all CPL columns exist and 16-byte loads are legal; 2: guarded scalar loads; 0: beyond d (zeros, nothing loaded)
    const int cmode = gcol + CPL <= p.d ? (aligned ? 1 : 2) : (gcol < p.d ? 2 : 0);
    const T* __restrict__ xcol = static_cast<const T*>(p.x) + gcol;
    const int ncol = min(CPL, p.d - gcol);                             // valid columns of this lane (cmode 2)
    const uint32_t st_off = lb * SC_BLOCK_BYTES + (cb >> 6) * 8192 + pw * 128 +
                            ((((cb & 63) >> 3) ^ (pw & 7)) << 4) + (cb & 7) * 2;
    const bool do_sums = p.csum != nullptr && blk_row == 0;
    const int4* __restrict__ mrow = p.meta + static_cast<long long>(kb0) * SC_KROWS + pw + 16 * (lane & 3);

    auto load_row = [&](T (&v)[CPL], long long row) {
      const T* src = xcol + row * p.ld;
      if (cmode == 1) {
        if constexpr (sizeof(T) == 4) {
#pragma unroll
          for (int j = 0; j < CPL / 4; ++j) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(src) + j);
            v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < CPL / 2; ++j) {
            const double2 t = __ldg(reinterpret_cast<const double2*>(src) + j);
            v[2 * j] = t.x; v[2 * j + 1] = t.y;
          }
        }
      } else if (cmode == 2) {
#pragma unroll
        for (int j = 0; j < CPL; ++j) v[j] = j < ncol ? __ldg(src + j) : T(0);
      } else {
#pragma unroll
        for (int j = 0; j < CPL; ++j) v[j] = T(0);
      }
    };
    // meta of a batch: lane r (and r + 4, ...) holds the record of the warp's r-th row of k-block i
    auto load_meta = [&](int i) -> int4 {
      return i < cnt ? __ldg(mrow + static_cast<long long>(i) * SC_KROWS) : make_int4(0, 0, 0, 0);
    };
    auto issue_loads = [&](T (&buf)[RB][CPL], const int4& m, int b) {
#pragma unroll
      for (int r = 0; r < RB; ++r) load_row(buf[r], __shfl_sync(0xffffffffu, m.x, b * RB + r));
    };
    T g[CPL];
    T acc[CPL];      // class sums in the rows' own precision: fp64 rows keep fp64 class means
#pragma unroll
    for (int c = 0; c < CPL; ++c) { g[c] = T(0); acc[c] = T(0); }
    int g_row = -1, cur_cls = -1;
    auto flush = [&]() {
      if (cur_cls >= 0 && cmode != 0) {
        T* dst = static_cast<T*>(p.csum) + static_cast<long long>(cur_cls) * p.csum_ld + gcol;
        if constexpr (sizeof(T) == 4) {
          red_add_v4(dst, acc[0], acc[1], acc[2], acc[3]);
          if (CPL == 8) red_add_v4(dst + 4, acc[CPL - 4], acc[CPL - 3], acc[CPL - 2], acc[CPL - 1]);
        } else {
#pragma unroll
          for (int c = 0; c < CPL; ++c) atomicAdd(dst + c, acc[c]);
        }
      }
#pragma unroll
      for (int c = 0; c < CPL; ++c) acc[c] = T(0);
    };
    int stage = 0;
    uint32_t phase = 0;
    auto convert = [&](T (&buf)[RB][CPL], const int4& m, int b) {
      if (b == 0) mbar_wait(&empty[stage], phase ^ 1);
      const uint32_t sbase = smem_u32(smem + stage * STAGE_BYTES) + st_off;
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        const int wr = b * RB + r;                                      // the warp's wr-th row: k row pw + 16 wr
        const int anc = __shfl_sync(0xffffffffu, m.y, wr);
        const float scl = __int_as_float(__shfl_sync(0xffffffffu, m.z, wr));
        const int cls = __shfl_sync(0xffffffffu, m.w, wr);
        if (anc != g_row) {                                             // new class: its anchor row (warp uniform)
          if (anc >= 0) {
            load_row(g, anc);
          } else {
#pragma unroll
            for (int c = 0; c < CPL; ++c) g[c] = T(0);
          }
          g_row = anc;
        }
        if (do_sums && cls != cur_cls) { flush(); cur_cls = cls; }
        float v[CPL];
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
          const T dv = buf[r][c] - g[c];                                // fp32 rows: exact to the ulp of a small value
          if (do_sums) acc[c] += dv;
          v[c] = static_cast<float>(dv * static_cast<T>(scl));
        }
        uint32_t hi[CPL / 2], lo[CPL / 2];
#pragma unroll
        for (int c = 0; c < CPL / 2; ++c) {
          uint32_t hh, ll;
          asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hh) : "f"(v[2 * c + 1]), "f"(v[2 * c]));
          const float r0 = v[2 * c] - __uint_as_float(hh << 16);
          const float r1 = v[2 * c + 1] - __uint_as_float(hh & 0xffff0000u);
          asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(ll) : "f"(r1), "f"(r0));
          hi[c] = hh;
          lo[c] = ll;
        }
        const uint32_t addr = sbase + wr * (16 * 128);
        if (CPL == 8) {
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(hi[0]), "r"(hi[1]), "r"(hi[CPL / 2 - 2]),
                       "r"(hi[CPL / 2 - 1]) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr + SC_TILE_BYTES), "r"(lo[0]), "r"(lo[1]),
                       "r"(lo[CPL / 2 - 2]), "r"(lo[CPL / 2 - 1]) : "memory");
        } else {
          asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(hi[0]), "r"(hi[1]) : "memory");
          asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr + SC_TILE_BYTES), "r"(lo[0]), "r"(lo[1]) : "memory");
        }
      }
      if (b == NB - 1) {
        fence_proxy_async_smem();      // generic-proxy stores -> visible to the tensor core's (async proxy) reads
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[stage]);
        if (++stage == SC_STAGES) { stage = 0; phase ^= 1; }
      }
    };
    T bufa[RB][CPL], bufb[RB][CPL];
    int4 m_cur = load_meta(0), m_nxt = load_meta(1);
    issue_loads(bufa, m_cur, 0);
    for (int i = 0; i < cnt; ++i) {
      const int4 m_nn = load_meta(i + 2);
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        // prefetch: the next batch of this k-block, or the first batch of the next one (NB is even: it is bufa's turn)
        if (b + 1 < NB) {
          if ((b + 1) & 1) issue_loads(bufb, m_cur, b + 1); else issue_loads(bufa, m_cur, b + 1);
        } else if (i + 1 < cnt) {
          issue_loads(bufa, m_nxt, 0);
        }
        if (b & 1) convert(bufb, m_cur, b); else convert(bufa, m_cur, b);
      }
      m_cur = m_nxt;
      m_nxt = m_nn;
    }
    if (do_sums) flush();

    // ===== epilogue: this CTA's 128 rows of the block row, all accumulators, -> partial[range] =====
    mbar_wait(tfull, 0);
    tc_fence_after();
    const int q = warp & 3;                      // TMEM lane quarter this warp may read
    const int h = pw >> 2;
    const int row = 256 * blk_row + 128 * static_cast<int>(rank) + 32 * q + lane;
    float* prow = p.partial + (static_cast<long long>(range) * p.dp + row) * p.dp;
    for (int c = h; c < 8 * NBLK; c += SC_PROD_WARPS / 4) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c * 32, r);
      tmem_ld_wait();
      if (row < p.dp) {
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4)
          *reinterpret_cast<float4*>(prow + c * 32 + 4 * j4) =
              make_float4(__uint_as_float(r[4 * j4]), __uint_as_float(r[4 * j4 + 1]), __uint_as_float(r[4 * j4 + 2]),
                          __uint_as_float(r[4 * j4 + 3]));
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc_2cta(tmem_base, NBLK * 256);
  }
}

// meta records of the sorted positions: source row, anchor row (first row of the class), sqrt(weight), class
__global__ void scatter_meta_kernel(const int32_t* __restrict__ order, const int32_t* __restrict__ seg_of_pos,
                                    const int32_t* __restrict__ seg_start, long long n, long long n_pad,
                                    int scale_by_count, int4* __restrict__ meta) {
  const long long p = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (p >= n_pad) return;
  // padding positions repeat the anchor row of the last class: (x - anchor) is exactly zero there, so they add
  // nothing to the scatter or to the class sums and the producer needs no branch for them
  const long long q = p < n ? p : n - 1;
  const int s = seg_of_pos[q];
  const int first = seg_start[s];
  const int cnt = seg_start[s + 1] - first;
  int4 m;
  m.x = p < n ? order[p] : order[first];
  m.y = order[first];
  m.z = __float_as_int(scale_by_count ? static_cast<float>(rsqrt(static_cast<double>(cnt))) : 1.f);
  m.w = s;
  meta[p] = m;
}

// plain SYRK of a row matrix (the correction term): no anchor, unit weight, no class sums
__global__ void plain_meta_kernel(long long n, long long n_pad, int4* __restrict__ meta) {
  const long long p = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (p >= n_pad) return;
  // padding: row 0 with weight 0
  meta[p] = p < n ? make_int4(static_cast<int>(p), -1, __float_as_int(1.f), 0) : make_int4(0, -1, __float_as_int(0.f), 0);
}

// delta_c = csum_c / n_c (scaled by sqrt(n_c w_c) for the correction SYRK);  mean_c = anchor_c + csum_c / n_c
template <typename T>
__global__ void scatter_means_kernel(const T* __restrict__ x, long long ld, int d, const int32_t* __restrict__ order,
                                     const int32_t* __restrict__ seg_start, const T* __restrict__ csum, int csum_ld,
                                     int scale_by_count, float* __restrict__ delta, int delta_ld,
                                     double* __restrict__ means, int32_t* __restrict__ counts) {
  const long long s = blockIdx.x;
  const int first = seg_start[s];
  const int cnt = seg_start[s + 1] - first;
  const double inv = 1.0 / static_cast<double>(cnt);
  // weight of the rank-one correction: n_c w_c = 1 (w = 1/n_c) or n_c (w = 1)
  const double cw = scale_by_count ? 1.0 : sqrt(static_cast<double>(cnt));
  const T* anchor = x + static_cast<long long>(order[first]) * ld;
  for (int c = threadIdx.x; c < delta_ld; c += blockDim.x) {
    double dl = 0.0;
    if (c < d) {
      dl = static_cast<double>(csum[s * csum_ld + c]) * inv;
      means[s * d + c] = static_cast<double>(anchor[c]) + dl;
    }
    delta[s * delta_ld + c] = static_cast<float>(dl * cw);
  }
  if (threadIdx.x == 0 && counts) counts[s] = cnt;
}

__global__ void scatter_reduce_kernel(const float* __restrict__ main_p, int ranges1, const float* __restrict__ corr_p,
                                      int ranges2, int dp, int d, double* __restrict__ out) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<long long>(d) * d) return;
  const int i = static_cast<int>(idx / d), j = static_cast<int>(idx % d);
  const long long plane = static_cast<long long>(dp) * dp;
  const long long a = static_cast<long long>(i) * dp + j, b = static_cast<long long>(j) * dp + i;
  double s = 0.0;
  for (int r = 0; r < ranges1; ++r)
    s += static_cast<double>(main_p[r * plane + a]) + static_cast<double>(main_p[r * plane + b]);
  for (int r = 0; r < ranges2; ++r)
    s -= static_cast<double>(corr_p[r * plane + a]) + static_cast<double>(corr_p[r * plane + b]);
  out[idx] = 0.5 * s;
}

template <typename T, int NBLK>
void launch_syrk(Context& ctx, const ScatterParams& p) {
  constexpr int smem = SC_STAGES * NBLK * SC_BLOCK_BYTES + SC_BAR_BYTES;
  static std::once_flag once;
  std::call_once(once, [] {
    cudaFuncSetAttribute(scatter_syrk_kernel<T, NBLK>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  });
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * NBLK * p.ranges);
  cfg.blockDim = dim3(SC_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = ctx.stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  PB_CUDA(cudaLaunchKernelEx(&cfg, scatter_syrk_kernel<T, NBLK>, p));
  ctx.count_launch();
}

// one pass of the fused kernel over `rows` rows described by w.meta; returns the number of ranges written
int run_syrk(Context& ctx, const void* x, bool is_f32, int64_t d, int64_t ld, int64_t rows, const int4* meta,
             void* csum, int csum_ld, float* partial, int dp) {
  const int nblk = dp / 256;
  ScatterParams p;
  p.x = x;
  p.ld = ld;
  p.d = static_cast<int>(d);
  p.meta = meta;
  p.csum = csum;
  p.csum_ld = csum_ld;
  p.partial = partial;
  p.dp = dp;
  p.nkb = static_cast<int>(ceil_div(rows, SC_KROWS));
  const int pairs = std::max(1, ctx.num_sms / 2);
  p.ranges = std::max(1, std::min(p.nkb, pairs / nblk));
  if (nblk == 1) {
    if (is_f32) launch_syrk<float, 1>(ctx, p); else launch_syrk<double, 1>(ctx, p);
  } else {
    if (is_f32) launch_syrk<float, 2>(ctx, p); else launch_syrk<double, 2>(ctx, p);
  }
  return p.ranges;
}

}  // namespace

int scatter_fused_max_dim() { return 512; }

void scatter_fused(Context& ctx, const void* x, bool is_f32, int64_t d, int64_t ld, const Segments& seg,
                   bool scale_by_count, double* scatter_out, double* means_out, int32_t* counts_out,
                   ScatterWork& w) {
  PB_CHECK(d > 0 && d <= 512, kInvalidArg, "scatter_fused: d <= 512");
  PB_CHECK(seg.n > 0 && seg.nseg > 0 && seg.n < (1ll << 31) - 64, kInvalidArg, "scatter_fused: bad segmentation");
  const int dp = d <= 256 ? 256 : 512;
  const int pairs = std::max(1, ctx.num_sms / 2);
  const int max_ranges = pairs / (dp / 256);
  const int64_t n_pad = round_up(seg.n, SC_KROWS), k_pad = round_up(seg.nseg, SC_KROWS);
  w.meta.reserve(static_cast<size_t>(std::max(n_pad, k_pad)));
  const size_t csum_bytes = static_cast<size_t>(seg.nseg) * dp * (is_f32 ? 4 : 8);
  w.csum.reserve(csum_bytes);
  w.delta.reserve(static_cast<size_t>(seg.nseg) * dp);
  w.partial.reserve(static_cast<size_t>(2) * max_ranges * dp * dp);
  float* part_main = w.partial.get();
  float* part_corr = part_main + static_cast<size_t>(max_ranges) * dp * dp;
  PB_CUDA(cudaMemsetAsync(w.csum.get(), 0, csum_bytes, ctx.stream));
  scatter_meta_kernel<<<static_cast<unsigned>(ceil_div(n_pad, 256)), 256, 0, ctx.stream>>>(
      seg.order.get(), seg.seg_of_pos.get(), seg.seg_start.get(), seg.n, n_pad, scale_by_count ? 1 : 0, w.meta.get());
  PB_CUDA(cudaGetLastError());
  ctx.count_launch();
  const int r1 = run_syrk(ctx, x, is_f32, d, ld, seg.n, w.meta.get(), w.csum.get(), dp, part_main, dp);
  if (is_f32)
    scatter_means_kernel<float><<<static_cast<unsigned>(seg.nseg), 128, 0, ctx.stream>>>(
        static_cast<const float*>(x), ld, static_cast<int>(d), seg.order.get(), seg.seg_start.get(),
        reinterpret_cast<const float*>(w.csum.get()), dp,
        scale_by_count ? 1 : 0, w.delta.get(), dp, means_out, counts_out);
  else
    scatter_means_kernel<double><<<static_cast<unsigned>(seg.nseg), 128, 0, ctx.stream>>>(
        static_cast<const double*>(x), ld, static_cast<int>(d), seg.order.get(), seg.seg_start.get(),
        reinterpret_cast<const double*>(w.csum.get()), dp,
        scale_by_count ? 1 : 0, w.delta.get(), dp, means_out, counts_out);
  PB_CUDA(cudaGetLastError());
  ctx.count_launch();
  // correction term: plain SYRK of the (weighted) delta rows
  plain_meta_kernel<<<static_cast<unsigned>(ceil_div(k_pad, 256)), 256, 0, ctx.stream>>>(seg.nseg, k_pad, w.meta.get());
  PB_CUDA(cudaGetLastError());
  ctx.count_launch();
  const int r2 = run_syrk(ctx, w.delta.get(), true, d, dp, seg.nseg, w.meta.get(), nullptr, 0, part_corr, dp);
  scatter_reduce_kernel<<<static_cast<unsigned>(ceil_div(d * d, 256)), 256, 0, ctx.stream>>>(
      part_main, r1, part_corr, r2, dp, static_cast<int>(d), scatter_out);
  PB_CUDA(cudaGetLastError());
  ctx.count_launch();
}

}  // namespace pb
