// Tensor-core contraction used by every dense step of the PLDA/LDA hot path:
//
//     C[M,N] = A[M,K] * B[N,K]^T          (both operands K-major, "NT" GEMM)
//
// with split-bf16 operands (hi + lo planes) and three tcgen05.mma per k-step
// (hi*hi + hi*lo + lo*hi, fp32 accumulation in TMEM): ~2^-17 relative error per
// product, which is what the parity bar (scores <= 1e-3, EER identical) needs --
// single-pass bf16/tf32 do not meet it.
//
// Replaces, behind the C ABI, the dense algebra Kaldi runs for the reference:
//   * enrol x test LLR grid      (Plda::LogLikelihoodRatio per pair, src/pldamodule.cpp:235,266)
//   * scatter SYRK               (PldaStats::AddSamples -> AddMat2, src/pldamodule.cpp:97)
//   * transform GEMM             (Plda::TransformIvector -> AddMatVec, src/pldamodule.cpp:171,224)
//   * EM projection + 2 SYRKs    (PldaEstimator::GetStatsFromClassMeans, src/pldamodule.cpp:106)
//   * LDA decision values        (python/liblda/lda.py:278)
//
// Kernel anatomy (persistent, warp-specialised, one CTA per SM, 384 threads = 3 warpgroups; CTA PAIRS with
// tcgen05 cta_group::2 -- pair tile 256 x BN x 64, each CTA owns 128 rows):
//   warp 0      TMA producer (both CTAs): 4 tile loads per k-block (own A_hi, A_lo; HALF of B_hi, B_lo), 128B
//               swizzle, 3-stage smem ring (64 KB/stage) guarded by full/empty mbarriers; the leader CTA's full
//               barrier collects the bytes of both CTAs; narrow boxes for a 16/32-column K tail
//   warp 1      TMEM allocator; in the leader CTA a single thread issues tcgen05.mma.cta_group::2 (M = 256);
//               multicast tcgen05.commit releases smem stages and publishes accumulators in both CTAs
//   warp 2      sharded B: watches the peers' ready flags (all ranks in parallel) for the TMA producer; else idle
//   warp 3      idle (warps 2-3 complete warpgroup 0, which donates registers via setmaxnreg)
//   warps 4..11 epilogue (two warps per TMEM lane quarter, interleaved 32-column chunks): all tcgen05.ld of a
//               tile in flight together -> accumulator stage released -> fused row/col/z-norm terms -> swizzled
//               smem staging -> TMA store (or fused row reductions, no store)
//   TMEM        2 accumulator stages x 256 fp32 columns (all 512 columns): the epilogue of tile i
//               overlaps the MMAs of tile i+1.
// A cta_group::1 instantiation (tile 128 x BN x 64, 2 x 96 KB stages) serves single-row-tile problems.
// Launch-level features: programmatic dependent launch behind the operand producer (the prologue above the
// griddepcontrol.wait overlaps the producer's tail); sharded B operand (GemmShard): the TMA producer polls the
// owner rank's ready flag before the first tile that touches its rows and the column tiles start at the rank's own.
#include <algorithm>
#include <cstdlib>

#include "runtime.h"

namespace pb {
namespace {

constexpr int BM = 128;
constexpr int BN_MAX = 256;
constexpr int BK = 64;  // bf16 elements: 128 bytes = one swizzle atom row
constexpr int STAGES = 2;
constexpr int UMMA_K = 16;
constexpr int A_TILE_BYTES = BM * BK * 2;                           // 16 KB
constexpr int B_TILE_BYTES = BN_MAX * BK * 2;                       // 32 KB
constexpr int STAGE_BYTES = 2 * A_TILE_BYTES + 2 * B_TILE_BYTES;    // 96 KB
constexpr int EPI_WARPS = 8;
constexpr int EPI_BUF_BYTES = 32 * 32 * 4;                          // one 32x32 fp32 box
constexpr int EPI_BYTES = EPI_WARPS * EPI_BUF_BYTES;                // 32 KB (one staging box per warp)
constexpr int COLC_BYTES = 2 * BN_MAX * 4;                          // column-term cache, one slot per accumulator stage
constexpr int BAR_BYTES = 256;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + COLC_BYTES + BAR_BYTES;
// warpgroup 0 = {TMA warp, MMA warp, 2 idle warps} gives its registers away (setmaxnreg.dec), warpgroups 1-2 =
// 8 epilogue warps take them (setmaxnreg.inc) so a whole 128-column accumulator slice fits in registers
constexpr int NUM_THREADS = 128 + 32 * EPI_WARPS;
constexpr int CTRL_REGS = 56;
constexpr int EPI_REGS = 224;
constexpr int TMEM_COLS = 512;

struct GemmParams {
  int m, n;
  int m_tiles, n_tiles, ksplit;
  int bn;             // tile width == UMMA N
  int nkb_total;      // 64-wide k-blocks in the whole reduction
  int kb_per_split;
  int last_ksteps;    // UMMA_K steps in the very last k-block (1..4)
  int group_m;
  int mpad;           // split-K: row pitch between partial planes
  int direct_store;   // 3 TMA store via swizzled smem staging (default), 0 coalesced LSU stores, 1 register stores, 2 none
  int tail_cols;      // K columns of the last k-block when it uses a narrow box (16 -> 32B swizzle, 32 -> 64B), else 0
  long long* dbg;     // optional stall counters of CTA 0/1 (env PLDA_B200_DBG=1): see plda_debug_counters
  int dbg_skip_a;     // timing experiment only (see Context::dbg_skip_a)
  int n_rot;          // column tiles are visited starting at this one (sharded B: the rank's own rows first)
  GemmEpilogue epi;
  GemmShard shard;    // flags == nullptr: B is complete before the launch
};

struct Work {
  int m_blk, n_blk, ks;
};

template <bool TWO>
struct Cfg {
  // cta_group::2: each CTA of the pair stages its own 128 rows of A and HALF of the B tile -> 64 KB per stage,
  // three stages; cta_group::1: full B tile per CTA -> 96 KB per stage, two stages.
  static constexpr int kStages = TWO ? 3 : 2;
  static constexpr int kBBytes = TWO ? B_TILE_BYTES / 2 : B_TILE_BYTES;
  static constexpr int kStageBytes = 2 * A_TILE_BYTES + 2 * kBBytes;
};
static_assert(Cfg<true>::kStages * Cfg<true>::kStageBytes == STAGES * STAGE_BYTES, "smem budget");

// Bounded poll of a peer's ready flag: a rank that never arrives is counted in *err (the host reports it) instead
// of hanging the GPU.  The error is sticky: once any wait of the session has timed out (2 s), later waits give up at
// once, so a broken exchange costs seconds, not seconds per tile.
__device__ __noinline__ void shard_wait(const unsigned* f, unsigned epoch, unsigned* err) {
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

template <bool TWO>
__device__ __forceinline__ Work decode(const GemmParams& p, int item, uint32_t rank) {
  // TWO: the unit of work is a PAIR of vertically adjacent 128-row tiles (one per CTA of the pair)
  const int m_units = TWO ? (p.m_tiles + 1) / 2 : p.m_tiles;
  Work w;
  const int tiles = m_units * p.n_tiles;
  w.ks = item / tiles;
  const int t = item - w.ks * tiles;
  const int per_group = p.group_m * p.n_tiles;
  const int g = t / per_group;
  const int first_m = g * p.group_m;
  const int gsz = min(p.group_m, m_units - first_m);
  const int r = t - g * per_group;
  const int unit = first_m + r % gsz;
  w.m_blk = TWO ? unit * 2 + static_cast<int>(rank) : unit;
  w.n_blk = r / gsz + p.n_rot;
  if (w.n_blk >= p.n_tiles) w.n_blk -= p.n_tiles;
  return w;
}

// FAST: the score-grid hot path only (TMA store; row term, uniform column term, z-norm affine) -- keeps the
// unrolled epilogue small enough for the instruction cache.  !FAST: every epilogue feature (per-row column
// groups, row moments, log-sum-exp partials, all store modes).
// LSE (EPI == 2): nothing stored, only the per-row online (max, sum exp) partials of the LDA log-softmax pass 1.
template <bool TWO, int EPI>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16x3_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                   const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                   const __grid_constant__ CUtensorMap tt_a_hi, const __grid_constant__ CUtensorMap tt_a_lo,
                   const __grid_constant__ CUtensorMap tt_b_hi, const __grid_constant__ CUtensorMap tt_b_lo,
                   const __grid_constant__ CUtensorMap tm_out, const GemmParams p) {
  constexpr bool RAG = EPI == 6;    // score-grid path with per-row column-term groups (ragged enrol counts)
  constexpr bool HYB = EPI == 7;    // score-grid path, every other chunk stored straight from registers (sector stores)
  constexpr bool FAST = EPI == 1 || RAG || HYB;
  constexpr bool LSE = EPI == 2;
  constexpr bool SECT = EPI == 3;   // score-grid path, accumulators stored straight from registers (no smem staging)
  constexpr bool MOM = EPI == 4;    // z-norm sink: per-row shifted moments of the scores, nothing stored
  constexpr bool HIST = EPI == 5;   // EER sink: target / tail non-target histograms of the scores, nothing stored
  constexpr int kStages = Cfg<TWO>::kStages;
  constexpr int kStageBytes = Cfg<TWO>::kStageBytes;
  constexpr int kBBytes = Cfg<TWO>::kBBytes;
  extern __shared__ __align__(1024) uint8_t smem[];   // 128B-swizzled tiles need 1024-byte alignment
  if ((smem_u32(smem) & 1023u) != 0) {
    if (threadIdx.x == 0) printf("plda_b200: dynamic shared memory is not 1024-byte aligned\n");
    __trap();
  }
  uint8_t* epi_base = smem + kStages * kStageBytes;
  float* colc = reinterpret_cast<float*>(epi_base + EPI_BYTES);
  uint64_t* full = reinterpret_cast<uint64_t*>(epi_base + EPI_BYTES + COLC_BYTES);
  uint64_t* empty = full + 3;
  uint64_t* tfull = empty + 3;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  // sharded B: per source rank, set by the flag-watcher warp once the rank's push of this epoch has been observed
  uint32_t* shard_seen = reinterpret_cast<uint32_t*>(epi_base + EPI_BYTES + COLC_BYTES + 128);   // [16] + "all seen"

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = TWO ? cluster_ctarank() : 0u;     // 0 = leader CTA of the pair (issues the MMAs)
  const int m_units = TWO ? (p.m_tiles + 1) / 2 : p.m_tiles;
  const int total_items = m_units * p.n_tiles * p.ksplit;
  const int first_item = TWO ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int item_stride = TWO ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const int bn_cta = TWO ? p.bn / 2 : p.bn;              // B rows staged by this CTA

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a_hi);
    tma_prefetch_desc(&tm_a_lo);
    tma_prefetch_desc(&tm_b_hi);
    tma_prefetch_desc(&tm_b_lo);
    if (p.tail_cols) {
      tma_prefetch_desc(&tt_a_hi);
      tma_prefetch_desc(&tt_a_lo);
      tma_prefetch_desc(&tt_b_hi);
      tma_prefetch_desc(&tt_b_lo);
    }
    if (p.direct_store == 3) tma_prefetch_desc(&tm_out);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) {
        // TWO: only the leader arms its full barrier (expect_tx covers the bytes of both CTAs); the peer's TMA
        // loads complete_tx on it remotely.  The peer cannot run a phase ahead: it refills a stage only after the
        // leader's multicast commit on empty[s], i.e. after the leader consumed the previous phase of full[s].
        // empty / tfull are signalled in both CTAs by the leader's multicast commits.
        mbar_init(&full[s], 1);
        mbar_init(&empty[s], 1);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(&tfull[a], 1);
        mbar_init(&tempty[a], TWO ? 2 * EPI_WARPS : EPI_WARPS);
      }
      mbar_fence_init();
    }
    if (lane < 17) shard_seen[lane] = 0u;
    __syncwarp();
    if (TWO) {
      tmem_alloc_2cta(tmem_slot, TMEM_COLS);
      tmem_relinquish_2cta();
    } else {
      tmem_alloc(tmem_slot, TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if (TWO) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // programmatic dependent launch: everything above overlapped the producer kernel's tail; its outputs (operands,
  // row / column terms) are visible after this wait.  No-op for a normal launch.
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(CTRL_REGS));
  if (warp == 0) {
    // ===================== TMA producer (both CTAs of a pair) ===================== //
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx_cta = 2 * A_TILE_BYTES + 2 * (bn_cta * BK * 2);
      // a short K tail (16 or 32 columns) is fetched with a narrow box: 1/4 or 1/2 of the bytes of a full k-block
      const uint32_t tx_tail_cta = 2 * (BM + bn_cta) * p.tail_cols * 2;
      long long dbg_prod_wait = 0;
      const long long t_start = clock64();
      uint32_t shard_ready = 0;
      uint32_t lead_full_addr[3] = {0, 0, 0};
      if (TWO) {
#pragma unroll
        for (int i = 0; i < 3; ++i) lead_full_addr[i] = mapa_shared(smem_u32(&full[i]), 0);
      }
      for (int item = first_item; item < total_items; item += item_stride) {
        const Work w = decode<TWO>(p, item, rank);
        const int kb0 = w.ks * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.nkb_total);
        const int brow = w.n_blk * p.bn + (TWO ? static_cast<int>(rank) * bn_cta : 0);
        if (p.shard.flags != nullptr) {
          // sharded B: rows of rank r are valid once flags[r] reaches this step's epoch
          const int b_lo = brow, b_hi = min(brow + bn_cta, p.n);
          bool waited = false;
          for (int r = 0; r < p.shard.world; ++r) {
            if ((shard_ready >> r) & 1u) continue;
            if (p.shard.bounds[r] >= b_hi || p.shard.bounds[r + 1] <= b_lo) continue;
            // the flag-watcher warp polls every rank's flag concurrently (system-scope acquire loads cost ~1 us each;
            // taken one after the other on this thread they stalled the operand ring once per source rank)
            uint32_t seen;
            do {
              asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(seen) : "r"(smem_u32(shard_seen + r)) : "memory");
            } while (seen == 0u);
            shard_ready |= 1u << r;
            waited = true;
          }
          // the peers wrote through the generic proxy; the TMA loads below read through the async proxy
          if (waited) asm volatile("fence.proxy.async;" ::: "memory");
        }
        for (int kb = kb0; kb < kb1; ++kb) {
          const long long t_w0 = clock64();
          mbar_wait(&empty[stage], phase ^ 1);
          dbg_prod_wait += clock64() - t_w0;
          uint8_t* s = smem + stage * kStageBytes;
          const bool tail = p.tail_cols != 0 && kb == p.nkb_total - 1;
          const bool skip_a = p.dbg_skip_a != 0 && item != first_item;
          uint32_t tx = tail ? tx_tail_cta : tx_cta;
          if (skip_a) tx -= tail ? 2u * BM * p.tail_cols * 2u : 2u * A_TILE_BYTES;
          const CUtensorMap* ma_hi = tail ? &tt_a_hi : &tm_a_hi;
          const CUtensorMap* ma_lo = tail ? &tt_a_lo : &tm_a_lo;
          const CUtensorMap* mb_hi = tail ? &tt_b_hi : &tm_b_hi;
          const CUtensorMap* mb_lo = tail ? &tt_b_lo : &tm_b_lo;
          if (TWO) {
            const uint32_t lead_full = stage == 0 ? lead_full_addr[0] : (stage == 1 ? lead_full_addr[1] : lead_full_addr[2]);
            if (rank == 0) mbar_arrive_expect_tx(&full[stage], 2 * tx);
            if (!skip_a) {
              tma_load_2d_2sm(s, ma_hi, lead_full, kb * BK, w.m_blk * BM);
              tma_load_2d_2sm(s + A_TILE_BYTES, ma_lo, lead_full, kb * BK, w.m_blk * BM);
            }
            tma_load_2d_2sm(s + 2 * A_TILE_BYTES, mb_hi, lead_full, kb * BK, brow);
            tma_load_2d_2sm(s + 2 * A_TILE_BYTES + kBBytes, mb_lo, lead_full, kb * BK, brow);
          } else {
            mbar_arrive_expect_tx(&full[stage], tx);
            tma_load_2d(s, ma_hi, &full[stage], kb * BK, w.m_blk * BM);
            tma_load_2d(s + A_TILE_BYTES, ma_lo, &full[stage], kb * BK, w.m_blk * BM);
            tma_load_2d(s + 2 * A_TILE_BYTES, mb_hi, &full[stage], kb * BK, brow);
            tma_load_2d(s + 2 * A_TILE_BYTES + kBBytes, mb_lo, &full[stage], kb * BK, brow);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
      if (p.dbg != nullptr && blockIdx.x < 2) {
        p.dbg[blockIdx.x * 16 + 0] = dbg_prod_wait;
        p.dbg[blockIdx.x * 16 + 1] = clock64() - t_start;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only when paired) ===================== //
    if (lane == 0 && rank == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint32_t idesc = umma_idesc_bf16_f32(TWO ? 2 * BM : BM, p.bn);
      long long dbg_full = 0, dbg_tempty = 0, dbg_tiles = 0;
      const long long t_start = clock64();
      for (int item = first_item; item < total_items; item += item_stride) {
        const Work w = decode<TWO>(p, item, rank);
        const int kb0 = w.ks * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.nkb_total);
        long long t_w0 = clock64();
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        dbg_tempty += clock64() - t_w0;
        ++dbg_tiles;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN_MAX;
        for (int kb = kb0; kb < kb1; ++kb) {
          t_w0 = clock64();
          mbar_wait(&full[stage], phase);
          dbg_full += clock64() - t_w0;
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * kStageBytes);
          const bool last_kb = kb == p.nkb_total - 1;
          // tile rows are (2 * tail_cols) bytes apart when the narrow tail box was used, 128 B otherwise
          const uint32_t swz = (last_kb && p.tail_cols != 0) ? 2u * p.tail_cols : 128u;
          const uint64_t a_hi = umma_desc_kmajor(sa, swz);
          const uint64_t a_lo = umma_desc_kmajor(sa + A_TILE_BYTES, swz);
          const uint64_t b_hi = umma_desc_kmajor(sa + 2 * A_TILE_BYTES, swz);
          const uint64_t b_lo = umma_desc_kmajor(sa + 2 * A_TILE_BYTES + kBBytes, swz);
          const int nks = last_kb ? p.last_ksteps : (BK / UMMA_K);
          for (int ks = 0; ks < nks; ++ks) {
            // advance 32 B (= UMMA_K bf16) inside the swizzle atom: +2 in the >>4 address field
            const uint64_t off = static_cast<uint64_t>(ks * 2);
            const uint32_t first = (kb == kb0 && ks == 0) ? 0u : 1u;
            if (TWO) {
              umma_bf16_ss_2cta(d_tmem, a_hi + off, b_hi + off, idesc, first);
              umma_bf16_ss_2cta(d_tmem, a_hi + off, b_lo + off, idesc, 1u);
              umma_bf16_ss_2cta(d_tmem, a_lo + off, b_hi + off, idesc, 1u);
            } else {
              umma_bf16_ss(d_tmem, a_hi + off, b_hi + off, idesc, first);
              umma_bf16_ss(d_tmem, a_hi + off, b_lo + off, idesc, 1u);
              umma_bf16_ss(d_tmem, a_lo + off, b_hi + off, idesc, 1u);
            }
          }
          // smem stage reusable (in both CTAs) once these MMAs retire
          if (TWO) umma_commit_2cta(&empty[stage], 3); else umma_commit(&empty[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        // accumulator complete -> epilogue warps (of both CTAs)
        if (TWO) umma_commit_2cta(&tfull[acc], 3); else umma_commit(&tfull[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      if (p.dbg != nullptr && blockIdx.x < 2) {
        p.dbg[blockIdx.x * 16 + 2] = dbg_full;
        p.dbg[blockIdx.x * 16 + 3] = dbg_tempty;
        p.dbg[blockIdx.x * 16 + 4] = clock64() - t_start;
        p.dbg[blockIdx.x * 16 + 5] = dbg_tiles;
      }
    }
  } else if (warp == 2) {
    // ===================== flag watcher (sharded B only) ===================== //
    // lane r waits (bounded, see shard_wait) for source rank r's push of this epoch -- all ranks in parallel -- and
    // publishes it to the TMA producer through shared memory: peer stores -> system fence -> flag (producer kernel)
    // -> acquire.sys here -> release.cta -> acquire.cta + fence.proxy.async in the TMA thread -> TMA loads.
    if (p.shard.flags != nullptr) {
      if (lane < p.shard.world) {
        shard_wait(p.shard.flags + lane, p.shard.epoch, p.shard.err);
        asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(shard_seen + lane)), "r"(1u) : "memory");
      }
      __syncwarp();
      // every rank has pushed: from here on the epilogue may fetch its column terms ahead of the accumulator
      if (lane == 0)
        asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(shard_seen + 16)), "r"(1u) : "memory");
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(EPI_REGS));
    // ===================== epilogue (warps 4..11) ===================== //
    // Two warps per TMEM lane quarter: warp handles the 32-column chunks c with c % 2 == h.
    const int ew = warp - 4;
    const int q = warp & 3;   // TMEM lane quarter this warp is allowed to read
    const int h = ew >> 2;
    uint8_t* sb = epi_base + ew * EPI_BUF_BYTES;
    const uint32_t sb_row = smem_u32(sb) + lane * 128;
    const GemmEpilogue& e = p.epi;
    int acc = 0;
    uint32_t acc_phase = 0;
    long long dbg_tfull = 0, dbg_store = 0, dbg_load = 0;
    const long long t_start = clock64();
    for (int item = first_item; item < total_items; item += item_stride) {
      const Work w = decode<TWO>(p, item, rank);
      const int m0 = w.m_blk * BM;
      const int n0 = w.n_blk * p.bn;
      const int m = m0 + q * 32 + lane;
      const bool mvalid = m < p.m;
      const bool warp_rows_valid = (m0 + q * 32) < p.m;
      float ra = 0.f, zm = 0.f, zi = 1.f;
      int g = 0;
      if (mvalid) {
        if (e.row_add) ra = __ldg(e.row_add + m);
        if (e.grp) g = __ldg(e.grp + m);
        if (e.zmean) { zm = __ldg(e.zmean + m); zi = __ldg(e.zinv + m); }
      }
      const float radd = ra - zm;
      int row_spk = 0;
      if (HIST && mvalid) row_spk = __ldg(e.row_spk + m);
      float rs = 0.f, rq = 0.f;
      float lmax = -INFINITY, lsum = 0.f;
      const int ncols = min(p.bn, p.n - n0);
      const int nchunks = (ncols + 31) >> 5;
      const long long out_row = static_cast<long long>(w.ks) * p.mpad + m;
      const float* colp = e.col_add ? e.col_add + static_cast<long long>(g) * e.col_ld + n0 : nullptr;
      // uniform column terms (no per-row groups): fetch this warp's chunks now (4 coalesced loads in flight
      // while the MMAs finish), park them in the per-stage smem cache once the accumulator is ready
      const bool col_cached = e.col_add != nullptr && e.grp == nullptr;
      // sharded B: the column terms travel with the operand rows, so they are only valid once the accumulator is
      // (producer saw the owner's flag -> TMA -> MMA -> tfull): fetched after the tfull wait, L1 bypassed
      bool col_late = p.shard.flags != nullptr;
      if (col_late) {
        // ... unless the flag watcher has already seen every rank's push: then they are as final as on one GPU
        uint32_t all_seen;
        asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(all_seen) : "r"(smem_u32(shard_seen + 16)) : "memory");
        if (all_seen != 0u) col_late = false;
      }
      float cpre[4] = {0.f, 0.f, 0.f, 0.f};
      if (col_cached && !col_late) {
        // (sharded: this SM has not read these words earlier in the launch -- the late path bypasses L1 -- and L1 is
        // invalidated at every launch boundary, so the read-only path cannot return a previous generation)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int cc = h + 2 * i;
          if (cc < nchunks) cpre[i] = __ldg(e.col_add + n0 + cc * 32 + lane);
        }
      }
      float* colslot = colc + acc * BN_MAX;
      // ragged counts: the column terms of a row come from its group's vector (rows of one warp may differ)
      float cbuf[RAG ? 32 : 1];
      auto load_cols = [&](int cc, float* dst) {
        const float4* cp = reinterpret_cast<const float4*>(colp + cc * 32);
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 t = __ldg(cp + j4);
          dst[4 * j4 + 0] = t.x; dst[4 * j4 + 1] = t.y; dst[4 * j4 + 2] = t.z; dst[4 * j4 + 3] = t.w;
        }
      };
      if (RAG && h < nchunks) load_cols(h, cbuf);
      // The accumulator stage is handed back BEFORE process() reads the column cache, so tfull of tile i+2 alone
      // does not prove that the three sibling warps (same chunk parity h, other lane quarters) are done reading
      // this slot for tile i: the four warps that share a slot half meet here once per tile (after this point
      // every one of them has finished tile i-1, hence its reads of tile i-2's slot).
      if (col_cached) asm volatile("bar.sync %0, 128;" ::"r"(1 + h) : "memory");
      const uint32_t tbase = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN_MAX;

      auto process = [&](uint32_t (&r)[32], int c, float* cadd, int next_c) {
        const int nbase = n0 + c * 32;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (col_cached) {
          const float4* cp = reinterpret_cast<const float4*>(colslot + c * 32);   // smem broadcast reads
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 t = cp[j4];
            v[4 * j4 + 0] += t.x; v[4 * j4 + 1] += t.y; v[4 * j4 + 2] += t.z; v[4 * j4 + 3] += t.w;
          }
        } else if (RAG) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += cadd[j];
          // the next chunk's terms travel while this chunk is staged and stored
          if (next_c >= 0) load_cols(next_c, cadd);
        } else if (!FAST && !LSE && colp != nullptr && mvalid) {
          const float4* cp = reinterpret_cast<const float4*>(colp + c * 32);
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 t = __ldg(cp + j4);
            v[4 * j4 + 0] += t.x; v[4 * j4 + 1] += t.y; v[4 * j4 + 2] += t.z; v[4 * j4 + 3] += t.w;
          }
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = (v[j] + radd) * zi;
        if (!FAST && !LSE && e.rsum != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (nbase + j < p.n) { rs += v[j]; rq += v[j] * v[j]; }
          }
        }
        if (LSE || (!FAST && e.lse_max != nullptr)) {
          float cmax = -INFINITY;
#pragma unroll
          for (int j = 0; j < 32; ++j) if (nbase + j < p.n) cmax = fmaxf(cmax, v[j]);
          const float nmax = fmaxf(lmax, cmax);
          float add = 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) if (nbase + j < p.n) add += __expf(v[j] - nmax);
          lsum = lsum * __expf(lmax - nmax) + add;
          lmax = nmax;
        }
        if (FAST) {
          if (warp_rows_valid) {
            if (lane == 0) tma_store_wait_read<0>();   // the previous TMA store of this warp has drained the buffer
            __syncwarp();
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const uint32_t addr = sb_row + ((j4 ^ (lane & 7)) << 4);   // 16-byte chunk ^= row % 8 (128B swizzle)
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v[4 * j4 + 0]),
                           "f"(v[4 * j4 + 1]), "f"(v[4 * j4 + 2]), "f"(v[4 * j4 + 3])
                           : "memory");
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tm_out, sb, nbase, w.ks * p.mpad + m0 + q * 32);
              tma_store_commit();
            }
          }
        } else if (!LSE && e.out != nullptr) {
          if (p.direct_store == 2) {
            // debug (PLDA_B200_EPI=skip): no store at all -> isolates the TMA/MMA main loop
          } else if (p.direct_store == 1) {
            if (mvalid) {
              float* op = e.out + out_row * e.ldo + nbase;
#pragma unroll
              for (int j = 0; j < 32; ++j) if (nbase + j < p.n) op[j] = v[j];
            }
          } else if (warp_rows_valid) {
            if (p.direct_store == 3) {
              const long long t_s0 = clock64();
              if (lane == 0) tma_store_wait_read<0>();   // the previous TMA store of this warp has drained the buffer
              dbg_store += clock64() - t_s0;
            }
            __syncwarp();
            // transpose through a 128B-swizzled 32x32 staging box: thread = row on the way in ...
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const uint32_t addr = sb_row + ((j4 ^ (lane & 7)) << 4);   // 16-byte chunk ^= row % 8
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v[4 * j4 + 0]),
                           "f"(v[4 * j4 + 1]), "f"(v[4 * j4 + 2]), "f"(v[4 * j4 + 3])
                           : "memory");
            }
            if (p.direct_store == 3) {
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                tma_store_2d(&tm_out, sb, nbase, w.ks * p.mpad + m0 + q * 32);
                tma_store_commit();
              }
            } else {
              // ... and 8 lanes per row on the way out: every warp store writes four full 128-byte rows
              __syncwarp();
              const int ch = lane & 7;
              const int col = nbase + ch * 4;
              const long long row_base = static_cast<long long>(w.ks) * p.mpad + m0 + q * 32;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int rr = i * 4 + (lane >> 3);
                float4 t;
                const uint32_t addr = smem_u32(sb) + rr * 128 + ((ch ^ (rr & 7)) << 4);
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w)
                             : "r"(addr));
                if (m0 + q * 32 + rr < p.m) {
                  float* op = e.out + (row_base + rr) * e.ldo + col;
                  if (col + 3 < p.n) {
                    *reinterpret_cast<float4*>(op) = t;
                  } else {
                    if (col < p.n) op[0] = t.x;
                    if (col + 1 < p.n) op[1] = t.y;
                    if (col + 2 < p.n) op[2] = t.z;
                  }
                }
              }
            }
          }
        }
      };

      if constexpr (HIST) {
        // speaker ids of this warp's columns, parked in the warp's own (otherwise unused) staging box
        int* sidx = reinterpret_cast<int*>(sb);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int cc = h + 2 * i;
          const int col = n0 + cc * 32 + lane;
          sidx[i * 32 + lane] = (cc < nchunks && col < p.n) ? __ldg(e.col_spk + col) : 0;
        }
        __syncwarp();
      }
      const long long t_w0 = clock64();
      mbar_wait(&tfull[acc], acc_phase);
      dbg_tfull += clock64() - t_w0;
      tc_fence_after();
      if (col_cached && col_late) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int cc = h + 2 * i;
          if (cc < nchunks) cpre[i] = __ldcg(e.col_add + n0 + cc * 32 + lane);
        }
      }
      if (col_cached) {
        // slot free (named barrier above); warps sharing chunks write identical values
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int cc = h + 2 * i;
          if (cc < nchunks) colslot[cc * 32 + lane] = cpre[i];
        }
        __syncwarp();
      }
      if constexpr (SECT) {
        // Register-direct epilogue: the 16x256b TMEM load shape hands four threads 8 consecutive columns of one row,
        // so every 8-byte store of a warp completes 32-byte sectors (8 rows x 32 B per instruction) -- no shared
        // memory staging, no TMA store: the shared-memory port is left to the operand fill and the MMA reads.
        uint32_t rr[4][2][16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int cc = h + 2 * i;
          if (cc < nchunks) {
            tmem_ld_16x256b_x4(tbase + cc * 32, rr[i][0]);
            tmem_ld_16x256b_x4(tbase + (16u << 16) + cc * 32, rr[i][1]);
          }
        }
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (TWO) mbar_arrive_cluster(mapa_shared(smem_u32(&tempty[acc]), 0));
          else mbar_arrive(&tempty[acc]);
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        // the four rows of this thread: lane/4 + {0, 8, 16, 24} inside the warp's 32-row quarter
        float radd4[4], zi4[4];
        float* orow[4];
        bool rok[4];
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4) {
          const int mr = m0 + q * 32 + s4 * 8 + (lane >> 2);
          rok[s4] = mr < p.m;
          float ra4 = 0.f, zm4 = 0.f;
          zi4[s4] = 1.f;
          if (rok[s4]) {
            if (e.row_add) ra4 = __ldg(e.row_add + mr);
            if (e.zmean) { zm4 = __ldg(e.zmean + mr); zi4[s4] = __ldg(e.zinv + mr); }
          }
          radd4[s4] = ra4 - zm4;
          orow[s4] = e.out + (static_cast<long long>(w.ks) * p.mpad + mr) * e.ldo + n0 + 2 * (lane & 3);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int cc = h + 2 * i;
          if (cc >= nchunks) continue;
#pragma unroll
          for (int rep = 0; rep < 4; ++rep) {
            const int cl = cc * 32 + rep * 8 + 2 * (lane & 3);     // column inside the tile
            float2 ct = make_float2(0.f, 0.f);
            if (col_cached) ct = *reinterpret_cast<const float2*>(colslot + cl);
#pragma unroll
            for (int s4 = 0; s4 < 4; ++s4) {
              const int h2 = s4 >> 1, half = s4 & 1;
              float2 v;
              v.x = (__uint_as_float(rr[i][h2][rep * 4 + half * 2 + 0]) + ct.x + radd4[s4]) * zi4[s4];
              v.y = (__uint_as_float(rr[i][h2][rep * 4 + half * 2 + 1]) + ct.y + radd4[s4]) * zi4[s4];
              if (rok[s4]) {
                const int col = n0 + cl;
                if (col + 1 < p.n) *reinterpret_cast<float2*>(orow[s4] + cc * 32 + rep * 8) = v;
                else if (col < p.n) orow[s4][cc * 32 + rep * 8] = v.x;
              }
            }
          }
        }
        continue;
      }
      if constexpr (HYB) {
        // Hybrid store path: the warp's 1st and 3rd chunk go through the swizzled staging box and a TMA store, the
        // 2nd and 4th are read in the 16x256b shape (four threads hold 8 consecutive columns of a row) and stored
        // straight from registers as full 32-byte sectors.  The shared-memory port -- the measured bound of this
        // kernel at d = 200 -- carries half of the epilogue traffic; the LSU / L2 request path, idle in the TMA
        // variant, carries the other half.
        uint32_t ra[2][32];
        uint32_t rs[2][2][16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int cc = h + 2 * i;
          if (cc >= nchunks) continue;
          if ((i & 1) == 0) {
            tmem_ld_32x32b_x32(tbase + cc * 32, ra[i >> 1]);
          } else {
            tmem_ld_16x256b_x4(tbase + cc * 32, rs[i >> 1][0]);
            tmem_ld_16x256b_x4(tbase + (16u << 16) + cc * 32, rs[i >> 1][1]);
          }
        }
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (TWO) mbar_arrive_cluster(mapa_shared(smem_u32(&tempty[acc]), 0));
          else mbar_arrive(&tempty[acc]);
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        // the four rows of this thread in the sector shape: lane/4 + {0, 8, 16, 24} inside the warp's 32-row quarter
        float radd4[4], zi4[4];
        float* orow[4];
        bool rok[4];
#pragma unroll
        for (int s4 = 0; s4 < 4; ++s4) {
          const int mr = m0 + q * 32 + s4 * 8 + (lane >> 2);
          rok[s4] = mr < p.m;
          float ra4 = 0.f, zm4 = 0.f;
          zi4[s4] = 1.f;
          if (rok[s4]) {
            if (e.row_add) ra4 = __ldg(e.row_add + mr);
            if (e.zmean) { zm4 = __ldg(e.zmean + mr); zi4[s4] = __ldg(e.zinv + mr); }
          }
          radd4[s4] = ra4 - zm4;
          orow[s4] = e.out + (static_cast<long long>(w.ks) * p.mpad + mr) * e.ldo + n0 + 2 * (lane & 3);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int cc = h + 2 * i;
          if (cc >= nchunks) continue;
          if ((i & 1) == 0) {
            process(ra[i >> 1], cc, nullptr, -1);
          } else {
#pragma unroll
            for (int rep = 0; rep < 4; ++rep) {
              const int cl = cc * 32 + rep * 8 + 2 * (lane & 3);     // column inside the tile
              float2 ct = make_float2(0.f, 0.f);
              if (col_cached) ct = *reinterpret_cast<const float2*>(colslot + cl);
#pragma unroll
              for (int s4 = 0; s4 < 4; ++s4) {
                const int h2 = s4 >> 1, half = s4 & 1;
                float2 v;
                v.x = (__uint_as_float(rs[i >> 1][h2][rep * 4 + half * 2 + 0]) + ct.x + radd4[s4]) * zi4[s4];
                v.y = (__uint_as_float(rs[i >> 1][h2][rep * 4 + half * 2 + 1]) + ct.y + radd4[s4]) * zi4[s4];
                if (rok[s4]) {
                  const int col = n0 + cl;
                  if (col + 1 < p.n) *reinterpret_cast<float2*>(orow[s4] + cc * 32 + rep * 8) = v;
                  else if (col < p.n) orow[s4][cc * 32 + rep * 8] = v.x;
                }
              }
            }
          }
        }
        continue;
      }
      // all of this warp's chunks are fetched with the TMEM loads in flight together (one ~1k-cycle latency per
      // tile instead of one per chunk), then the accumulator stage is handed back BEFORE the post-processing
      // and the stores, so the MMA warp never waits for the store path
      uint32_t r[4][32];
      const long long t_l0 = clock64();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int cc = h + 2 * i;
        if (cc < nchunks) tmem_ld_32x32b_x32(tbase + cc * 32, r[i]);
      }
      tmem_ld_wait();
      dbg_load += clock64() - t_l0;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (TWO) mbar_arrive_cluster(mapa_shared(smem_u32(&tempty[acc]), 0));   // the leader CTA's MMA warp waits for both
        else mbar_arrive(&tempty[acc]);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      if constexpr (MOM) {
        // Shifted moments of this thread's row over the warp's chunks: pivot = the slot's first score, so the
        // fp32 sums carry deviations of the size of the cohort spread, not of the score itself (no cancellation
        // when |mean| >> std); the slots of a row are merged in fp64 (moments_reduce, Chan's update).
        float s1 = 0.f, s2 = 0.f, piv = 0.f;
        int cnt = 0;
        if (h < nchunks) {
          piv = (__uint_as_float(r[0][0]) + (col_cached ? colslot[h * 32] : 0.f) + radd) * zi;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int cc = h + 2 * i;
            if (cc >= nchunks) continue;
            const int nbase = n0 + cc * 32;
            const float4* cp = reinterpret_cast<const float4*>(colslot + cc * 32);
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
              if (col_cached) t = cp[j4];
              const float tv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
              for (int jj = 0; jj < 4; ++jj) {
                const int j = 4 * j4 + jj;
                if (nbase + j < p.n) {
                  const float dv = (__uint_as_float(r[i][j]) + tv[jj] + radd) * zi - piv;
                  s1 += dv;
                  s2 = fmaf(dv, dv, s2);
                  ++cnt;
                }
              }
            }
          }
        }
        if (mvalid)
          e.mom[(static_cast<long long>(m) * p.n_tiles + w.n_blk) * 2 + h] = make_float4(s1, s2, piv, static_cast<float>(cnt));
        continue;
      }
      if constexpr (HIST) {
        const int* sidx = reinterpret_cast<const int*>(sb);
        unsigned below = 0;
        const int top = e.nbins - 1;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int cc = h + 2 * i;
          if (cc >= nchunks) continue;
          const int nbase = n0 + cc * 32;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (!mvalid || nbase + j >= p.n) continue;
            const float v = (__uint_as_float(r[i][j]) + (col_cached ? colslot[cc * 32 + j] : 0.f) + radd) * zi;
            const bool tgt = sidx[i * 32 + j] == row_spk;
            if (tgt || v >= e.theta_lo) {
              int b = __float2int_rd((v - e.hist_lo) * e.hist_scale);
              b = b < 0 ? 0 : (b > top ? top : b);
              atomicAdd((tgt ? e.hist_t : e.hist_n) + b, 1ull);
            } else {
              ++below;
            }
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) below += __shfl_xor_sync(0xffffffffu, below, o);
        if (lane == 0 && below != 0) atomicAdd(e.below, static_cast<unsigned long long>(below));
        __syncwarp();     // the next item's speaker ids overwrite the staging box
        continue;
      }
      if constexpr (RAG) {
        // the thread's own group vector, one chunk ahead of the chunk being stored (first chunk fetched above,
        // behind the accumulator wait)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int cc = h + 2 * i;
          if (cc < nchunks) process(r[i], cc, cbuf, (i + 1 < 4 && cc + 2 < nchunks) ? cc + 2 : -1);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int cc = h + 2 * i;
          if (cc < nchunks) process(r[i], cc, nullptr, -1);
        }
      }

      if (!FAST && mvalid) {
        if (!LSE && e.rsum != nullptr) {
          atomicAdd(e.rsum + m, static_cast<double>(rs));
          atomicAdd(e.rsq + m, static_cast<double>(rq));
        }
        if (e.lse_max != nullptr) {
          const long long slot = (static_cast<long long>(m) * p.n_tiles + w.n_blk) * 2 + h;
          e.lse_max[slot] = lmax;
          e.lse_sum[slot] = lsum;
        }
      }
    }
    if (!SECT && (FAST || p.direct_store == 3) && lane == 0) tma_store_wait<0>();
    if (p.dbg != nullptr && blockIdx.x < 2 && lane == 0 && (ew == 0 || ew == 7)) {
      const int o = blockIdx.x * 16 + (ew == 0 ? 6 : 9);
      p.dbg[o + 0] = dbg_tfull;
      p.dbg[o + 1] = dbg_store;
      p.dbg[o + 2] = clock64() - t_start;
      if (ew == 0) p.dbg[blockIdx.x * 16 + 12] = dbg_load;
    }
  }

  tc_fence_before();
  if (TWO) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  if (warp == 1) {
    __syncwarp();
    if (TWO) tmem_dealloc_2cta(tmem_base, TMEM_COLS); else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------- //
// split-K partial reduction in fp64
// ------------------------------------------------------------------------- //
__global__ void reduce_partials_kernel(const float* __restrict__ partial, int ksplit, int m, int n, int mpad,
                                       int npad, double* __restrict__ out, long long ldo, double alpha,
                                       int symmetrise) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<long long>(m) * n) return;
  const int i = static_cast<int>(idx / n);
  const int j = static_cast<int>(idx % n);
  double s = 0.0;
  for (int k = 0; k < ksplit; ++k) s += static_cast<double>(partial[(static_cast<long long>(k) * mpad + i) * npad + j]);
  if (symmetrise) {
    double t = 0.0;
    for (int k = 0; k < ksplit; ++k) t += static_cast<double>(partial[(static_cast<long long>(k) * mpad + j) * npad + i]);
    s = 0.5 * (s + t);
  }
  out[i * ldo + j] = alpha * s;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  PB_CHECK(fn != nullptr, kCudaError, "cuTensorMapEncodeTiled not available from the driver");
  return fn;
}

struct Plan {
  GemmParams p;
  int grid;
  bool two_cta;
};

Plan make_plan(const Context& ctx, int64_t m, int64_t n, int64_t k, int ksplit) {
  PB_CHECK(m > 0 && n > 0 && k > 0, kInvalidArg, "gemm: empty problem");
  PB_CHECK(m < (1ll << 31) && n < (1ll << 31) && k < (1ll << 31), kInvalidArg, "gemm: dimension too large");
  Plan pl{};
  GemmParams& p = pl.p;
  p.m = static_cast<int>(m);
  p.n = static_cast<int>(n);
  p.bn = n >= BN_MAX ? BN_MAX : static_cast<int>(round_up(n, 16));
  p.m_tiles = static_cast<int>(ceil_div(m, BM));
  p.n_tiles = static_cast<int>(ceil_div(n, p.bn));
  const int64_t k16 = round_up(k, 16);
  p.nkb_total = static_cast<int>(ceil_div(k16, BK));
  p.last_ksteps = static_cast<int>((k16 - static_cast<int64_t>(BK) * (p.nkb_total - 1)) / UMMA_K);
  if (ksplit < 1) ksplit = 1;
  if (ksplit > p.nkb_total) ksplit = p.nkb_total;
  p.kb_per_split = static_cast<int>(ceil_div(p.nkb_total, ksplit));
  p.ksplit = static_cast<int>(ceil_div(p.nkb_total, p.kb_per_split));
  p.group_m = 16;
  p.mpad = p.m_tiles * BM;
  const long long items = static_cast<long long>(p.m_tiles) * p.n_tiles * p.ksplit;
  PB_CHECK(items < (1ll << 31), kInvalidArg, "gemm: too many tiles");
  pl.grid = static_cast<int>(items < ctx.num_sms ? items : ctx.num_sms);
  // CTA pairs (cta_group::2): worth it once there are at least two row tiles to pair up
  pl.two_cta = ctx.gemm_two_cta && p.m_tiles >= 2;
  if (pl.two_cta) {
    const long long pair_items = static_cast<long long>((p.m_tiles + 1) / 2) * p.n_tiles * p.ksplit;
    const long long clusters = std::min<long long>(pair_items, ctx.num_sms / 2);
    pl.grid = static_cast<int>(2 * clusters);
  }
  return pl;
}

struct TmapSet {
  CUtensorMap a_hi, a_lo, b_hi, b_lo, ta_hi, ta_lo, tb_hi, tb_lo, out;
};

template <bool TWO, int EPI>
void launch_one(Context& ctx, const Plan& pl, const TmapSet& tm, bool pdl) {
  static std::once_flag attr_once;
  std::call_once(attr_once, [] {
    cudaFuncSetAttribute(gemm_bf16x3_kernel<TWO, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  });
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(pl.grid);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = ctx.stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (TWO) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl && TWO) {     // the single-CTA kernel keeps plain stream order (as before)
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  PB_CUDA(cudaLaunchKernelEx(&cfg, gemm_bf16x3_kernel<TWO, EPI>, tm.a_hi, tm.a_lo, tm.b_hi, tm.b_lo, tm.ta_hi, tm.ta_lo,
                             tm.tb_hi, tm.tb_lo, tm.out, pl.p));
}

template <int EPI>
void launch_epi(Context& ctx, const Plan& pl, const TmapSet& tm, bool pdl) {
  if (pl.two_cta) launch_one<true, EPI>(ctx, pl, tm, pdl);
  else launch_one<false, EPI>(ctx, pl, tm, pdl);
}

void launch(Context& ctx, const SplitOperand& a, const SplitOperand& b, Plan& pl, float* out, int64_t ldo,
            int64_t out_rows) {
  GemmParams& p = pl.p;
  PB_CHECK(a.ld % 8 == 0 && b.ld % 8 == 0, kInvalidArg, "gemm: operand row pitch must be a multiple of 8 bf16");
  PB_CHECK((reinterpret_cast<uintptr_t>(a.hi) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.lo) & 15) == 0 &&
               (reinterpret_cast<uintptr_t>(b.hi) & 15) == 0 && (reinterpret_cast<uintptr_t>(b.lo) & 15) == 0,
           kInvalidArg, "gemm: operands must be 16-byte aligned");
  const uint64_t kext = static_cast<uint64_t>(p.nkb_total - 1) * BK + static_cast<uint64_t>(p.last_ksteps) * UMMA_K;
  PB_CHECK(static_cast<int64_t>(kext) <= a.ld && static_cast<int64_t>(kext) <= b.ld, kInvalidArg,
           "gemm: operand pitch smaller than round_up(k,16)");
  CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo, tta_hi, tta_lo, ttb_hi, ttb_lo, tout;
  encode_tmap_2d(&ta_hi, TmaType::BF16, a.hi, kext, a.rows, a.ld * 2, BK, BM, 128);
  encode_tmap_2d(&ta_lo, TmaType::BF16, a.lo, kext, a.rows, a.ld * 2, BK, BM, 128);
  const uint32_t b_box = pl.two_cta ? p.bn / 2 : p.bn;      // a CTA pair stages half of the B tile per CTA
  encode_tmap_2d(&tb_hi, TmaType::BF16, b.hi, kext, b.rows, b.ld * 2, BK, b_box, 128);
  encode_tmap_2d(&tb_lo, TmaType::BF16, b.lo, kext, b.rows, b.ld * 2, BK, b_box, 128);
  // K tail of 16 / 32 columns: narrow boxes (32 B / 64 B swizzle) instead of a zero-padded 64-column block
  const int tail = p.last_ksteps * UMMA_K;
  p.tail_cols = (p.nkb_total > 1 && (tail == 16 || tail == 32) && ctx.k_tail_boxes) ? tail : 0;
  if (p.tail_cols) {
    encode_tmap_2d(&tta_hi, TmaType::BF16, a.hi, kext, a.rows, a.ld * 2, tail, BM, tail * 2);
    encode_tmap_2d(&tta_lo, TmaType::BF16, a.lo, kext, a.rows, a.ld * 2, tail, BM, tail * 2);
    encode_tmap_2d(&ttb_hi, TmaType::BF16, b.hi, kext, b.rows, b.ld * 2, tail, b_box, tail * 2);
    encode_tmap_2d(&ttb_lo, TmaType::BF16, b.lo, kext, b.rows, b.ld * 2, tail, b_box, tail * 2);
  } else {
    tta_hi = ta_hi; tta_lo = ta_lo; ttb_hi = tb_hi; ttb_lo = tb_lo;
  }
  const bool aligned_out = out != nullptr && (ldo % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  // 3: TMA store from the swizzled staging box (default); 0: coalesced LSU stores from the same box
  // (PLDA_B200_EPI=lsu); 1: register stores (unaligned outputs / PLDA_B200_EPI=direct); 2: no store (debug)
  p.direct_store = ctx.epi_mode == 2 ? 2 : ((ctx.epi_mode == 1 || !aligned_out) ? 1 : (ctx.epi_mode == 3 ? 0 : 3));
  if (out != nullptr && p.direct_store == 3) {
    encode_tmap_2d(&tout, TmaType::F32, out, p.n, out_rows, ldo * 4, 32, 32, 128);
  } else {
    tout = ta_hi;  // never dereferenced
  }
  p.dbg = ctx.gemm_dbg.size() >= 32 ? ctx.gemm_dbg.get() : nullptr;
  p.dbg_skip_a = ctx.dbg_skip_a;
  const GemmEpilogue& ep = p.epi;
  // 1: score-grid hot path, 2: LDA log-sum-exp pass (nothing stored), 0: everything else
  int epi = 0;
  if (out != nullptr && p.direct_store == 3 && ep.rsum == nullptr && ep.lse_max == nullptr && ep.grp == nullptr) epi = 1;
  // register-direct sector stores (PLDA_B200_EPI=sector): same preconditions as the TMA path (8-byte aligned rows)
  static const bool ragged_generic = getenv("PLDA_B200_RAGGED_EPI") != nullptr;   // A/B switch: generic epilogue
  if (out != nullptr && p.direct_store == 3 && ep.rsum == nullptr && ep.lse_max == nullptr && ep.grp != nullptr &&
      ep.col_add != nullptr && !ragged_generic)
    epi = 6;
  if (epi == 1 && ctx.epi_sector) epi = 3;
  else if (epi == 1 && ctx.epi_hybrid) epi = 7;
  else if (out == nullptr && ep.lse_max != nullptr && ep.rsum == nullptr && ep.grp == nullptr) epi = 2;
  else if (out == nullptr && ep.mom != nullptr) epi = 4;
  else if (out == nullptr && ep.hist_n != nullptr) epi = 5;
  if (epi == 4 || epi == 5)
    PB_CHECK(ep.grp == nullptr && ep.rsum == nullptr && ep.lse_max == nullptr, kInvalidArg,
             "gemm: the moments / histogram sinks need uniform column terms and no other reduction");
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (ctx.profile_gemm) {
    PB_CUDA(cudaEventCreate(&e0));
    PB_CUDA(cudaEventCreate(&e1));
    PB_CUDA(cudaEventRecord(e0, ctx.stream));
  }
  // the events of the profiling pass sit between the producer and this launch: no overlap to declare then
  const bool pdl = ctx.pdl_pending && !ctx.profile_gemm;
  ctx.pdl_pending = false;
  const TmapSet tm{ta_hi, ta_lo, tb_hi, tb_lo, tta_hi, tta_lo, ttb_hi, ttb_lo, tout};
  switch (epi) {
    case 1: launch_epi<1>(ctx, pl, tm, pdl); break;
    case 2: launch_epi<2>(ctx, pl, tm, pdl); break;
    case 3: launch_epi<3>(ctx, pl, tm, pdl); break;
    case 4: launch_epi<4>(ctx, pl, tm, pdl); break;
    case 5: launch_epi<5>(ctx, pl, tm, pdl); break;
    case 6: launch_epi<6>(ctx, pl, tm, pdl); break;
    case 7: launch_epi<7>(ctx, pl, tm, pdl); break;
    default: launch_epi<0>(ctx, pl, tm, pdl); break;
  }
  PB_CUDA(cudaGetLastError());
  if (ctx.profile_gemm) {
    PB_CUDA(cudaEventRecord(e1, ctx.stream));
    ctx.gemm_events.emplace_back(e0, e1);
  }
  ctx.count_launch();
}

}  // namespace

void encode_tmap_2d(CUtensorMap* out, TmaType type, const void* base, uint64_t inner, uint64_t outer,
                    uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer, uint32_t swizzle_bytes) {
  EncodeTiledFn fn = get_encode_fn();
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstride[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt = type == TmaType::BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : (swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUresult r = fn(out, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    throw Error(kCudaError, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r)) +
                                " (inner=" + std::to_string(inner) + " outer=" + std::to_string(outer) +
                                " stride=" + std::to_string(row_stride_bytes) + " box=" + std::to_string(box_inner) +
                                "x" + std::to_string(box_outer) + ")");
  }
}

void gemm_bf16x3(Context& ctx, const SplitOperand& a, const SplitOperand& b, int64_t m, int64_t n, int64_t k,
                 const GemmEpilogue& epi, const GemmShard* shard) {
  if (shard != nullptr && shard->flags != nullptr) {
    PB_CHECK(shard->world >= 1 && shard->world <= 16 && shard->rank >= 0 && shard->rank < shard->world &&
                 shard->err != nullptr && shard->bounds[0] == 0 && shard->bounds[shard->world] == n,
             kInvalidArg, "gemm: bad shard description");
  }
  if (epi.col_add != nullptr) {
    PB_CHECK(epi.col_ld % 4 == 0 && epi.col_ld >= round_up(n, 32) &&
                 (reinterpret_cast<uintptr_t>(epi.col_add) & 15) == 0,
             kInvalidArg, "gemm: col_add must be 16B aligned with pitch >= round_up(n,32)");
  }
  // small-K score grids: enrol operand in tensor memory (no A traffic through shared memory)
  if (gemm_ts_score(ctx, a, b, m, n, k, epi, shard)) return;
  Plan pl = make_plan(ctx, m, n, k, 1);
  pl.p.epi = epi;
  if (shard != nullptr && shard->flags != nullptr) {
    PB_CHECK(shard->world >= 1 && shard->world <= 16 && shard->rank >= 0 && shard->rank < shard->world &&
                 shard->err != nullptr && shard->bounds[0] == 0 && shard->bounds[shard->world] == n,
             kInvalidArg, "gemm: bad shard description");
    PB_CHECK(epi.grp == nullptr, kInvalidArg, "gemm: sharded B needs uniform column terms");
    pl.p.shard = *shard;
    // start with the column tile that holds this rank's first row (empty shard: tile 0)
    const int first = shard->bounds[shard->rank] < n ? shard->bounds[shard->rank] : 0;
    pl.p.n_rot = first / pl.p.bn;
  }
  if (epi.col_add != nullptr) {
    PB_CHECK(epi.col_ld % 4 == 0 && epi.col_ld >= round_up(n, 32) &&
                 (reinterpret_cast<uintptr_t>(epi.col_add) & 15) == 0,
             kInvalidArg, "gemm: col_add must be 16B aligned with pitch >= round_up(n,32)");
  }
  launch(ctx, a, b, pl, epi.out, epi.ldo, m);
}

namespace {
// A rank without enrol rows launches no GEMM; it still has to observe every peer's flag for the epoch before it may
// push again (the protocol's only back-pressure), or it could overwrite a generation a peer is still reading.
__global__ void shard_wait_all_kernel(GemmShard sh) {
  if (static_cast<int>(threadIdx.x) < sh.world) shard_wait(sh.flags + threadIdx.x, sh.epoch, sh.err);
}
}  // namespace

void shard_wait_all(Context& ctx, const GemmShard& shard) {
  PB_CHECK(shard.flags != nullptr && shard.err != nullptr && shard.world >= 1 && shard.world <= 16, kInvalidArg,
           "shard_wait_all: bad shard description");
  shard_wait_all_kernel<<<1, 32, 0, ctx.stream>>>(shard);
  PB_CUDA(cudaGetLastError());
  ctx.count_launch();
}

int gemm_n_tiles(int64_t n) {
  const int64_t bn = n >= BN_MAX ? BN_MAX : round_up(n, 16);
  return static_cast<int>(ceil_div(n, bn));
}

int choose_ksplit(const Context& ctx, int64_t m, int64_t n, int64_t k) {
  const int64_t bn = n >= BN_MAX ? BN_MAX : round_up(n, 16);
  const int64_t tiles = ceil_div(m, BM) * ceil_div(n, bn);
  const int64_t nkb = ceil_div(round_up(k, 16), BK);
  int64_t s = (2 * static_cast<int64_t>(ctx.num_sms)) / tiles;
  if (s < 1) s = 1;
  const int64_t max_by_work = nkb / 4 > 0 ? nkb / 4 : 1;   // at least 4 k-blocks per split
  if (s > max_by_work) s = max_by_work;
  return static_cast<int>(s);
}

void gemm_bf16x3_splitk(Context& ctx, const SplitOperand& a, const SplitOperand& b, int64_t m, int64_t n, int64_t k,
                        int ksplit, float* partial) {
  Plan pl = make_plan(ctx, m, n, k, ksplit);
  GemmEpilogue epi;
  epi.out = partial;
  epi.ldo = round_up(n, 4);
  pl.p.epi = epi;
  launch(ctx, a, b, pl, partial, epi.ldo, static_cast<int64_t>(pl.p.ksplit) * pl.p.mpad);
}

// number of partial planes gemm_bf16x3_splitk will actually write for a requested ksplit
int effective_ksplit(const Context& ctx, int64_t m, int64_t n, int64_t k, int ksplit) {
  return make_plan(ctx, m, n, k, ksplit).p.ksplit;
}

void reduce_partials_f64(Context& ctx, const float* partial, int ksplit, int64_t m, int64_t n, double* out,
                         int64_t ldo, double alpha, bool symmetrise) {
  const int mpad = static_cast<int>(round_up(m, BM));
  const int npad = static_cast<int>(round_up(n, 4));
  const long long total = m * n;
  const int threads = 256;
  reduce_partials_kernel<<<static_cast<unsigned>(ceil_div(total, threads)), threads, 0, ctx.stream>>>(
      partial, ksplit, static_cast<int>(m), static_cast<int>(n), mpad, npad, out, ldo, alpha, symmetrise ? 1 : 0);
  PB_CUDA(cudaGetLastError());
  ctx.count_launch();
}

}  // namespace pb
