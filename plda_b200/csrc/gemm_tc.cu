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
// Kernel anatomy (persistent, warp-specialised, one CTA per SM, 192 threads):
//   warp 0      TMA producer: 4 tile loads per k-block (A_hi, A_lo, B_hi, B_lo), 128B swizzle,
//               2-stage smem ring (96 KB/stage) guarded by full/empty mbarriers
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer; tcgen05.commit releases smem
//               stages and publishes finished accumulators
//   warps 2..5  epilogue: tcgen05.ld (32 lanes x 32 columns) -> fused row/col/z-norm terms ->
//               swizzled smem staging -> TMA store (or fused row reductions, no store)
//   TMEM        2 accumulator stages x 256 fp32 columns (all 512 columns): the epilogue of tile i
//               overlaps the MMAs of tile i+1.
// Tile = 128 x BN (BN <= 256, multiple of 16) x 64.
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
constexpr int EPI_WARPS = 4;
constexpr int EPI_BUF_BYTES = 32 * 32 * 4;                          // one 32x32 fp32 box
constexpr int EPI_BYTES = EPI_WARPS * 2 * EPI_BUF_BYTES;            // 32 KB
constexpr int BAR_BYTES = 256;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES + 1024;
constexpr int NUM_THREADS = 192;
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
  int direct_store;
  GemmEpilogue epi;
};

struct Work {
  int m_blk, n_blk, ks;
};

__device__ __forceinline__ Work decode(const GemmParams& p, int item) {
  Work w;
  const int tiles = p.m_tiles * p.n_tiles;
  w.ks = item / tiles;
  const int t = item - w.ks * tiles;
  const int per_group = p.group_m * p.n_tiles;
  const int g = t / per_group;
  const int first_m = g * p.group_m;
  const int gsz = min(p.group_m, p.m_tiles - first_m);
  const int r = t - g * per_group;
  w.m_blk = first_m + r % gsz;
  w.n_blk = r / gsz;
  return w;
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16x3_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                   const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                   const __grid_constant__ CUtensorMap tm_out, const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi_base = smem + STAGES * STAGE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(epi_base + EPI_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_items = p.m_tiles * p.n_tiles * p.ksplit;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a_hi);
    tma_prefetch_desc(&tm_a_lo);
    tma_prefetch_desc(&tm_b_hi);
    tma_prefetch_desc(&tm_b_lo);
    tma_prefetch_desc(&tm_out);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&full[s], 1);
        mbar_init(&empty[s], 1);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(&tfull[a], 1);
        mbar_init(&tempty[a], EPI_WARPS);
      }
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer ===================== //
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx_bytes = 2 * A_TILE_BYTES + 2 * (p.bn * BK * 2);
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const Work w = decode(p, item);
        const int kb0 = w.ks * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.nkb_total);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full[stage], tx_bytes);
          uint8_t* s = smem + stage * STAGE_BYTES;
          tma_load_2d(s, &tm_a_hi, &full[stage], kb * BK, w.m_blk * BM);
          tma_load_2d(s + A_TILE_BYTES, &tm_a_lo, &full[stage], kb * BK, w.m_blk * BM);
          tma_load_2d(s + 2 * A_TILE_BYTES, &tm_b_hi, &full[stage], kb * BK, w.n_blk * p.bn);
          tma_load_2d(s + 2 * A_TILE_BYTES + B_TILE_BYTES, &tm_b_lo, &full[stage], kb * BK, w.n_blk * p.bn);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer ===================== //
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint32_t idesc = umma_idesc_bf16_f32(BM, p.bn);
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const Work w = decode(p, item);
        const int kb0 = w.ks * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.nkb_total);
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN_MAX;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint64_t a_hi = umma_desc_kmajor_sw128(sa);
          const uint64_t a_lo = umma_desc_kmajor_sw128(sa + A_TILE_BYTES);
          const uint64_t b_hi = umma_desc_kmajor_sw128(sa + 2 * A_TILE_BYTES);
          const uint64_t b_lo = umma_desc_kmajor_sw128(sa + 2 * A_TILE_BYTES + B_TILE_BYTES);
          const int nks = (kb == p.nkb_total - 1) ? p.last_ksteps : (BK / UMMA_K);
          for (int ks = 0; ks < nks; ++ks) {
            // advance 32 B (= UMMA_K bf16) inside the 128 B swizzle atom: +2 in the >>4 address field
            const uint64_t off = static_cast<uint64_t>(ks * 2);
            const uint32_t first = (kb == kb0 && ks == 0) ? 0u : 1u;
            umma_bf16_ss(d_tmem, a_hi + off, b_hi + off, idesc, first);
            umma_bf16_ss(d_tmem, a_hi + off, b_lo + off, idesc, 1u);
            umma_bf16_ss(d_tmem, a_lo + off, b_hi + off, idesc, 1u);
          }
          umma_commit(&empty[stage]);          // smem stage reusable once these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[acc]);              // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) ===================== //
    const int q = warp & 3;   // TMEM lane quarter this warp is allowed to read
    uint8_t* my_epi = epi_base + (warp - 2) * 2 * EPI_BUF_BYTES;
    const GemmEpilogue& e = p.epi;
    int acc = 0;
    uint32_t acc_phase = 0;
    int buf = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const Work w = decode(p, item);
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
      float rs = 0.f, rq = 0.f;
      float lmax = -INFINITY, lsum = 0.f;
      const int ncols = min(p.bn, p.n - n0);
      const int nchunks = (ncols + 31) >> 5;
      const long long out_row = static_cast<long long>(w.ks) * p.mpad + m;

      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      for (int c = 0; c < nchunks; ++c) {
        uint32_t r[32];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN_MAX + c * 32;
        tmem_ld_32x32b_x32(taddr, r);
        tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        const int nbase = n0 + c * 32;
        if (e.col_add != nullptr && mvalid) {
          const float4* cp = reinterpret_cast<const float4*>(e.col_add + static_cast<long long>(g) * e.col_ld + nbase);
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 t = __ldg(cp + j4);
            v[4 * j4 + 0] += t.x; v[4 * j4 + 1] += t.y; v[4 * j4 + 2] += t.z; v[4 * j4 + 3] += t.w;
          }
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = (v[j] + ra - zm) * zi;
        if (e.rsum != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (nbase + j < p.n) { rs += v[j]; rq += v[j] * v[j]; }
          }
        }
        if (e.lse_max != nullptr) {
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
        if (e.out != nullptr) {
          if (p.direct_store) {
            if (mvalid) {
              float* op = e.out + out_row * e.ldo + nbase;
#pragma unroll
              for (int j = 0; j < 32; ++j) if (nbase + j < p.n) op[j] = v[j];
            }
          } else if (warp_rows_valid) {
            uint8_t* sb = my_epi + buf * EPI_BUF_BYTES;
            if (lane == 0) tma_store_wait_read<1>();   // the store that last used this buffer has drained
            __syncwarp();
            const uint32_t row_addr = smem_u32(sb) + lane * 128;
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const uint32_t addr = row_addr + ((j4 ^ (lane & 7)) << 4);   // 128B-swizzle: chunk ^= row%8
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
            buf ^= 1;
          }
        }
      }
      // accumulator stage drained -> hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }

      if (mvalid) {
        if (e.rsum != nullptr) {
          atomicAdd(e.rsum + m, static_cast<double>(rs));
          atomicAdd(e.rsq + m, static_cast<double>(rq));
        }
        if (e.lse_max != nullptr) {
          e.lse_max[static_cast<long long>(m) * p.n_tiles + w.n_blk] = lmax;
          e.lse_sum[static_cast<long long>(m) * p.n_tiles + w.n_blk] = lsum;
        }
      }
    }
    if (lane == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, TMEM_COLS);
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
  return pl;
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
  CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo, tout;
  encode_tmap_2d(&ta_hi, TmaType::BF16, a.hi, kext, a.rows, a.ld * 2, BK, BM);
  encode_tmap_2d(&ta_lo, TmaType::BF16, a.lo, kext, a.rows, a.ld * 2, BK, BM);
  encode_tmap_2d(&tb_hi, TmaType::BF16, b.hi, kext, b.rows, b.ld * 2, BK, p.bn);
  encode_tmap_2d(&tb_lo, TmaType::BF16, b.lo, kext, b.rows, b.ld * 2, BK, p.bn);
  const bool can_tma_store = out != nullptr && (ldo % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  p.direct_store = (ctx.epi_direct || !can_tma_store) ? 1 : 0;
  if (out != nullptr && !p.direct_store) {
    encode_tmap_2d(&tout, TmaType::F32, out, p.n, out_rows, ldo * 4, 32, 32);
  } else {
    tout = ta_hi;  // never dereferenced
  }
  static std::once_flag attr_once;
  std::call_once(attr_once, [] {
    cudaFuncSetAttribute(gemm_bf16x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  });
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (ctx.profile_gemm) {
    PB_CUDA(cudaEventCreate(&e0));
    PB_CUDA(cudaEventCreate(&e1));
    PB_CUDA(cudaEventRecord(e0, ctx.stream));
  }
  gemm_bf16x3_kernel<<<pl.grid, NUM_THREADS, SMEM_BYTES, ctx.stream>>>(ta_hi, ta_lo, tb_hi, tb_lo, tout, p);
  PB_CUDA(cudaGetLastError());
  if (ctx.profile_gemm) {
    PB_CUDA(cudaEventRecord(e1, ctx.stream));
    ctx.gemm_events.emplace_back(e0, e1);
  }
  ctx.count_launch();
}

}  // namespace

void encode_tmap_2d(CUtensorMap* out, TmaType type, const void* base, uint64_t inner, uint64_t outer,
                    uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer) {
  EncodeTiledFn fn = get_encode_fn();
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstride[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt = type == TmaType::BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = fn(out, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    throw Error(kCudaError, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r)) +
                                " (inner=" + std::to_string(inner) + " outer=" + std::to_string(outer) +
                                " stride=" + std::to_string(row_stride_bytes) + " box=" + std::to_string(box_inner) +
                                "x" + std::to_string(box_outer) + ")");
  }
}

void gemm_bf16x3(Context& ctx, const SplitOperand& a, const SplitOperand& b, int64_t m, int64_t n, int64_t k,
                 const GemmEpilogue& epi) {
  Plan pl = make_plan(ctx, m, n, k, 1);
  pl.p.epi = epi;
  if (epi.col_add != nullptr) {
    PB_CHECK(epi.col_ld % 4 == 0 && epi.col_ld >= round_up(n, 32) &&
                 (reinterpret_cast<uintptr_t>(epi.col_add) & 15) == 0,
             kInvalidArg, "gemm: col_add must be 16B aligned with pitch >= round_up(n,32)");
  }
  launch(ctx, a, b, pl, epi.out, epi.ldo, m);
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
