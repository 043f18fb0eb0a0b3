// Context lifetime + the small generic kernels every path shares (operand splitting,
// fp64 SIMT GEMM used for d x d algebra and as the exact-precision mode).
#include "runtime.h"

#include <stdlib.h>
#include <string.h>

namespace pb {

Context::Context(int dev) : device(dev) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    throw Error(kCudaError, std::string("no CUDA device available (") + cudaGetErrorString(e) +
                                "); plda_b200 has no CPU fallback");
  }
  PB_CHECK(dev >= 0 && dev < count, kInvalidArg, "device index out of range");
  PB_CUDA(cudaSetDevice(dev));
  cudaDeviceProp prop;
  PB_CUDA(cudaGetDeviceProperties(&prop, dev));
  PB_CHECK(prop.major == 10, kCudaError,
           std::string("plda_b200 is built for sm_100a only; device is sm_") + std::to_string(prop.major) +
               std::to_string(prop.minor));
  num_sms = prop.multiProcessorCount;
  PB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  const char* epi = getenv("PLDA_B200_EPI");
  epi_mode = epi == nullptr ? 0
             : (strcmp(epi, "direct") == 0 ? 1 : (strcmp(epi, "skip") == 0 ? 2 : (strcmp(epi, "lsu") == 0 ? 3 : 0)));
  epi_sector = epi != nullptr && strcmp(epi, "sector") == 0;
  epi_hybrid = epi != nullptr && strcmp(epi, "hybrid") == 0;
  const char* kt = getenv("PLDA_B200_KTAIL");
  k_tail_boxes = !(kt != nullptr && strcmp(kt, "0") == 0);
  const char* dbg = getenv("PLDA_B200_DBG");
  if (dbg != nullptr && strcmp(dbg, "1") == 0) {
    gemm_dbg.alloc(32);
    PB_CUDA(cudaMemset(gemm_dbg.get(), 0, 32 * sizeof(long long)));
  }
  const char* sk = getenv("PLDA_B200_DBGSKIPA");
  dbg_skip_a = (sk != nullptr && sk[0] == '1') ? 1 : 0;
  const char* pdl = getenv("PLDA_B200_PDL");
  if (pdl && pdl[0] == '0') pdl_enabled = false;
  const char* gm = getenv("PLDA_B200_GEMM");
  gemm_two_cta = !(gm != nullptr && strcmp(gm, "1cta") == 0);
  const char* ts = getenv("PLDA_B200_TS");
  if (ts != nullptr) gemm_ts = ts[0] == '1';
}

Context::~Context() {
  profile_reset();
  if (owns_stream && stream) cudaStreamDestroy(stream);
}

void Context::profile_reset() {
  for (auto& pr : gemm_events) {
    cudaEventDestroy(pr.first);
    cudaEventDestroy(pr.second);
  }
  gemm_events.clear();
}

void Context::profile_collect(double* total_ms, int64_t* count) {
  sync();
  double tot = 0.0;
  for (auto& pr : gemm_events) {
    float ms = 0.f;
    PB_CUDA(cudaEventElapsedTime(&ms, pr.first, pr.second));
    tot += ms;
  }
  *total_ms = tot;
  *count = static_cast<int64_t>(gemm_events.size());
}

// ------------------------------------------------------------------------- //
// fp64 SIMT GEMM (64x64x16 tiles, 4x4 register micro-tile)
// ------------------------------------------------------------------------- //
namespace {

// One problem of a (possibly batched) launch: blockIdx.z selects it.
struct GemmF64Problem {
  const double* a;
  const double* b;
  double* c;
};
struct GemmF64Batch {
  GemmF64Problem p[2];
};

template <bool TA, bool TB, int TILE>   // TILE = 64 (4x4 per thread) or 32 (2x2 per thread: more CTAs for d x d work)
__global__ void __launch_bounds__(256)
gemm_f64_kernel(int m, int n, int k, double alpha, const GemmF64Batch batch, long long lda, long long ldb, double beta,
                long long ldc) {
  constexpr int MT = TILE / 16;
  constexpr int NL = TILE * 16 / 256;        // elements per thread per operand and k-chunk
  __shared__ double sa[16][TILE + 1];
  __shared__ double sb[16][TILE + 1];
  const double* __restrict__ a = batch.p[blockIdx.z].a;
  const double* __restrict__ b = batch.p[blockIdx.z].b;
  double* __restrict__ c = batch.p[blockIdx.z].c;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * TILE, n0 = blockIdx.x * TILE;
  double acc[MT][MT] = {};
  // the next k-chunk travels global -> registers while the current one is consumed from shared memory
  double ra[NL], rb[NL];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int l = 0; l < NL; ++l) {
      const int i = threadIdx.x + l * 256;
      int mm, kk;
      if (TA) { mm = i % TILE; kk = i / TILE; } else { kk = i & 15; mm = i >> 4; }
      const int gm = m0 + mm, gk = k0 + kk;
      ra[l] = (gm < m && gk < k) ? (TA ? a[gk * lda + gm] : a[gm * lda + gk]) : 0.0;
      int nn, kb;
      if (TB) { kb = i & 15; nn = i >> 4; } else { nn = i % TILE; kb = i / TILE; }
      const int gn = n0 + nn, gkb = k0 + kb;
      rb[l] = (gn < n && gkb < k) ? (TB ? b[gn * ldb + gkb] : b[gkb * ldb + gn]) : 0.0;
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < k; k0 += 16) {
#pragma unroll
    for (int l = 0; l < NL; ++l) {
      const int i = threadIdx.x + l * 256;
      int mm, kk;
      if (TA) { mm = i % TILE; kk = i / TILE; } else { kk = i & 15; mm = i >> 4; }
      sa[kk][mm] = ra[l];
      int nn, kb;
      if (TB) { kb = i & 15; nn = i >> 4; } else { nn = i % TILE; kb = i / TILE; }
      sb[kb][nn] = rb[l];
    }
    __syncthreads();
    if (k0 + 16 < k) fetch(k0 + 16);
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      double va[MT], vb[MT];
#pragma unroll
      for (int i = 0; i < MT; ++i) va[i] = sa[kk][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < MT; ++j) vb[j] = sb[kk][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < MT; ++j) acc[i][j] = fma(va[i], vb[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < MT; ++i) {
    const int gm = m0 + ty + 16 * i;
    if (gm >= m) continue;
#pragma unroll
    for (int j = 0; j < MT; ++j) {
      const int gn = n0 + tx + 16 * j;
      if (gn >= n) continue;
      double* p = c + gm * ldc + gn;
      *p = beta == 0.0 ? alpha * acc[i][j] : alpha * acc[i][j] + beta * (*p);
    }
  }
}

template <int TILE>
void launch_gemm_f64(Context& ctx, bool ta, bool tb, int m, int n, int k, double alpha, const GemmF64Batch& batch,
                     int count, int64_t lda, int64_t ldb, double beta, int64_t ldc) {
  dim3 grid(static_cast<unsigned>(ceil_div(n, TILE)), static_cast<unsigned>(ceil_div(m, TILE)), static_cast<unsigned>(count));
  if (ta && tb) gemm_f64_kernel<true, true, TILE><<<grid, 256, 0, ctx.stream>>>(m, n, k, alpha, batch, lda, ldb, beta, ldc);
  else if (ta) gemm_f64_kernel<true, false, TILE><<<grid, 256, 0, ctx.stream>>>(m, n, k, alpha, batch, lda, ldb, beta, ldc);
  else if (tb) gemm_f64_kernel<false, true, TILE><<<grid, 256, 0, ctx.stream>>>(m, n, k, alpha, batch, lda, ldb, beta, ldc);
  else gemm_f64_kernel<false, false, TILE><<<grid, 256, 0, ctx.stream>>>(m, n, k, alpha, batch, lda, ldb, beta, ldc);
}

void gemm_f64_any(Context& ctx, bool ta, bool tb, int64_t m, int64_t n, int64_t k, double alpha, const GemmF64Batch& batch,
                  int count, int64_t lda, int64_t ldb, double beta, int64_t ldc) {
  PB_CHECK(m > 0 && n > 0 && k > 0, kInvalidArg, "gemm_f64: empty problem");
  const int mi = static_cast<int>(m), ni = static_cast<int>(n), ki = static_cast<int>(k);
  // d x d algebra (a handful of 64x64 tiles) is latency bound: use 32x32 tiles to spread over more SMs
  if (ceil_div(m, 64) * ceil_div(n, 64) * count < 2 * ctx.num_sms)
    launch_gemm_f64<32>(ctx, ta, tb, mi, ni, ki, alpha, batch, count, lda, ldb, beta, ldc);
  else
    launch_gemm_f64<64>(ctx, ta, tb, mi, ni, ki, alpha, batch, count, lda, ldb, beta, ldc);
  PB_CUDA(cudaGetLastError());
  ctx.count_launch();
}

}  // namespace

void gemm_f64(Context& ctx, bool ta, bool tb, int64_t m, int64_t n, int64_t k, double alpha, const double* a,
              int64_t lda, const double* b, int64_t ldb, double beta, double* c, int64_t ldc) {
  GemmF64Batch batch;
  batch.p[0] = GemmF64Problem{a, b, c};
  batch.p[1] = batch.p[0];
  gemm_f64_any(ctx, ta, tb, m, n, k, alpha, batch, 1, lda, ldb, beta, ldc);
}

void gemm_f64_pair(Context& ctx, bool ta, bool tb, int64_t m, int64_t n, int64_t k, double alpha, const double* a0,
                   const double* b0, double* c0, const double* a1, const double* b1, double* c1, int64_t lda,
                   int64_t ldb, double beta, int64_t ldc) {
  GemmF64Batch batch;
  batch.p[0] = GemmF64Problem{a0, b0, c0};
  batch.p[1] = GemmF64Problem{a1, b1, c1};
  gemm_f64_any(ctx, ta, tb, m, n, k, alpha, batch, 2, lda, ldb, beta, ldc);
}

}  // namespace pb
