// Sinks of the enrol x test score grid other than "the whole fp32 matrix":
//   * z-norm moments       (src/pldamodule.cpp:240-253: mean and population std per enrol model) from the per-slot
//                           shifted partials the MOM epilogue of the tensor GEMM writes (gemm_tc.cu)
//   * listed trials        (scoring/scorePLDA.py:302-318: the reference scores the trials of a list one by one) --
//                           either straight from the transformed vectors (sparse lists) or gathered from a grid slab
// The EER histogram sink lives in the GEMM epilogue itself (gemm_tc.cu, EPI 5); its host side is in engine_sinks.cu.
#include <algorithm>

#include "kernels.h"

namespace pb {
namespace {

__global__ void __launch_bounds__(256)
moments_reduce_kernel(const float4* __restrict__ mom, long long ne, int slots, float* __restrict__ zmean,
                      float* __restrict__ zinv, double* __restrict__ mean_out, double* __restrict__ std_out) {
  const long long m = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (m >= ne) return;
  double n = 0.0, mean = 0.0, m2 = 0.0;
  const float4* row = mom + m * slots;
  for (int s = 0; s < slots; ++s) {
    const float4 t = row[s];
    const double c = static_cast<double>(t.w);
    if (c <= 0.0) continue;
    const double s1 = static_cast<double>(t.x), s2 = static_cast<double>(t.y);
    const double mt = static_cast<double>(t.z) + s1 / c;     // slot mean
    double m2t = s2 - s1 * s1 / c;                            // slot sum of squared deviations
    if (m2t < 0.0) m2t = 0.0;
    const double tot = n + c;
    const double delta = mt - mean;
    mean += delta * (c / tot);
    m2 += m2t + delta * delta * (n * c / tot);
    n = tot;
  }
  const double var = n > 0.0 ? m2 / n : 0.0;                  // population variance (:246-250)
  const double sd = sqrt(var > 0.0 ? var : 0.0);
  if (zmean) zmean[m] = static_cast<float>(mean);
  if (zinv) zinv[m] = static_cast<float>(1.0 / sd);
  if (mean_out) mean_out[m] = mean;
  if (std_out) std_out[m] = sd;
}

// One warp per trial.  LLR = 1/2 logdet(n) - 1/2 sum esq_i e_i^2 + sum sc_i e_i t_i + sum tsq_i t_i^2
// (the constants come from the per-count table).
template <typename T>
__global__ void __launch_bounds__(256)
score_trials_kernel(const T* __restrict__ enrol, long long ld_e, const T* __restrict__ test, long long ld_t, int dim,
                    const double* __restrict__ tables, const int32_t* __restrict__ grp,
                    const float* __restrict__ zmean, const float* __restrict__ zinv, const int32_t* __restrict__ te,
                    const int32_t* __restrict__ tt, long long n_trials, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long warps = static_cast<long long>(gridDim.x) * (blockDim.x >> 5);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x >> 5) + (threadIdx.x >> 5); i < n_trials;
       i += warps) {
    const int e = __ldg(te + i), t = __ldg(tt + i);
    const double* tab = tables + (grp ? static_cast<long long>(__ldg(grp + e)) * kScoreConstsSize : 0);
    const T* er = enrol + static_cast<long long>(e) * ld_e;
    const T* tr = test + static_cast<long long>(t) * ld_t;
    double acc = 0.0;
    for (int c = lane; c < dim; c += 32) {
      const double ev = static_cast<double>(__ldg(er + c)), tv = static_cast<double>(__ldg(tr + c));
      acc = fma(__ldg(tab + kScoreConstsScale + c) * ev, tv, acc);
      acc = fma(-0.5 * __ldg(tab + kScoreConstsEnrolSq + c) * ev, ev, acc);
      acc = fma(__ldg(tab + kScoreConstsTestSq + c) * tv, tv, acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      double s = 0.5 * __ldg(tab + kScoreConstsLogdet) + acc;
      if (zmean) s = (s - static_cast<double>(__ldg(zmean + e))) * static_cast<double>(__ldg(zinv + e));
      out[i] = static_cast<float>(s);
    }
  }
}

__global__ void __launch_bounds__(256)
gather_trials_kernel(const float* __restrict__ slab, long long ld, int r0, int rows, const int32_t* __restrict__ te,
                     const int32_t* __restrict__ tt, long long n_trials, float* __restrict__ out) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n_trials; i += stride) {
    const int e = __ldg(te + i) - r0;
    if (e >= 0 && e < rows) out[i] = __ldg(slab + static_cast<long long>(e) * ld + __ldg(tt + i));
  }
}

}  // namespace

void moments_reduce(Context& ctx, const float4* mom, int64_t ne, int n_tiles, float* zmean, float* zinv,
                    double* mean_out, double* std_out) {
  if (ne == 0) return;
  moments_reduce_kernel<<<static_cast<unsigned>(ceil_div(ne, 256)), 256, 0, ctx.stream>>>(mom, ne, 2 * n_tiles, zmean,
                                                                                         zinv, mean_out, std_out);
  PB_CUDA(cudaGetLastError());
  ctx.count_launch();
}

void score_trials_direct(Context& ctx, const void* enrol, int64_t ld_e, const void* test, int64_t ld_t, bool is_f32,
                         int64_t dim, const double* tables, const int32_t* grp, const float* zmean, const float* zinv,
                         const int32_t* te, const int32_t* tt, int64_t n_trials, float* out) {
  if (n_trials == 0) return;
  const int64_t want = ceil_div(n_trials, 8);
  const unsigned blocks = static_cast<unsigned>(std::min<int64_t>(want, 16ll * ctx.num_sms));
  if (is_f32)
    score_trials_kernel<float><<<blocks, 256, 0, ctx.stream>>>(static_cast<const float*>(enrol), ld_e,
                                                               static_cast<const float*>(test), ld_t,
                                                               static_cast<int>(dim), tables, grp, zmean, zinv, te, tt,
                                                               n_trials, out);
  else
    score_trials_kernel<double><<<blocks, 256, 0, ctx.stream>>>(static_cast<const double*>(enrol), ld_e,
                                                                static_cast<const double*>(test), ld_t,
                                                                static_cast<int>(dim), tables, grp, zmean, zinv, te, tt,
                                                                n_trials, out);
  PB_CUDA(cudaGetLastError());
  ctx.count_launch();
}

void gather_trials(Context& ctx, const float* slab, int64_t ld, int64_t r0, int64_t rows, const int32_t* te,
                   const int32_t* tt, int64_t n_trials, float* out) {
  if (n_trials == 0 || rows == 0) return;
  const unsigned blocks = static_cast<unsigned>(std::min<int64_t>(ceil_div(n_trials, 256), 8ll * ctx.num_sms));
  gather_trials_kernel<<<blocks, 256, 0, ctx.stream>>>(slab, ld, static_cast<int>(r0), static_cast<int>(rows), te, tt,
                                                       n_trials, out);
  PB_CUDA(cudaGetLastError());
  ctx.count_launch();
}

}  // namespace pb
