// HBM-bound operand producers and row-wise epilogues of the scoring / transform path.
// All are one-warp-per-row kernels with coalesced loads along the feature axis, fp64
// arithmetic on the way in (inputs are the caller's fp64/fp32 rows), split-bf16 or fp32 out.
#include "kernels.h"

namespace pb {
namespace {

constexpr int kWarpsPerBlock = 8;

template <typename T>
__device__ __forceinline__ double load_as_f64(const T* p) { return static_cast<double>(*p); }

__device__ __forceinline__ void store_split(__nv_bfloat16* hi, __nv_bfloat16* lo, long long idx, double v) {
  __nv_bfloat16 h, l;
  split_bf16(v, h, l);
  hi[idx] = h;
  lo[idx] = l;
}

// ------------------------------------------------------------------------- //
template <typename T>
__global__ void split_rows_kernel(const T* __restrict__ in, long long rows, int cols, long long ld_in,
                                  const double* __restrict__ sub, const double* __restrict__ col_scale,
                                  const double* __restrict__ row_scale, __nv_bfloat16* __restrict__ hi,
                                  __nv_bfloat16* __restrict__ lo, int ld_out) {
  const long long r = blockIdx.x * static_cast<long long>(kWarpsPerBlock) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const double rs = row_scale ? row_scale[r] : 1.0;
  const T* src = in + r * ld_in;
  for (int c = lane; c < ld_out; c += 32) {
    double v = 0.0;
    if (c < cols) {
      v = static_cast<double>(src[c]);
      if (sub) v -= sub[c];
      if (col_scale) v *= col_scale[c];
      v *= rs;
    }
    store_split(hi, lo, r * ld_out + c, v);
  }
}

template <typename T>
__global__ void convert_kernel(const T* __restrict__ in, long long rows, int cols, long long ld_in,
                               const double* __restrict__ sub, double* __restrict__ out, long long ld_out) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= rows * cols) return;
  const long long r = idx / cols;
  const int c = static_cast<int>(idx - r * cols);
  double v = static_cast<double>(in[r * ld_in + c]);
  if (sub) v -= sub[c];
  out[r * ld_out + c] = v;
}

__global__ void f64_to_f32_kernel(const double* __restrict__ in, long long rows, int cols, long long ld_in,
                                  float* __restrict__ out, long long ld_out) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= rows * cols) return;
  const long long r = idx / cols;
  const int c = static_cast<int>(idx - r * cols);
  out[r * ld_out + c] = static_cast<float>(in[r * ld_in + c]);
}

// ------------------------------------------------------------------------- //
// enrol-side LLR operand:  a = n psi/(n psi+1), v = 1 + psi/(n psi+1)
//   L[e,i] = e_i a_i / v_i ;  row[e] = 1/2 sum_i [log(1+psi_i) - log v_i - a_i^2 e_i^2 / v_i]
template <typename T>
__global__ void score_prep_enrol_kernel(const T* __restrict__ enrol, long long ne, int d, long long ld,
                                        const int32_t* __restrict__ counts, int const_count, const double* __restrict__ psi,
                                        __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int ld_out,
                                        double* __restrict__ l_f64, float* __restrict__ row_term,
                                        double* __restrict__ row_term_f64) {
  const long long r = blockIdx.x * static_cast<long long>(kWarpsPerBlock) + (threadIdx.x >> 5);
  if (r >= ne) return;
  const int lane = threadIdx.x & 31;
  const double n = static_cast<double>(counts ? counts[r] : const_count);
  const T* src = enrol + r * ld;
  double acc = 0.0;
  const int cmax = hi ? ld_out : d;
  for (int c = lane; c < cmax; c += 32) {
    double lv = 0.0;
    if (c < d) {
      const double p = psi[c];
      const double den = n * p + 1.0;
      const double a = n * p / den;
      const double v = 1.0 + p / den;
      const double e = static_cast<double>(src[c]);
      lv = e * a / v;
      acc += log1p(p) - log(v) - a * a * e * e / v;
      if (l_f64) l_f64[r * d + c] = lv;
    }
    if (hi) store_split(hi, lo, r * ld_out + c, lv);
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    if (row_term) row_term[r] = static_cast<float>(0.5 * acc);
    if (row_term_f64) row_term_f64[r] = 0.5 * acc;
  }
}

// test-side: R = T ; col[g][t] = sum_i q_i(n_g) t_i^2,  q_i(n) = 1/2 (1/(1+psi_i) - 1/v_i(n))
template <typename T>
__global__ void score_prep_test_kernel(const T* __restrict__ test, long long nt, int d, long long ld,
                                       const int32_t* __restrict__ group_counts, int ngroups, int const_count,
                                       const double* __restrict__ psi, __nv_bfloat16* __restrict__ hi,
                                       __nv_bfloat16* __restrict__ lo, int ld_out, float* __restrict__ col_term,
                                       long long col_ld, double* __restrict__ col_term_f64) {
  const long long r = blockIdx.x * static_cast<long long>(kWarpsPerBlock) + (threadIdx.x >> 5);
  if (r >= nt) return;
  const int lane = threadIdx.x & 31;
  const T* src = test + r * ld;
  if (hi) {
    for (int c = lane; c < ld_out; c += 32) {
      const double v = c < d ? static_cast<double>(src[c]) : 0.0;
      store_split(hi, lo, r * ld_out + c, v);
    }
  }
  for (int g = 0; g < ngroups; ++g) {
    const double n = static_cast<double>(group_counts ? group_counts[g] : const_count);
    double acc = 0.0;
    for (int c = lane; c < d; c += 32) {
      const double p = psi[c];
      const double v = 1.0 + p / (n * p + 1.0);
      const double t = static_cast<double>(src[c]);
      acc += 0.5 * (1.0 / (1.0 + p) - 1.0 / v) * t * t;
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      if (col_term) col_term[g * col_ld + r] = static_cast<float>(acc);
      if (col_term_f64) col_term_f64[g * col_ld + r] = acc;
    }
  }
}

// Fused operand producer of the score grid for the common case of ONE enrol count n (tensor path):
// blocks [0, eb) handle 32 enrol rows each, blocks [eb, eb+tb) 32 test rows each.  The per-column constants
// (a/v, a^2/v, q and the log-determinant term) depend only on (n, psi) and are computed once per block in smem
// instead of once per element; the zero padding of the column-term row is written here too (no memset).
template <typename T>
__global__ void __launch_bounds__(256)
score_prep_uniform_kernel(const T* __restrict__ enrol, long long ne, long long ld_e, const T* __restrict__ test,
                          long long nt, long long ld_t, int d, int count, const double* __restrict__ psi,
                          __nv_bfloat16* __restrict__ l_hi, __nv_bfloat16* __restrict__ l_lo,
                          __nv_bfloat16* __restrict__ r_hi, __nv_bfloat16* __restrict__ r_lo, int ld_out,
                          float* __restrict__ row_term, float* __restrict__ col_term, long long col_ld,
                          unsigned enrol_blocks) {
  __shared__ double s_s[1024], s_w[1024], s_q[1024];
  __shared__ double s_red[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool is_enrol = blockIdx.x < enrol_blocks;
  const double n = static_cast<double>(count);
  double cpart = 0.0;
  for (int c = threadIdx.x; c < d; c += 256) {
    const double p = psi[c];
    const double den = n * p + 1.0;
    const double a = n * p / den;
    const double v = 1.0 + p / den;
    s_s[c] = a / v;
    s_w[c] = a * a / v;
    s_q[c] = 0.5 * (1.0 / (1.0 + p) - 1.0 / v);
    if (is_enrol) cpart += log1p(p) - log(v);
  }
  cpart = warp_sum(cpart);
  if (lane == 0) s_red[warp] = cpart;
  __syncthreads();
  double cst = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) cst += s_red[i];
  if (is_enrol) {
    const long long r0 = static_cast<long long>(blockIdx.x) * 32;
    for (int i = warp; i < 32; i += 8) {
      const long long r = r0 + i;
      if (r >= ne) break;
      const T* src = enrol + r * ld_e;
      double acc = 0.0;
      for (int c = lane; c < ld_out; c += 32) {
        double lv = 0.0;
        if (c < d) {
          const double e = static_cast<double>(src[c]);
          lv = e * s_s[c];
          acc += s_w[c] * e * e;
        }
        store_split(l_hi, l_lo, r * ld_out + c, lv);
      }
      acc = warp_sum(acc);
      if (lane == 0) row_term[r] = static_cast<float>(0.5 * (cst - acc));
    }
  } else {
    const long long r0 = static_cast<long long>(blockIdx.x - enrol_blocks) * 32;
    for (int i = warp; i < 32; i += 8) {
      const long long r = r0 + i;
      if (r >= nt) {
        if (lane == 0 && r < col_ld) col_term[r] = 0.f;   // padding read (never stored) by the GEMM epilogue
        continue;
      }
      const T* src = test + r * ld_t;
      double acc = 0.0;
      for (int c = lane; c < ld_out; c += 32) {
        double tv = 0.0;
        if (c < d) {
          tv = static_cast<double>(src[c]);
          acc += s_q[c] * tv * tv;
        }
        store_split(r_hi, r_lo, r * ld_out + c, tv);
      }
      acc = warp_sum(acc);
      if (lane == 0) col_term[r] = static_cast<float>(acc);
    }
  }
}

__global__ void score_epilogue_f64_kernel(const double* __restrict__ gram, long long ne, long long nt,
                                          const double* __restrict__ row_term, const double* __restrict__ col_term,
                                          long long col_ld, const int32_t* __restrict__ grp,
                                          const float* __restrict__ zmean, const float* __restrict__ zinv,
                                          float* __restrict__ out, long long ldo, double* __restrict__ rsum,
                                          double* __restrict__ rsq) {
  // one block per enrol row; threads stride the columns
  const long long m = blockIdx.x;
  const int g = grp ? grp[m] : 0;
  const double ra = row_term[m];
  const double zm = zmean ? static_cast<double>(zmean[m]) : 0.0;
  const double zi = zinv ? static_cast<double>(zinv[m]) : 1.0;
  double s1 = 0.0, s2 = 0.0;
  for (long long n = threadIdx.x; n < nt; n += blockDim.x) {
    const double v = (gram[m * nt + n] + ra + col_term[g * col_ld + n] - zm) * zi;
    if (out) out[m * ldo + n] = static_cast<float>(v);
    s1 += v;
    s2 += v * v;
  }
  if (rsum) {
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(rsum + m, s1);
      atomicAdd(rsq + m, s2);
    }
  }
}

__global__ void znorm_finalize_kernel(const double* __restrict__ rsum, const double* __restrict__ rsq, long long ne,
                                      double m, float* __restrict__ zmean, float* __restrict__ zinv,
                                      double* __restrict__ mean_out, double* __restrict__ std_out) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= ne) return;
  const double mean = rsum[i] / m;
  double var = rsq[i] / m - mean * mean;     // population variance (src/pldamodule.cpp:246-250)
  if (var < 0.0) var = 0.0;
  const double sd = sqrt(var);
  if (zmean) zmean[i] = static_cast<float>(mean);
  if (zinv) zinv[i] = static_cast<float>(1.0 / sd);
  if (mean_out) mean_out[i] = mean;
  if (std_out) std_out[i] = sd;
}

// y *= sqrt(dim / sum y_i^2/(psi_i + 1/n))   (Plda::GetNormalizationFactor)
template <typename T>
__global__ void length_normalise_kernel(const T* __restrict__ y, long long rows, int dim, long long ld_y,
                                        const double* __restrict__ psi, const int32_t* __restrict__ counts,
                                        int const_count, double* __restrict__ out64, long long ld64,
                                        float* __restrict__ out32, long long ld32) {
  const long long r = blockIdx.x * static_cast<long long>(kWarpsPerBlock) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const double inv_n = 1.0 / static_cast<double>(counts ? counts[r] : const_count);
  const T* src = y + r * ld_y;
  double acc = 0.0;
  for (int c = lane; c < dim; c += 32) {
    const double v = static_cast<double>(src[c]);
    acc += v * v / (psi[c] + inv_n);
  }
  acc = warp_sum(acc);
  const double f = sqrt(static_cast<double>(dim) / acc);
  for (int c = lane; c < dim; c += 32) {
    const double v = static_cast<double>(src[c]) * f;
    if (out64) out64[r * ld64 + c] = v;
    if (out32) out32[r * ld32 + c] = static_cast<float>(v);
  }
}

inline unsigned row_blocks(int64_t rows) { return static_cast<unsigned>(ceil_div(rows, kWarpsPerBlock)); }

}  // namespace

void split_rows(Context& ctx, const void* in, bool is_f32, int64_t rows, int64_t cols, int64_t ld_in,
                const double* sub, const double* col_scale, const double* row_scale, SplitBuf& out) {
  out.reserve(rows, cols);
  if (rows == 0) return;
  if (is_f32)
    split_rows_kernel<float><<<row_blocks(rows), kWarpsPerBlock * 32, 0, ctx.stream>>>(
        static_cast<const float*>(in), rows, static_cast<int>(cols), ld_in, sub, col_scale, row_scale, out.hi.get(),
        out.lo.get(), static_cast<int>(out.ld));
  else
    split_rows_kernel<double><<<row_blocks(rows), kWarpsPerBlock * 32, 0, ctx.stream>>>(
        static_cast<const double*>(in), rows, static_cast<int>(cols), ld_in, sub, col_scale, row_scale, out.hi.get(),
        out.lo.get(), static_cast<int>(out.ld));
  PB_CUDA(cudaGetLastError());
  ctx.count_launch();
}

void convert_to_f64(Context& ctx, const void* in, bool is_f32, int64_t rows, int64_t cols, int64_t ld_in, double* out,
                    int64_t ld_out, const double* sub) {
  if (rows == 0) return;
  const long long total = rows * cols;
  const unsigned blocks = static_cast<unsigned>(ceil_div(total, 256));
  if (is_f32)
    convert_kernel<float><<<blocks, 256, 0, ctx.stream>>>(static_cast<const float*>(in), rows, static_cast<int>(cols),
                                                          ld_in, sub, out, ld_out);
  else
    convert_kernel<double><<<blocks, 256, 0, ctx.stream>>>(static_cast<const double*>(in), rows,
                                                           static_cast<int>(cols), ld_in, sub, out, ld_out);
  PB_CUDA(cudaGetLastError());
  ctx.count_launch();
}

void convert_f64_to_f32(Context& ctx, const double* in, int64_t rows, int64_t cols, int64_t ld_in, float* out,
                        int64_t ld_out) {
  if (rows == 0) return;
  f64_to_f32_kernel<<<static_cast<unsigned>(ceil_div(rows * cols, 256)), 256, 0, ctx.stream>>>(
      in, rows, static_cast<int>(cols), ld_in, out, ld_out);
  PB_CUDA(cudaGetLastError());
  ctx.count_launch();
}

void score_prep_enrol(Context& ctx, const void* enrol, bool is_f32, int64_t ne, int64_t d, int64_t ld,
                      const int32_t* counts, int const_count, const double* psi, SplitBuf* l_out, double* l_f64, float* row_term,
                      double* row_term_f64) {
  if (l_out) l_out->reserve(ne, d);
  __nv_bfloat16* hi = l_out ? l_out->hi.get() : nullptr;
  __nv_bfloat16* lo = l_out ? l_out->lo.get() : nullptr;
  const int ldo = l_out ? static_cast<int>(l_out->ld) : 0;
  if (is_f32)
    score_prep_enrol_kernel<float><<<row_blocks(ne), kWarpsPerBlock * 32, 0, ctx.stream>>>(
        static_cast<const float*>(enrol), ne, static_cast<int>(d), ld, counts, const_count, psi, hi, lo, ldo, l_f64, row_term,
        row_term_f64);
  else
    score_prep_enrol_kernel<double><<<row_blocks(ne), kWarpsPerBlock * 32, 0, ctx.stream>>>(
        static_cast<const double*>(enrol), ne, static_cast<int>(d), ld, counts, const_count, psi, hi, lo, ldo, l_f64, row_term,
        row_term_f64);
  PB_CUDA(cudaGetLastError());
  ctx.count_launch();
}

void score_prep_test(Context& ctx, const void* test, bool is_f32, int64_t nt, int64_t d, int64_t ld,
                     const int32_t* group_counts, int ngroups, int const_count, const double* psi, SplitBuf* r_out, float* col_term,
                     int64_t col_ld, double* col_term_f64) {
  if (r_out) r_out->reserve(nt, d);
  __nv_bfloat16* hi = r_out ? r_out->hi.get() : nullptr;
  __nv_bfloat16* lo = r_out ? r_out->lo.get() : nullptr;
  const int ldo = r_out ? static_cast<int>(r_out->ld) : 0;
  if (is_f32)
    score_prep_test_kernel<float><<<row_blocks(nt), kWarpsPerBlock * 32, 0, ctx.stream>>>(
        static_cast<const float*>(test), nt, static_cast<int>(d), ld, group_counts, ngroups, const_count, psi, hi, lo, ldo,
        col_term, col_ld, col_term_f64);
  else
    score_prep_test_kernel<double><<<row_blocks(nt), kWarpsPerBlock * 32, 0, ctx.stream>>>(
        static_cast<const double*>(test), nt, static_cast<int>(d), ld, group_counts, ngroups, const_count, psi, hi, lo, ldo,
        col_term, col_ld, col_term_f64);
  PB_CUDA(cudaGetLastError());
  ctx.count_launch();
}

void score_prep_uniform(Context& ctx, const void* enrol, int64_t ne, int64_t ld_e, const void* test, int64_t nt,
                        int64_t ld_t, bool is_f32, int64_t d, int count, const double* psi, SplitBuf& l_out,
                        SplitBuf& r_out, float* row_term, float* col_term, int64_t col_ld) {
  PB_CHECK(d <= 1024, kInvalidArg, "score: dimension above 1024 is not supported");
  l_out.reserve(ne, d);
  r_out.reserve(nt, d);
  const unsigned eb = static_cast<unsigned>(ceil_div(ne, 32));
  const unsigned tb = static_cast<unsigned>(ceil_div(col_ld, 32));
  if (is_f32)
    score_prep_uniform_kernel<float><<<eb + tb, 256, 0, ctx.stream>>>(
        static_cast<const float*>(enrol), ne, ld_e, static_cast<const float*>(test), nt, ld_t, static_cast<int>(d),
        count, psi, l_out.hi.get(), l_out.lo.get(), r_out.hi.get(), r_out.lo.get(), static_cast<int>(l_out.ld),
        row_term, col_term, col_ld, eb);
  else
    score_prep_uniform_kernel<double><<<eb + tb, 256, 0, ctx.stream>>>(
        static_cast<const double*>(enrol), ne, ld_e, static_cast<const double*>(test), nt, ld_t, static_cast<int>(d),
        count, psi, l_out.hi.get(), l_out.lo.get(), r_out.hi.get(), r_out.lo.get(), static_cast<int>(l_out.ld),
        row_term, col_term, col_ld, eb);
  PB_CUDA(cudaGetLastError());
  ctx.count_launch();
}

void score_epilogue_f64(Context& ctx, const double* gram, int64_t ne, int64_t nt, const double* row_term,
                        const double* col_term, int64_t col_ld, const int32_t* grp, const float* zmean,
                        const float* zinv, float* out, int64_t ldo, double* rsum, double* rsq) {
  if (ne == 0 || nt == 0) return;
  score_epilogue_f64_kernel<<<static_cast<unsigned>(ne), 256, 0, ctx.stream>>>(gram, ne, nt, row_term, col_term, col_ld,
                                                                              grp, zmean, zinv, out, ldo, rsum, rsq);
  PB_CUDA(cudaGetLastError());
  ctx.count_launch();
}

void znorm_finalize(Context& ctx, const double* rsum, const double* rsq, int64_t ne, int64_t m, float* zmean,
                    float* zinv, double* mean_out, double* std_out) {
  if (ne == 0) return;
  znorm_finalize_kernel<<<static_cast<unsigned>(ceil_div(ne, 256)), 256, 0, ctx.stream>>>(
      rsum, rsq, ne, static_cast<double>(m), zmean, zinv, mean_out, std_out);
  PB_CUDA(cudaGetLastError());
  ctx.count_launch();
}

void length_normalise(Context& ctx, const void* y, bool y_is_f32, int64_t rows, int64_t dim, int64_t ld_y,
                      const double* psi, const int32_t* counts, int32_t const_count, double* out64, int64_t ld64,
                      float* out32, int64_t ld32) {
  if (rows == 0) return;
  if (y_is_f32)
    length_normalise_kernel<float><<<row_blocks(rows), kWarpsPerBlock * 32, 0, ctx.stream>>>(
        static_cast<const float*>(y), rows, static_cast<int>(dim), ld_y, psi, counts, const_count, out64, ld64, out32,
        ld32);
  else
    length_normalise_kernel<double><<<row_blocks(rows), kWarpsPerBlock * 32, 0, ctx.stream>>>(
        static_cast<const double*>(y), rows, static_cast<int>(dim), ld_y, psi, counts, const_count, out64, ld64, out32,
        ld32);
  PB_CUDA(cudaGetLastError());
  ctx.count_launch();
}

}  // namespace pb
