// K7: element-wise part of the diagonalised EM iteration.
//
// Kaldi's PldaEstimator::GetStatsFromClassMeans (called through src/pldamodule.cpp:106) loops over the
// classes with d x d packed-matrix updates.  In the basis A that jointly diagonalises (W, B)
// (A W A^T = I, A B A^T = diag(psi)) the same statistics are (proved equal to
// Kaldi's loop to 1e-11 in tests/test_oracle_plda.py::test_diagonalised_em_equals_kaldi_loop):
//     u_s = A (m_s - mu);  r = psi/(1 + n_s psi);  g = n_s r
//     A B_stats A^T        = diag(sum_s w_s r)      + sum_s w_s     (g u_s)(g u_s)^T
//     A (W_stats - S) A^T  = diag(sum_s w_s n_s r)  + sum_s w_s n_s ((1-g) u_s)((1-g) u_s)^T
// with the reference's class weights w_s = 1/n_s (src/pldamodule.cpp:97).
// This file produces the two SYRK operands P = sqrt(w) g u, Q = sqrt(w n)(1-g) u (transposed split-bf16,
// K-major along the class axis) and the two diagonals; the SYRKs themselves run on tcgen05 (gemm_tc.cu).
#include "kernels.h"

namespace pb {
namespace {

// tile: 32 classes x 32 dims, transposed through smem (same pattern as center_scale_split_t)
__global__ void __launch_bounds__(256)
em_posterior_t_kernel(const float* __restrict__ u, long long ldu, long long k, int d,
                      const int32_t* __restrict__ counts, const double* __restrict__ psi,
                      __nv_bfloat16* __restrict__ p_hi, __nv_bfloat16* __restrict__ p_lo,
                      __nv_bfloat16* __restrict__ q_hi, __nv_bfloat16* __restrict__ q_lo, long long ld_out,
                      double* __restrict__ db, double* __restrict__ dw) {
  __shared__ float sp_hi[32][33], sp_lo[32][33], sq_hi[32][33], sq_lo[32][33];
  __shared__ double s_db[8][33], s_dw[8][33];
  const long long s0 = blockIdx.x * 32ll;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = c0 + tx;
  const double ps = c < d ? psi[c] : 0.0;
  double acc_b = 0.0, acc_w = 0.0;
  for (int i = ty; i < 32; i += 8) {
    const long long s = s0 + i;
    double pv = 0.0, qv = 0.0;
    if (s < k && c < d) {
      const double n = static_cast<double>(counts[s]);
      const double w = 1.0 / n;
      const double r = ps / (1.0 + n * ps);
      const double g = n * r;
      const double uv = static_cast<double>(u[s * ldu + c]);
      pv = sqrt(w) * g * uv;
      qv = (1.0 - g) * uv;            // sqrt(w n) = 1 with w = 1/n
      acc_b += w * r;
      acc_w += r;                      // w n r = r
    }
    __nv_bfloat16 h, l;
    split_bf16(pv, h, l);
    sp_hi[i][tx] = __bfloat162float(h);
    sp_lo[i][tx] = __bfloat162float(l);
    split_bf16(qv, h, l);
    sq_hi[i][tx] = __bfloat162float(h);
    sq_lo[i][tx] = __bfloat162float(l);
  }
  s_db[ty][tx] = acc_b;
  s_dw[ty][tx] = acc_w;
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int cc = c0 + i;
    const long long s = s0 + tx;
    if (cc < d && s < ld_out) {
      p_hi[cc * ld_out + s] = __float2bfloat16_rn(sp_hi[tx][i]);
      p_lo[cc * ld_out + s] = __float2bfloat16_rn(sp_lo[tx][i]);
      q_hi[cc * ld_out + s] = __float2bfloat16_rn(sq_hi[tx][i]);
      q_lo[cc * ld_out + s] = __float2bfloat16_rn(sq_lo[tx][i]);
    }
  }
  if (ty == 0 && c < d) {
    double tb = 0.0, tw = 0.0;
    for (int i = 0; i < 8; ++i) { tb += s_db[i][tx]; tw += s_dw[i][tx]; }
    atomicAdd(db + c, tb);
    atomicAdd(dw + c, tw);
  }
}

__global__ void __launch_bounds__(256)
em_posterior_f64_kernel(const double* __restrict__ u, long long k, int d, const int32_t* __restrict__ counts,
                        const double* __restrict__ psi, double* __restrict__ p, double* __restrict__ q,
                        double* __restrict__ db, double* __restrict__ dw) {
  __shared__ double s_db[8][33], s_dw[8][33];
  const long long s0 = blockIdx.x * 32ll;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = c0 + tx;
  const double ps = c < d ? psi[c] : 0.0;
  double acc_b = 0.0, acc_w = 0.0;
  for (int i = ty; i < 32; i += 8) {
    const long long s = s0 + i;
    if (s < k && c < d) {
      const double n = static_cast<double>(counts[s]);
      const double w = 1.0 / n;
      const double r = ps / (1.0 + n * ps);
      const double g = n * r;
      const double uv = u[s * d + c];
      p[s * d + c] = sqrt(w) * g * uv;
      q[s * d + c] = (1.0 - g) * uv;
      acc_b += w * r;
      acc_w += r;
    }
  }
  s_db[ty][tx] = acc_b;
  s_dw[ty][tx] = acc_w;
  __syncthreads();
  if (ty == 0 && c < d) {
    double tb = 0.0, tw = 0.0;
    for (int i = 0; i < 8; ++i) { tb += s_db[i][tx]; tw += s_dw[i][tx]; }
    atomicAdd(db + c, tb);
    atomicAdd(dw + c, tw);
  }
}

__global__ void add_diag_scale_kernel(double* __restrict__ x, const double* __restrict__ dg, int d, double scale,
                                      const double* __restrict__ base) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<long long>(d) * d) return;
  const int i = static_cast<int>(idx / d), j = static_cast<int>(idx % d);
  double v = x[idx];
  if (dg && i == j) v += dg[i];
  v *= scale;
  if (base) v += base[idx];
  x[idx] = v;
}

// B_stats / W_stats in the diagonalising basis from the split-K partials of ONE stacked SYRK ([P ; Q] [P ; Q]^T,
// 2d x 2d: the diagonal blocks are P^T P and Q^T Q), symmetrised, plus the diagonal terms -- one launch instead of
// two reductions and two diagonal updates.
__global__ void em_stats_reduce_kernel(const float* __restrict__ partial, int ksplit, int d, long long mpad,
                                       long long npad, const double* __restrict__ db, const double* __restrict__ dw,
                                       double* __restrict__ bs, double* __restrict__ ws) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<long long>(d) * d) return;
  const int i = static_cast<int>(idx / d), j = static_cast<int>(idx % d);
  double sb = 0.0, sw = 0.0;
  for (int k = 0; k < ksplit; ++k) {
    const float* pl = partial + k * mpad * npad;
    sb += static_cast<double>(pl[i * npad + j]) + static_cast<double>(pl[j * npad + i]);
    sw += static_cast<double>(pl[(d + i) * npad + d + j]) + static_cast<double>(pl[(d + j) * npad + d + i]);
  }
  sb *= 0.5;
  sw *= 0.5;
  if (i == j) { sb += db[i]; sw += dw[i]; }
  bs[idx] = sb;
  ws[idx] = sw;
}

// EstimateFromStats on the back-transformed statistics: B = sym(B) / B_count ; W = (sym(W) + S) / W_count
__global__ void em_finalize_kernel(double* __restrict__ between, double* __restrict__ within,
                                   const double* __restrict__ scatter, double inv_b, double inv_w, int d) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<long long>(d) * d) return;
  const int i = static_cast<int>(idx / d), j = static_cast<int>(idx % d);
  if (j > i) return;
  const long long tr = static_cast<long long>(j) * d + i;
  const double b = 0.5 * (between[idx] + between[tr]) * inv_b;
  const double w = (0.5 * (within[idx] + within[tr]) + 0.5 * (scatter[idx] + scatter[tr])) * inv_w;
  between[idx] = b;
  between[tr] = b;
  within[idx] = w;
  within[tr] = w;
}

__global__ void set_identity_kernel(double* __restrict__ a, int d) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<long long>(d) * d) return;
  a[idx] = (idx / d == idx % d) ? 1.0 : 0.0;
}

}  // namespace

void em_posterior_t(Context& ctx, const float* u, int64_t ldu, int64_t k, int64_t d, const int32_t* counts,
                    const double* psi, SplitBuf& pt, SplitBuf& qt, double* db, double* dw) {
  const int64_t kpad = round_up(k, 64);
  for (SplitBuf* b : {&pt, &qt}) {
    b->rows = d;
    b->k = k;
    b->ld = kpad;
    b->hi.reserve(static_cast<size_t>(d) * kpad);
    b->lo.reserve(static_cast<size_t>(d) * kpad);
  }
  PB_CUDA(cudaMemsetAsync(db, 0, d * sizeof(double), ctx.stream));
  PB_CUDA(cudaMemsetAsync(dw, 0, d * sizeof(double), ctx.stream));
  dim3 grid(static_cast<unsigned>(kpad / 32), static_cast<unsigned>(ceil_div(d, 32)));
  em_posterior_t_kernel<<<grid, 256, 0, ctx.stream>>>(u, ldu, k, static_cast<int>(d), counts, psi, pt.hi.get(),
                                                      pt.lo.get(), qt.hi.get(), qt.lo.get(), kpad, db, dw);
  PB_CUDA(cudaGetLastError());
  ctx.count_launch();
}

void em_posterior_stacked(Context& ctx, const float* u, int64_t ldu, int64_t k, int64_t d, const int32_t* counts,
                          const double* psi, SplitBuf& pq, double* db, double* dw) {
  const int64_t kpad = round_up(k, 64);
  pq.rows = 2 * d;
  pq.k = k;
  pq.ld = kpad;
  pq.hi.reserve(static_cast<size_t>(2 * d) * kpad);
  pq.lo.reserve(static_cast<size_t>(2 * d) * kpad);
  PB_CUDA(cudaMemsetAsync(db, 0, d * sizeof(double), ctx.stream));
  PB_CUDA(cudaMemsetAsync(dw, 0, d * sizeof(double), ctx.stream));
  dim3 grid(static_cast<unsigned>(kpad / 32), static_cast<unsigned>(ceil_div(d, 32)));
  em_posterior_t_kernel<<<grid, 256, 0, ctx.stream>>>(u, ldu, k, static_cast<int>(d), counts, psi, pq.hi.get(),
                                                      pq.lo.get(), pq.hi.get() + d * kpad, pq.lo.get() + d * kpad, kpad,
                                                      db, dw);
  PB_CUDA(cudaGetLastError());
  ctx.count_launch();
}

void em_stats_reduce(Context& ctx, const float* partial, int ksplit, int64_t d, const double* db, const double* dw,
                     double* bs, double* ws) {
  const long long mpad = round_up(2 * d, 128), npad = round_up(2 * d, 4);
  em_stats_reduce_kernel<<<static_cast<unsigned>(ceil_div(d * d, 256)), 256, 0, ctx.stream>>>(
      partial, ksplit, static_cast<int>(d), mpad, npad, db, dw, bs, ws);
  PB_CUDA(cudaGetLastError());
  ctx.count_launch();
}

void em_finalize(Context& ctx, double* between, double* within, const double* scatter, double inv_b, double inv_w,
                 int64_t d) {
  em_finalize_kernel<<<static_cast<unsigned>(ceil_div(d * d, 256)), 256, 0, ctx.stream>>>(between, within, scatter, inv_b,
                                                                                         inv_w, static_cast<int>(d));
  PB_CUDA(cudaGetLastError());
  ctx.count_launch();
}

void em_posterior_f64(Context& ctx, const double* u, int64_t k, int64_t d, const int32_t* counts, const double* psi,
                      double* p, double* q, double* db, double* dw) {
  PB_CUDA(cudaMemsetAsync(db, 0, d * sizeof(double), ctx.stream));
  PB_CUDA(cudaMemsetAsync(dw, 0, d * sizeof(double), ctx.stream));
  dim3 grid(static_cast<unsigned>(ceil_div(k, 32)), static_cast<unsigned>(ceil_div(d, 32)));
  em_posterior_f64_kernel<<<grid, 256, 0, ctx.stream>>>(u, k, static_cast<int>(d), counts, psi, p, q, db, dw);
  PB_CUDA(cudaGetLastError());
  ctx.count_launch();
}

void add_diag_scale(Context& ctx, double* x, const double* dg, int64_t d, double scale, const double* base) {
  add_diag_scale_kernel<<<static_cast<unsigned>(ceil_div(d * d, 256)), 256, 0, ctx.stream>>>(x, dg, static_cast<int>(d),
                                                                                           scale, base);
  PB_CUDA(cudaGetLastError());
  ctx.count_launch();
}

void set_identity(Context& ctx, double* a, int64_t d) {
  set_identity_kernel<<<static_cast<unsigned>(ceil_div(d * d, 256)), 256, 0, ctx.stream>>>(a, static_cast<int>(d));
  PB_CUDA(cudaGetLastError());
  ctx.count_launch();
}

}  // namespace pb
