// K8: fp64 d x d factorizations for d <= 1024 -- Cholesky, lower-triangular inverse and a
// symmetric eigensolver (one-sided Jacobi).  They replace Kaldi's TpMatrix::Cholesky /
// TpMatrix::Invert / SpMatrix::Eig inside PldaEstimator::GetOutput / ComputeNormalizingTransform
// (reached from src/pldamodule.cpp:106) and are also what lets the EM iteration run in the jointly
// diagonalising basis.  They are latency-bound (no meaningful roofline); the design goal is a small,
// fixed number of device-wide synchronisations, all matrices L2-resident.
#include <cooperative_groups.h>

#include "kernels.h"

namespace pb {
namespace {

// ------------------------------------------------------------------------- //
// Cholesky: one CTA, right-looking, column j staged in smem, trailing rows updated warp-per-row.
// ------------------------------------------------------------------------- //
__global__ void __launch_bounds__(1024)
cholesky_kernel(double* __restrict__ a, int d, int* __restrict__ info) {
  extern __shared__ double col[];   // d doubles
  __shared__ double s_piv;
  const int tid = threadIdx.x, nthreads = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31, nwarps = nthreads >> 5;
  if (tid == 0) *info = 0;
  for (int j = 0; j < d; ++j) {
    if (tid == 0) s_piv = a[static_cast<long long>(j) * d + j];
    __syncthreads();
    const double ajj = s_piv;
    if (!(ajj > 0.0)) {
      if (tid == 0) *info = j + 1;
      return;
    }
    const double ljj = sqrt(ajj);
    const double inv = 1.0 / ljj;
    for (int i = j + tid; i < d; i += nthreads) {
      const double v = (i == j) ? ljj : a[static_cast<long long>(i) * d + j] * inv;
      a[static_cast<long long>(i) * d + j] = v;
      col[i] = v;
    }
    __syncthreads();
    // a[i][k] -= l[i] * l[k]  for j < k <= i
    for (int i = j + 1 + warp; i < d; i += nwarps) {
      const double li = col[i];
      double* row = a + static_cast<long long>(i) * d;
      for (int k = j + 1 + lane; k <= i; k += 32) row[k] -= li * col[k];
    }
    __syncthreads();
  }
  // zero the strict upper triangle
  for (long long idx = tid; idx < static_cast<long long>(d) * d; idx += nthreads) {
    const int i = static_cast<int>(idx / d), k = static_cast<int>(idx % d);
    if (k > i) a[idx] = 0.0;
  }
}

// ------------------------------------------------------------------------- //
// inv = L^-1: warp j solves L x = e_j by forward substitution (x in smem), writes column j.
// ------------------------------------------------------------------------- //
__global__ void __launch_bounds__(256)
tri_inverse_kernel(const double* __restrict__ l, double* __restrict__ inv, int d) {
  extern __shared__ double xs[];   // 8 warps x d
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * 8 + warp;
  if (j >= d) return;
  double* x = xs + static_cast<long long>(warp) * d;
  for (int i = lane; i < j; i += 32) inv[static_cast<long long>(i) * d + j] = 0.0;
  for (int i = j; i < d; ++i) {
    const double* row = l + static_cast<long long>(i) * d;
    double acc = 0.0;
    for (int k = j + lane; k < i; k += 32) acc += row[k] * x[k];
    acc = warp_sum(acc);
    const double v = ((i == j ? 1.0 : 0.0) - acc) / row[i];
    if (lane == 0) x[i] = v;
    __syncwarp();
  }
  for (int i = j + lane; i < d; i += 32) inv[static_cast<long long>(i) * d + j] = x[i];
}

// ------------------------------------------------------------------------- //
// One-sided (Hestenes) Jacobi on a symmetric matrix.
//   gt[p,:] = column p of G = B V ;  vt[p,:] = column p of V   (rows are contiguous)
// Each round pairs rows by the round-robin tournament; a warp owns one pair.  Rounds are
// separated by a device-wide barrier (cooperative launch guarantees co-residency).
// ------------------------------------------------------------------------- //
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int& target, unsigned int nblocks) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += nblocks;
    __threadfence();
    atomicAdd(counter, 1u);
    const uint64_t t0 = global_timer_ns();
    unsigned int spins = 0;
    while (atomicAdd(counter, 0u) < target) {
      if ((++spins & 0xfff) == 0 && global_timer_ns() - t0 > 5000000000ull) {
        printf("plda_b200: grid barrier timeout (block %d)\n", blockIdx.x);
        __trap();
      }
    }
    __threadfence();
  }
  __syncthreads();
}

template <int EPL>   // elements per lane: d <= 32*EPL
__global__ void __launch_bounds__(256)
jacobi_kernel(double* __restrict__ gt, double* __restrict__ vt, int d, double tol, double abs_tol, int max_sweeps,
              unsigned int* __restrict__ barrier_counter, int* __restrict__ rotated, int* __restrict__ sweeps_done) {
  const int warp_in_grid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int m = d + (d & 1);
  const int npairs = m >> 1;
  unsigned int target = 0;
  int sweep = 0;
  for (; sweep < max_sweeps; ++sweep) {
    for (int r = 0; r < m - 1; ++r) {
      if (warp_in_grid < npairs) {
        int p, q;
        if (warp_in_grid == 0) { p = m - 1; q = r; }
        else { p = (r + warp_in_grid) % (m - 1); q = (r - warp_in_grid + (m - 1)) % (m - 1); }
        if (p > q) { const int t = p; p = q; q = t; }
        if (q < d) {   // (odd d: the padded index sits out)
          double gp[EPL], gq[EPL];
          double alpha = 0.0, beta = 0.0, gamma = 0.0;
          double* rp = gt + static_cast<long long>(p) * d;
          double* rq = gt + static_cast<long long>(q) * d;
#pragma unroll
          for (int j = 0; j < EPL; ++j) {
            const int i = lane + 32 * j;
            gp[j] = i < d ? __ldcg(rp + i) : 0.0;
            gq[j] = i < d ? __ldcg(rq + i) : 0.0;
            alpha += gp[j] * gp[j];
            beta += gq[j] * gq[j];
            gamma += gp[j] * gq[j];
          }
          alpha = warp_sum(alpha);
          beta = warp_sum(beta);
          gamma = warp_sum(gamma);
          const double lim = tol * sqrt(alpha * beta);
          if (fabs(gamma) > lim && fabs(gamma) > abs_tol) {
            const double zeta = (beta - alpha) / (2.0 * gamma);
            const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
            const double c = 1.0 / sqrt(1.0 + t * t);
            const double s = c * t;
#pragma unroll
            for (int j = 0; j < EPL; ++j) {
              const int i = lane + 32 * j;
              if (i < d) {
                __stcg(rp + i, c * gp[j] - s * gq[j]);
                __stcg(rq + i, s * gp[j] + c * gq[j]);
              }
            }
            double* vp = vt + static_cast<long long>(p) * d;
            double* vq = vt + static_cast<long long>(q) * d;
#pragma unroll
            for (int j = 0; j < EPL; ++j) {
              const int i = lane + 32 * j;
              if (i < d) {
                const double a = __ldcg(vp + i), b = __ldcg(vq + i);
                __stcg(vp + i, c * a - s * b);
                __stcg(vq + i, s * a + c * b);
              }
            }
            if (lane == 0) atomicOr(rotated + sweep, 1);
          }
        }
      }
      grid_barrier(barrier_counter, target, gridDim.x);
    }
    // every block reads the flag after the barrier that closed the sweep
    const int any = *reinterpret_cast<volatile int*>(rotated + sweep);
    if (!any) { ++sweep; break; }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *sweeps_done = sweep;
}

// lam[p] = v_p . g_p (Rayleigh quotient), rank by descending value, write sorted rows.
__global__ void __launch_bounds__(256)
eig_lambda_kernel(const double* __restrict__ gt, const double* __restrict__ vt, int d, double* __restrict__ lam) {
  const int p = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (p >= d) return;
  double acc = 0.0;
  for (int i = lane; i < d; i += 32) acc += gt[static_cast<long long>(p) * d + i] * vt[static_cast<long long>(p) * d + i];
  acc = warp_sum(acc);
  if (lane == 0) lam[p] = acc;
}

__global__ void eig_sort_kernel(const double* __restrict__ lam, const double* __restrict__ vt, int d,
                                double* __restrict__ evals, double* __restrict__ evecs_t) {
  // one block per source row p: rank = #{q : lam[q] > lam[p] or (== and q < p)}
  __shared__ int s_rank;
  const int p = blockIdx.x;
  if (threadIdx.x == 0) s_rank = 0;
  __syncthreads();
  const double lp = lam[p];
  int cnt = 0;
  for (int q = threadIdx.x; q < d; q += blockDim.x) {
    const double lq = lam[q];
    if (lq > lp || (lq == lp && q < p)) ++cnt;
  }
  atomicAdd(&s_rank, cnt);
  __syncthreads();
  const int rk = s_rank;
  if (threadIdx.x == 0) evals[rk] = lp > 0.0 ? lp : 0.0;      // ApplyFloor(0.0)
  for (int i = threadIdx.x; i < d; i += blockDim.x)
    evecs_t[static_cast<long long>(rk) * d + i] = vt[static_cast<long long>(p) * d + i];
}

__global__ void frob2_kernel(const double* __restrict__ a, long long n, double* __restrict__ out) {
  double acc = 0.0;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    acc += a[i] * a[i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

}  // namespace

void cholesky_lower(Context& ctx, double* a, int64_t d, int* info_dev) {
  PB_CHECK(d > 0 && d <= 4096, kInvalidArg, "cholesky: dimension out of range");
  cholesky_kernel<<<1, 1024, d * sizeof(double), ctx.stream>>>(a, static_cast<int>(d), info_dev);
  PB_CUDA(cudaGetLastError());
  ctx.count_launch();
}

void tri_inverse_lower(Context& ctx, const double* l, double* inv, int64_t d) {
  const size_t smem = 8 * d * sizeof(double);
  static std::once_flag once;
  std::call_once(once, [] {
    cudaFuncSetAttribute(tri_inverse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 1024 * sizeof(double));
  });
  PB_CHECK(d <= 1024, kInvalidArg, "tri_inverse: d <= 1024");
  tri_inverse_kernel<<<static_cast<unsigned>(ceil_div(d, 8)), 256, smem, ctx.stream>>>(l, inv, static_cast<int>(d));
  PB_CUDA(cudaGetLastError());
  ctx.count_launch();
}

void eig_sym_jacobi(Context& ctx, const double* b, int64_t d, const double* v0_t, double* evals, double* evecs_t,
                    EigWork& w, int* sweeps_out) {
  PB_CHECK(d > 0 && d <= 1024, kInvalidArg, "eig: d <= 1024");
  const int max_sweeps = 40;
  w.g.reserve(d * d);
  w.v.reserve(d * d);
  w.lam.reserve(d);
  w.tmp.reserve(1);
  w.flags.reserve(max_sweeps + 4);   // [0] barrier counter, [1] sweeps_done, [2..] rotated flags
  PB_CUDA(cudaMemsetAsync(w.flags.get(), 0, (max_sweeps + 4) * sizeof(int), ctx.stream));
  if (v0_t != nullptr) {
    // warm start: V = V0, G^T = V0^T B  (B symmetric)
    PB_CUDA(cudaMemcpyAsync(w.v.get(), v0_t, d * d * sizeof(double), cudaMemcpyDeviceToDevice, ctx.stream));
    gemm_f64(ctx, false, false, d, d, d, 1.0, v0_t, d, b, d, 0.0, w.g.get(), d);
  } else {
    PB_CUDA(cudaMemcpyAsync(w.g.get(), b, d * d * sizeof(double), cudaMemcpyDeviceToDevice, ctx.stream));
    set_identity(ctx, w.v.get(), d);
  }
  // absolute floor on |g_p . g_q| relative to ||B||_F^2 (noise level of near-null directions)
  PB_CUDA(cudaMemsetAsync(w.tmp.get(), 0, sizeof(double), ctx.stream));
  frob2_kernel<<<32, 256, 0, ctx.stream>>>(b, d * d, w.tmp.get());
  ctx.count_launch();
  double frob2 = 0.0;
  PB_CUDA(cudaMemcpyAsync(&frob2, w.tmp.get(), sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
  ctx.sync();
  double tol = 1e-14, abs_tol = 1e-26 * frob2;
  int di = static_cast<int>(d);
  const int m = di + (di & 1);
  const int npairs = m / 2;
  int blocks = static_cast<int>(ceil_div(npairs, 8));
  PB_CHECK(blocks <= ctx.num_sms, kInvalidArg, "eig: too many blocks for a cooperative launch");
  double* gp = w.g.get();
  double* vp = w.v.get();
  unsigned int* counter = reinterpret_cast<unsigned int*>(w.flags.get());
  int* sweeps_done = w.flags.get() + 1;
  int* rotated = w.flags.get() + 2;
  int ms = max_sweeps;
  void* args[] = {&gp, &vp, &di, &tol, &abs_tol, &ms, &counter, &rotated, &sweeps_done};
  const void* fn = d <= 256 ? reinterpret_cast<const void*>(jacobi_kernel<8>)
                 : d <= 512 ? reinterpret_cast<const void*>(jacobi_kernel<16>)
                            : reinterpret_cast<const void*>(jacobi_kernel<32>);
  PB_CUDA(cudaLaunchCooperativeKernel(fn, dim3(blocks), dim3(256), args, 0, ctx.stream));
  ctx.count_launch();
  eig_lambda_kernel<<<static_cast<unsigned>(ceil_div(d, 8)), 256, 0, ctx.stream>>>(w.g.get(), w.v.get(), di,
                                                                                  w.lam.get());
  eig_sort_kernel<<<static_cast<unsigned>(d), 128, 0, ctx.stream>>>(w.lam.get(), w.v.get(), di, evals, evecs_t);
  PB_CUDA(cudaGetLastError());
  ctx.count_launch(2);
  if (sweeps_out) {
    PB_CUDA(cudaMemcpyAsync(sweeps_out, sweeps_done, sizeof(int), cudaMemcpyDeviceToHost, ctx.stream));
    ctx.sync();
  }
}

}  // namespace pb
