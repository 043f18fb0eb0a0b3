// K8: fp64 d x d factorizations for d <= 1024 -- Cholesky, lower-triangular inverse and a
// symmetric eigensolver (one-sided Jacobi).  They replace Kaldi's TpMatrix::Cholesky /
// TpMatrix::Invert / SpMatrix::Eig inside PldaEstimator::GetOutput / ComputeNormalizingTransform
// (reached from src/pldamodule.cpp:106) and are also what lets the EM iteration run in the jointly
// diagonalising basis.  They are latency-bound (no meaningful roofline); the design goal is a small,
// fixed number of device-wide synchronisations, all matrices L2-resident.
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "kernels.h"

namespace pb {
namespace {

// ------------------------------------------------------------------------- //
// Blocked (left-looking) Cholesky, panel width 32: per panel
//   (1) gemm_f64:  A[k0:, panel] -= L[k0:, :k0] L[panel, :k0]^T          (all SMs)
//   (2) chol_diag_kernel:  factor the 32 x 32 diagonal block in shared memory (one CTA, 1024 threads)
//   (3) chol_solve_kernel: rows below, X Lkk^T = A[rows, panel]            (32 rows per CTA)
// ------------------------------------------------------------------------- //
__global__ void __launch_bounds__(1024)
chol_diag_kernel(double* __restrict__ a, int d, int k0, int nb, int* __restrict__ info) {
  __shared__ double s[32][33];
  __shared__ int s_fail;
  const int i = threadIdx.x >> 5, k = threadIdx.x & 31;   // thread (i, k) owns s[i][k]
  if (threadIdx.x == 0) s_fail = 0;
  s[i][k] = (i < nb && k < nb) ? a[static_cast<long long>(k0 + i) * d + k0 + k] : 0.0;
  __syncthreads();
  for (int j = 0; j < nb; ++j) {
    const double ajj = s[j][j];
    if (!(ajj > 0.0)) {
      if (threadIdx.x == 0) { *info = k0 + j + 1; }
      return;                       // uniform: every thread read the same pivot
    }
    const double inv = rsqrt(ajj);
    __syncthreads();
    if (k == j && i >= j) s[i][j] = (i == j) ? sqrt(ajj) : s[i][j] * inv;
    __syncthreads();
    if (k > j && i >= k) s[i][k] -= s[i][j] * s[k][j];
    __syncthreads();
  }
  if (i < nb && k < nb) a[static_cast<long long>(k0 + i) * d + k0 + k] = k <= i ? s[i][k] : 0.0;
}

__global__ void __launch_bounds__(32)
chol_solve_kernel(double* __restrict__ a, int d, int k0, int nb) {
  __shared__ double lkk[32][33];
  const int t = threadIdx.x;
  for (int j = 0; j < nb; ++j) lkk[j][t] = t < nb ? a[static_cast<long long>(k0 + j) * d + k0 + t] : 0.0;
  __syncwarp();
  const int r = k0 + nb + blockIdx.x * 32 + t;
  if (r >= d) return;
  double* row = a + static_cast<long long>(r) * d + k0;
  double x[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    if (j < nb) {
      double v = row[j];
#pragma unroll
      for (int q = 0; q < 32; ++q)
        if (q < j) v -= x[q] * lkk[j][q];
      x[j] = v / lkk[j][j];
    }
  }
#pragma unroll
  for (int j = 0; j < 32; ++j)
    if (j < nb) row[j] = x[j];
}

__global__ void zero_upper_kernel(double* __restrict__ a, int d) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<long long>(d) * d) return;
  if (idx % d > idx / d) a[idx] = 0.0;
}

// ------------------------------------------------------------------------- //
// Cholesky + triangular inverse of a d x d SPD matrix (d <= 512) in ONE launch: a single thread-block cluster, CTA i
// owns the 32-row block i.  Left-looking by block column k:
//     every CTA i >= k :  S_i = A[i,k] - sum_{p<k} L[i,p] L[k,p]^T            (32 x 32 tile, kept in shared memory)
//     CTA k            :  L[k,k] = chol(S_k)                                    -> cluster barrier
//     every CTA i >  k :  L[i,k] = S_i L[k,k]^-T                                -> cluster barrier
// then the inverse by block columns: X[j,j] = L[j,j]^-1 (barrier), X[i,j] = -X[i,i] sum_{j<=k<i} L[i,k] X[k,j].
// Replaces ~3 launches per panel + a column-sequential inverse (22 + 1 launches at d = 200) on the EM critical path.
// ------------------------------------------------------------------------- //
constexpr int CB = 32;          // block size
constexpr int CKC = 64;         // K chunk of the tile products
constexpr int CP = CKC + 1;     // shared-memory row pitch (doubles): conflict-free strided reads

__device__ __forceinline__ void cluster_barrier_all() {
  __threadfence();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// 32 x 32 x K tile products of the cluster kernel run on 512 of the 1024 threads as 4 x 2 register tiles over four
// K slices (6 shared-memory loads per 8 FMAs; a thread-per-output loop needs 2 loads per FMA and is bound by the
// shared-memory port): thread t < 512 -> slice t >> 7, rows {ti + 8 a}, columns {tj + 16 b}.
struct Tile42 {
  double acc[4][2];
  int ks, ti, tj;
  bool active;
  __device__ __forceinline__ void init() {
    active = threadIdx.x < 512;
    ks = threadIdx.x >> 7;
    const int tt = threadIdx.x & 127;
    ti = tt >> 4;
    tj = tt & 15;
#pragma unroll
    for (int a = 0; a < 4; ++a) { acc[a][0] = 0.0; acc[a][1] = 0.0; }
  }
  // acc += A[row][kk] * B(kk, col) for kk in this thread's quarter of [0, klen): A row-major with pitch lda, B element
  // (kk, col) at b[kk * sbk + col * sbc]
  __device__ __forceinline__ void fma_slice(const double* a, int lda, const double* b, int sbk, int sbc, int klen) {
    if (!active) return;
    const int q = klen >> 2;
    const double* ap = a + ti * lda;
    const double* bp = b + tj * sbc;
#pragma unroll 4
    for (int kk = ks * q; kk < (ks + 1) * q; ++kk) {
      const double y0 = bp[kk * sbk], y1 = bp[kk * sbk + 16 * sbc];
#pragma unroll
      for (int a4 = 0; a4 < 4; ++a4) {
        const double x = ap[a4 * 8 * lda + kk];
        acc[a4][0] = fma(x, y0, acc[a4][0]);
        acc[a4][1] = fma(x, y1, acc[a4][1]);
      }
    }
  }
  // sum of the four slices for output (r, c) of thread (r = threadIdx.x >> 5, c = lane); red: [4][32][32] scratch that
  // no thread is still reading; ends with a block barrier (red reusable)
  __device__ __forceinline__ double reduce(double* red) {
    if (active) {
#pragma unroll
      for (int a4 = 0; a4 < 4; ++a4) {
        red[(ks * 32 + ti + 8 * a4) * 32 + tj] = acc[a4][0];
        red[(ks * 32 + ti + 8 * a4) * 32 + tj + 16] = acc[a4][1];
      }
    }
    __syncthreads();
    const int o = (threadIdx.x >> 5) * 32 + (threadIdx.x & 31);
    const double v = (red[o] + red[1024 + o]) + (red[2048 + o] + red[3072 + o]);
    __syncthreads();
    return v;
  }
};

// acc(r, c) = sum_kk P[r][kk] * Q[c][kk] over K columns of two row-major 32-row panels in global memory (L2), staged
// through shared memory in chunks of CKC columns; the next chunk travels global -> registers while the current one
// is consumed.  Thread (r = threadIdx.x >> 5, c = lane); 1024 threads.
__device__ __forceinline__ double tile_dot(const double* __restrict__ p, const double* __restrict__ q, long long ld,
                                           int rows_p, int rows_q, int k, double* sp, double* sq) {
  // element (rr, kk) of a chunk: thread t loads (t >> 6, t & 63) and (t >> 6) + 16
  const int lr = threadIdx.x >> 6, lk = threadIdx.x & 63;
  double pa[2], qa[2];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int rr = lr + 16 * h;
      const bool kin = k0 + lk < k;
      pa[h] = (rr < rows_p && kin) ? __ldcg(p + rr * ld + k0 + lk) : 0.0;
      qa[h] = (rr < rows_q && kin) ? __ldcg(q + rr * ld + k0 + lk) : 0.0;
    }
  };
  if (k <= 0) return 0.0;
  Tile42 tl;
  tl.init();
  fetch(0);
  for (int k0 = 0; k0 < k; k0 += CKC) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      sp[(lr + 16 * h) * CP + lk] = pa[h];
      sq[(lr + 16 * h) * CP + lk] = qa[h];
    }
    __syncthreads();
    if (k0 + CKC < k) fetch(k0 + CKC);
    tl.fma_slice(sp, CP, sq, 1, CP, CKC);
    __syncthreads();
  }
  return tl.reduce(sp);            // sp and sq are contiguous: [4][32][32] fits in their 2 x 32 x 65 doubles
}

__global__ void __launch_bounds__(1024, 1)
chol_inverse_cluster_kernel(double* __restrict__ a, double* __restrict__ inv, int d, int* __restrict__ info,
                            long long* __restrict__ dbg) {
  extern __shared__ double csm[];
  double* sp = csm;                  // [32][CP]
  double* sq = sp + CB * CP;         // [32][CP]
  double* st = sq + CB * CP;         // [32][33] the CTA's current tile
  double* sl = st + CB * 33;         // [32][33] a diagonal block's inverse
  __shared__ double s_rd[CB];        // reciprocal diagonal of the factored block
  const int nb = gridDim.x;
  const int bi = blockIdx.x;                                   // row block this CTA owns
  const int r = threadIdx.x >> 5, c = threadIdx.x & 31;
  const int rows_i = min(CB, d - bi * CB);
  const long long ld = d;
  // PLDA_B200_DBG=1: SM cycles of the LAST row block's CTA per phase (it takes part in every panel)
  const bool prof = dbg != nullptr && bi == nb - 1 && threadIdx.x == 0;
  long long t_ph[5] = {0, 0, 0, 0, 0};
  long long t0 = prof ? clock64() : 0;
  auto lap = [&](int i) {
    if (prof) { const long long t1 = clock64(); t_ph[i] += t1 - t0; t0 = t1; }
  };
  // ---------------- Cholesky (left-looking by block column k) ----------------
  for (int k = 0; k < nb; ++k) {
    const int rows_k = min(CB, d - k * CB);
    if (bi >= k) {
      // (the tile's own element travels from L2 while the products run)
      const bool in_tile = r < rows_i && c < rows_k;
      const double a_rc = in_tile ? __ldcg(a + (static_cast<long long>(bi) * CB + r) * ld + k * CB + c) : 0.0;
      const double dot = tile_dot(a + static_cast<long long>(bi) * CB * ld, a + static_cast<long long>(k) * CB * ld, ld,
                                  rows_i, rows_k, k * CB, sp, sq);
      double v = (r == c && r >= rows_k) ? 1.0 : 0.0;             // identity padding keeps the factor regular
      if (in_tile) v = a_rc - dot;
      st[r * 33 + c] = v;
      __syncthreads();
      if (bi == k) {
        // (a) factor the diagonal tile: thread (r, c) keeps S[r][c] in a register; per step the pivot column is
        //     published through a double-buffered shared-memory vector -> ONE block barrier per step, and every
        //     thread derives rsqrt(pivot) itself (no fp64 sqrt / divide, no serial trailing update)
        {
          double v = st[r * 33 + c];
          double* colbuf = sp;                                     // [2][32], the K-chunk staging is idle here
          if (c == 0) colbuf[r] = v;
          __syncthreads();
          bool failed = false;
          for (int j = 0; j < CB; ++j) {
            const double* col = colbuf + (j & 1) * CB;
            const double piv2 = col[j];
            if (!(piv2 > 0.0)) {                                   // uniform across the block
              if (threadIdx.x == 0) *info = k * CB + j + 1;
              failed = true;
              break;
            }
            const double rp = rsqrt(piv2);
            const double lr = col[r] * rp, lc = col[c] * rp;       // L[r][j], L[c][j] for r, c > j
            if (c == j) {
              if (r == j) { v = piv2 * rp; s_rd[j] = rp; }
              else if (r > j) v = lr;
            } else if (c > j && r >= c) {
              v = fma(-lr, lc, v);
            }
            if (c == j + 1) colbuf[((j + 1) & 1) * CB + r] = v;
            __syncthreads();
          }
          st[r * 33 + c] = v;
          if (failed && r == c) s_rd[r] = 1.0;                     // keep what follows finite; info reports the failure
        }
        __syncthreads();
        // (b) its inverse: warp r solves L x = e_r column-wise (lane c holds the running right-hand side b_c): one
        //     broadcast + one FMA per step instead of a 5-level shuffle reduction
        {
          double b = c == r ? 1.0 : 0.0, x = 0.0;
          for (int j = 0; j < CB; ++j) {
            const double xj = __shfl_sync(0xffffffffu, b, j) * s_rd[j];
            if (c == j) x = xj;
            else if (c > j) b = fma(-st[c * 33 + j], xj, b);
          }
          sl[c * 33 + r] = x;                                      // entry (row c, column r) of L_kk^-1
        }
        __syncthreads();
        if (r < rows_k && c < rows_k) {
          const long long o = (static_cast<long long>(k) * CB + r) * ld + k * CB + c;
          __stcg(a + o, c <= r ? st[r * 33 + c] : 0.0);
          __stcg(inv + o, c <= r ? sl[r * 33 + c] : 0.0);
        }
      }
    }
    lap(0);                 // tile update (+ the diagonal factor and its inverse in the last panel)
    cluster_barrier_all();
    lap(1);                 // barrier: the diagonal CTA's factor + inverse, then the cluster barrier itself
    if (bi > k) {
      // L[i,k] = S_i L_kk^-T : a 32 x 32 x 32 product against the inverse CTA k just published
      sl[r * 33 + c] = (r < rows_k && c < rows_k) ? __ldcg(inv + (static_cast<long long>(k) * CB + r) * ld + k * CB + c) : 0.0;
      __syncthreads();
      Tile42 tl;
      tl.init();
      tl.fma_slice(st, 33, sl, 1, 33, CB);
      const double x = tl.reduce(sp);
      if (r < rows_i && c < rows_k) __stcg(a + (static_cast<long long>(bi) * CB + r) * ld + k * CB + c, x);
    } else if (bi < k) {
      // blocks above the diagonal of column k: zero (clean triangular factor and inverse for the caller)
      if (r < rows_i && c < rows_k) {
        const long long o = (static_cast<long long>(bi) * CB + r) * ld + k * CB + c;
        __stcg(a + o, 0.0);
        __stcg(inv + o, 0.0);
      }
    }
    lap(2);                 // panel product
    cluster_barrier_all();
    lap(3);
  }
  if (prof) {
    for (int i = 0; i < 4; ++i) dbg[i] = t_ph[i];
  }
  const long long t_inv0 = (dbg != nullptr && bi == 0 && threadIdx.x == 0) ? clock64() : 0;
  // ---------------- inverse: CTA bi computes block column bi of X = L^-1 (X[bi,bi] is already there) ----------------
  // The X tiles of the column stay in shared memory (xs[k] = X[k,bi]); the L tiles (final since the last barrier) and
  // the diagonal inverses travel global -> registers one step ahead of their use.
  {
    double* xs = sl + CB * 33;                                     // [nb][32][33]
    xs[(bi * CB + r) * 33 + c] = (r < rows_i && c < rows_i) ? __ldcg(inv + (static_cast<long long>(bi) * CB + r) * ld + bi * CB + c) : 0.0;
    auto l_tile = [&](int i, int k0) -> double {                   // element (r, c) of L[i,k0]
      return (i * CB + r < d && k0 * CB + c < d) ? __ldcg(a + (static_cast<long long>(i) * CB + r) * ld + k0 * CB + c) : 0.0;
    };
    auto x_diag = [&](int i) -> double {                           // element (r, c) of X[i,i]
      return (i * CB + r < d && i * CB + c < d) ? __ldcg(inv + (static_cast<long long>(i) * CB + r) * ld + i * CB + c) : 0.0;
    };
    double l_next = bi + 1 < nb ? l_tile(bi + 1, bi) : 0.0;
    for (int i = bi + 1; i < nb; ++i) {
      const int rows_u = min(CB, d - i * CB);
      const double xd = x_diag(i);
      // T = sum_{bi <= k < i} L[i,k] X[k,bi]   (accumulated in the register tiles across the k loop)
      Tile42 tl;
      tl.init();
      for (int k0 = bi; k0 < i; ++k0) {
        sq[r * 33 + c] = l_next;
        __syncthreads();
        if (k0 + 1 < i) l_next = l_tile(i, k0 + 1);
        else if (i + 1 < nb) l_next = l_tile(i + 1, bi);
        tl.fma_slice(sq, 33, xs + k0 * CB * 33, 33, 1, CB);
        __syncthreads();
      }
      st[r * 33 + c] = tl.reduce(sp);
      sl[r * 33 + c] = xd;                 // (operands outside the reduction scratch, which spans sp and sq)
      __syncthreads();
      Tile42 tx;
      tx.init();
      tx.fma_slice(sl, 33, st, 33, 1, CB);
      double x = tx.reduce(sp);
      x = (r < rows_u && c < rows_i) ? -x : 0.0;
      xs[(i * CB + r) * 33 + c] = x;
      if (r < rows_u && c < rows_i) __stcg(inv + (static_cast<long long>(i) * CB + r) * ld + bi * CB + c, x);
      __syncthreads();
    }
  }
  if (dbg != nullptr && bi == 0 && threadIdx.x == 0) dbg[4] = clock64() - t_inv0;   // the longest inverse column
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
// Device-wide barrier used by the eigensolver (cooperative launch guarantees co-residency).
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

// ------------------------------------------------------------------------- //
// Block one-sided (Hestenes) Jacobi on a symmetric matrix, Gram formulation.
//   gt[p,:] = column p of G = B V (rows contiguous).  Columns are grouped in blocks of `bw`; a CTA owns one
//   PAIR of blocks per global round (round-robin tournament over blocks, one device-wide barrier per round)
//   and orthogonalises its 2*bw columns through their small Gram matrix:
//   1. Gl = C^T C            (m2 x m2; DMMA.8x8x4 fragments out of shared memory, a warp per 8 x 8 tile)
//   2. one two-sided Jacobi tournament on Gl (m2 - 1 rounds of m2/2 disjoint rotations, rows then columns),
//      accumulating the rotations in Q -- rotation angles come from 3 Gram entries, no d-length reductions
//   3. C <- C Q              (written straight back to global memory)
// Steps 1 and 3 are GEMM-shaped and use every thread; only step 2 is sequential, on a 32 x 32 (16 x 16) matrix.
// ------------------------------------------------------------------------- //
// CLUSTER: the whole grid is ONE thread-block cluster (<= 16 CTAs): the per-round barrier is the hardware cluster
// barrier (release / acquire at cluster scope orders the __stcg / __ldcg column exchange through L2) instead of an
// atomic counter polled through L2.  big2: a sweep whose largest rotation satisfied gamma^2 <= big2 alpha beta is
// the last one (quadratic convergence: the remaining couplings are ~big2 relative).
//
// One sweep = one INTRA round (CTA c orthogonalises the columns inside its blocks 2c and 2c+1: bw - 1 steps of bw
// disjoint rotations) followed by nblk_pad - 1 CROSS rounds of the round-robin block tournament, in which only the
// bw x bw pairs (x in block I, y in block J) are rotated (bw steps of bw disjoint rotations): every column pair is
// visited exactly once per sweep -- a cyclic-by-blocks Jacobi sweep -- instead of re-rotating the intra-block pairs
// in every round (2 bw - 1 steps per round).
// D(8x8) += A(8x4) B(4x8) on the fp64 tensor cores.  Fragments: a = A[lane/4][lane%4], b = B[lane%4][lane/4],
// (c0, c1) = D[lane/4][2*(lane%4) + {0, 1}].
__device__ __forceinline__ void dmma_8x8x4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// TEAM (cluster form only): 1 or 2 CTAs share a block pair.  With 2, each CTA holds HALF of the column length: it
// loads, forms a partial Gram matrix of, and applies the rotations to its own rows only; the partials are exchanged
// through distributed shared memory (one more cluster barrier per round) and both CTAs run the identical rotation
// tournament on the identical sum.  The d-proportional phases of a round are fp64-FMA bound on one SM.
template <bool CLUSTER, int BW, int TEAM>
__global__ void __launch_bounds__(512, 1)
block_jacobi_gram_kernel(double* __restrict__ gt, int d, int /*bw*/, int nblk_pad, double tol, double big2, int max_sweeps,
                         unsigned int* __restrict__ barrier_counter, int* __restrict__ rotated,
                         int* __restrict__ sweeps_done, long long* __restrict__ dbg, int variant) {
  extern __shared__ double sm[];
  constexpr int bw = BW;                      // compile-time block width: the tournament's index arithmetic (% bw,
                                              // % (bw - 1)) sits on the critical path of every step
  constexpr int m2 = 2 * bw;
  const int dp = d | 1;                       // odd pitch: strided column reads are bank-conflict free
  constexpr int gp = m2 + 1;
  double* cols = sm;                          // [m2][dp]
  double* gl = cols + m2 * dp;                // [m2][gp]
  double* qm = gl + m2 * gp;                  // [m2][gp]
  __shared__ int s_rot;
  __shared__ int s_big;                       // a rotation above the stop threshold was seen
  __shared__ double s_abs;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nthreads = blockDim.x, nwarps = blockDim.x >> 5;
  constexpr int half = m2 >> 1;               // == bw: rotations per tournament step
  unsigned int target = 0;
  int sweep = 0;
  long long t_ph[6] = {0, 0, 0, 0, 0, 0};
  // CTA 0's team holds the padding block: profile the next team
  const bool prof = dbg != nullptr && blockIdx.x == (gridDim.x > TEAM ? static_cast<unsigned>(TEAM) : 0u) && threadIdx.x == 0;
  // rows of the columns this CTA works on
  const int team_h = TEAM == 2 ? static_cast<int>(blockIdx.x & 1u) : 0;
  const int dh = TEAM == 2 ? (((d + 1) / 2 + 7) & ~7) : d;
  const int row0 = team_h * dh;
  const int nloc = max(0, min(d, row0 + dh) - row0);
  for (; sweep < max_sweeps; ++sweep) {
    for (int r = -1; r < nblk_pad - 1; ++r) {
      long long t0 = prof ? clock64() : 0;
      int bi, bj;
      const int k = blockIdx.x / TEAM;
      if (r < 0) { bi = 2 * k; bj = 2 * k + 1; }                                  // intra round
      else if (k == 0) { bi = nblk_pad - 1; bj = r; }
      else { bi = (r + k) % (nblk_pad - 1); bj = (r - k + (nblk_pad - 1)) % (nblk_pad - 1); }
      // ---- 0. load the 2 bw columns (a warp per column, lanes along it)
      for (int slot = warp; slot < m2; slot += nwarps) {
        const int col = (slot < bw ? bi : bj) * bw + (slot < bw ? slot : slot - bw);
        const double* src = gt + static_cast<long long>(col) * d + row0;
        double* dst = cols + slot * dp;
        if (col < d) { for (int i = lane; i < nloc; i += 32) dst[i] = __ldcg(src + i); }
        else { for (int i = lane; i < nloc; i += 32) dst[i] = 0.0; }
      }
      for (int idx = threadIdx.x; idx < m2 * m2; idx += nthreads) {
        const int i = idx / m2, j = idx - i * m2;
        qm[i * gp + j] = i == j ? 1.0 : 0.0;
      }
      if (threadIdx.x == 0) { s_rot = 0; s_big = 0; }
      __syncthreads();
      if (prof) { const long long t1 = clock64(); t_ph[0] += t1 - t0; t0 = t1; }
      // ---- 1. Gram on the fp64 tensor cores: a warp owns one 8 x 8 tile of C^T C and walks the column length in
      //         DMMA.8x8x4 steps (A and B fragments are the same access pattern: lane -> column i0 + lane/4, row
      //         r + lane%4).  The plain-FMA form of this phase was bound by the fp64 FMA rate of the SM.
      {
        const int nt8 = m2 >> 3;
        const int lr = lane >> 2, lk = lane & 3;
        double* gdst = TEAM == 2 ? qm + m2 * gp : gl;   // team: partial Gram in the scratch behind qm
        for (int t = warp; t < nt8 * nt8; t += nwarps) {
          const int ti = t / nt8, tj = t - ti * nt8;
          if (tj < ti) continue;                       // symmetric: mirrored below
          const double* ap = cols + (ti * 8 + lr) * dp + lk;
          const double* bp = cols + (tj * 8 + lr) * dp + lk;
          double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;   // two accumulator pairs: independent DMMA chains
          int r = 0;
          for (; r + 8 <= nloc; r += 8) {
            dmma_8x8x4(c0, c1, ap[r], bp[r]);
            dmma_8x8x4(e0, e1, ap[r + 4], bp[r + 4]);
          }
          for (; r < nloc; r += 4) {
            const bool in = r + lk < nloc;
            dmma_8x8x4(c0, c1, in ? ap[r] : 0.0, in ? bp[r] : 0.0);
          }
          c0 += e0;
          c1 += e1;
          const int gi = ti * 8 + lr, gj = tj * 8 + 2 * lk;
          gdst[gi * gp + gj] = c0;
          gdst[gi * gp + gj + 1] = c1;
          if (ti != tj) {
            gdst[gj * gp + gi] = c0;
            gdst[(gj + 1) * gp + gi] = c1;
          }
        }
        if (TEAM == 2) {
          // exchange: own partial + the partner's (read through distributed shared memory); a + b == b + a exactly,
          // so both CTAs of the team continue with bit-identical Gram matrices
          asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
          asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
          const uint32_t remote = mapa_shared(smem_u32(gdst), blockIdx.x ^ 1u);
          for (int idx = threadIdx.x; idx < m2 * m2; idx += nthreads) {
            const int i = idx / m2, j = idx - i * m2;
            double other;
            asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(other) : "r"(remote + (i * gp + j) * 8) : "memory");
            gl[i * gp + j] = gdst[i * gp + j] + other;
          }
        }
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        double mx = 0.0;
        for (int i = 0; i < m2; ++i) mx = fmax(mx, gl[i * gp + i]);
        // rounding noise of a Gram entry is relative to |g_x||g_y| (covered by `tol`); the absolute floor only
        // guards columns that vanished entirely
        s_abs = 1e-290 * mx;
      }
      __syncthreads();
      const double abs_tol = s_abs;
      if (prof) { const long long t1 = clock64(); t_ph[1] += t1 - t0; t0 = t1; }
      // ---- 2. two-sided Jacobi on the Gram matrix: intra pairs (r < 0) or cross pairs
      const int nsteps = r < 0 ? bw - 1 : bw;
      for (int lr = 0; lr < nsteps; ++lr) {
        int x = 0, y = 0;
        double c = 1.0, sn = 0.0;
        bool rot = false;
        if (warp < half) {
          if (r < 0) {
            // round-robin over the bw columns of one block; warps [0, bw/2) take block I, the rest block J
            const int hb = bw >> 1;
            const int w = warp < hb ? warp : warp - hb;
            const int off = warp < hb ? 0 : bw;
            if (w == 0) { x = bw - 1; y = lr; }
            else { x = (lr + w) % (bw - 1); y = (lr - w + (bw - 1)) % (bw - 1); }
            x += off;
            y += off;
          } else {
            x = warp;
            y = bw + (warp + lr) % bw;
          }
          const double alpha = gl[x * gp + x], beta = gl[y * gp + y], gamma = gl[x * gp + y];
          if (lane == 0 && gamma * gamma > big2 * alpha * beta) s_big = 1;
          if (gamma * gamma > tol * tol * alpha * beta && fabs(gamma) > abs_tol) {
            // rotation angle in fp32 (a 1e-7 relative error in the angle only leaves a 1e-7 * gamma residual),
            // but (c, s) exactly orthonormal in fp64: c = rsqrt(1 + t^2) by two Newton steps, s = c t
            // (approximate divide / square root: they sit on the critical path of every tournament step)
            const float zf = __fdividef(static_cast<float>(beta - alpha), 2.0f * static_cast<float>(gamma));
            float rt;
            asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rt) : "f"(fmaf(zf, zf, 1.0f)));
            const float tf = copysignf(__fdividef(1.0f, fabsf(zf) + rt), zf);
            const double t = static_cast<double>(tf);
            const double xx = fma(t, t, 1.0);
            double c0 = static_cast<double>(rsqrtf(static_cast<float>(xx)));
            c0 = c0 * fma(-0.5 * xx, c0 * c0, 1.5);
            c0 = c0 * fma(-0.5 * xx, c0 * c0, 1.5);
            c = c0;
            sn = c * t;
            rot = tf != 0.0f;
          }
        }
        if (rot) {
          for (int j = lane; j < m2; j += 32) {   // rows x, y  (J^T G)
            const double gx = gl[x * gp + j], gy = gl[y * gp + j];
            gl[x * gp + j] = c * gx - sn * gy;
            gl[y * gp + j] = sn * gx + c * gy;
          }
          if (lane == 0) s_rot = 1;
        }
        __syncthreads();
        if (rot) {
          for (int i = lane; i < m2; i += 32) {   // columns x, y  (G J) and Q <- Q J
            const double gx = gl[i * gp + x], gy = gl[i * gp + y];
            gl[i * gp + x] = c * gx - sn * gy;
            gl[i * gp + y] = sn * gx + c * gy;
            const double qx = qm[i * gp + x], qy = qm[i * gp + y];
            qm[i * gp + x] = c * qx - sn * qy;
            qm[i * gp + y] = sn * qx + c * qy;
          }
        }
        __syncthreads();
      }
      if (prof) { const long long t1 = clock64(); t_ph[2] += t1 - t0; t0 = t1; }
      // ---- 3. C <- C Q, written back to global: new column x' = sum_x Q[x][x'] * old column x
      if (s_rot) {
        if ((variant & 2) != 0) {
        // fp64 tensor cores again: a warp owns 8 rows x 8 output columns per tile, K = the m2 old columns
        {
          const int nt8 = m2 >> 3;
          const int mt8 = (nloc + 7) >> 3;
          const int lr = lane >> 2, lk = lane & 3;
          for (int t = warp; t < mt8 * nt8; t += nwarps) {
            const int rt = t / nt8, ct = t - rt * nt8;
            // rows past d (last tile only) read the next column / the Gram area: those outputs are never stored
            const double* ap = cols + lk * dp + rt * 8 + lr;          // A[i][k] = old column k0 + k, row rt*8 + i
            const double* bp = qm + lk * gp + ct * 8 + lr;            // B[k][j] = Q[k0 + k][ct*8 + j]
            double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;
            for (int k0 = 0; k0 < m2; k0 += 8) {
              dmma_8x8x4(c0, c1, ap[k0 * dp], bp[k0 * gp]);
              dmma_8x8x4(e0, e1, ap[(k0 + 4) * dp], bp[(k0 + 4) * gp]);
            }
            c0 += e0;
            c1 += e1;
            const int i = rt * 8 + lr;
            if (i < nloc) {
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                const int slot = ct * 8 + 2 * lk + j;
                const int col = (slot < bw ? bi : bj) * bw + (slot < bw ? slot : slot - bw);
                if (col < d) __stcg(gt + static_cast<long long>(col) * d + row0 + i, j == 0 ? c0 : c1);
              }
            }
          }
        }
        } else {
        // a warp owns 4 output columns x 128 rows (lane: rows i0 + 32 m): 4 column loads + 4 broadcast loads of Q per
        // 16 FMAs
        const int groups = m2 >> 2;
        const int rparts = (nloc + 127) >> 7;
        for (int wi = warp; wi < groups * rparts; wi += nwarps) {
          const int xg = wi % groups, part = wi / groups;
          const double* qrow = qm + xg * 4;
          const int i0 = part * 128 + lane;
          double acc[4][4];
#pragma unroll
          for (int m = 0; m < 4; ++m)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[m][j] = 0.0;
#pragma unroll 2
          for (int x = 0; x < m2; ++x) {
            // rows past d read the next column / the Gram area: finite or not, those accumulators are never stored
            double v[4], qv[4];
#pragma unroll
            for (int m = 0; m < 4; ++m) v[m] = cols[x * dp + i0 + 32 * m];
#pragma unroll
            for (int j = 0; j < 4; ++j) qv[j] = qrow[x * gp + j];
#pragma unroll
            for (int m = 0; m < 4; ++m)
#pragma unroll
              for (int j = 0; j < 4; ++j) acc[m][j] = fma(v[m], qv[j], acc[m][j]);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int slot = xg * 4 + j;
            const int col = (slot < bw ? bi : bj) * bw + (slot < bw ? slot : slot - bw);
            if (col >= d) continue;
#pragma unroll
            for (int m = 0; m < 4; ++m) {
              const int i = i0 + 32 * m;
              if (i < nloc) __stcg(gt + static_cast<long long>(col) * d + row0 + i, acc[m][j]);
            }
          }
        }
        }
        // bit 0: something rotated; bit 1: a rotation above the stop threshold happened.  Jacobi converges
        // quadratically, so a sweep whose largest rotation was below the threshold leaves every pair below its
        // square: converged without paying for a verification sweep.
        if (threadIdx.x == 0) atomicOr(rotated + sweep, s_big ? 3 : 1);
      }
      if (prof) { const long long t1 = clock64(); t_ph[3] += t1 - t0; t0 = t1; }
      if (CLUSTER) {
        __threadfence();
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
      } else {
        grid_barrier(barrier_counter, target, gridDim.x);
      }
      if (prof) { t_ph[4] += clock64() - t0; t_ph[5] += 1; }
    }
    const int flags = *reinterpret_cast<volatile int*>(rotated + sweep);
    if ((flags & 2) == 0) { ++sweep; break; }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *sweeps_done = sweep;
  if (prof)
    for (int i = 0; i < 6; ++i) dbg[i] = t_ph[i];
}

// lambda_p = |g_p| ; v_p = g_p / |g_p|  (zero vector if the column vanished)
__global__ void __launch_bounds__(256)
eig_normalise_kernel(const double* __restrict__ gt, int d, double* __restrict__ lam, double* __restrict__ vt) {
  const int p = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (p >= d) return;
  double acc = 0.0;
  for (int i = lane; i < d; i += 32) {
    const double g = gt[static_cast<long long>(p) * d + i];
    acc += g * g;
  }
  acc = warp_sum(acc);
  const double nrm = sqrt(acc);
  const double inv = nrm > 0.0 ? 1.0 / nrm : 0.0;
  for (int i = lane; i < d; i += 32) vt[static_cast<long long>(p) * d + i] = gt[static_cast<long long>(p) * d + i] * inv;
  if (lane == 0) lam[p] = nrm;
}

}  // namespace

void cholesky_lower(Context& ctx, double* a, int64_t d, int* info_dev) {
  PB_CHECK(d > 0 && d <= 4096, kInvalidArg, "cholesky: dimension out of range");
  PB_CUDA(cudaMemsetAsync(info_dev, 0, sizeof(int), ctx.stream));
  const int di = static_cast<int>(d);
  for (int k0 = 0; k0 < di; k0 += 32) {
    const int nb = std::min(32, di - k0);
    if (k0 > 0)   // A[k0:, k0:k0+nb] -= L[k0:, :k0] L[k0:k0+nb, :k0]^T
      gemm_f64(ctx, false, true, di - k0, nb, k0, -1.0, a + static_cast<int64_t>(k0) * d, d,
               a + static_cast<int64_t>(k0) * d, d, 1.0, a + static_cast<int64_t>(k0) * d + k0, d);
    chol_diag_kernel<<<1, 1024, 0, ctx.stream>>>(a, di, k0, nb, info_dev);
    const int below = di - k0 - nb;
    if (below > 0) chol_solve_kernel<<<static_cast<unsigned>(ceil_div(below, 32)), 32, 0, ctx.stream>>>(a, di, k0, nb);
    ctx.count_launch(below > 0 ? 2 : 1);
  }
  zero_upper_kernel<<<static_cast<unsigned>(ceil_div(d * d, 256)), 256, 0, ctx.stream>>>(a, di);
  PB_CUDA(cudaGetLastError());
  ctx.count_launch();
}

bool cholesky_inverse_fused(Context& ctx, double* a, double* inv, int64_t d, int* info_dev) {
  static const char* mode = getenv("PLDA_B200_CHOL");
  if (mode != nullptr && strcmp(mode, "legacy") == 0) return false;
  const int nb = static_cast<int>(ceil_div(d, 32));
  if (d < 1 || nb > 16) return false;
  const size_t smem = (2 * 32 * 65 + 2 * 32 * 33 + static_cast<size_t>(nb) * 32 * 33) * sizeof(double);
  static std::once_flag once;
  std::call_once(once, [&] {
    cudaFuncSetAttribute(chol_inverse_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         static_cast<int>((2 * 32 * 65 + 2 * 32 * 33 + 16 * 32 * 33) * sizeof(double)));
    cudaFuncSetAttribute(chol_inverse_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  });
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  cfg.gridDim = dim3(nb);
  cfg.blockDim = dim3(1024);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = ctx.stream;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = nb;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int max_clusters = 0;
  if (cudaOccupancyMaxActiveClusters(&max_clusters, chol_inverse_cluster_kernel, &cfg) != cudaSuccess || max_clusters < 1) {
    cudaGetLastError();
    return false;
  }
  PB_CUDA(cudaMemsetAsync(info_dev, 0, sizeof(int), ctx.stream));
  static const bool dbg_on = getenv("PLDA_B200_DBG") != nullptr;
  static long long* dbgp = nullptr;
  if (dbg_on && dbgp == nullptr) PB_CUDA(cudaMalloc(reinterpret_cast<void**>(&dbgp), 8 * sizeof(long long)));
  PB_CUDA(cudaLaunchKernelEx(&cfg, chol_inverse_cluster_kernel, a, inv, static_cast<int>(d), info_dev, dbgp));
  ctx.count_launch();
  if (dbgp != nullptr) {
    long long h[5];
    PB_CUDA(cudaMemcpyAsync(h, dbgp, sizeof(h), cudaMemcpyDeviceToHost, ctx.stream));
    ctx.sync();
    fprintf(stderr, "plda_b200 cholesky d=%d (%d panels), SM cycles of the last row block's CTA: tile update %lld  wait for the "
            "diagonal + barrier %lld  panel product %lld  barrier %lld | inverse (column 0) %lld\n",
            static_cast<int>(d), nb, h[0], h[1], h[2], h[3], h[4]);
  }
  return true;
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
                    EigWork& w, int* sweeps_out, double stop_rotation) {
  PB_CHECK(d > 0 && d <= 1024, kInvalidArg, "eig: d <= 1024");
  const int max_sweeps = 40;
  w.g.reserve(d * d);
  w.v.reserve(d * d);
  w.lam.reserve(d);
  w.tmp.reserve(1);
  w.flags.reserve(max_sweeps + 4);   // [0] barrier counter, [1] sweeps_done, [2..] rotated flags
  PB_CUDA(cudaMemsetAsync(w.flags.get(), 0, (max_sweeps + 4) * sizeof(int), ctx.stream));
  if (v0_t != nullptr) {
    // warm start from an orthogonal basis V0 (rows): G^T = V0^T B  (B symmetric)
    gemm_f64(ctx, false, false, d, d, d, 1.0, v0_t, d, b, d, 0.0, w.g.get(), d);
  } else {
    PB_CUDA(cudaMemcpyAsync(w.g.get(), b, d * d * sizeof(double), cudaMemcpyDeviceToDevice, ctx.stream));
  }
  // |g_p.g_q| / (|g_p||g_q|) cannot be driven below ~d*eps in fp64 (rounding of the d-term dot products);
  // a tighter threshold makes every sweep "rotate" forever (measured: 40 sweeps instead of ~8)
  double tol = 4.0 * static_cast<double>(d) * 2.220446049250313e-16;
  int di = static_cast<int>(d);
  // block width: 16 columns per block up to d = 512 (32 x 32 Gram per CTA), 8 above
  int bw = d <= 512 ? 16 : 8;
  int nblk = static_cast<int>(ceil_div(d, bw));
  if (nblk < 2) nblk = 2;
  int nblk_pad = nblk + (nblk & 1);
  const int blocks = nblk_pad / 2;
  const int m2 = 2 * bw;
  // columns + Gram + rotation accumulator + scratch (the team partner's view of the partial Gram matrix)
  // (with two CTAs per pair each CTA stages only its half of the column length; sized for the one-CTA form)
  const size_t smem = (static_cast<size_t>(m2) * (d | 1) + 3 * static_cast<size_t>(m2) * (m2 + 1)) * sizeof(double);
  PB_CHECK(smem <= 200 * 1024, kInvalidArg, "eig: dimension too large");
  PB_CHECK(blocks <= ctx.num_sms, kInvalidArg, "eig: too many blocks for a cooperative launch");
  static std::once_flag once;
  std::call_once(once, [] {
    cudaFuncSetAttribute(block_jacobi_gram_kernel<false, 16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(block_jacobi_gram_kernel<false, 8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(block_jacobi_gram_kernel<true, 16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(block_jacobi_gram_kernel<true, 16, 1>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaFuncSetAttribute(block_jacobi_gram_kernel<true, 16, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(block_jacobi_gram_kernel<true, 16, 2>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  });
  double* gp = w.g.get();
  // A/B switch (PLDA_B200_JACOBI): 2 = apply phase on DMMA as well (default: Gram on DMMA, apply on plain FMAs --
  // measured at d = 200: DMMA.8x8x4 sustains ~20 FMA/clk/SM against 64 for DFMA, but the Gram phase's plain-FMA form
  // is load-bound and gains from the tensor-core fragments, the apply phase does not)
  static const char* var_env = getenv("PLDA_B200_JACOBI");
  int variant = var_env != nullptr ? atoi(var_env) : 0;
  static const bool dbg_on = getenv("PLDA_B200_DBG") != nullptr;
  long long* dbgp = nullptr;
  if (dbg_on) {
    w.dbg.reserve(8);
    dbgp = w.dbg.get();
  }
  unsigned int* counter = reinterpret_cast<unsigned int*>(w.flags.get());
  int* sweeps_done = w.flags.get() + 1;
  int* rotated = w.flags.get() + 2;
  int ms = max_sweeps;
  double big2 = stop_rotation * stop_rotation;
  // one cluster for the whole solve when the hardware can place it (<= 16 CTAs, one per SM in a GPC)
  bool use_cluster = blocks <= 16 && !(getenv("PLDA_B200_EIG") != nullptr && strcmp(getenv("PLDA_B200_EIG"), "grid") == 0);
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  // two CTAs per block pair (each half of the column length) when the doubled cluster still fits: d <= 256
  static const char* team_env = getenv("PLDA_B200_JACOBI_TEAM");
  int team = (use_cluster && bw == 16 && 2 * blocks <= 16 && d >= 64 && !(team_env != nullptr && atoi(team_env) == 1)) ? 2 : 1;
  auto configure = [&](int t) {
    cfg.gridDim = dim3(blocks * t);
    cfg.blockDim = dim3(512);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ctx.stream;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = blocks * t;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  };
  if (use_cluster) {
    int max_clusters = 0;
    if (bw != 16) use_cluster = false;
    if (use_cluster && team == 2) {
      configure(2);
      if (cudaOccupancyMaxActiveClusters(&max_clusters, block_jacobi_gram_kernel<true, 16, 2>, &cfg) != cudaSuccess ||
          max_clusters < 1) {
        cudaGetLastError();
        team = 1;
      }
    }
    if (use_cluster && team == 1) {
      configure(1);
      if (cudaOccupancyMaxActiveClusters(&max_clusters, block_jacobi_gram_kernel<true, 16, 1>, &cfg) != cudaSuccess ||
          max_clusters < 1) {
        cudaGetLastError();
        use_cluster = false;
      }
    }
  }
  if (use_cluster && team == 2) {
    PB_CUDA(cudaLaunchKernelEx(&cfg, block_jacobi_gram_kernel<true, 16, 2>, gp, di, bw, nblk_pad, tol, big2, ms, counter,
                               rotated, sweeps_done, dbgp, variant));
  } else if (use_cluster) {
    PB_CUDA(cudaLaunchKernelEx(&cfg, block_jacobi_gram_kernel<true, 16, 1>, gp, di, bw, nblk_pad, tol, big2, ms, counter,
                               rotated, sweeps_done, dbgp, variant));
  } else {
    void* args[] = {&gp, &di, &bw, &nblk_pad, &tol, &big2, &ms, &counter, &rotated, &sweeps_done, &dbgp, &variant};
    const void* fn = bw == 16 ? reinterpret_cast<const void*>(block_jacobi_gram_kernel<false, 16, 1>)
                              : reinterpret_cast<const void*>(block_jacobi_gram_kernel<false, 8, 1>);
    PB_CUDA(cudaLaunchCooperativeKernel(fn, dim3(blocks), dim3(512), args, smem, ctx.stream));
  }
  ctx.count_launch();
  eig_normalise_kernel<<<static_cast<unsigned>(ceil_div(d, 8)), 256, 0, ctx.stream>>>(w.g.get(), di, w.lam.get(),
                                                                                     w.v.get());
  eig_sort_kernel<<<static_cast<unsigned>(d), 128, 0, ctx.stream>>>(w.lam.get(), w.v.get(), di, evals, evecs_t);
  PB_CUDA(cudaGetLastError());
  ctx.count_launch(2);
  if (sweeps_out) {
    PB_CUDA(cudaMemcpyAsync(sweeps_out, sweeps_done, sizeof(int), cudaMemcpyDeviceToHost, ctx.stream));
    ctx.sync();
  }
  if (dbgp != nullptr) {
    long long h[6];
    PB_CUDA(cudaMemcpyAsync(h, dbgp, sizeof(h), cudaMemcpyDeviceToHost, ctx.stream));
    ctx.sync();
    const double n = h[5] > 0 ? static_cast<double>(h[5]) : 1.0;
    fprintf(stderr, "plda_b200 jacobi d=%d: %lld rounds; SM cycles per round: load %.0f  gram %.0f  tournament %.0f  apply %.0f  barrier %.0f\n",
            di, h[5], h[0] / n, h[1] / n, h[2] / n, h[3] / n, h[4] / n);
  }
}

}  // namespace pb
