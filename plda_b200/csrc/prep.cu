// HBM-bound operand producers and row-wise epilogues of the scoring / transform path.
// Warp-per-row kernels with coalesced loads along the feature axis; inputs are the caller's fp64 / fp32 rows,
// outputs split-bf16 planes or fp32.  The score-grid producer for one enrol count (score_prep_uniform_kernel) is
// the tuned one: table-driven constants, 16-byte loads and stores, balanced warp-stride rows, optional fan-out of
// the test side to every rank's operand buffer (sharded grid); the per-row-count variants serve ragged enrol counts
// and the exact fp64 mode.
#include <algorithm>

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
//
// A lane owns groups of 8 consecutive columns: two 16-byte loads in (fp32 rows), one 16-byte store per bf16 plane
// out.  Every output (both planes and the row / column term) is written to `ndst` destinations: 1 on a single GPU;
// on a sharded grid the test-side destinations are the operand buffers of EVERY rank, mapped over
// NVLink peer memory, so this kernel is the producer AND the all-gather of the transformed test operand.  The
// last block to finish then publishes `epoch` in each rank's ready flag (release at system scope); the consuming
// GEMM polls those flags tile by tile (gemm_tc.cu, GemmShard).
// Raw 8-column group of a row: two 16-byte loads when the row allows it, guarded scalar loads otherwise.
template <typename T>
__device__ __forceinline__ void load8(const T* __restrict__ src, int c, int d, bool vec, T (&v)[8]);
template <>
__device__ __forceinline__ void load8<float>(const float* __restrict__ src, int c, int d, bool vec, float (&v)[8]) {
  if (vec && c + 8 <= d) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(src + c));
    const float4 b = __ldg(reinterpret_cast<const float4*>(src + c + 4));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = c + j < d ? __ldg(src + c + j) : 0.f;
  }
}
template <>
__device__ __forceinline__ void load8<double>(const double* __restrict__ src, int c, int d, bool vec, double (&v)[8]) {
  if (vec && c + 8 <= d) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const double2 t = __ldg(reinterpret_cast<const double2*>(src + c + 2 * j));
      v[2 * j] = t.x; v[2 * j + 1] = t.y;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = c + j < d ? __ldg(src + c + j) : 0.0;
  }
}

// One 8-column group of one row: accumulate the weighted squares, scale, split into bf16 hi / lo planes.
// fp64 rows are scaled and split in fp64 (reference contract).  fp32 rows (the resident hot path) carry 24 bits: the
// operand value is formed and split in fp32 -- the split of an fp32 value is exact (x - hi is representable) and the
// packed cvt.rn.bf16x2.f32 does two elements per instruction; the row / column terms (sums of d weighted squares)
// are accumulated in fp64.
struct Group8F32 {
  double sq[8];
  float scale[8];
  double part;
  __device__ __forceinline__ void load_consts(const double* c_sq, const double* c_scale, bool scaled, int c, int d) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const bool in = c + j < d;
      sq[j] = in ? __ldg(c_sq + c + j) : 0.0;
      scale[j] = in ? (scaled ? static_cast<float>(__ldg(c_scale + c + j)) : 1.f) : 0.f;
    }
  }
  __device__ __forceinline__ void run(float (&v)[8], uint4& hi, uint4& lo) {
    double p = 0.0;
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      p = fma(sq[j], static_cast<double>(v[j] * v[j]), p);   // squares rounded to fp32 one by one, summed in fp64
      v[j] *= scale[j];                 // test side: x 1 (exact); padding columns: x 0
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t hh, ll;
      asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hh) : "f"(v[2 * j + 1]), "f"(v[2 * j]));
      const float r0 = v[2 * j] - __uint_as_float(hh << 16);
      const float r1 = v[2 * j + 1] - __uint_as_float(hh & 0xffff0000u);
      asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(ll) : "f"(r1), "f"(r0));
      h[j] = hh;
      l[j] = ll;
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
    part = p;
  }
  // sum_j w[j] v[j]^2 over the first `valid` columns, in the arithmetic of run()
  __device__ __forceinline__ double weighted_squares(const double* w, const float (&v)[8], int valid) const {
    double p = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (j < valid) p = fma(__ldg(w + j), static_cast<double>(v[j] * v[j]), p);
    return p;
  }
};
struct Group8F64 {
  double sq[8], scale[8];
  double part;
  __device__ __forceinline__ void load_consts(const double* c_sq, const double* c_scale, bool scaled, int c, int d) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const bool in = c + j < d;
      sq[j] = in ? __ldg(c_sq + c + j) : 0.0;
      scale[j] = in ? (scaled ? __ldg(c_scale + c + j) : 1.0) : 0.0;
    }
  }
  __device__ __forceinline__ void run(double (&v)[8], uint4& hi, uint4& lo) {
    double p = 0.0;
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      p += sq[j] * v[j] * v[j];
      v[j] *= scale[j];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      __nv_bfloat16 h0, l0, h1, l1;
      split_bf16(v[2 * j], h0, l0);
      split_bf16(v[2 * j + 1], h1, l1);
      h[j] = static_cast<uint32_t>(__bfloat16_as_ushort(h0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(h1)) << 16);
      l[j] = static_cast<uint32_t>(__bfloat16_as_ushort(l0)) | (static_cast<uint32_t>(__bfloat16_as_ushort(l1)) << 16);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
    part = p;
  }
  __device__ __forceinline__ double weighted_squares(const double* w, const double (&v)[8], int valid) const {
    double p = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (j < valid) p += __ldg(w + j) * v[j] * v[j];
    return p;
  }
};
template <typename T> struct Group8Of;
template <> struct Group8Of<float> { typedef Group8F32 type; };
template <> struct Group8Of<double> { typedef Group8F64 type; };

// `consts` (kScoreConsts* layout, built once per (model, enrol count) by the engine): a/v, a^2/v, q and the
// log-determinant term.  Grid: the first `test_blocks` blocks stride over the test rows, the rest over the enrol rows -- a
// warp takes rows gw, gw + W, gw + 2W, gw + 3W (W = warps of its side) per round with the four row loads issued
// together; a lane owns 8-column groups and keeps the constants of its columns in registers.
template <typename T>
__global__ void __launch_bounds__(256, 2)
score_prep_uniform_kernel(const T* __restrict__ enrol, long long ne, long long ld_e, const T* __restrict__ test,
                          long long nt, long long ld_t, long long test_row0, long long test_pad_end, int d,
                          const double* __restrict__ consts, __nv_bfloat16* __restrict__ l_hi,
                          __nv_bfloat16* __restrict__ l_lo, float* __restrict__ row_term, const PrepDst tdst,
                          int ld_out, unsigned test_blocks, int vec_e, int vec_t, const PrepSignal sig) {
  // the dependent GEMM may be scheduled as soon as every block of this grid is running: its prologue overlaps
  // this kernel's tail, its griddepcontrol.wait returns when this grid has completed (and flushed)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // test blocks come first in the grid: on a sharded grid their stores travel over NVLink while the enrol blocks run
  const bool is_enrol = blockIdx.x >= test_blocks;
  const double* __restrict__ c_scale = consts + kScoreConstsScale;
  const double* __restrict__ c_sq = consts + (is_enrol ? kScoreConstsEnrolSq : kScoreConstsTestSq);
  const double cst = __ldg(consts + kScoreConstsLogdet);
  const long long nrows = is_enrol ? ne : nt;
  const T* __restrict__ base = is_enrol ? enrol : test;
  const long long ld_in = is_enrol ? ld_e : ld_t;
  const bool vec = (is_enrol ? vec_e : vec_t) != 0;
  const long long blk = is_enrol ? blockIdx.x - test_blocks : blockIdx.x;
  const long long w_side = static_cast<long long>(is_enrol ? gridDim.x - test_blocks : test_blocks) * 8;
  // destination 0 (the only one on a single GPU) lives in registers
  __nv_bfloat16* __restrict__ o_hi = is_enrol ? l_hi : tdst.hi[0] + test_row0 * ld_out;
  __nv_bfloat16* __restrict__ o_lo = is_enrol ? l_lo : tdst.lo[0] + test_row0 * ld_out;
  float* __restrict__ o_term = is_enrol ? row_term : tdst.term[0] + test_row0;
  const int extra = is_enrol ? 0 : tdst.n - 1;
  const bool one_group = ld_out <= 256;        // d <= 256: the lane's constants are loaded once
  typename Group8Of<T>::type g;
  if (one_group && lane * 8 < ld_out) g.load_consts(c_sq, c_scale, is_enrol, lane * 8, d);
  for (long long rb = blk * 8 + warp; rb < nrows; rb += 4 * w_side) {
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int c = lane * 8; c < ld_out; c += 256) {
      if (!one_group) g.load_consts(c_sq, c_scale, is_enrol, c, d);
      T v[4][8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const long long r = rb + i * w_side;
        if (r < nrows) load8<T>(base + r * ld_in, c, d, vec, v[i]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const long long r = rb + i * w_side;
        if (r >= nrows) continue;
        uint4 hi, lo;
        g.run(v[i], hi, lo);
        acc[i] += static_cast<double>(g.part);
        const long long o = r * ld_out + c;
        *reinterpret_cast<uint4*>(o_hi + o) = hi;
        *reinterpret_cast<uint4*>(o_lo + o) = lo;
        for (int w = 1; w <= extra; ++w) {       // sharded grid: the other ranks' operand buffers (peer memory)
          *reinterpret_cast<uint4*>(tdst.hi[w] + test_row0 * ld_out + o) = hi;
          *reinterpret_cast<uint4*>(tdst.lo[w] + test_row0 * ld_out + o) = lo;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long r = rb + i * w_side;
      const double a = warp_sum(acc[i]);
      if (lane != 0 || r >= nrows) continue;
      const float t = is_enrol ? static_cast<float>(0.5 * (cst - a)) : static_cast<float>(a);
      o_term[r] = t;
      for (int w = 1; w <= extra; ++w) tdst.term[w][test_row0 + r] = t;
    }
  }
  if (!is_enrol && blk == 0) {
    // zero padding of the column-term row: read (never stored) by the GEMM epilogue
    for (long long gr = test_row0 + nt + threadIdx.x; gr < test_pad_end; gr += 256)
      for (int w = 0; w < tdst.n; ++w) tdst.term[w][gr] = 0.f;
  }
  if (sig.counter != nullptr && !is_enrol) {
    // publish: every thread's peer stores are performed system-wide, then the LAST test block raises the ready flag
    // of this source rank in every destination region -- one thread per destination, so the NVLink round trips of
    // the flag stores overlap instead of adding up (block-uniform branch: the barriers are safe)
    __shared__ int s_last;
    // Release pattern of the block: the barrier orders every thread's peer stores before thread 0, whose ONE
    // system-scope fence (cumulative) then publishes them ahead of the counter / flag -- the same shape NCCL's
    // primitives use -- instead of a system fence in each of the 256 threads.  (sig.fence_per_thread: A/B switch.)
    if (sig.fence_per_thread) __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence_system();
      const unsigned prev = atomicAdd(sig.counter, 1u);
      s_last = prev == test_blocks - 1 ? 1 : 0;
      if (s_last) atomicExch(sig.counter, 0u);
    }
    __syncthreads();
    if (s_last && static_cast<int>(threadIdx.x) < sig.n) {
      __threadfence_system();     // orders the counter observation (all blocks fenced before adding) before the flag
      asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(sig.flag[threadIdx.x]), "r"(sig.epoch) : "memory");
    }
  }
}

// Ragged enrol counts (tensor path): the same producer with the per-column constants read from one table per
// distinct count (`tables`: [ng][kScoreConstsSize], built on the host once per call) instead of a log1p / log / three
// divisions per element.  Blocks [0, eb) take 8 enrol rows each (row r uses the table of grp[r]); the rest take 8
// test rows each and write the split row once plus one column term per group.  One warp per row.
template <typename T>
__global__ void __launch_bounds__(256)
score_prep_grouped_kernel(const T* __restrict__ enrol, long long ne, long long ld_e, const int32_t* __restrict__ grp,
                          const T* __restrict__ test, long long nt, long long ld_t, int d, int ng,
                          const double* __restrict__ tables, __nv_bfloat16* __restrict__ l_hi,
                          __nv_bfloat16* __restrict__ l_lo, __nv_bfloat16* __restrict__ r_hi,
                          __nv_bfloat16* __restrict__ r_lo, int ld_out, float* __restrict__ row_term,
                          float* __restrict__ col_term, long long col_ld, unsigned enrol_blocks) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (blockIdx.x < enrol_blocks) {
    const long long r = static_cast<long long>(blockIdx.x) * kWarpsPerBlock + warp;
    if (r >= ne) return;
    const double* tab = tables + static_cast<long long>(grp[r]) * kScoreConstsSize;
    const T* src = enrol + r * ld_e;
    double acc = 0.0;
    for (int c = lane; c < ld_out; c += 32) {
      double lv = 0.0;
      if (c < d) {
        const double e = static_cast<double>(src[c]);
        lv = e * __ldg(tab + kScoreConstsScale + c);
        acc += __ldg(tab + kScoreConstsEnrolSq + c) * e * e;
      }
      store_split(l_hi, l_lo, r * ld_out + c, lv);
    }
    acc = warp_sum(acc);
    if (lane == 0) row_term[r] = static_cast<float>(0.5 * (__ldg(tab + kScoreConstsLogdet) - acc));
  } else {
    const long long r = static_cast<long long>(blockIdx.x - enrol_blocks) * kWarpsPerBlock + warp;
    if (r >= nt) return;
    const T* src = test + r * ld_t;
    for (int c = lane; c < ld_out; c += 32)
      store_split(r_hi, r_lo, r * ld_out + c, c < d ? static_cast<double>(src[c]) : 0.0);
    for (int g = 0; g < ng; ++g) {
      const double* tab = tables + static_cast<long long>(g) * kScoreConstsSize + kScoreConstsTestSq;
      double acc = 0.0;
      for (int c = lane; c < d; c += 32) {
        const double t = static_cast<double>(src[c]);
        acc += __ldg(tab + c) * t * t;
      }
      acc = warp_sum(acc);
      if (lane == 0) col_term[g * col_ld + r] = static_cast<float>(acc);
    }
  }
}

// Vectorised form of the ragged-count producer: one warp per row, a lane owns 8-column groups (16-byte loads, one
// 16-byte store per plane), constants read from the row's table (L1 / L2 resident: 24 KB per distinct count).
// Test rows: the row is split once (unit scale) and every group's column term is taken from the same registers.
// Test blocks come FIRST in the grid and write to every destination of `tdst` at row offset test_row0 (one
// destination on a single GPU; on a sharded grid the peers' operand buffers, followed by the ready flags of `sig`);
// the enrol rows (pitch ld_l) stay local.
template <typename T>
__global__ void __launch_bounds__(256)
score_prep_grouped_vec_kernel(const T* __restrict__ enrol, long long ne, long long ld_e,
                              const int32_t* __restrict__ grp, const T* __restrict__ test, long long nt,
                              long long ld_t, int d, int ng, const double* __restrict__ tables,
                              __nv_bfloat16* __restrict__ l_hi, __nv_bfloat16* __restrict__ l_lo, int ld_l,
                              const PrepDst tdst, long long test_row0, int ld_out, float* __restrict__ row_term,
                              float* __restrict__ col_term, long long col_ld, unsigned test_blocks, int vec_e,
                              int vec_t, int embed, const PrepSignal sig) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  typename Group8Of<T>::type g;
  const bool is_enrol = blockIdx.x >= test_blocks;
  if (is_enrol) {
    const long long r = static_cast<long long>(blockIdx.x - test_blocks) * kWarpsPerBlock + warp;
    if (r >= ne) return;
    const double* tab = tables + static_cast<long long>(__ldg(grp + r)) * kScoreConstsSize;
    const T* src = enrol + r * ld_e;
    double acc = 0.0;
    for (int c = lane * 8; c < ld_l; c += 256) {
      g.load_consts(tab + kScoreConstsEnrolSq, tab + kScoreConstsScale, true, c, d);
      T v[8];
      load8<T>(src, c, d, vec_e != 0, v);
      uint4 hi, lo;
      g.run(v, hi, lo);
      acc += g.part;
      *reinterpret_cast<uint4*>(l_hi + r * ld_l + c) = hi;
      *reinterpret_cast<uint4*>(l_lo + r * ld_l + c) = lo;
    }
    acc = warp_sum(acc);
    if (lane == 0) row_term[r] = static_cast<float>(0.5 * (__ldg(tab + kScoreConstsLogdet) - acc));
    if (embed) {
      // column terms inside the product: a one-hot pair of extra K columns selects the row's group (see below)
      __syncwarp();                       // the zero fill of the padding columns above is by other lanes
      if (lane == 0) {
        const long long o = r * ld_l + d + 2 * __ldg(grp + r);
        l_hi[o] = __float2bfloat16(1.0f);
        l_hi[o + 1] = __float2bfloat16(1.0f);
      }
    }
    return;
  }
  const long long r = static_cast<long long>(blockIdx.x) * kWarpsPerBlock + warp;
  if (r < nt) {
    const T* src = test + r * ld_t;
    const long long orow = (test_row0 + r) * ld_out;
    // one pass over the row: it is split once (unit scale) and every group's weighted sum of squares is taken from
    // the same registers, eight groups per pass
    for (int g0 = 0; g0 < ng; g0 += 8) {
      double acc[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
      for (int c = lane * 8; c < ld_out; c += 256) {
        const double* tab = tables + static_cast<long long>(g0) * kScoreConstsSize;
        g.load_consts(tab + kScoreConstsTestSq, tab + kScoreConstsScale, false, c, d);
        T v[8];
        load8<T>(src, c, d, vec_t != 0, v);
        uint4 hi, lo;
        g.run(v, hi, lo);                 // leaves v unchanged (x 1) inside the row, 0 in the padding columns
        acc[0] += g.part;
        if (g0 == 0) {
          for (int w = 0; w < tdst.n; ++w) {
            *reinterpret_cast<uint4*>(tdst.hi[w] + orow + c) = hi;
            *reinterpret_cast<uint4*>(tdst.lo[w] + orow + c) = lo;
          }
        }
#pragma unroll
        for (int gi = 1; gi < 8; ++gi) {
          if (g0 + gi >= ng) break;
          const double* tq = tables + static_cast<long long>(g0 + gi) * kScoreConstsSize + kScoreConstsTestSq + c;
          acc[gi] += g.weighted_squares(tq, v, d - c);
        }
      }
      if (g0 == 0 && embed) __syncwarp();   // the zero fill of the padding columns above is by other lanes
#pragma unroll
      for (int gi = 0; gi < 8; ++gi) {
        if (g0 + gi >= ng) break;
        const double a = warp_sum(acc[gi]);
        if (lane != 0) continue;
        if (col_term != nullptr) col_term[(g0 + gi) * col_ld + r] = static_cast<float>(a);
        if (embed) {
          // The group's column term rides in two extra K columns of the test operand, as four bf16 pieces that add
          // up to the fp32 value (hi/lo of the term, hi/lo of what those two left): against the enrol row's one-hot
          // pair the hi*hi + hi*lo products of the bf16x3 scheme deliver exactly that sum into the fp32 accumulator,
          // so the GEMM epilogue needs no per-row column vector (ragged counts run the uniform-count kernel).
          const float cf = static_cast<float>(a);
          const __nv_bfloat16 h1 = __float2bfloat16(cf);
          const float r1 = cf - __bfloat162float(h1);
          const __nv_bfloat16 l1 = __float2bfloat16(r1);
          const float r2 = r1 - __bfloat162float(l1);
          const __nv_bfloat16 h2 = __float2bfloat16(r2);
          const __nv_bfloat16 l2 = __float2bfloat16(r2 - __bfloat162float(h2));
          const long long o = orow + d + 2 * (g0 + gi);
          for (int w = 0; w < tdst.n; ++w) {
            tdst.hi[w][o] = h1; tdst.lo[w][o] = l1;
            tdst.hi[w][o + 1] = h2; tdst.lo[w][o + 1] = l2;
          }
        }
      }
    }
  }
  if (sig.counter != nullptr) {
    // publish (same release pattern as score_prep_uniform_kernel): block barrier, one system fence, block counter;
    // the LAST test block raises this source rank's ready flag in every destination region
    __shared__ int s_last;
    if (sig.fence_per_thread) __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence_system();
      const unsigned prev = atomicAdd(sig.counter, 1u);
      s_last = prev == test_blocks - 1 ? 1 : 0;
      if (s_last) atomicExch(sig.counter, 0u);
    }
    __syncthreads();
    if (s_last && static_cast<int>(threadIdx.x) < sig.n) {
      __threadfence_system();
      asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(sig.flag[threadIdx.x]), "r"(sig.epoch) : "memory");
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

namespace {
inline bool rows_vectorisable(const void* p, int64_t ld, bool is_f32) {
  // 8-column groups start 16-byte aligned: fp32 rows need ld % 4 == 0, fp64 rows ld % 2 == 0
  return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && ld % (is_f32 ? 4 : 2) == 0;
}
}  // namespace

void score_prep_uniform_multi(Context& ctx, const void* enrol, int64_t ne, int64_t ld_e, SplitBuf* l_out,
                              float* row_term, const void* test, int64_t nt, int64_t ld_t, int64_t test_row0,
                              int64_t test_pad_end, const PrepDst& tdst, int64_t ld_out, bool is_f32, int64_t d,
                              const double* consts, const PrepSignal& sig) {
  PB_CHECK(d <= 1024, kInvalidArg, "score: dimension above 1024 is not supported");
  PB_CHECK(ld_out % 16 == 0 && ld_out >= d, kInvalidArg, "score prep: operand pitch must be a multiple of 16");
  PB_CHECK(tdst.n >= 0 && tdst.n <= kMaxPeers && sig.n <= kMaxPeers, kInvalidArg, "score prep: too many destinations");
  if (l_out) l_out->reserve(ne, d, ld_out);      // one pitch for both sides (a session may leave room for group columns)
  const int ldo = static_cast<int>(l_out ? l_out->ld : ld_out);
  // two resident blocks per SM, split between the sides in proportion to their rows; a small side gets one block
  // per 32 rows (one round of four rows per warp).  When the test rows travel to peers (tdst.n > 1) the test side
  // instead gets one block per 8 rows, up to one per SM: the blocks are first in the grid, short, and many warps in
  // flight hide the NVLink store latency; the enrol blocks take their slots as they retire.
  const int64_t e_rows = l_out ? ne : 0;
  const int64_t t_rows = tdst.n > 0 ? nt : 0;
  const int64_t slots = 2ll * ctx.num_sms;
  auto side_blocks = [&](int64_t rows) -> unsigned {
    if (rows <= 0) return 0u;
    const int64_t share = std::max<int64_t>(1, slots * rows / (e_rows + t_rows));
    return static_cast<unsigned>(std::min<int64_t>(ceil_div(rows, 32), share));
  };
  const unsigned eb = side_blocks(e_rows);
  unsigned tb = side_blocks(t_rows);
  if (tdst.n > 1 && t_rows > 0)
    tb = static_cast<unsigned>(std::max<int64_t>(tb, std::min<int64_t>(ceil_div(t_rows, 8), ctx.num_sms)));
  // an empty shard still has to raise its ready flag / write the padding of the column-term row
  if (tb == 0 && tdst.n > 0 && (sig.counter != nullptr || test_pad_end > test_row0 + nt)) tb = 1;
  if (eb + tb == 0) return;
  const int vec_e = enrol && rows_vectorisable(enrol, ld_e, is_f32) ? 1 : 0;
  const int vec_t = test && rows_vectorisable(test, ld_t, is_f32) ? 1 : 0;
  __nv_bfloat16* lhi = l_out ? l_out->hi.get() : nullptr;
  __nv_bfloat16* llo = l_out ? l_out->lo.get() : nullptr;
  if (is_f32)
    score_prep_uniform_kernel<float><<<eb + tb, 256, 0, ctx.stream>>>(
        static_cast<const float*>(enrol), l_out ? ne : 0, ld_e, static_cast<const float*>(test), nt, ld_t, test_row0,
        test_pad_end, static_cast<int>(d), consts, lhi, llo, row_term, tdst, ldo, tb, vec_e, vec_t, sig);
  else
    score_prep_uniform_kernel<double><<<eb + tb, 256, 0, ctx.stream>>>(
        static_cast<const double*>(enrol), l_out ? ne : 0, ld_e, static_cast<const double*>(test), nt, ld_t, test_row0,
        test_pad_end, static_cast<int>(d), consts, lhi, llo, row_term, tdst, ldo, tb, vec_e, vec_t, sig);
  PB_CUDA(cudaGetLastError());
  ctx.count_launch();
  ctx.pdl_pending = ctx.pdl_enabled;
}

void score_prep_uniform(Context& ctx, const void* enrol, int64_t ne, int64_t ld_e, const void* test, int64_t nt,
                        int64_t ld_t, bool is_f32, int64_t d, const double* consts, SplitBuf& l_out,
                        SplitBuf& r_out, float* row_term, float* col_term, int64_t col_ld) {
  r_out.reserve(nt, d);
  PrepDst tdst;
  tdst.n = 1;
  tdst.hi[0] = r_out.hi.get();
  tdst.lo[0] = r_out.lo.get();
  tdst.term[0] = col_term;
  score_prep_uniform_multi(ctx, enrol, ne, ld_e, &l_out, row_term, test, nt, ld_t, 0, col_ld, tdst, r_out.ld, is_f32, d,
                           consts, PrepSignal{});
}

void score_prep_grouped(Context& ctx, const void* enrol, int64_t ne, int64_t ld_e, const int32_t* grp_dev,
                        const void* test, int64_t nt, int64_t ld_t, bool is_f32, int64_t d, int ng,
                        const double* tables_dev, SplitBuf& l_out, SplitBuf& r_out, float* row_term, float* col_term,
                        int64_t col_ld) {
  PB_CHECK(d <= 1024 && ng >= 1, kInvalidArg, "score: dimension above 1024 is not supported");
  l_out.reserve(ne, d);
  r_out.reserve(nt, d);
  const unsigned eb = row_blocks(ne), tb = row_blocks(nt);
  if (eb + tb == 0) return;
  if (is_f32)
    score_prep_grouped_kernel<float><<<eb + tb, kWarpsPerBlock * 32, 0, ctx.stream>>>(
        static_cast<const float*>(enrol), ne, ld_e, grp_dev, static_cast<const float*>(test), nt, ld_t,
        static_cast<int>(d), ng, tables_dev, l_out.hi.get(), l_out.lo.get(), r_out.hi.get(), r_out.lo.get(),
        static_cast<int>(l_out.ld), row_term, col_term, col_ld, eb);
  else
    score_prep_grouped_kernel<double><<<eb + tb, kWarpsPerBlock * 32, 0, ctx.stream>>>(
        static_cast<const double*>(enrol), ne, ld_e, grp_dev, static_cast<const double*>(test), nt, ld_t,
        static_cast<int>(d), ng, tables_dev, l_out.hi.get(), l_out.lo.get(), r_out.hi.get(), r_out.lo.get(),
        static_cast<int>(l_out.ld), row_term, col_term, col_ld, eb);
  PB_CUDA(cudaGetLastError());
  ctx.count_launch();
}

void score_prep_grouped_vec(Context& ctx, const void* enrol, int64_t ne, int64_t ld_e, const int32_t* grp_dev,
                            const void* test, int64_t nt, int64_t ld_t, bool is_f32, int64_t d, int ng,
                            const double* tables_dev, SplitBuf& l_out, SplitBuf& r_out, float* row_term,
                            float* col_term, int64_t col_ld, bool embed, const PrepDst* shard_dst, int64_t test_row0,
                            int64_t shard_ld, const PrepSignal* sig) {
  PB_CHECK(d <= 1024 && ng >= 1, kInvalidArg, "score: dimension above 1024 is not supported");
  const int64_t kk = embed ? d + 2 * ng : d;      // embed: two extra K columns per distinct count
  l_out.reserve(ne, kk);
  PrepDst dst;
  int64_t ld_out = 0;
  if (shard_dst != nullptr) {
    // sharded grid: the test rows go to the peers' operand buffers (pitch shard_ld), not to r_out
    PB_CHECK(embed && shard_ld >= round_up(kk, 16) && shard_ld % 8 == 0, kInvalidArg,
             "score: the sharded operand pitch has no room for the group columns");
    dst = *shard_dst;
    ld_out = shard_ld;
  } else {
    r_out.reserve(nt, kk);
    dst.n = 1;
    dst.hi[0] = r_out.hi.get();
    dst.lo[0] = r_out.lo.get();
    dst.term[0] = nullptr;
    ld_out = r_out.ld;
  }
  const unsigned eb = row_blocks(ne), tb = row_blocks(nt);
  // a sharded producer with no test rows still has to raise its flags: one (empty) test block
  const unsigned tb_launch = (sig != nullptr && sig->counter != nullptr && tb == 0) ? 1u : tb;
  if (eb + tb_launch == 0) return;
  const int vec_e = enrol && rows_vectorisable(enrol, ld_e, is_f32) ? 1 : 0;
  const int vec_t = test && rows_vectorisable(test, ld_t, is_f32) ? 1 : 0;
  const PrepSignal sg = sig != nullptr ? *sig : PrepSignal{};
  if (is_f32)
    score_prep_grouped_vec_kernel<float><<<eb + tb_launch, kWarpsPerBlock * 32, 0, ctx.stream>>>(
        static_cast<const float*>(enrol), ne, ld_e, grp_dev, static_cast<const float*>(test), nt, ld_t,
        static_cast<int>(d), ng, tables_dev, l_out.hi.get(), l_out.lo.get(), static_cast<int>(l_out.ld), dst, test_row0,
        static_cast<int>(ld_out), row_term, col_term, col_ld, tb_launch, vec_e, vec_t, embed ? 1 : 0, sg);
  else
    score_prep_grouped_vec_kernel<double><<<eb + tb_launch, kWarpsPerBlock * 32, 0, ctx.stream>>>(
        static_cast<const double*>(enrol), ne, ld_e, grp_dev, static_cast<const double*>(test), nt, ld_t,
        static_cast<int>(d), ng, tables_dev, l_out.hi.get(), l_out.lo.get(), static_cast<int>(l_out.ld), dst, test_row0,
        static_cast<int>(ld_out), row_term, col_term, col_ld, tb_launch, vec_e, vec_t, embed ? 1 : 0, sg);
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
