// K1 (label segmentation) and K2/K4 (per-speaker segmented sums), plus the centred /
// scaled operand producers of the scatter SYRK (K3's input).
//
// Reference behaviour replaced: the std::set / std::vector bucketing of
// src/pldamodule.cpp:76-92 (fit) and the std::map accumulation of :147-156 (transform).
// Labels are 8 bytes per row against d*8 bytes of features, so the label sort uses the
// toolkit's cub::DeviceRadixSort; everything touching the feature rows is hand-written.
#include <algorithm>

#include <cub/cub.cuh>

#include "kernels.h"

namespace pb {
namespace {

__global__ void iota_kernel(int32_t* __restrict__ v, long long n) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i < n) v[i] = static_cast<int32_t>(i);
}

__global__ void head_flags_kernel(const uint64_t* __restrict__ keys, long long n, int32_t* __restrict__ flags) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i < n) flags[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

// Head flags on the rows as given + the identity order; *unsorted is raised when a label is smaller than its
// predecessor (the caller then falls back to the radix sort).
__global__ void head_flags_order_kernel(const uint64_t* __restrict__ keys, long long n, int32_t* __restrict__ flags,
                                        int32_t* __restrict__ order, int32_t* __restrict__ unsorted) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const uint64_t k = keys[i];
  const uint64_t prev = i == 0 ? k : keys[i - 1];
  flags[i] = (i == 0 || k != prev) ? 1 : 0;
  order[i] = static_cast<int32_t>(i);
  if (k < prev) *unsorted = 1;
}

// seg_of_pos holds inclusive_scan(flags): this makes it zero-based (each thread its own element) and scatters the
// heads (label and first position of every segment).
__global__ void scatter_heads_kernel(const uint64_t* __restrict__ keys, const int32_t* __restrict__ flags,
                                     int32_t* __restrict__ seg_of_pos, long long n, uint64_t* __restrict__ seg_label,
                                     int32_t* __restrict__ seg_start, long long nseg) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const int32_t s = seg_of_pos[i] - 1;
  seg_of_pos[i] = s;
  if (flags[i] != 0) {
    seg_label[s] = keys[i];
    seg_start[s] = static_cast<int32_t>(i);
  }
  if (i == n - 1) seg_start[nseg] = static_cast<int32_t>(n);
}

// Balanced segmented sum: each warp owns kRowsPerWarp consecutive SORTED positions, keeps a running
// fp64 partial per column in registers and flushes it with atomicAdd at every segment change.
// Work is independent of the segment-size distribution (2 speakers or 50k speakers).
constexpr int kRowsPerWarp = 32;
constexpr int kMaxColsPerLane = 32;   // d <= 1024

template <typename T>
__global__ void __launch_bounds__(256)
segment_sums_kernel(const T* __restrict__ x, int d, long long ld, const int32_t* __restrict__ order,
                    const int32_t* __restrict__ seg_of_pos, long long n, double* __restrict__ sums) {
  const long long warp = blockIdx.x * static_cast<long long>(blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const long long p0 = warp * kRowsPerWarp;
  if (p0 >= n) return;
  const long long p1 = min(p0 + kRowsPerWarp, n);
  const int ncol = (d + 31) >> 5;
  double acc[kMaxColsPerLane];
#pragma unroll
  for (int j = 0; j < kMaxColsPerLane; ++j) acc[j] = 0.0;
  int cur = seg_of_pos[p0];
  for (long long p = p0; p < p1; ++p) {
    const int s = seg_of_pos[p];
    if (s != cur) {
#pragma unroll
      for (int j = 0; j < kMaxColsPerLane; ++j) {
        if (j < ncol) {
          const int c = lane + 32 * j;
          if (c < d) atomicAdd(sums + static_cast<long long>(cur) * d + c, acc[j]);
          acc[j] = 0.0;
        }
      }
      cur = s;
    }
    const T* row = x + static_cast<long long>(order[p]) * ld;
#pragma unroll
    for (int j = 0; j < kMaxColsPerLane; ++j) {
      if (j < ncol) {
        const int c = lane + 32 * j;
        if (c < d) acc[j] += static_cast<double>(row[c]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < kMaxColsPerLane; ++j) {
    if (j < ncol) {
      const int c = lane + 32 * j;
      if (c < d) atomicAdd(sums + static_cast<long long>(cur) * d + c, acc[j]);
    }
  }
}

// 16-byte vectorised variant (float4 / double2 loads): used when d, the row pitch and the base address allow it.
template <typename T, int VEC>
__global__ void __launch_bounds__(256)
segment_sums_vec_kernel(const T* __restrict__ x, int d, long long ld, const int32_t* __restrict__ order,
                        const int32_t* __restrict__ seg_of_pos, long long n, double* __restrict__ sums) {
  constexpr int kMaxVec = 1024 / (32 * VEC);      // vector slots per lane for d <= 1024
  struct alignas(16) Vec { T v[VEC]; };
  const long long warp = blockIdx.x * static_cast<long long>(blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const long long p0 = warp * kRowsPerWarp;
  if (p0 >= n) return;
  const long long p1 = min(p0 + kRowsPerWarp, n);
  const int nvec = d / VEC;                        // vectors per row
  double acc[kMaxVec][VEC];
#pragma unroll
  for (int j = 0; j < kMaxVec; ++j)
#pragma unroll
    for (int e = 0; e < VEC; ++e) acc[j][e] = 0.0;
  auto flush = [&](int seg) {
#pragma unroll
    for (int j = 0; j < kMaxVec; ++j) {
      const int vi = lane + 32 * j;
      if (vi < nvec) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          atomicAdd(sums + static_cast<long long>(seg) * d + vi * VEC + e, acc[j][e]);
          acc[j][e] = 0.0;
        }
      }
    }
  };
  int cur = seg_of_pos[p0];
  for (long long p = p0; p < p1; ++p) {
    const int sg = seg_of_pos[p];
    if (sg != cur) { flush(cur); cur = sg; }
    const Vec* row = reinterpret_cast<const Vec*>(x + static_cast<long long>(order[p]) * ld);
    Vec tmp[kMaxVec];
#pragma unroll
    for (int j = 0; j < kMaxVec; ++j) {            // all loads of the row first (memory-level parallelism)
      const int vi = lane + 32 * j;
      if (vi < nvec) tmp[j] = row[vi];
    }
#pragma unroll
    for (int j = 0; j < kMaxVec; ++j) {
      const int vi = lane + 32 * j;
      if (vi < nvec) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[j][e] += static_cast<double>(tmp[j].v[e]);
      }
    }
  }
  flush(cur);
}

__global__ void finalize_means_kernel(double* __restrict__ sums, int d, const int32_t* __restrict__ seg_start,
                                      long long nseg, int32_t* __restrict__ counts) {
  const long long s = blockIdx.x;
  const int cnt = seg_start[s + 1] - seg_start[s];
  const double inv = 1.0 / static_cast<double>(cnt);
  for (int c = threadIdx.x; c < d; c += blockDim.x) sums[s * d + c] *= inv;
  if (threadIdx.x == 0 && counts) counts[s] = cnt;
}

// Transposing producer: tile of 32 sorted positions x 32 columns through smem so that both the gathered
// row reads (along d) and the transposed split-bf16 writes (along the position axis) are coalesced.
template <typename T>
__global__ void __launch_bounds__(256)
center_scale_split_t_kernel(const T* __restrict__ x, int d, long long ld, const int32_t* __restrict__ order,
                            const int32_t* __restrict__ seg_of_pos, const int32_t* __restrict__ seg_start,
                            long long n, const double* __restrict__ means, int scale_by_count,
                            __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, long long ld_out) {
  __shared__ float s_hi[32][33];
  __shared__ float s_lo[32][33];
  const long long p0 = blockIdx.x * 32ll;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const long long p = p0 + i;
    const int c = c0 + tx;
    double v = 0.0;
    if (p < n && c < d) {
      const int s = seg_of_pos[p];
      const double sc = scale_by_count ? rsqrt(static_cast<double>(seg_start[s + 1] - seg_start[s])) : 1.0;
      v = (static_cast<double>(x[static_cast<long long>(order[p]) * ld + c]) - means[static_cast<long long>(s) * d + c]) * sc;
    }
    __nv_bfloat16 h, l;
    split_bf16(v, h, l);
    s_hi[i][tx] = __bfloat162float(h);
    s_lo[i][tx] = __bfloat162float(l);
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i;
    const long long p = p0 + tx;
    if (c < d && p < ld_out) {
      hi[c * ld_out + p] = __float2bfloat16_rn(s_hi[tx][i]);
      lo[c * ld_out + p] = __float2bfloat16_rn(s_lo[tx][i]);
    }
  }
}

// Wider version of the same producer for aligned rows: tile = 64 sorted positions x 64 columns.
//   in   16 threads per row, 4 consecutive columns each (one 16-byte load for fp32 rows, two for fp64), 16 rows per
//        pass; the per-row scalars (source row, class, 1/sqrt(n_class)) are resolved once per row in shared memory
//        instead of once per element
//   out  every warp store writes one full 128-byte line: 64 positions x bf16 of ONE column of one plane
// Same arithmetic as the 32 x 32 kernel (fp64 centring and scaling before the split), so the operand bits are equal.
template <typename T>
__global__ void __launch_bounds__(256)
center_scale_split_t64_kernel(const T* __restrict__ x, int d, long long ld, const int32_t* __restrict__ order,
                              const int32_t* __restrict__ seg_of_pos, const int32_t* __restrict__ seg_start,
                              long long n, const double* __restrict__ means, int scale_by_count,
                              __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, long long ld_out) {
  constexpr int TP = 64, TC = 64, PITCH = TP + 2;       // bf16 elements per smem row (pad: 2-way conflicts at most)
  __shared__ __align__(16) unsigned short s_hi[TC][PITCH];
  __shared__ __align__(16) unsigned short s_lo[TC][PITCH];
  __shared__ long long s_src[TP];
  __shared__ long long s_mean[TP];
  __shared__ double s_scale[TP];
  const long long p0 = blockIdx.x * static_cast<long long>(TP);
  const int c0 = blockIdx.y * TC;
  if (threadIdx.x < TP) {
    const long long p = p0 + threadIdx.x;
    long long src = -1, mrow = 0;
    double sc = 0.0;
    if (p < n) {
      const int sg = seg_of_pos[p];
      src = static_cast<long long>(order[p]) * ld;
      mrow = static_cast<long long>(sg) * d;
      sc = scale_by_count ? rsqrt(static_cast<double>(seg_start[sg + 1] - seg_start[sg])) : 1.0;
    }
    s_src[threadIdx.x] = src;
    s_mean[threadIdx.x] = mrow;
    s_scale[threadIdx.x] = sc;
  }
  __syncthreads();
  const int t16 = threadIdx.x & 15, r16 = threadIdx.x >> 4;
  const int c = c0 + 4 * t16;
#pragma unroll
  for (int pass = 0; pass < TP / 16; ++pass) {
    const int row = pass * 16 + r16;
    const long long src = s_src[row];
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    if (src >= 0 && c < d) {
      if (c + 4 <= d) {
        if (sizeof(T) == 4) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(x + src + c));
          v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
        } else {
          const double2 a = __ldg(reinterpret_cast<const double2*>(x + src + c));
          const double2 b = __ldg(reinterpret_cast<const double2*>(x + src + c + 2));
          v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) if (c + j < d) v[j] = static_cast<double>(x[src + c + j]);
      }
      const double sc = s_scale[row];
      const double* m = means + s_mean[row] + c;
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = c + j < d ? (v[j] - m[j]) * sc : 0.0;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      __nv_bfloat16 h, l;
      split_bf16(v[j], h, l);
      s_hi[4 * t16 + j][row] = __bfloat16_as_ushort(h);
      s_lo[4 * t16 + j][row] = __bfloat16_as_ushort(l);
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // 128 line stores (64 columns x 2 planes), 16 per warp; a lane carries two positions
  for (int i = warp; i < 2 * TC; i += 8) {
    const int cl = i >> 1;
    const int cg = c0 + cl;
    if (cg >= d) continue;
    const unsigned short* srow = (i & 1) ? s_lo[cl] : s_hi[cl];
    const uint32_t pair = static_cast<uint32_t>(srow[2 * lane]) | (static_cast<uint32_t>(srow[2 * lane + 1]) << 16);
    __nv_bfloat16* dst = ((i & 1) ? lo : hi) + static_cast<long long>(cg) * ld_out + p0;
    if (p0 + 2 * lane + 1 < ld_out) *reinterpret_cast<uint32_t*>(dst + 2 * lane) = pair;
  }
}

template <typename T>
__global__ void center_scale_f64_kernel(const T* __restrict__ x, int d, long long ld, const int32_t* __restrict__ order,
                                        const int32_t* __restrict__ seg_of_pos, const int32_t* __restrict__ seg_start,
                                        long long n, const double* __restrict__ means, int scale_by_count,
                                        double* __restrict__ out) {
  const long long p = blockIdx.x;
  const int s = seg_of_pos[p];
  const double sc = scale_by_count ? rsqrt(static_cast<double>(seg_start[s + 1] - seg_start[s])) : 1.0;
  const T* row = x + static_cast<long long>(order[p]) * ld;
  for (int c = threadIdx.x; c < d; c += blockDim.x)
    out[p * d + c] = (static_cast<double>(row[c]) - means[static_cast<long long>(s) * d + c]) * sc;
}

// sum_out[c] += sum_s means[s,c]/n_s ; class_weight += sum_s 1/n_s.  Grid: 32 columns x a slice of the classes
// per block (fp64 atomics across slices); outputs are zeroed by the launcher.
__global__ void __launch_bounds__(256)
class_weighted_sum_kernel(const double* __restrict__ means, const int32_t* __restrict__ counts, long long k, int d,
                          double* __restrict__ sum_out, double* __restrict__ class_weight_out) {
  __shared__ double red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  double acc = 0.0, wacc = 0.0;
  for (long long s = static_cast<long long>(blockIdx.y) * 8 + ty; s < k; s += 8ll * gridDim.y) {
    const double w = 1.0 / static_cast<double>(counts[s]);
    if (c < d) acc += w * means[s * d + c];
    wacc += w;
  }
  red[ty][tx] = acc;
  __syncthreads();
  if (ty == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += red[i][tx];
    if (c < d) atomicAdd(sum_out + c, t);
  }
  if (blockIdx.x == 0 && class_weight_out) {
    __syncthreads();
    red[ty][tx] = wacc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int i = 0; i < 8; ++i) t += red[i][0];
      atomicAdd(class_weight_out, t);
    }
  }
}

inline unsigned blocks_for(long long n, int t) { return static_cast<unsigned>(ceil_div(n, t)); }

}  // namespace

void build_segments(Context& ctx, const uint64_t* labels_dev, int64_t n, Segments& seg) {
  PB_CHECK(n > 0 && n < (1ll << 31), kInvalidArg, "labels: need 1 <= n < 2^31 rows");
  seg.n = n;
  seg.vals_in.reserve(n);
  seg.order.reserve(n);
  seg.seg_of_pos.reserve(n);
  seg.flags.reserve(2);
  size_t tmp_bytes = 0, tmp2 = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, labels_dev, seg.keys_out.get(), seg.vals_in.get(),
                                  seg.order.get(), static_cast<int>(n), 0, 64, ctx.stream);
  cub::DeviceScan::InclusiveSum(nullptr, tmp2, seg.vals_in.get(), seg.seg_of_pos.get(), static_cast<int>(n), ctx.stream);
  seg.cub_tmp.reserve(std::max(tmp_bytes, tmp2));
  // Pass 1 assumes what callers nearly always hand over -- rows already grouped by ascending label: head flags and
  // the identity order in one kernel, which also records whether any label is smaller than its predecessor.
  PB_CUDA(cudaMemsetAsync(seg.flags.get(), 0, 2 * sizeof(int32_t), ctx.stream));
  head_flags_order_kernel<<<blocks_for(n, 256), 256, 0, ctx.stream>>>(labels_dev, n, seg.vals_in.get(),
                                                                     seg.order.get(), seg.flags.get());
  size_t avail = seg.cub_tmp.size();
  PB_CUDA(cub::DeviceScan::InclusiveSum(seg.cub_tmp.get(), avail, seg.vals_in.get(), seg.seg_of_pos.get(),
                                        static_cast<int>(n), ctx.stream));
  ctx.count_launch(2);
  int32_t last = 0, unsorted = 0;
  PB_CUDA(cudaMemcpyAsync(&last, seg.seg_of_pos.get() + (n - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, ctx.stream));
  PB_CUDA(cudaMemcpyAsync(&unsorted, seg.flags.get(), sizeof(int32_t), cudaMemcpyDeviceToHost, ctx.stream));
  ctx.sync();
  const uint64_t* keys = labels_dev;
  if (unsorted != 0) {
    // general case: stable radix sort of (label, row), then the same flags / scan on the sorted keys
    seg.keys_out.reserve(n);
    iota_kernel<<<blocks_for(n, 256), 256, 0, ctx.stream>>>(seg.vals_in.get(), n);
    avail = seg.cub_tmp.size();
    PB_CUDA(cub::DeviceRadixSort::SortPairs(seg.cub_tmp.get(), avail, labels_dev, seg.keys_out.get(),
                                            seg.vals_in.get(), seg.order.get(), static_cast<int>(n), 0, 64, ctx.stream));
    head_flags_kernel<<<blocks_for(n, 256), 256, 0, ctx.stream>>>(seg.keys_out.get(), n, seg.vals_in.get());
    avail = seg.cub_tmp.size();
    PB_CUDA(cub::DeviceScan::InclusiveSum(seg.cub_tmp.get(), avail, seg.vals_in.get(), seg.seg_of_pos.get(),
                                          static_cast<int>(n), ctx.stream));
    ctx.count_launch(4);
    PB_CUDA(cudaMemcpyAsync(&last, seg.seg_of_pos.get() + (n - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, ctx.stream));
    ctx.sync();
    keys = seg.keys_out.get();
  }
  seg.nseg = last;
  seg.seg_label.reserve(seg.nseg);
  seg.seg_start.reserve(seg.nseg + 1);
  scatter_heads_kernel<<<blocks_for(n, 256), 256, 0, ctx.stream>>>(keys, seg.vals_in.get(), seg.seg_of_pos.get(), n,
                                                                  seg.seg_label.get(), seg.seg_start.get(), seg.nseg);
  PB_CUDA(cudaGetLastError());
  ctx.count_launch();
}

void segment_sums(Context& ctx, const void* x, bool is_f32, int64_t d, int64_t ld, const Segments& seg, double* sums) {
  PB_CHECK(d <= 32 * kMaxColsPerLane, kInvalidArg, "feature dimension above 1024 is not supported");
  PB_CUDA(cudaMemsetAsync(sums, 0, seg.nseg * d * sizeof(double), ctx.stream));
  const long long warps = ceil_div(seg.n, kRowsPerWarp);
  const unsigned blocks = static_cast<unsigned>(ceil_div(warps, 8));
  const int vec = is_f32 ? 4 : 2;
  const bool can_vec = d % vec == 0 && ld % vec == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  if (can_vec && is_f32)
    segment_sums_vec_kernel<float, 4><<<blocks, 256, 0, ctx.stream>>>(static_cast<const float*>(x), static_cast<int>(d),
                                                                      ld, seg.order.get(), seg.seg_of_pos.get(), seg.n,
                                                                      sums);
  else if (can_vec)
    segment_sums_vec_kernel<double, 2><<<blocks, 256, 0, ctx.stream>>>(static_cast<const double*>(x),
                                                                       static_cast<int>(d), ld, seg.order.get(),
                                                                       seg.seg_of_pos.get(), seg.n, sums);
  else if (is_f32)
    segment_sums_kernel<float><<<blocks, 256, 0, ctx.stream>>>(static_cast<const float*>(x), static_cast<int>(d), ld,
                                                               seg.order.get(), seg.seg_of_pos.get(), seg.n, sums);
  else
    segment_sums_kernel<double><<<blocks, 256, 0, ctx.stream>>>(static_cast<const double*>(x), static_cast<int>(d), ld,
                                                                seg.order.get(), seg.seg_of_pos.get(), seg.n, sums);
  PB_CUDA(cudaGetLastError());
  ctx.count_launch();
}

void segment_finalize_means(Context& ctx, double* sums, int64_t d, const Segments& seg, int32_t* counts_out) {
  finalize_means_kernel<<<static_cast<unsigned>(seg.nseg), 128, 0, ctx.stream>>>(sums, static_cast<int>(d),
                                                                                 seg.seg_start.get(), seg.nseg,
                                                                                 counts_out);
  PB_CUDA(cudaGetLastError());
  ctx.count_launch();
}

void center_scale_split_t(Context& ctx, const void* x, bool is_f32, int64_t d, int64_t ld, const Segments& seg,
                          const double* means, bool scale_by_count, SplitBuf& xt) {
  // operand [d rows x n cols]; pad the reduction axis to a multiple of 64 and zero the tail
  const int64_t npad = round_up(seg.n, 64);
  xt.rows = d;
  xt.k = seg.n;
  xt.ld = npad;
  xt.hi.reserve(static_cast<size_t>(d) * npad);
  xt.lo.reserve(static_cast<size_t>(d) * npad);
  // aligned rows take the 64 x 64 tile kernel (16-byte loads, full-line stores); npad is a multiple of 64
  const bool wide = (reinterpret_cast<uintptr_t>(x) & 15) == 0 && ld % (is_f32 ? 4 : 2) == 0;
  if (wide) {
    dim3 grid64(static_cast<unsigned>(npad / 64), static_cast<unsigned>(ceil_div(d, 64)));
    if (is_f32)
      center_scale_split_t64_kernel<float><<<grid64, 256, 0, ctx.stream>>>(
          static_cast<const float*>(x), static_cast<int>(d), ld, seg.order.get(), seg.seg_of_pos.get(),
          seg.seg_start.get(), seg.n, means, scale_by_count ? 1 : 0, xt.hi.get(), xt.lo.get(), npad);
    else
      center_scale_split_t64_kernel<double><<<grid64, 256, 0, ctx.stream>>>(
          static_cast<const double*>(x), static_cast<int>(d), ld, seg.order.get(), seg.seg_of_pos.get(),
          seg.seg_start.get(), seg.n, means, scale_by_count ? 1 : 0, xt.hi.get(), xt.lo.get(), npad);
    PB_CUDA(cudaGetLastError());
    ctx.count_launch();
    return;
  }
  dim3 grid(static_cast<unsigned>(npad / 32), static_cast<unsigned>(ceil_div(d, 32)));
  if (is_f32)
    center_scale_split_t_kernel<float><<<grid, 256, 0, ctx.stream>>>(
        static_cast<const float*>(x), static_cast<int>(d), ld, seg.order.get(), seg.seg_of_pos.get(),
        seg.seg_start.get(), seg.n, means, scale_by_count ? 1 : 0, xt.hi.get(), xt.lo.get(), npad);
  else
    center_scale_split_t_kernel<double><<<grid, 256, 0, ctx.stream>>>(
        static_cast<const double*>(x), static_cast<int>(d), ld, seg.order.get(), seg.seg_of_pos.get(),
        seg.seg_start.get(), seg.n, means, scale_by_count ? 1 : 0, xt.hi.get(), xt.lo.get(), npad);
  PB_CUDA(cudaGetLastError());
  ctx.count_launch();
}

void center_scale_f64(Context& ctx, const void* x, bool is_f32, int64_t d, int64_t ld, const Segments& seg,
                      const double* means, bool scale_by_count, double* out) {
  if (is_f32)
    center_scale_f64_kernel<float><<<static_cast<unsigned>(seg.n), 128, 0, ctx.stream>>>(
        static_cast<const float*>(x), static_cast<int>(d), ld, seg.order.get(), seg.seg_of_pos.get(),
        seg.seg_start.get(), seg.n, means, scale_by_count ? 1 : 0, out);
  else
    center_scale_f64_kernel<double><<<static_cast<unsigned>(seg.n), 128, 0, ctx.stream>>>(
        static_cast<const double*>(x), static_cast<int>(d), ld, seg.order.get(), seg.seg_of_pos.get(),
        seg.seg_start.get(), seg.n, means, scale_by_count ? 1 : 0, out);
  PB_CUDA(cudaGetLastError());
  ctx.count_launch();
}

void class_weighted_sum(Context& ctx, const double* means, const int32_t* counts, int64_t k, int64_t d,
                        double* sum_out, double* class_weight_out) {
  PB_CUDA(cudaMemsetAsync(sum_out, 0, d * sizeof(double), ctx.stream));
  if (class_weight_out) PB_CUDA(cudaMemsetAsync(class_weight_out, 0, sizeof(double), ctx.stream));
  const unsigned slices = static_cast<unsigned>(std::min<int64_t>(64, std::max<int64_t>(1, k / 64)));
  class_weighted_sum_kernel<<<dim3(static_cast<unsigned>(ceil_div(d, 32)), slices), 256, 0, ctx.stream>>>(
      means, counts, k, static_cast<int>(d), sum_out, class_weight_out);
  PB_CUDA(cudaGetLastError());
  ctx.count_launch();
}

}  // namespace pb
