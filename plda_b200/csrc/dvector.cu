// d-vector pooling: frame-level network activations -> one vector per utterance.
//
// Replaces scoring/extractdvector.py:19-58 for a BATCH of utterances: getnormalizedvector (each frame divided by its
// L2 norm, :19-29) followed by the mean / max / (population) variance over the frames of the utterance (:32-46), and
// the *_nol2 variants (:49-58) that pool the raw frames.  The reference loops over utterances in Python
// (extractvectors, :160-168); here the frames of all utterances are one [N x d] matrix with CSR offsets per utterance.
//
// HBM-bound segmented reduction: every frame is read from HBM once.  A block owns one work item (an utterance, or a
// chunk of at most kChunk frames of a long one) and walks it in batches of 32 frames:
//   phase A  one warp per frame: coalesced row read, warp-shuffle sum of squares -> 1/norm in shared memory
//   phase B  one thread per column (columns t, t+256, ...): the 32 frames of the batch are re-read from L1/L2 with
//            coalesced rows and accumulated (sum, sum of squares, max) in fp64 registers
// Items of one utterance are merged by a second kernel (deterministic, no atomics).
#include "kernels.h"

namespace pb {
namespace {

constexpr int kThreads = 256;
constexpr int kBatch = 32;
constexpr int kMaxColsPerThread = 4;       // d <= 1024
constexpr long long kChunk = 2048;

template <typename T>
__global__ void __launch_bounds__(kThreads)
dvector_pool_kernel(const T* __restrict__ frames, long long ld, int d, const long long* __restrict__ item_start,
                    const long long* __restrict__ item_end, int l2norm, double* __restrict__ part_sum,
                    double* __restrict__ part_sq, double* __restrict__ part_max) {
  __shared__ double s_nrm[kBatch];
  const long long item = blockIdx.x;
  const long long f0 = item_start[item], f1 = item_end[item];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double sum[kMaxColsPerThread], sq[kMaxColsPerThread], mx[kMaxColsPerThread];
#pragma unroll
  for (int j = 0; j < kMaxColsPerThread; ++j) { sum[j] = 0.0; sq[j] = 0.0; mx[j] = -INFINITY; }
  for (long long b0 = f0; b0 < f1; b0 += kBatch) {
    const int nb = static_cast<int>(min(static_cast<long long>(kBatch), f1 - b0));
    if (l2norm) {
      __syncthreads();                       // the previous batch is done with s_nrm
      for (int i = warp; i < nb; i += kThreads / 32) {
        const T* row = frames + (b0 + i) * ld;
        double a = 0.0;
        for (int c = lane; c < d; c += 32) {
          const double v = static_cast<double>(row[c]);
          a += v * v;
        }
        a = warp_sum(a);
        if (lane == 0) s_nrm[i] = sqrt(a);   // np.linalg.norm(utt, axis=1), extractdvector.py:28
      }
      __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < kMaxColsPerThread; ++j) {
      const int c = threadIdx.x + j * kThreads;
      if (c >= d) break;
      for (int i = 0; i < nb; ++i) {
        double v = static_cast<double>(frames[(b0 + i) * ld + c]);
        if (l2norm) v = v / s_nrm[i];        // uttvec / denom[:, np.newaxis], :29 (a zero frame gives nan, as in numpy)
        sum[j] += v;
        sq[j] += v * v;
        mx[j] = (v > mx[j] || v != v) ? v : mx[j];     // nan propagates like np.max
      }
    }
  }
#pragma unroll
  for (int j = 0; j < kMaxColsPerThread; ++j) {
    const int c = threadIdx.x + j * kThreads;
    if (c >= d) break;
    part_sum[item * d + c] = sum[j];
    part_sq[item * d + c] = sq[j];
    part_max[item * d + c] = mx[j];
  }
}

// mode 0: mean (np.mean, :39), 1: max (np.max, :34), 2: population variance (np.var, :46)
__global__ void dvector_finish_kernel(const double* __restrict__ part_sum, const double* __restrict__ part_sq,
                                      const double* __restrict__ part_max, const long long* __restrict__ utt_item0,
                                      const long long* __restrict__ utt_frames, int d, int mode,
                                      double* __restrict__ out, long long ldo) {
  const long long u = blockIdx.x;
  const long long i0 = utt_item0[u], i1 = utt_item0[u + 1];
  const double n = static_cast<double>(utt_frames[u]);
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    double s = 0.0, q = 0.0, m = -INFINITY;
    for (long long i = i0; i < i1; ++i) {
      s += part_sum[i * d + c];
      q += part_sq[i * d + c];
      const double v = part_max[i * d + c];
      m = (v > m || v != v) ? v : m;
    }
    const double mean = s / n;
    double r = mean;
    if (mode == 1) r = m;
    else if (mode == 2) r = fmax(q / n - mean * mean, 0.0);
    out[u * ldo + c] = r;
  }
}

}  // namespace

void dvector_pool(Context& ctx, const void* frames, bool is_f32, int64_t n_frames, int64_t d, int64_t ld,
                  const int64_t* offsets_host, int64_t n_utts, int mode, bool l2norm, double* out_dev, int64_t ldo) {
  PB_CHECK(d > 0 && d <= kThreads * kMaxColsPerThread, kInvalidArg, "dvector_pool: need 0 < d <= 1024");
  PB_CHECK(mode >= 0 && mode <= 2, kInvalidArg, "dvector_pool: mode must be 0 (mean), 1 (max) or 2 (var)");
  PB_CHECK(n_utts >= 0 && offsets_host != nullptr && ldo >= d && ld >= d, kInvalidArg, "dvector_pool: bad arguments");
  if (n_utts == 0) return;
  PB_CHECK(offsets_host[0] >= 0 && offsets_host[n_utts] <= n_frames, kInvalidArg, "dvector_pool: offsets out of range");
  // work items: an utterance, or kChunk-frame pieces of a long one
  std::vector<long long> h_start, h_end, h_item0(n_utts + 1), h_frames(n_utts);
  for (int64_t u = 0; u < n_utts; ++u) {
    const int64_t a = offsets_host[u], b = offsets_host[u + 1];
    // np.max of an empty utterance raises, np.mean warns and returns nan: refuse both
    PB_CHECK(b > a, kValueError, "dvector_pool: every utterance needs at least one frame");
    h_item0[u] = static_cast<long long>(h_start.size());
    h_frames[u] = b - a;
    for (int64_t s = a; s < b; s += kChunk) {
      h_start.push_back(s);
      h_end.push_back(std::min<int64_t>(s + kChunk, b));
    }
  }
  h_item0[n_utts] = static_cast<long long>(h_start.size());
  const int64_t n_items = static_cast<int64_t>(h_start.size());
  DevBuf<long long> d_start(n_items), d_end(n_items), d_item0(n_utts + 1), d_frames(n_utts);
  DevBuf<double> part(static_cast<size_t>(3) * n_items * d);
  PB_CUDA(cudaMemcpyAsync(d_start.get(), h_start.data(), n_items * sizeof(long long), cudaMemcpyHostToDevice, ctx.stream));
  PB_CUDA(cudaMemcpyAsync(d_end.get(), h_end.data(), n_items * sizeof(long long), cudaMemcpyHostToDevice, ctx.stream));
  PB_CUDA(cudaMemcpyAsync(d_item0.get(), h_item0.data(), (n_utts + 1) * sizeof(long long), cudaMemcpyHostToDevice, ctx.stream));
  PB_CUDA(cudaMemcpyAsync(d_frames.get(), h_frames.data(), n_utts * sizeof(long long), cudaMemcpyHostToDevice, ctx.stream));
  double* ps = part.get();
  double* pq = ps + n_items * d;
  double* pm = pq + n_items * d;
  if (is_f32)
    dvector_pool_kernel<float><<<static_cast<unsigned>(n_items), kThreads, 0, ctx.stream>>>(
        static_cast<const float*>(frames), ld, static_cast<int>(d), d_start.get(), d_end.get(), l2norm ? 1 : 0, ps, pq, pm);
  else
    dvector_pool_kernel<double><<<static_cast<unsigned>(n_items), kThreads, 0, ctx.stream>>>(
        static_cast<const double*>(frames), ld, static_cast<int>(d), d_start.get(), d_end.get(), l2norm ? 1 : 0, ps, pq, pm);
  PB_CUDA(cudaGetLastError());
  dvector_finish_kernel<<<static_cast<unsigned>(n_utts), 256, 0, ctx.stream>>>(ps, pq, pm, d_item0.get(), d_frames.get(),
                                                                            static_cast<int>(d), mode, out_dev, ldo);
  PB_CUDA(cudaGetLastError());
  ctx.count_launch(2);
  ctx.sync();   // the host-side item tables and the partial buffers die with this frame
}

}  // namespace pb
