// PldaEngine: model state, transform, pair / grid scoring, z-norm.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <numeric>
#include <random>

#include "engine.h"

namespace pb {
namespace {

// Plda::LogLikelihoodRatio for one pair, in the operation order Kaldi uses; one warp.
__global__ void score_pair_kernel(const double* __restrict__ psi, const double* __restrict__ e,
                                  const double* __restrict__ t, int dim, double n, double* __restrict__ out) {
  const int lane = threadIdx.x;
  double logdet_g = 0.0, quad_g = 0.0, logdet_w = 0.0, quad_w = 0.0;
  for (int i = lane; i < dim; i += 32) {
    const double p = psi[i];
    const double mean = n * p / (n * p + 1.0) * e[i];
    const double var = 1.0 + p / (n * p + 1.0);
    const double diff = t[i] - mean;
    logdet_g += log(var);
    quad_g += diff * diff * (1.0 / var);
    const double var2 = p + 1.0;
    logdet_w += log(var2);
    quad_w += t[i] * t[i] * (1.0 / var2);
  }
  logdet_g = warp_sum(logdet_g);
  quad_g = warp_sum(quad_g);
  logdet_w = warp_sum(logdet_w);
  quad_w = warp_sum(quad_w);
  if (lane == 0) {
    const double kLog2Pi = 1.8378770664093454835606594728112;
    const double given = -0.5 * (logdet_g + kLog2Pi * dim + quad_g);
    const double without = -0.5 * (logdet_w + kLog2Pi * dim + quad_w);
    *out = given - without;
  }
}

// Plda::SmoothWithinClassCovariance: c = 1 + f psi; psi /= c; row i of A *= c^-1/2
__global__ void smooth_kernel(double* __restrict__ transform, double* __restrict__ psi, int d, double f) {
  const int i = blockIdx.x;
  const double c = 1.0 + f * psi[i];
  const double s = rsqrt(c);
  for (int j = threadIdx.x; j < d; j += blockDim.x) transform[static_cast<long long>(i) * d + j] *= s;
  __syncthreads();
  if (threadIdx.x == 0) psi[i] = psi[i] / c;
}

__global__ void gather_rows_kernel(const uint8_t* __restrict__ in, long long row_bytes_in, const int32_t* __restrict__ idx,
                                   long long rows, long long copy_bytes, uint8_t* __restrict__ out) {
  // one block per output row, 16-byte vector copies when aligned
  const long long r = blockIdx.x;
  const uint8_t* src = in + static_cast<long long>(idx[r]) * row_bytes_in;
  uint8_t* dst = out + r * copy_bytes;
  if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst) | copy_bytes) & 15) == 0) {
    const int4* s4 = reinterpret_cast<const int4*>(src);
    int4* d4 = reinterpret_cast<int4*>(dst);
    for (long long i = threadIdx.x; i < copy_bytes / 16; i += blockDim.x) d4[i] = s4[i];
  } else {
    for (long long i = threadIdx.x; i < copy_bytes; i += blockDim.x) dst[i] = src[i];
  }
}

inline size_t elem_size(int dtype) { return dtype == 1 ? 4 : 8; }

}  // namespace

// ------------------------------------------------------------------------- //
void PldaEngine::ensure_copy_stream() {
  if (copy_stream != nullptr) return;
  PB_CUDA(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    PB_CUDA(cudaEventCreateWithFlags(&ev_done[i], cudaEventDisableTiming));
    PB_CUDA(cudaEventCreateWithFlags(&ev_free[i], cudaEventDisableTiming));
  }
}

void PldaEngine::stage(const void* p, int64_t rows, int64_t cols, int64_t ld, int dtype, int loc, Staged& s,
                       DevBuf<uint8_t>* keep) {
  PB_CHECK(dtype == 0 || dtype == 1, kInvalidArg, "dtype must be PLDA_F64 or PLDA_F32");
  PB_CHECK(loc == 0 || loc == 1, kInvalidArg, "loc must be PLDA_HOST or PLDA_DEVICE");
  PB_CHECK(p != nullptr || rows == 0, kInvalidArg, "null matrix pointer");
  PB_CHECK(ld >= cols, kInvalidArg, "row pitch smaller than the number of columns");
  s.is_f32 = dtype == 1;
  if (loc == 1) {
    s.ptr = p;
    s.ld = ld;
    return;
  }
  const size_t es = elem_size(dtype);
  DevBuf<uint8_t>& buf = keep ? *keep : s.own;
  buf.reserve(static_cast<size_t>(rows > 0 ? rows : 1) * cols * es);
  if (rows > 0 && ld == cols)
    PB_CUDA(cudaMemcpyAsync(buf.get(), p, static_cast<size_t>(rows) * cols * es, cudaMemcpyHostToDevice, ctx.stream));
  else if (rows > 0)
    PB_CUDA(cudaMemcpy2DAsync(buf.get(), cols * es, p, ld * es, cols * es, rows, cudaMemcpyHostToDevice, ctx.stream));
  s.ptr = buf.get();
  s.ld = cols;
}

// SURVEY App. A.7:  a = n psi/(n psi+1), v = 1 + psi/(n psi+1), q = 1/2 (1/(1+psi) - 1/v)
void PldaEngine::fill_score_consts(const double* psi, int64_t dim, int count, double* out) {
  const double n = static_cast<double>(count);
  double logdet = 0.0;
  for (int64_t i = 0; i < dim; ++i) {
    const double p = psi[i];
    const double den = n * p + 1.0;
    const double a = n * p / den;
    const double v = 1.0 + p / den;
    out[kScoreConstsScale + i] = a / v;
    out[kScoreConstsEnrolSq + i] = a * a / v;
    out[kScoreConstsTestSq + i] = 0.5 * (1.0 / (1.0 + p) - 1.0 / v);
    logdet += log1p(p) - log(v);
  }
  out[kScoreConstsLogdet] = logdet;
}

const double* PldaEngine::score_consts_for(int count, int64_t dim) {
  for (auto& c : score_consts)
    if (c.count == count && c.dim == dim) return c.dev.get();
  PB_CHECK(dim > 0 && dim <= 1024 && dim <= static_cast<int64_t>(model.h_psi.size()), kInvalidArg,
           "score: dimension above 1024 is not supported");
  std::vector<double> h(kScoreConstsSize, 0.0);
  fill_score_consts(model.h_psi.data(), dim, count, h.data());
  if (score_consts.size() >= 16) score_consts.erase(score_consts.begin());
  score_consts.emplace_back();
  ScoreConsts& c = score_consts.back();
  c.count = count;
  c.dim = dim;
  c.dev.alloc(kScoreConstsSize);
  // pageable source: the runtime copies it to its staging buffer before the call returns (no sync needed)
  PB_CUDA(cudaMemcpyAsync(c.dev.get(), h.data(), kScoreConstsSize * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
  return c.dev.get();
}

void PldaEngine::refresh_model_operands() {
  score_consts.clear();
  ragged_key.clear();
  last_counts.clear();
  const int64_t d = model.d;
  split_rows(ctx, model.transform.get(), false, d, d, d, nullptr, nullptr, nullptr, model.a_split);
  model.h_psi.resize(d);
  PB_CUDA(cudaMemcpyAsync(model.h_psi.data(), model.psi.get(), d * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
  ctx.sync();
  model.ready = true;
}

void PldaEngine::set_model(int64_t d, const double* mean, const double* transform, const double* psi) {
  PB_CHECK(d > 0 && mean && transform && psi, kInvalidArg, "set_model: bad arguments");
  model.d = d;
  model.mean.reserve(d);
  model.transform.reserve(d * d);
  model.psi.reserve(d);
  PB_CUDA(cudaMemcpyAsync(model.mean.get(), mean, d * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
  PB_CUDA(cudaMemcpyAsync(model.transform.get(), transform, d * d * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
  PB_CUDA(cudaMemcpyAsync(model.psi.get(), psi, d * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
  refresh_model_operands();
}

void PldaEngine::get_model(double* mean, double* transform, double* psi) {
  require_model();
  const int64_t d = model.d;
  if (mean) PB_CUDA(cudaMemcpyAsync(mean, model.mean.get(), d * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
  if (transform)
    PB_CUDA(cudaMemcpyAsync(transform, model.transform.get(), d * d * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
  if (psi) PB_CUDA(cudaMemcpyAsync(psi, model.psi.get(), d * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
  ctx.sync();
}

void PldaEngine::get_covariances(double* within, double* between) {
  require_model();
  PB_CHECK(model.within.size() >= static_cast<size_t>(model.d * model.d), kNotFitted,
           "covariances are only available after fit()");
  const size_t bytes = model.d * model.d * sizeof(double);
  if (within) PB_CUDA(cudaMemcpyAsync(within, model.within.get(), bytes, cudaMemcpyDeviceToHost, ctx.stream));
  if (between) PB_CUDA(cudaMemcpyAsync(between, model.between.get(), bytes, cudaMemcpyDeviceToHost, ctx.stream));
  ctx.sync();
}

void PldaEngine::smooth(double factor) {
  require_model();
  PB_CHECK(factor >= 0.0 && factor <= 1.0, kInvalidArg, "smoothing factor must be in [0,1] (Kaldi asserts this)");
  smooth_kernel<<<static_cast<unsigned>(model.d), 128, 0, ctx.stream>>>(model.transform.get(), model.psi.get(),
                                                                        static_cast<int>(model.d), factor);
  PB_CUDA(cudaGetLastError());
  ctx.count_launch();
  refresh_model_operands();
}

// ------------------------------------------------------------------------- //
// transform
// ------------------------------------------------------------------------- //
void PldaEngine::transform_device_rows(const void* x, bool is_f32, int64_t n, int64_t d, int64_t ld,
                                       const double* sub, const int32_t* counts_dev, int32_t const_count, int64_t dim,
                                       double* out64, int64_t ld64, float* out32, int64_t ld32) {
  if (n == 0) return;
  if (precision == 1) {
    // exact mode: centre in fp64, fp64 GEMM  y = (x - mu) A[:dim]^T
    ws_f64a.reserve(n * d);
    convert_to_f64(ctx, x, is_f32, n, d, ld, ws_f64a.get(), d, sub);
    ws_f64b.reserve(n * dim);
    gemm_f64(ctx, false, true, n, dim, d, 1.0, ws_f64a.get(), d, model.transform.get(), d, 0.0, ws_f64b.get(), dim);
    length_normalise(ctx, ws_f64b.get(), false, n, dim, dim, model.psi.get(), counts_dev, const_count, out64, ld64,
                     out32, ld32);
    return;
  }
  // tensor path: centre in fp64 while splitting, one NT GEMM against the first `dim` rows of A
  split_rows(ctx, x, is_f32, n, d, ld, sub, nullptr, nullptr, ws_x);
  SplitOperand a = model.a_split.view();
  a.rows = dim;
  const int64_t ldy = round_up(dim, 4);
  ws_y.reserve(n * ldy);
  GemmEpilogue epi;
  epi.out = ws_y.get();
  epi.ldo = ldy;
  gemm_bf16x3(ctx, ws_x.view(), a, n, dim, d, epi);
  length_normalise(ctx, ws_y.get(), true, n, dim, ldy, model.psi.get(), counts_dev, const_count, out64, ld64, out32,
                   ld32);
}

void PldaEngine::transform_rows(const void* x, int64_t n, int64_t d, int64_t ldx, int dtype, int loc,
                                const int32_t* counts, int32_t const_count, int64_t targetdim, void* out, int64_t ldo,
                                int out_dtype, int out_loc) {
  require_model();
  PB_CHECK(d == model.d, kInvalidArg, "transform: feature dimension does not match the model");
  PB_CHECK(targetdim >= 0 && targetdim <= d, kInvalidArg, "transform: targetdim out of range");
  const int64_t dim = targetdim == 0 ? d : targetdim;
  PB_CHECK(ldo >= dim, kInvalidArg, "transform: output pitch too small");
  PB_CHECK(counts != nullptr || const_count > 0, kInvalidArg, "transform: counts must be positive");
  if (n == 0) return;
  Staged sx;
  stage(x, n, d, ldx, dtype, loc, sx);
  const int32_t* counts_dev = nullptr;
  if (counts) {
    for (int64_t i = 0; i < n; ++i) PB_CHECK(counts[i] > 0, kInvalidArg, "transform: counts must be positive");
    ws_counts.reserve(n);
    PB_CUDA(cudaMemcpyAsync(ws_counts.get(), counts, n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx.stream));
    counts_dev = ws_counts.get();
  }
  const bool out_f32 = out_dtype == 1;
  if (out_loc == 1) {
    transform_device_rows(sx.ptr, sx.is_f32, n, d, sx.ld, model.mean.get(), counts_dev, const_count, dim,
                          out_f32 ? nullptr : static_cast<double*>(out), ldo,
                          out_f32 ? static_cast<float*>(out) : nullptr, ldo);
    ctx.sync();
    return;
  }
  const size_t es = out_f32 ? 4 : 8;
  DevBuf<uint8_t> tmp(static_cast<size_t>(n) * dim * es);
  transform_device_rows(sx.ptr, sx.is_f32, n, d, sx.ld, model.mean.get(), counts_dev, const_count, dim,
                        out_f32 ? nullptr : reinterpret_cast<double*>(tmp.get()), dim,
                        out_f32 ? reinterpret_cast<float*>(tmp.get()) : nullptr, dim);
  PB_CUDA(cudaMemcpy2DAsync(out, ldo * es, tmp.get(), dim * es, dim * es, n, cudaMemcpyDeviceToHost, ctx.stream));
  ctx.sync();
}

void PldaEngine::transform_grouped(const void* x, int64_t n, int64_t d, int64_t ldx, int dtype, int loc,
                                   const uint64_t* labels, int64_t targetdim, uint64_t* out_labels,
                                   int64_t* out_counts, double* out_vecs, int64_t* n_out) {
  require_model();
  PB_CHECK(d == model.d, kInvalidArg, "transform: feature dimension does not match the model");
  PB_CHECK(targetdim >= 0 && targetdim <= d, kInvalidArg, "transform: targetdim out of range");
  PB_CHECK(labels != nullptr || n == 0, kInvalidArg, "transform: labels are required");
  const int64_t dim = targetdim == 0 ? d : targetdim;
  *n_out = 0;
  if (n == 0) return;
  Staged sx;
  stage(x, n, d, ldx, dtype, loc, sx);
  DevBuf<uint64_t> lab(n);
  PB_CUDA(cudaMemcpyAsync(lab.get(), labels, n * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx.stream));
  build_segments(ctx, lab.get(), n, segs);                            // K1
  const int64_t r = segs.nseg;
  ws_f64a.reserve(r * d);
  segment_sums(ctx, sx.ptr, sx.is_f32, d, sx.ld, segs, ws_f64a.get());   // K4
  ws_counts.reserve(r);
  segment_finalize_means(ctx, ws_f64a.get(), d, segs, ws_counts.get());
  DevBuf<double> outv(static_cast<size_t>(r) * dim);
  transform_device_rows(ws_f64a.get(), false, r, d, d, model.mean.get(), ws_counts.get(), 0, dim, outv.get(), dim,
                        nullptr, 0);                                  // K5
  std::vector<int32_t> hc(r);
  PB_CUDA(cudaMemcpyAsync(hc.data(), ws_counts.get(), r * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx.stream));
  PB_CUDA(cudaMemcpyAsync(out_labels, segs.seg_label.get(), r * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx.stream));
  PB_CUDA(cudaMemcpyAsync(out_vecs, outv.get(), r * dim * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
  ctx.sync();
  for (int64_t i = 0; i < r; ++i) out_counts[i] = hc[i];
  *n_out = r;
}

// ------------------------------------------------------------------------- //
// scoring
// ------------------------------------------------------------------------- //
void PldaEngine::score_pair(uint64_t id, int64_t n_enrol, const double* enrol, const double* test, int64_t dim,
                            float* out) {
  require_model();
  PB_CHECK(dim > 0 && dim <= model.d, kInvalidArg, "score: vector dimension does not match the model");
  PB_CHECK(n_enrol > 0, kInvalidArg, "score: enrol count must be positive");
  ws_f64a.reserve(2 * dim + 1);
  double* e = ws_f64a.get();
  double* t = e + dim;
  double* res = t + dim;
  PB_CUDA(cudaMemcpyAsync(e, enrol, dim * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
  PB_CUDA(cudaMemcpyAsync(t, test, dim * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
  score_pair_kernel<<<1, 32, 0, ctx.stream>>>(model.psi.get(), e, t, static_cast<int>(dim),
                                             static_cast<double>(n_enrol), res);
  PB_CUDA(cudaGetLastError());
  ctx.count_launch();
  double s = 0.0;
  PB_CUDA(cudaMemcpyAsync(&s, res, sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
  ctx.sync();
  auto it = znorm.find(id);
  if (!znorm.empty() && it != znorm.end()) s = (s - it->second.first) / it->second.second;   // pldamodule.cpp:269-273
  *out = static_cast<float>(s);                                                                 // "f" :276
}

// Enrol counts -> groups of equal count.  Uniform counts (the common case) use the single cached table of that
// count; ragged counts get one table per distinct count, built on the host from the psi mirror and kept on the
// device until the set of counts (or the model) changes.
PldaEngine::ScoreGroups PldaEngine::prepare_groups(const int32_t* counts, int64_t ne, int64_t dim) {
  PB_CHECK(counts != nullptr, kInvalidArg, "score: enrol counts are required");
  PB_CHECK(dim > 0 && dim <= 1024, kInvalidArg, "score: dimension above 1024 is not supported");
  ScoreGroups g;
  g.uniform = true;
  for (int64_t i = 0; i < ne; ++i) {
    PB_CHECK(counts[i] > 0, kInvalidArg, "score: enrol counts must be positive");
    g.uniform = g.uniform && counts[i] == counts[0];
  }
  g.uniform_count = counts[0];
  if (g.uniform) {
    g.tables = score_consts_for(g.uniform_count, dim);
    return g;
  }
  // the device copies (counts, group index per row) are reused while the caller keeps passing the same counts
  const bool same_counts = static_cast<int64_t>(last_counts.size()) == ne &&
                           memcmp(last_counts.data(), counts, ne * sizeof(int32_t)) == 0 && !ragged_key.empty();
  std::vector<int32_t> gcounts;
  if (same_counts) {
    gcounts = ragged_key;
  } else {
    // distinct counts in ascending order + group index per row: a direct table when the counts are small (they
    // are utterance counts), a sort otherwise
    int32_t mx = 0;
    for (int64_t i = 0; i < ne; ++i) mx = std::max(mx, counts[i]);
    std::vector<int32_t> grp(ne);
    if (mx <= (1 << 20)) {
      std::vector<int32_t> tab(static_cast<size_t>(mx) + 1, -1);
      for (int64_t i = 0; i < ne; ++i) tab[counts[i]] = 0;
      for (int32_t c = 1; c <= mx; ++c)
        if (tab[c] == 0) { tab[c] = static_cast<int32_t>(gcounts.size()); gcounts.push_back(c); }
      for (int64_t i = 0; i < ne; ++i) grp[i] = tab[counts[i]];
    } else {
      gcounts.assign(counts, counts + ne);
      std::sort(gcounts.begin(), gcounts.end());
      gcounts.erase(std::unique(gcounts.begin(), gcounts.end()), gcounts.end());
      for (int64_t i = 0; i < ne; ++i)
        grp[i] = static_cast<int32_t>(std::lower_bound(gcounts.begin(), gcounts.end(), counts[i]) - gcounts.begin());
    }
    rg_counts.reserve(ne);
    rg_grp.reserve(ne);
    rg_gcounts.reserve(gcounts.size());
    PB_CUDA(cudaMemcpyAsync(rg_counts.get(), counts, ne * sizeof(int32_t), cudaMemcpyHostToDevice, ctx.stream));
    PB_CUDA(cudaMemcpyAsync(rg_grp.get(), grp.data(), ne * sizeof(int32_t), cudaMemcpyHostToDevice, ctx.stream));
    PB_CUDA(cudaMemcpyAsync(rg_gcounts.get(), gcounts.data(), gcounts.size() * sizeof(int32_t), cudaMemcpyHostToDevice,
                            ctx.stream));
    last_counts.assign(counts, counts + ne);
    if (ragged_key != gcounts) ragged_key.clear();       // forces the tables to be rebuilt below
  }
  g.ng = static_cast<int>(gcounts.size());
  g.counts_dev = rg_counts.get();
  g.grp_dev = rg_grp.get();
  g.gcounts_dev = rg_gcounts.get();
  if (ragged_key != gcounts || ragged_dim != dim) {
    std::vector<double> tabs(static_cast<size_t>(g.ng) * kScoreConstsSize, 0.0);
    for (int i = 0; i < g.ng; ++i)
      fill_score_consts(model.h_psi.data(), dim, gcounts[i], tabs.data() + static_cast<size_t>(i) * kScoreConstsSize);
    ws_tables.reserve(tabs.size());
    PB_CUDA(cudaMemcpyAsync(ws_tables.get(), tabs.data(), tabs.size() * sizeof(double), cudaMemcpyHostToDevice,
                            ctx.stream));
    ragged_key = gcounts;
    ragged_dim = dim;
  }
  g.tables = ws_tables.get();
  return g;
}

namespace {
__global__ void znorm_affine_kernel(const double* __restrict__ mean, const double* __restrict__ stdv, long long n,
                                    float* __restrict__ zmean, float* __restrict__ zinv) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  zmean[i] = static_cast<float>(mean[i]);
  zinv[i] = static_cast<float>(1.0 / stdv[i]);
}
}  // namespace

// Per-enrol-row z-norm affine (s - mean) / std as fp32 device vectors: from caller arrays (host or device fp64) when
// given, else from the id table filled by norm() (rows whose id is unknown stay unnormalised, pldamodule.cpp:269-273).
void PldaEngine::znorm_affine(const uint64_t* ids, int64_t ne, const double* zmean_in, const double* zstd_in, int z_loc,
                              const float** zmean, const float** zinv) {
  *zmean = nullptr;
  *zinv = nullptr;
  if (zmean_in != nullptr && zstd_in != nullptr) {
    ws_zmean.reserve(2 * ne);
    if (z_loc == 1) {
      znorm_affine_kernel<<<static_cast<unsigned>(ceil_div(ne, 256)), 256, 0, ctx.stream>>>(
          zmean_in, zstd_in, ne, ws_zmean.get(), ws_zmean.get() + ne);
      PB_CUDA(cudaGetLastError());
      ctx.count_launch();
    } else {
      std::vector<float> hz(2 * ne);
      for (int64_t i = 0; i < ne; ++i) {
        hz[i] = static_cast<float>(zmean_in[i]);
        hz[ne + i] = static_cast<float>(1.0 / zstd_in[i]);
      }
      PB_CUDA(cudaMemcpyAsync(ws_zmean.get(), hz.data(), 2 * ne * sizeof(float), cudaMemcpyHostToDevice, ctx.stream));
    }
    *zmean = ws_zmean.get();
    *zinv = ws_zmean.get() + ne;
    return;
  }
  if (ids == nullptr || znorm.empty()) return;
  std::vector<float> hz(2 * ne);
  for (int64_t i = 0; i < ne; ++i) {
    auto it = znorm.find(ids[i]);
    hz[i] = it == znorm.end() ? 0.f : static_cast<float>(it->second.first);
    hz[ne + i] = it == znorm.end() ? 1.f : static_cast<float>(1.0 / it->second.second);
  }
  ws_zmean.reserve(2 * ne);
  PB_CUDA(cudaMemcpyAsync(ws_zmean.get(), hz.data(), 2 * ne * sizeof(float), cudaMemcpyHostToDevice, ctx.stream));
  *zmean = ws_zmean.get();
  *zinv = ws_zmean.get() + ne;
}

// Tensor-path operands of one grid: ws_l / ws_row (enrol side), ws_r / ws_col (test side, one column-term row per
// distinct enrol count).
void PldaEngine::produce_score_operands(const Staged& se, int64_t ne, const Staged& st, int64_t nt, int64_t dim,
                                        const ScoreGroups& g, int64_t col_ld) {
  ws_row.reserve(ne);
  ws_col.reserve(static_cast<size_t>(g.ng) * col_ld);
  score_k = dim;
  score_cols_embedded = false;
  if (g.uniform) {
    score_prep_uniform(ctx, se.ptr, ne, se.ld, st.ptr, nt, st.ld, se.is_f32, dim, g.tables, ws_l, ws_r, ws_row.get(),
                       ws_col.get(), col_ld);
    return;
  }
  static const char* mode = getenv("PLDA_B200_RAGGED");
  // producers below write every column term they own; the [nt, col_ld) padding is only read for columns that are
  // never stored
  if (mode != nullptr)
    PB_CUDA(cudaMemsetAsync(ws_col.get(), 0, static_cast<size_t>(g.ng) * col_ld * sizeof(float), ctx.stream));
  if (mode != nullptr && strcmp(mode, "old") == 0) {            // A/B switch: per-element log / divide producers
    score_prep_enrol(ctx, se.ptr, se.is_f32, ne, dim, se.ld, g.counts_dev, g.uniform_count, model.psi.get(), &ws_l,
                     nullptr, ws_row.get(), nullptr);
    score_prep_test(ctx, st.ptr, st.is_f32, nt, dim, st.ld, g.gcounts_dev, g.ng, g.uniform_count, model.psi.get(), &ws_r,
                    ws_col.get(), col_ld, nullptr);
  } else if (mode != nullptr && strcmp(mode, "scalar") == 0) {  // A/B switch: table-driven, 2-byte stores
    score_prep_grouped(ctx, se.ptr, ne, se.ld, g.grp_dev, st.ptr, nt, st.ld, se.is_f32, dim, g.ng, g.tables, ws_l, ws_r,
                       ws_row.get(), ws_col.get(), col_ld);
  } else {
    // up to 8 distinct counts: their column terms travel inside the operands (at most one more 16-wide k-step) and
    // the grid runs the uniform-count kernel; more: per-row group vectors in the epilogue
    const bool embed = g.ng <= 8 && !(mode != nullptr && strcmp(mode, "epilogue") == 0);
    score_prep_grouped_vec(ctx, se.ptr, ne, se.ld, g.grp_dev, st.ptr, nt, st.ld, se.is_f32, dim, g.ng, g.tables, ws_l,
                           ws_r, ws_row.get(), ws_col.get(), col_ld, embed);
    if (embed) {
      score_k = dim + 2 * g.ng;
      score_cols_embedded = true;
    }
  }
}

void PldaEngine::bind_col_terms(GemmEpilogue& epi, const ScoreGroups& g, int64_t r0, int64_t col_ld) const {
  if (score_cols_embedded) return;
  epi.col_add = ws_col.get();
  epi.col_ld = col_ld;
  epi.grp = g.grp_dev ? g.grp_dev + r0 : nullptr;
}

void PldaEngine::score_grid(const void* enrol, int64_t ne, int64_t ld_enrol, const int32_t* counts,
                            const uint64_t* ids, const void* test, int64_t nt, int64_t ld_test, int64_t dim,
                            int dtype, int loc, float* out, int64_t ldo, int out_loc, const double* zmean_in,
                            const double* zstd_in, int z_loc) {
  require_model();
  PB_CHECK(dim > 0 && dim <= model.d, kInvalidArg, "score_grid: vector dimension does not match the model");
  PB_CHECK(ne >= 0 && nt >= 0, kInvalidArg, "score_grid: negative size");
  PB_CHECK(out != nullptr || ne * nt == 0, kInvalidArg, "score_grid: null output");
  PB_CHECK(ldo >= nt, kInvalidArg, "score_grid: output pitch too small");
  if (ne == 0 || nt == 0) return;
  const ScoreGroups g = prepare_groups(counts, ne, dim);

  const bool trace = getenv("PLDA_B200_TRACE") != nullptr;
  const auto t_begin = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!trace) return;
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
    fprintf(stderr, "plda_b200 score_grid: %-28s +%.3f ms\n", what, ms);
  };
  Staged se, st;
  stage(enrol, ne, dim, ld_enrol, dtype, loc, se, &ws_stage[0]);
  stage(test, nt, dim, ld_test, dtype, loc, st, &ws_stage[1]);
  lap("inputs staged (enqueued)");

  // optional z-norm affine per enrol row
  const float* zmean = nullptr;
  const float* zinv = nullptr;
  znorm_affine(ids, ne, zmean_in, zstd_in, z_loc, &zmean, &zinv);

  const int64_t col_ld = round_up(nt, 32);
  const bool exact = precision == 1;
  const int64_t ldo_dev = out_loc == 1 ? ldo : round_up(nt, 4);

  if (exact) {
    ws_f64a.reserve(ne * dim);                       // L (fp64)
    ws_row64.reserve(ne);
    ws_col64.reserve(static_cast<size_t>(g.ng) * col_ld);
    score_prep_enrol(ctx, se.ptr, se.is_f32, ne, dim, se.ld, g.counts_dev, g.uniform_count, model.psi.get(), nullptr,
                     ws_f64a.get(), nullptr, ws_row64.get());
    score_prep_test(ctx, st.ptr, st.is_f32, nt, dim, st.ld, g.gcounts_dev, g.ng, g.uniform_count, model.psi.get(), nullptr,
                    nullptr, col_ld, ws_col64.get());
    ws_f64b.reserve(nt * dim);
    convert_to_f64(ctx, st.ptr, st.is_f32, nt, dim, st.ld, ws_f64b.get(), dim);
  } else {
    produce_score_operands(se, ne, st, nt, dim, g, col_ld);
  }

  // enrol-row chunks: bounded staging when the result goes back to the host, one chunk otherwise
  int64_t chunk = ne;
  if (out_loc == 0 || exact) {
    const int64_t budget = exact ? (512ll << 20) / 8 : (1ll << 30) / 4;   // elements per staging buffer
    chunk = std::max<int64_t>(128, (budget / std::max<int64_t>(ldo_dev, 1)) / 128 * 128);
    chunk = std::min(chunk, ne);
  }
  if (out_loc == 0) {
    ensure_copy_stream();
    const int nbuf = chunk < ne ? 2 : 1;   // the second staging buffer is only needed when the grid is chunked
    for (int i = 0; i < nbuf; ++i) ws_out[i].reserve(static_cast<size_t>(chunk) * ldo_dev);
    // the events may still carry a record of an earlier call: this call's chain starts clean
    for (int i = 0; i < 2; ++i) PB_CUDA(cudaEventRecord(ev_free[i], copy_stream));
  }
  lap("workspaces ready");
  auto launch_chunk = [&](int64_t r0, int b) {
    const int64_t rows = std::min(chunk, ne - r0);
    float* dst = out_loc == 1 ? out + r0 * ldo : ws_out[b].get();
    if (out_loc == 0) PB_CUDA(cudaStreamWaitEvent(ctx.stream, ev_free[b], 0));
    if (exact) {
      ws_gram.reserve(static_cast<size_t>(rows) * nt);
      gemm_f64(ctx, false, true, rows, nt, dim, 1.0, ws_f64a.get() + r0 * dim, dim, ws_f64b.get(), dim, 0.0,
               ws_gram.get(), nt);
      score_epilogue_f64(ctx, ws_gram.get(), rows, nt, ws_row64.get() + r0, ws_col64.get(), col_ld,
                         g.grp_dev ? g.grp_dev + r0 : nullptr, zmean ? zmean + r0 : nullptr, zinv ? zinv + r0 : nullptr, dst,
                         ldo_dev, nullptr, nullptr);
    } else {
      SplitOperand a = ws_l.view();
      a.hi += r0 * a.ld;
      a.lo += r0 * a.ld;
      a.rows = rows;
      GemmEpilogue epi;
      epi.out = dst;
      epi.ldo = ldo_dev;
      epi.row_add = ws_row.get() + r0;
      bind_col_terms(epi, g, r0, col_ld);
      epi.zmean = zmean ? zmean + r0 : nullptr;
      epi.zinv = zinv ? zinv + r0 : nullptr;
      gemm_bf16x3(ctx, a, ws_r.view(), rows, nt, score_k, epi);
    }
    if (out_loc == 0) PB_CUDA(cudaEventRecord(ev_done[b], ctx.stream));
  };
  auto copy_chunk = [&](int64_t r0, int b) {
    const int64_t rows = std::min(chunk, ne - r0);
    PB_CUDA(cudaStreamWaitEvent(copy_stream, ev_done[b], 0));
    if (ldo == nt && ldo_dev == nt)   // contiguous on both sides: one linear copy (full-rate DMA)
      PB_CUDA(cudaMemcpyAsync(out + r0 * ldo, ws_out[b].get(), static_cast<size_t>(rows) * nt * sizeof(float),
                              cudaMemcpyDeviceToHost, copy_stream));
    else
      PB_CUDA(cudaMemcpy2DAsync(out + r0 * ldo, ldo * sizeof(float), ws_out[b].get(), ldo_dev * sizeof(float),
                                nt * sizeof(float), rows, cudaMemcpyDeviceToHost, copy_stream));
    PB_CUDA(cudaEventRecord(ev_free[b], copy_stream));
  };
  if (out_loc == 1) {
    for (int64_t r0 = 0; r0 < ne; r0 += chunk) launch_chunk(r0, 0);
    // on a caller-provided stream the result is stream-ordered with the caller's work: no host sync
    if (ctx.owns_stream) ctx.sync();
  } else {
    // software pipeline: the GEMM of chunk i+1 is in flight while chunk i drains over PCIe
    int b = 0;
    launch_chunk(0, 0);
    for (int64_t r0 = 0; r0 < ne; r0 += chunk) {
      if (r0 + chunk < ne) launch_chunk(r0 + chunk, b ^ 1);
      copy_chunk(r0, b);
      b ^= 1;
    }
    lap("all work enqueued");
    if (trace) { ctx.sync(); lap("compute stream drained"); }
    PB_CUDA(cudaStreamSynchronize(copy_stream));
    ctx.sync();
    lap("device->host drained");
  }
}

// ------------------------------------------------------------------------- //
// z-norm statistics
// ------------------------------------------------------------------------- //
// Rows of the background set a z-norm pass uses: all of them (numutts == 0 or m), else the first `numutts` entries
// of a Fisher-Yates shuffle of 0..m-1 driven by splitmix64(seed) -- a DEFINED sequence (the reference's
// std::random_shuffle is unseeded, src/pldamodule.cpp:204-213), so a caller or a test can reproduce the subset.
void norm_selection(int64_t m, int64_t numutts, uint64_t seed, int32_t* out) {
  std::vector<int32_t> idx(m);
  std::iota(idx.begin(), idx.end(), 0);
  uint64_t state = seed;
  auto next = [&state]() {
    uint64_t z = (state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  };
  for (int64_t i = m - 1; i > 0; --i) {
    const int64_t j = static_cast<int64_t>(next() % static_cast<uint64_t>(i + 1));
    std::swap(idx[i], idx[j]);
  }
  std::copy(idx.begin(), idx.begin() + numutts, out);
}

void PldaEngine::norm(const void* bkg, int64_t m, int64_t d, int64_t ldb, int dtype, int loc,
                      const uint64_t* enrol_ids, const void* enrol, int64_t ne, int64_t ld_enrol, int64_t dim,
                      int enrol_dtype, int enrol_loc, int64_t numutts, uint64_t seed, double* mean_out,
                      double* std_out, int out_loc) {
  require_model();
  PB_CHECK(d == model.d, kInvalidArg, "norm: background dimension does not match the model");
  PB_CHECK(dim > 0 && dim <= model.d, kInvalidArg, "norm: enrol dimension does not match the model");
  PB_CHECK(m > 0 && m < (1ll << 31), kInvalidArg, "norm: no background vectors");
  PB_CHECK(numutts >= 0 && numutts <= m, kInvalidArg, "norm: numutts out of range");
  PB_CHECK(enrol_ids != nullptr || (mean_out != nullptr && std_out != nullptr) || ne == 0, kInvalidArg,
           "norm: enrol ids or output arrays are required");
  if (ne == 0) return;
  if (numutts == 0) numutts = m;

  Staged sb, se;
  stage(bkg, m, d, ldb, dtype, loc, sb);
  const void* rows_ptr = sb.ptr;
  int64_t rows_ld = sb.ld;
  if (numutts < m) {
    std::vector<int32_t> idx(numutts);
    norm_selection(m, numutts, seed, idx.data());
    ws_te.reserve(numutts);
    PB_CUDA(cudaMemcpyAsync(ws_te.get(), idx.data(), numutts * sizeof(int32_t), cudaMemcpyHostToDevice, ctx.stream));
    const size_t es = sb.is_f32 ? 4 : 8;
    ws_gather.reserve(static_cast<size_t>(numutts) * d * es);
    gather_rows_kernel<<<static_cast<unsigned>(numutts), 128, 0, ctx.stream>>>(
        static_cast<const uint8_t*>(sb.ptr), sb.ld * es, ws_te.get(), numutts, d * es, ws_gather.get());
    PB_CUDA(cudaGetLastError());
    ctx.count_launch();
    rows_ptr = ws_gather.get();
    rows_ld = d;
  }
  // background rows transformed with num_examples = m  (src/pldamodule.cpp:224)
  ws_bt.reserve(static_cast<size_t>(numutts) * dim);
  transform_device_rows(rows_ptr, sb.is_f32, numutts, d, rows_ld, model.mean.get(), nullptr,
                        static_cast<int32_t>(std::min<int64_t>(m, INT32_MAX)), dim, ws_bt.get(), dim, nullptr, 0);
  stage(enrol, ne, dim, ld_enrol, enrol_dtype, enrol_loc, se);

  // S[e, b] = LLR(train = bkg_b, n = 1, test = enrol_e): symmetric in (e,b) for n = 1, so enrol rows are the
  // M side (row reduction over the cohort happens in the GEMM epilogue; the grid is never materialised).
  ws_zstat.reserve(2 * ne);
  double* dmean = ws_zstat.get();
  double* dstd = ws_zstat.get() + ne;
  const int64_t col_ld = round_up(numutts, 32);
  if (precision == 1) {
    ws_rsum.reserve(ne);
    ws_rsq.reserve(ne);
    PB_CUDA(cudaMemsetAsync(ws_rsum.get(), 0, ne * sizeof(double), ctx.stream));
    PB_CUDA(cudaMemsetAsync(ws_rsq.get(), 0, ne * sizeof(double), ctx.stream));
    ws_f64a.reserve(ne * dim);
    ws_row64.reserve(ne);
    ws_col64.reserve(col_ld);
    score_prep_enrol(ctx, se.ptr, se.is_f32, ne, dim, se.ld, nullptr, 1, model.psi.get(), nullptr, ws_f64a.get(),
                     nullptr, ws_row64.get());
    score_prep_test(ctx, ws_bt.get(), false, numutts, dim, dim, nullptr, 1, 1, model.psi.get(), nullptr, nullptr,
                    col_ld, ws_col64.get());
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(ne, (512ll << 20) / 8 / numutts));
    ws_gram.reserve(static_cast<size_t>(chunk) * numutts);
    for (int64_t r0 = 0; r0 < ne; r0 += chunk) {
      const int64_t rows = std::min(chunk, ne - r0);
      gemm_f64(ctx, false, true, rows, numutts, dim, 1.0, ws_f64a.get() + r0 * dim, dim, ws_bt.get(), dim, 0.0,
               ws_gram.get(), numutts);
      score_epilogue_f64(ctx, ws_gram.get(), rows, numutts, ws_row64.get() + r0, ws_col64.get(), col_ld, nullptr,
                         nullptr, nullptr, nullptr, 0, ws_rsum.get() + r0, ws_rsq.get() + r0);
    }
    znorm_finalize(ctx, ws_rsum.get(), ws_rsq.get(), ne, numutts, nullptr, nullptr, dmean, dstd);
  } else {
    // one enrol count (n = 1): the table-driven producer (no per-element log / divide); the two sides may differ in
    // dtype (caller's enrol rows vs the fp64 transformed cohort), hence one launch per side
    ws_row.reserve(ne);
    ws_col.reserve(col_ld);
    const double* consts = score_consts_for(1, dim);
    PrepDst none, cohort;
    score_prep_uniform_multi(ctx, se.ptr, ne, se.ld, &ws_l, ws_row.get(), nullptr, 0, 0, 0, 0, none, round_up(dim, 16),
                             se.is_f32, dim, consts, PrepSignal{});
    ws_r.reserve(numutts, dim);
    cohort.n = 1;
    cohort.hi[0] = ws_r.hi.get();
    cohort.lo[0] = ws_r.lo.get();
    cohort.term[0] = ws_col.get();
    score_prep_uniform_multi(ctx, nullptr, 0, 0, nullptr, nullptr, ws_bt.get(), numutts, dim, 0, col_ld, cohort, ws_r.ld,
                             false, dim, consts, PrepSignal{});
    ctx.pdl_pending = false;     // the GEMM below does not directly follow a producer it may overlap
    // moments sink: per-(row, column tile, epilogue half) shifted fp32 partials, merged in fp64 -- no atomics, and
    // no cancellation when |mean| >> std
    const int n_tiles = gemm_n_tiles(numutts);
    ws_mom.reserve(static_cast<size_t>(ne) * n_tiles * 2);
    GemmEpilogue epi;
    epi.row_add = ws_row.get();
    epi.col_add = ws_col.get();
    epi.col_ld = col_ld;
    epi.mom = ws_mom.get();
    gemm_bf16x3(ctx, ws_l.view(), ws_r.view(), ne, numutts, dim, epi);
    moments_reduce(ctx, ws_mom.get(), ne, n_tiles, nullptr, nullptr, dmean, dstd);
  }
  if (mean_out != nullptr && std_out != nullptr && out_loc == 1) {
    PB_CUDA(cudaMemcpyAsync(mean_out, dmean, ne * sizeof(double), cudaMemcpyDeviceToDevice, ctx.stream));
    PB_CUDA(cudaMemcpyAsync(std_out, dstd, ne * sizeof(double), cudaMemcpyDeviceToDevice, ctx.stream));
  }
  const bool to_host = enrol_ids != nullptr || (mean_out != nullptr && out_loc == 0);
  if (!to_host) {
    if (ctx.owns_stream) ctx.sync();
    return;
  }
  std::vector<double> hm(2 * ne);
  PB_CUDA(cudaMemcpyAsync(hm.data(), ws_zstat.get(), 2 * ne * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
  ctx.sync();
  if (mean_out != nullptr && std_out != nullptr && out_loc == 0) {
    std::copy(hm.begin(), hm.begin() + ne, mean_out);
    std::copy(hm.begin() + ne, hm.end(), std_out);
  }
  if (enrol_ids != nullptr) {
    znorm.reserve(znorm.size() + ne);
    for (int64_t i = 0; i < ne; ++i) znorm.emplace(enrol_ids[i], std::make_pair(hm[i], hm[ne + i]));   // insert: first wins
  }
}

// ------------------------------------------------------------------------- //
// kernel-level test hooks
// ------------------------------------------------------------------------- //
void PldaEngine::test_gemm(const double* a, const double* b, int64_t m, int64_t n, int64_t k, int ksplit, float* out) {
  DevBuf<double> da(m * k), db(n * k);
  PB_CUDA(cudaMemcpyAsync(da.get(), a, m * k * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
  PB_CUDA(cudaMemcpyAsync(db.get(), b, n * k * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
  SplitBuf sa, sb;
  split_rows(ctx, da.get(), false, m, k, k, nullptr, nullptr, nullptr, sa);
  split_rows(ctx, db.get(), false, n, k, k, nullptr, nullptr, nullptr, sb);
  const int64_t ldo = round_up(n, 4);
  if (ksplit <= 1) {
    DevBuf<float> dout(m * ldo);
    PB_CUDA(cudaMemsetAsync(dout.get(), 0xff, m * ldo * sizeof(float), ctx.stream));   // NaN canary
    GemmEpilogue epi;
    epi.out = dout.get();
    epi.ldo = ldo;
    gemm_bf16x3(ctx, sa.view(), sb.view(), m, n, k, epi);
    PB_CUDA(cudaMemcpy2DAsync(out, n * sizeof(float), dout.get(), ldo * sizeof(float), n * sizeof(float), m,
                              cudaMemcpyDeviceToHost, ctx.stream));
    ctx.sync();
    return;
  }
  const int eff = effective_ksplit(ctx, m, n, k, ksplit);
  const int64_t mpad = round_up(m, 128);
  DevBuf<float> part(static_cast<size_t>(eff) * mpad * ldo);
  PB_CUDA(cudaMemsetAsync(part.get(), 0, part.size() * sizeof(float), ctx.stream));
  gemm_bf16x3_splitk(ctx, sa.view(), sb.view(), m, n, k, ksplit, part.get());
  DevBuf<double> d64(m * n);
  reduce_partials_f64(ctx, part.get(), eff, m, n, d64.get(), n, 1.0, false);
  std::vector<double> h(m * n);
  PB_CUDA(cudaMemcpyAsync(h.data(), d64.get(), m * n * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
  ctx.sync();
  for (int64_t i = 0; i < m * n; ++i) out[i] = static_cast<float>(h[i]);
}

}  // namespace pb
