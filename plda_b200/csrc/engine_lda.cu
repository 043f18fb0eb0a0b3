// LdaEngine: the reference's pure-Python LDA (python/liblda/lda.py) on the device.
//   fit_svd  <- LDA.fit + _solve_svd (lda.py:106-138, 178-221).  The two SVDs are replaced by
//              symmetric eigendecompositions of the corresponding Gram matrices (X^T X = V S^2 V^T):
//              identical scalings up to column signs, which coef/intercept do not depend on.
//   predict  <- decision_function (lda.py:253-279) and predict_log_proba (lda.py:306-325).
#include <algorithm>
#include <cmath>

#include <stdlib.h>
#include <string.h>

#include "engine.h"

namespace pb {
namespace {

__global__ void lse_combine_kernel(const float* __restrict__ lmax, const float* __restrict__ lsum, long long m,
                                   int n_tiles, float* __restrict__ neg_lse) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= m) return;
  // two partials per (row, column tile): one per epilogue warp half; an unused half holds (-inf, 0)
  const int np = 2 * n_tiles;
  float mx = -INFINITY;
  for (int t = 0; t < np; ++t) mx = fmaxf(mx, lmax[i * np + t]);
  float s = 0.f;
  for (int t = 0; t < np; ++t) {
    const float lm = lmax[i * np + t];
    if (lm > -INFINITY) s += lsum[i * np + t] * __expf(lm - mx);
  }
  neg_lse[i] = -(mx + __logf(s));
}

// exact-mode row log-softmax on an fp64 grid (+ intercept), fp32 out; one block per row
__global__ void __launch_bounds__(256)
lda_epilogue_f64_kernel(const double* __restrict__ z, long long nt, long long k, const double* __restrict__ intercept,
                        int log_proba, float* __restrict__ out, long long ldo) {
  __shared__ double red[256];
  const long long r = blockIdx.x;
  const double* row = z + r * k;
  double mx = -INFINITY;
  for (long long c = threadIdx.x; c < k; c += blockDim.x) mx = fmax(mx, row[c] + intercept[c]);
  red[threadIdx.x] = mx;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + s]);
    __syncthreads();
  }
  mx = red[0];
  __syncthreads();
  double sum = 0.0;
  for (long long c = threadIdx.x; c < k; c += blockDim.x) sum += exp(row[c] + intercept[c] - mx);
  red[threadIdx.x] = sum;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  const double lse = log_proba ? mx + log(red[0]) : 0.0;
  for (long long c = threadIdx.x; c < k; c += blockDim.x)
    out[r * ldo + c] = static_cast<float>(row[c] + intercept[c] - lse);
}

}  // namespace

// predict_proba (lda.py:281-304): p = 1 / (1 + exp(-z)) then OvR normalisation by the row sum, in place on fp32 rows
__global__ void __launch_bounds__(256)
proba_rows_kernel(float* __restrict__ z, long long nt, long long k, long long ld, int normalise) {
  __shared__ float red[256];
  const long long r = blockIdx.x;
  float* row = z + r * ld;
  float sum = 0.f;
  for (long long c = threadIdx.x; c < k; c += blockDim.x) {
    const float p = 1.0f / (1.0f + __expf(-row[c]));
    row[c] = p;
    sum += p;
  }
  if (!normalise) return;
  red[threadIdx.x] = sum;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  const float inv = 1.0f / red[0];
  for (long long c = threadIdx.x; c < k; c += blockDim.x) row[c] *= inv;
}

void lse_combine(Context& ctx, const float* lmax, const float* lsum, int64_t m, int n_tiles, float* neg_lse) {
  lse_combine_kernel<<<static_cast<unsigned>(ceil_div(m, 256)), 256, 0, ctx.stream>>>(lmax, lsum, m, n_tiles, neg_lse);
  PB_CUDA(cudaGetLastError());
  ctx.count_launch();
}

void LdaEngine::refresh_operands() {
  coef.reserve(k * d);
  intercept.reserve(k);
  PB_CUDA(cudaMemcpyAsync(coef.get(), h_coef.data(), k * d * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
  PB_CUDA(cudaMemcpyAsync(intercept.get(), h_intercept.data(), k * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
  split_rows(ctx, coef.get(), false, k, d, d, nullptr, nullptr, nullptr, coef_split);
  const int64_t col_ld = round_up(k, 32);
  std::vector<float> f(col_ld, 0.f);
  for (int64_t i = 0; i < k; ++i) f[i] = static_cast<float>(h_intercept[i]);
  intercept_f32.reserve(col_ld);
  PB_CUDA(cudaMemcpyAsync(intercept_f32.get(), f.data(), col_ld * sizeof(float), cudaMemcpyHostToDevice, ctx.stream));
  ctx.sync();
  ready = true;
}

void LdaEngine::set_coef(int64_t k_, int64_t d_, const double* c, const double* b) {
  PB_CHECK(k_ > 0 && d_ > 0 && c && b, kInvalidArg, "lda_set_coef: bad arguments");
  k = k_;
  d = d_;
  h_coef.assign(c, c + k * d);
  h_intercept.assign(b, b + k);
  h_classes.resize(k);
  for (int64_t i = 0; i < k; ++i) h_classes[i] = i;
  refresh_operands();
}

void LdaEngine::class_stats(const void* x, int64_t n, int64_t d_, int64_t ldx, int dtype, int loc,
                            const int64_t* labels, const double* priors_in, int64_t n_priors, ClassStats& out,
                            bool allow_single_class) {
  PB_CHECK((n > 1 || allow_single_class) && n > 0 && d_ > 0 && d_ <= 1024, kInvalidArg, "lda_fit: need n > 1 and 0 < d <= 1024");
  PB_CHECK(labels != nullptr && x != nullptr, kInvalidArg, "lda_fit: null input");
  PB_CHECK(dtype == 0 || dtype == 1, kInvalidArg, "dtype must be PLDA_F64 or PLDA_F32");
  const bool is_f32 = dtype == 1;
  const size_t es = is_f32 ? 4 : 8;
  const void* xd = x;
  int64_t ld = ldx;
  if (loc == 0) {
    ws_in.reserve(static_cast<size_t>(n) * d_ * es);
    PB_CUDA(cudaMemcpy2DAsync(ws_in.get(), d_ * es, x, ldx * es, d_ * es, n, cudaMemcpyHostToDevice, ctx.stream));
    xd = ws_in.get();
    ld = d_;
  }
  // order-preserving map of signed labels to uint64 keys (classes = np.unique(labels), lda.py:118)
  std::vector<uint64_t> keys(n);
  for (int64_t i = 0; i < n; ++i) keys[i] = static_cast<uint64_t>(labels[i]) ^ (1ull << 63);
  DevBuf<uint64_t> dkeys(n);
  PB_CUDA(cudaMemcpyAsync(dkeys.get(), keys.data(), n * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx.stream));
  Segments segs;
  build_segments(ctx, dkeys.get(), n, segs);
  const int64_t kk = segs.nseg;
  if (!allow_single_class) {
    PB_CHECK(kk >= 2, kValueError, "lda_fit: at least two classes are required");
    PB_CHECK(n > kk, kValueError, "lda_fit: need more samples than classes");
  }
  DevBuf<double> means(static_cast<size_t>(kk) * d_);
  DevBuf<int32_t> counts(kk);
  // within scatter  Sw = xc^T xc  (unscaled)
  const size_t dd = static_cast<size_t>(d_) * d_;
  DevBuf<double> sw(dd);
  static const char* stats_mode = getenv("PLDA_B200_STATS");    // "legacy": round-1 materialised-operand path (A/B)
  const bool fused = precision != 1 && d_ <= scatter_fused_max_dim() &&
                     !(stats_mode != nullptr && strcmp(stats_mode, "legacy") == 0);
  if (fused) {
    // class means and the class-centred Gram (lda.py:183-191) in ONE read of the rows (csrc/scatter_tc.cu)
    scatter_fused(ctx, xd, is_f32, d_, ld, segs, /*scale_by_count=*/false, sw.get(), means.get(), counts.get(), scat);
  } else {
    segment_sums(ctx, xd, is_f32, d_, ld, segs, means.get());
    segment_finalize_means(ctx, means.get(), d_, segs, counts.get());
    if (precision == 1) {
      ws_gram.reserve(static_cast<size_t>(n) * d_);
      center_scale_f64(ctx, xd, is_f32, d_, ld, segs, means.get(), false, ws_gram.get());
      gemm_f64(ctx, true, false, d_, d_, n, 1.0, ws_gram.get(), d_, ws_gram.get(), d_, 0.0, sw.get(), d_);
    } else {
      SplitBuf xt;
      center_scale_split_t(ctx, xd, is_f32, d_, ld, segs, means.get(), false, xt);
      const int ks = choose_ksplit(ctx, d_, d_, n);
      const int eff = effective_ksplit(ctx, d_, d_, n, ks);
      DevBuf<float> partial(static_cast<size_t>(eff) * round_up(d_, 128) * round_up(d_, 4));
      gemm_bf16x3_splitk(ctx, xt.view(), xt.view(), d_, d_, n, ks, partial.get());
      reduce_partials_f64(ctx, partial.get(), eff, d_, d_, sw.get(), d_, 1.0, true);
    }
  }
  out.k = kk;
  out.sw.resize(dd);
  out.means.resize(static_cast<size_t>(kk) * d_);
  out.counts.resize(kk);
  std::vector<uint64_t> h_keys(kk);
  PB_CUDA(cudaMemcpyAsync(out.sw.data(), sw.get(), dd * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
  PB_CUDA(cudaMemcpyAsync(out.means.data(), means.get(), kk * d_ * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
  PB_CUDA(cudaMemcpyAsync(out.counts.data(), counts.get(), kk * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx.stream));
  PB_CUDA(cudaMemcpyAsync(h_keys.data(), segs.seg_label.get(), kk * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx.stream));
  ctx.sync();
  out.classes.resize(kk);
  for (int64_t c = 0; c < kk; ++c) out.classes[c] = static_cast<int64_t>(h_keys[c] ^ (1ull << 63));
  out.n = n;
  out.d = d_;
  finish_priors(out, priors_in, n_priors);
}

// priors (lda.py:119-127): given or empirical n_k / n, renormalised when they do not sum to one
void LdaEngine::finish_priors(ClassStats& out, const double* priors_in, int64_t n_priors) {
  const int64_t kk = out.k;
  out.priors.resize(kk);
  if (priors_in != nullptr) {
    PB_CHECK(n_priors == kk, kInvalidArg, "lda_fit: priors length does not match the number of classes");
    for (int64_t i = 0; i < kk; ++i) out.priors[i] = priors_in[i];
  } else {
    for (int64_t i = 0; i < kk; ++i) out.priors[i] = static_cast<double>(out.counts[i]) / static_cast<double>(out.n);
  }
  double psum = 0.0;
  for (double v : out.priors) psum += v;
  if (psum != 1.0) for (double& v : out.priors) v /= psum;
}

// Stats of this rank's rows only (sharded fit, SURVEY 8e "LDA fit"): the caller all-reduces the scatter and
// all-gathers the per-class rows, then calls fit_from_stats on every rank.
void LdaEngine::local_class_stats(const void* x, int64_t n, int64_t d_, int64_t ldx, int dtype, int loc,
                                  const int64_t* labels) {
  class_stats(x, n, d_, ldx, dtype, loc, labels, nullptr, 0, shard_stats, /*allow_single_class=*/true);
}

void LdaEngine::fit_from_stats(int solver, int64_t n, int64_t kk, int64_t d_, const double* sw, const double* means,
                               const int64_t* counts, const int64_t* classes, const double* priors_in,
                               int64_t n_priors) {
  PB_CHECK(kk >= 2, kValueError, "lda_fit: at least two classes are required");
  PB_CHECK(n > kk, kValueError, "lda_fit: need more samples than classes");
  PB_CHECK(d_ > 0 && d_ <= 1024, kInvalidArg, "lda_fit: need 0 < d <= 1024");
  PB_CHECK(sw && means && counts && classes, kInvalidArg, "lda_fit_from_stats: null input");
  ClassStats cs;
  cs.k = kk;
  cs.n = n;
  cs.d = d_;
  cs.sw.assign(sw, sw + static_cast<size_t>(d_) * d_);
  cs.means.assign(means, means + static_cast<size_t>(kk) * d_);
  cs.counts.resize(kk);
  cs.classes.assign(classes, classes + kk);
  int64_t total = 0;
  for (int64_t c = 0; c < kk; ++c) {
    PB_CHECK(counts[c] > 0, kInvalidArg, "lda_fit_from_stats: class counts must be positive");
    PB_CHECK(c == 0 || classes[c] > classes[c - 1], kInvalidArg, "lda_fit_from_stats: classes must be strictly increasing");
    cs.counts[c] = static_cast<int32_t>(counts[c]);
    total += counts[c];
  }
  PB_CHECK(total == n, kInvalidArg, "lda_fit_from_stats: class counts do not add up to n");
  finish_priors(cs, priors_in, n_priors);
  PB_CHECK(solver >= 0 && solver <= 2, kInvalidArg, "lda_fit_from_stats: solver must be 0 (svd), 1 (lsqr) or 2 (eigen)");
  if (solver == 0) solve_svd(cs); else if (solver == 1) solve_lsqr(cs); else solve_eigen(cs);
}

void LdaEngine::fit_svd(const void* x, int64_t n, int64_t d_, int64_t ldx, int dtype, int loc, const int64_t* labels,
                        const double* priors_in, int64_t n_priors) {
  ClassStats cs;
  class_stats(x, n, d_, ldx, dtype, loc, labels, priors_in, n_priors, cs, false);
  solve_svd(cs);
}

// lda.py:178-221 (_solve_svd) from the class statistics: both SVDs as symmetric eigenproblems of Gram matrices
void LdaEngine::solve_svd(const ClassStats& cs) {
  const double tol = 1e-4;
  const int64_t kk = cs.k, n = cs.n, d_ = cs.d;
  const size_t dd = static_cast<size_t>(d_) * d_;
  const std::vector<double>& h_sw = cs.sw;
  const std::vector<double>& h_means = cs.means;
  const std::vector<double>& pri = cs.priors;
  // xbar = priors . means ; std_j = sqrt(Sw_jj / n)  (xc has zero column mean) ; fac = 1/(n - K)
  std::vector<double> xbar(d_, 0.0), stdv(d_);
  for (int64_t c = 0; c < kk; ++c)
    for (int64_t j = 0; j < d_; ++j) xbar[j] += pri[c] * h_means[c * d_ + j];
  for (int64_t j = 0; j < d_; ++j) {
    double s = std::sqrt(h_sw[j * d_ + j] / static_cast<double>(n));
    stdv[j] = s == 0.0 ? 1.0 : s;
  }
  const double fac = 1.0 / static_cast<double>(n - kk);
  // G1 = fac * D^-1 Sw D^-1  -> eig on device
  std::vector<double> g1(dd);
  for (int64_t i = 0; i < d_; ++i)
    for (int64_t j = 0; j < d_; ++j) g1[i * d_ + j] = fac * h_sw[i * d_ + j] / (stdv[i] * stdv[j]);
  DevBuf<double> dg(dd), dvt(dd), dlam(d_);
  EigWork ew;
  PB_CUDA(cudaMemcpyAsync(dg.get(), g1.data(), dd * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
  eig_sym_jacobi(ctx, dg.get(), d_, nullptr, dlam.get(), dvt.get(), ew, nullptr);
  std::vector<double> lam(d_), vt(dd);
  PB_CUDA(cudaMemcpyAsync(lam.data(), dlam.get(), d_ * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
  PB_CUDA(cudaMemcpyAsync(vt.data(), dvt.get(), dd * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
  ctx.sync();
  int64_t rank = 0;
  for (int64_t i = 0; i < d_; ++i) if (std::sqrt(lam[i]) > tol) ++rank;     // S > tol (lda.py:202)
  PB_CHECK(rank > 0, kValueError, "lda_fit: within-class scatter has rank 0");
  // scalings[j, r] = V[r, j] / std_j / S_r     [d x rank]
  std::vector<double> scal(static_cast<size_t>(d_) * rank);
  for (int64_t r = 0; r < rank; ++r) {
    const double s = std::sqrt(lam[r]);
    for (int64_t j = 0; j < d_; ++j) scal[j * rank + r] = vt[r * d_ + j] / stdv[j] / s;
  }
  // X2 = diag(sqrt(n p fac)) (means - xbar) scalings      [K x rank]   (device GEMM)
  std::vector<double> mcw(static_cast<size_t>(kk) * d_), mc(static_cast<size_t>(kk) * d_);
  for (int64_t c = 0; c < kk; ++c) {
    const double w = std::sqrt(static_cast<double>(n) * pri[c] * fac);
    for (int64_t j = 0; j < d_; ++j) {
      mc[c * d_ + j] = h_means[c * d_ + j] - xbar[j];
      mcw[c * d_ + j] = w * mc[c * d_ + j];
    }
  }
  DevBuf<double> dmcw(kk * d_), dmc(kk * d_), dscal(d_ * rank), dx2(kk * rank), dg2(rank * rank);
  PB_CUDA(cudaMemcpyAsync(dmcw.get(), mcw.data(), kk * d_ * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
  PB_CUDA(cudaMemcpyAsync(dmc.get(), mc.data(), kk * d_ * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
  PB_CUDA(cudaMemcpyAsync(dscal.get(), scal.data(), d_ * rank * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
  gemm_f64(ctx, false, false, kk, rank, d_, 1.0, dmcw.get(), d_, dscal.get(), rank, 0.0, dx2.get(), rank);
  gemm_f64(ctx, true, false, rank, rank, kk, 1.0, dx2.get(), rank, dx2.get(), rank, 0.0, dg2.get(), rank);   // X2^T X2
  DevBuf<double> dvt2(rank * rank), dlam2(rank);
  eig_sym_jacobi(ctx, dg2.get(), rank, nullptr, dlam2.get(), dvt2.get(), ew, nullptr);
  std::vector<double> lam2(rank);
  PB_CUDA(cudaMemcpyAsync(lam2.data(), dlam2.get(), rank * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
  ctx.sync();
  const double s0 = std::sqrt(lam2[0]);
  int64_t rank2 = 0;
  for (int64_t i = 0; i < rank; ++i) if (std::sqrt(lam2[i]) > tol * s0) ++rank2;   // lda.py:212
  PB_CHECK(rank2 > 0, kValueError, "lda_fit: between-class scatter has rank 0");
  // _scalings = scalings V2[:, :rank2]  (rows of dvt2 are eigenvectors) -> [d x rank2]
  DevBuf<double> dscal2(d_ * rank2), dcoefp(kk * rank2), dcoef(kk * d_);
  gemm_f64(ctx, false, true, d_, rank2, rank, 1.0, dscal.get(), rank, dvt2.get(), rank, 0.0, dscal2.get(), rank2);
  // coef_proj = (means - xbar) _scalings ; coef = coef_proj _scalings^T
  gemm_f64(ctx, false, false, kk, rank2, d_, 1.0, dmc.get(), d_, dscal2.get(), rank2, 0.0, dcoefp.get(), rank2);
  gemm_f64(ctx, false, true, kk, d_, rank2, 1.0, dcoefp.get(), rank2, dscal2.get(), rank2, 0.0, dcoef.get(), d_);
  std::vector<double> coefp(static_cast<size_t>(kk) * rank2);
  h_coef.resize(static_cast<size_t>(kk) * d_);
  PB_CUDA(cudaMemcpyAsync(coefp.data(), dcoefp.get(), kk * rank2 * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
  PB_CUDA(cudaMemcpyAsync(h_coef.data(), dcoef.get(), kk * d_ * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
  ctx.sync();
  h_intercept.resize(kk);
  h_classes.resize(kk);
  for (int64_t c = 0; c < kk; ++c) {
    double s = 0.0;
    for (int64_t r = 0; r < rank2; ++r) s += coefp[c * rank2 + r] * coefp[c * rank2 + r];
    double dot = 0.0;
    for (int64_t j = 0; j < d_; ++j) dot += xbar[j] * h_coef[c * d_ + j];
    h_intercept[c] = -0.5 * s + std::log(pri[c]) - dot;
    h_classes[c] = cs.classes[c];
  }
  k = kk;
  d = d_;
  // keep xbar / scalings for transform()
  this->rank = rank2;   // (a local `rank` holds the first SVD rank above)
  h_xbar = xbar;
  h_scalings.resize(static_cast<size_t>(d_) * rank2);
  PB_CUDA(cudaMemcpyAsync(h_scalings.data(), dscal2.get(), d_ * rank2 * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
  ctx.sync();
  h_evals.clear();
  refresh_transform_operands();
  refresh_operands();
}

// device copies used by transform(): scalings^T [rank x d] as the B operand of the projection GEMM, xbar
void LdaEngine::refresh_transform_operands() {
  const int64_t d_ = d, r_ = rank;
  std::vector<double> st(static_cast<size_t>(r_) * d_);
  for (int64_t j = 0; j < d_; ++j)
    for (int64_t r = 0; r < r_; ++r) st[r * d_ + j] = h_scalings[j * r_ + r];
  DevBuf<double> dst(r_ * d_);
  PB_CUDA(cudaMemcpyAsync(dst.get(), st.data(), r_ * d_ * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
  split_rows(ctx, dst.get(), false, r_, d_, d_, nullptr, nullptr, nullptr, scal_split);
  xbar_dev.reserve(d_);
  PB_CUDA(cudaMemcpyAsync(xbar_dev.get(), h_xbar.data(), d_ * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
  ctx.sync();
}

void LdaEngine::fit_lsqr(const void* x, int64_t n, int64_t d_, int64_t ldx, int dtype, int loc, const int64_t* labels,
                         const double* priors_in, int64_t n_priors) {
  ClassStats cs;
  class_stats(x, n, d_, ldx, dtype, loc, labels, priors_in, n_priors, cs, false);
  solve_lsqr(cs);
}

void LdaEngine::solve_lsqr(const ClassStats& cs) {
  const int64_t kk = cs.k, n = cs.n, d_ = cs.d;
  const size_t dd = static_cast<size_t>(d_) * d_;
  // cov = sum_k priors_k * cov_k,  cov_k = (1/n_k) sum_{i in k} (x - m_k)(x - m_k)^T   (_class_cov, lda.py:10-16).
  // With empirical priors n_k/n this is Sw/n exactly; general priors need the per-class scatters, which the fused
  // scatter pass does not keep -> only the empirical case runs on the fused path.
  bool empirical = true;
  for (int64_t c = 0; c < kk; ++c)
    if (std::fabs(cs.priors[c] - static_cast<double>(cs.counts[c]) / static_cast<double>(n)) > 1e-12) empirical = false;
  PB_CHECK(empirical, kInvalidArg, "lda lsqr: only empirical priors (priors=None) are supported on the device path");
  std::vector<double> cov(dd);
  for (size_t i = 0; i < dd; ++i) cov[i] = cs.sw[i] / static_cast<double>(n);
  // coef = means cov^-1 via Cholesky: cov = L L^T, T1 = L^-1, cov^-1 = T1^T T1
  DevBuf<double> dcov(dd), dt1(dd), dmeans(kk * d_), dtmp(kk * d_), dcoef(kk * d_);
  DevBuf<int> info(1);
  PB_CUDA(cudaMemcpyAsync(dcov.get(), cov.data(), dd * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
  PB_CUDA(cudaMemcpyAsync(dmeans.get(), cs.means.data(), kk * d_ * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
  cholesky_lower(ctx, dcov.get(), d_, info.get());
  tri_inverse_lower(ctx, dcov.get(), dt1.get(), d_);
  gemm_f64(ctx, false, true, kk, d_, d_, 1.0, dmeans.get(), d_, dt1.get(), d_, 0.0, dtmp.get(), d_);    // means T1^T
  gemm_f64(ctx, false, false, kk, d_, d_, 1.0, dtmp.get(), d_, dt1.get(), d_, 0.0, dcoef.get(), d_);    // (.) T1
  int h_info = 0;
  h_coef.resize(static_cast<size_t>(kk) * d_);
  PB_CUDA(cudaMemcpyAsync(&h_info, info.get(), sizeof(int), cudaMemcpyDeviceToHost, ctx.stream));
  PB_CUDA(cudaMemcpyAsync(h_coef.data(), dcoef.get(), kk * d_ * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
  ctx.sync();
  PB_CHECK(h_info == 0, kValueError, "lda lsqr: the pooled class covariance is singular (need more samples than dims)");
  h_intercept.resize(kk);
  h_classes = cs.classes;
  for (int64_t c = 0; c < kk; ++c) {
    double dot = 0.0;
    for (int64_t j = 0; j < d_; ++j) dot += cs.means[c * d_ + j] * h_coef[c * d_ + j];
    h_intercept[c] = -0.5 * dot + std::log(cs.priors[c]);
  }
  k = kk;
  d = d_;
  rank = 0;
  refresh_operands();
}

void LdaEngine::fit_eigen(const void* x, int64_t n, int64_t d_, int64_t ldx, int dtype, int loc, const int64_t* labels,
                          const double* priors_in, int64_t n_priors) {
  ClassStats cs;
  class_stats(x, n, d_, ldx, dtype, loc, labels, priors_in, n_priors, cs, false);
  solve_eigen(cs);
}

// lda.py:140-176 (_solve_eigen): Sw = sum_k p_k cov_k, St = cov(X), Sb = St - Sw, generalised symmetric
// eigenproblem Sb v = lambda Sw v (scipy eigh(Sb, Sw)), eigenvectors sorted by descending eigenvalue and rescaled
// to unit 2-norm, coef = means V V^T, intercept = -1/2 diag(means coef^T) + log priors.
// eigh(Sb, Sw) is done the way PLDA's GetOutput does it: Sw = C C^T, T1 = C^-1,
// B' = T1 Sb T1^T = U diag(lambda) U^T, V = T1^T U  (so V^T Sw V = I, what LAPACK's dsygvd returns).
// With empirical priors p_k = n_k/n:  Sw = S/n (S = pooled within scatter) and Sb = sum_k p_k (m_k - xbar)(.)^T.
// When K - 1 < d the eigenvalue 0 is degenerate and the basis of its eigenspace is arbitrary (for LAPACK too);
// that part of V V^T only moves every decision value of a sample by the same amount, so the log-softmax and the
// predictions do not depend on it, raw coef / decision values (and the sigmoid-based predict_proba) do.
void LdaEngine::solve_eigen(const ClassStats& cs) {
  const int64_t kk = cs.k, n = cs.n, d_ = cs.d;
  const size_t dd = static_cast<size_t>(d_) * d_;
  bool empirical = true;
  for (int64_t c = 0; c < kk; ++c)
    if (std::fabs(cs.priors[c] - static_cast<double>(cs.counts[c]) / static_cast<double>(n)) > 1e-12) empirical = false;
  PB_CHECK(empirical, kInvalidArg, "lda eigen: only empirical priors (priors=None) are supported on the device path");
  std::vector<double> sw(dd), xbar(d_, 0.0), mcw(static_cast<size_t>(kk) * d_);
  for (size_t i = 0; i < dd; ++i) sw[i] = cs.sw[i] / static_cast<double>(n);
  for (int64_t c = 0; c < kk; ++c)
    for (int64_t j = 0; j < d_; ++j) xbar[j] += cs.priors[c] * cs.means[c * d_ + j];
  for (int64_t c = 0; c < kk; ++c) {
    const double w = std::sqrt(cs.priors[c]);
    for (int64_t j = 0; j < d_; ++j) mcw[c * d_ + j] = w * (cs.means[c * d_ + j] - xbar[j]);
  }
  DevBuf<double> dsw(dd), dt1(dd), dsb(dd), dtmp(dd), dbp(dd), dut(dd), dvt(dd), dlam(d_);
  DevBuf<double> dmcw(kk * d_), dmeans(kk * d_), dproj(kk * d_), dcoef(kk * d_);
  DevBuf<int> info(1);
  EigWork ew;
  PB_CUDA(cudaMemcpyAsync(dsw.get(), sw.data(), dd * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
  PB_CUDA(cudaMemcpyAsync(dmcw.get(), mcw.data(), kk * d_ * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
  PB_CUDA(cudaMemcpyAsync(dmeans.get(), cs.means.data(), kk * d_ * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
  gemm_f64(ctx, true, false, d_, d_, kk, 1.0, dmcw.get(), d_, dmcw.get(), d_, 0.0, dsb.get(), d_);       // Sb
  cholesky_lower(ctx, dsw.get(), d_, info.get());                                                         // C
  tri_inverse_lower(ctx, dsw.get(), dt1.get(), d_);                                                       // T1
  gemm_f64(ctx, false, false, d_, d_, d_, 1.0, dt1.get(), d_, dsb.get(), d_, 0.0, dtmp.get(), d_);        // T1 Sb
  gemm_f64(ctx, false, true, d_, d_, d_, 1.0, dtmp.get(), d_, dt1.get(), d_, 0.0, dbp.get(), d_);         // (.) T1^T
  // The Jacobi solver recovers eigenvectors as normalised columns of B'V, which is noise for an eigenvalue 0 -- and
  // B' has d - K + 1 of them when K - 1 < d.  Solve the shifted problem B' + sigma I (sigma = mean eigenvalue): same
  // eigenvectors, all eigenvalues >= sigma > 0, the null space becomes an ordinary degenerate eigenspace.
  std::vector<double> hbp(dd);
  int h_info = 0;
  PB_CUDA(cudaMemcpyAsync(&h_info, info.get(), sizeof(int), cudaMemcpyDeviceToHost, ctx.stream));
  PB_CUDA(cudaMemcpyAsync(hbp.data(), dbp.get(), dd * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
  ctx.sync();
  PB_CHECK(h_info == 0, kValueError, "lda eigen: the pooled class covariance is singular (need more samples than dims)");
  double sigma = 0.0;
  for (int64_t i = 0; i < d_; ++i) sigma += hbp[i * d_ + i];
  sigma = sigma / static_cast<double>(d_);
  PB_CHECK(sigma > 0.0 && std::isfinite(sigma), kValueError, "lda eigen: the class means coincide (between scatter is zero)");
  for (int64_t i = 0; i < d_; ++i) hbp[i * d_ + i] += sigma;
  PB_CUDA(cudaMemcpyAsync(dbp.get(), hbp.data(), dd * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
  eig_sym_jacobi(ctx, dbp.get(), d_, nullptr, dlam.get(), dut.get(), ew, nullptr);     // rows of dut: eigenvectors
  gemm_f64(ctx, false, false, d_, d_, d_, 1.0, dut.get(), d_, dt1.get(), d_, 0.0, dvt.get(), d_);         // V^T = U^T T1
  std::vector<double> vt(dd), lam(d_);
  PB_CUDA(cudaMemcpyAsync(vt.data(), dvt.get(), dd * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
  PB_CUDA(cudaMemcpyAsync(lam.data(), dlam.get(), d_ * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
  ctx.sync();
  for (int64_t i = 0; i < d_; ++i) lam[i] = std::max(lam[i] - sigma, 0.0);
  // unit 2-norm eigenvectors (lda.py:171); scalings[:, r] = r-th vector
  for (int64_t r = 0; r < d_; ++r) {
    double nrm = 0.0;
    for (int64_t j = 0; j < d_; ++j) nrm += vt[r * d_ + j] * vt[r * d_ + j];
    nrm = std::sqrt(nrm);
    for (int64_t j = 0; j < d_; ++j) vt[r * d_ + j] /= nrm;
  }
  PB_CUDA(cudaMemcpyAsync(dvt.get(), vt.data(), dd * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
  gemm_f64(ctx, false, true, kk, d_, d_, 1.0, dmeans.get(), d_, dvt.get(), d_, 0.0, dproj.get(), d_);     // means V
  gemm_f64(ctx, false, false, kk, d_, d_, 1.0, dproj.get(), d_, dvt.get(), d_, 0.0, dcoef.get(), d_);     // (.) V^T
  h_coef.resize(static_cast<size_t>(kk) * d_);
  PB_CUDA(cudaMemcpyAsync(h_coef.data(), dcoef.get(), kk * d_ * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
  ctx.sync();
  h_intercept.resize(kk);
  h_classes = cs.classes;
  for (int64_t c = 0; c < kk; ++c) {
    double dot = 0.0;
    for (int64_t j = 0; j < d_; ++j) dot += cs.means[c * d_ + j] * h_coef[c * d_ + j];
    h_intercept[c] = -0.5 * dot + std::log(cs.priors[c]);
  }
  // transform for this solver is X scalings (no centring, lda.py:345-346)
  h_scalings.assign(dd, 0.0);
  for (int64_t r = 0; r < d_; ++r)
    for (int64_t j = 0; j < d_; ++j) h_scalings[j * d_ + r] = vt[r * d_ + j];
  h_xbar.assign(d_, 0.0);
  h_evals = lam;
  k = kk;
  d = d_;
  rank = d_;
  refresh_operands();
  refresh_transform_operands();
}

void LdaEngine::transform(const void* x, int64_t nt, int64_t d_, int64_t ldx, int dtype, int loc, int64_t n_components,
                          float* out, int64_t ldo, int out_loc) {
  PB_CHECK(ready && rank > 0, kNotFitted, "transform needs a model fitted with the 'svd' or 'eigen' solver");
  PB_CHECK(d_ == d, kValueError, "X has a different number of features than the model");
  PB_CHECK(n_components > 0 && n_components <= rank && ldo >= n_components, kInvalidArg, "transform: bad n_components");
  if (nt == 0) return;
  const bool is_f32 = dtype == 1;
  const size_t es = is_f32 ? 4 : 8;
  const void* xd = x;
  int64_t ld = ldx;
  if (loc == 0) {
    ws_in.reserve(static_cast<size_t>(nt) * d * es);
    PB_CUDA(cudaMemcpy2DAsync(ws_in.get(), d * es, x, ldx * es, d * es, nt, cudaMemcpyHostToDevice, ctx.stream));
    xd = ws_in.get();
    ld = d;
  }
  const int64_t ldo_dev = out_loc == 1 ? ldo : round_up(n_components, 4);
  float* dst = out;
  if (out_loc == 0) {
    ws_out[0].reserve(static_cast<size_t>(nt) * ldo_dev);
    dst = ws_out[0].get();
  }
  if (precision == 1) {
    ws_gram.reserve(static_cast<size_t>(nt) * (d + n_components) + static_cast<size_t>(d) * rank);
    double* xf = ws_gram.get();
    double* z = xf + nt * d;
    double* sc = z + nt * n_components;
    convert_to_f64(ctx, xd, is_f32, nt, d, ld, xf, d, xbar_dev.get());
    PB_CUDA(cudaMemcpyAsync(sc, h_scalings.data(), d * rank * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
    gemm_f64(ctx, false, false, nt, n_components, d, 1.0, xf, d, sc, rank, 0.0, z, n_components);
    convert_f64_to_f32(ctx, z, nt, n_components, n_components, dst, ldo_dev);
  } else {
    split_rows(ctx, xd, is_f32, nt, d, ld, xbar_dev.get(), nullptr, nullptr, ws_x);
    SplitOperand b = scal_split.view();
    b.rows = n_components;
    GemmEpilogue epi;
    epi.out = dst;
    epi.ldo = ldo_dev;
    gemm_bf16x3(ctx, ws_x.view(), b, nt, n_components, d, epi);
  }
  if (out_loc == 0)
    PB_CUDA(cudaMemcpy2DAsync(out, ldo * sizeof(float), dst, ldo_dev * sizeof(float), n_components * sizeof(float), nt,
                              cudaMemcpyDeviceToHost, ctx.stream));
  ctx.sync();
}

void LdaEngine::predict(const void* x, int64_t nt, int64_t d_, int64_t ldx, int dtype, int loc, int log_proba,
                        float* out, int64_t ldo, int out_loc) {
  PB_CHECK(ready, kNotFitted, "This LDA instance is not fitted yet");
  if (d_ != d)
    throw Error(kValueError, "X has " + std::to_string(d_) + " features per sample; expecting " + std::to_string(d));
  PB_CHECK(ldo >= k, kInvalidArg, "lda_predict: output pitch too small");
  PB_CHECK(dtype == 0 || dtype == 1, kInvalidArg, "dtype must be PLDA_F64 or PLDA_F32");
  if (nt == 0) return;
  const bool is_f32 = dtype == 1;
  const size_t es = is_f32 ? 4 : 8;
  const int64_t ldo_dev = out_loc == 1 ? ldo : round_up(k, 4);
  // row chunks keep the staging buffers bounded (the full 1M x 5k fp32 grid is 20 GB)
  int64_t chunk = nt;
  if (out_loc == 0 || loc == 0 || precision == 1) {
    const int64_t budget = (precision == 1 ? (256ll << 20) : (1ll << 30)) / 4;
    chunk = std::min<int64_t>(nt, std::max<int64_t>(128, (budget / ldo_dev) / 128 * 128));
  }
  const int64_t col_ld = round_up(k, 32);
  for (int64_t r0 = 0; r0 < nt; r0 += chunk) {
    const int64_t rows = std::min(chunk, nt - r0);
    const void* xd;
    int64_t ld;
    if (loc == 0) {
      ws_in.reserve(static_cast<size_t>(rows) * d * es);
      PB_CUDA(cudaMemcpy2DAsync(ws_in.get(), d * es, static_cast<const uint8_t*>(x) + r0 * ldx * es, ldx * es, d * es,
                                rows, cudaMemcpyHostToDevice, ctx.stream));
      xd = ws_in.get();
      ld = d;
    } else {
      xd = static_cast<const uint8_t*>(x) + r0 * ldx * es;
      ld = ldx;
    }
    float* dst = out_loc == 1 ? out + r0 * ldo : nullptr;
    if (out_loc == 0) {
      ws_out[0].reserve(static_cast<size_t>(rows) * ldo_dev);
      dst = ws_out[0].get();
    }
    if (precision == 1) {
      ws_gram.reserve(static_cast<size_t>(rows) * (d + k));
      double* xf = ws_gram.get();
      double* z = xf + rows * d;
      convert_to_f64(ctx, xd, is_f32, rows, d, ld, xf, d);
      gemm_f64(ctx, false, true, rows, k, d, 1.0, xf, d, coef.get(), d, 0.0, z, k);
      lda_epilogue_f64_kernel<<<static_cast<unsigned>(rows), 256, 0, ctx.stream>>>(z, rows, k, intercept.get(),
                                                                                  log_proba == 1 ? 1 : 0, dst, ldo_dev);
      PB_CUDA(cudaGetLastError());
      ctx.count_launch();
    } else {
      split_rows(ctx, xd, is_f32, rows, d, ld, nullptr, nullptr, nullptr, ws_x);
      GemmEpilogue epi;
      epi.col_add = intercept_f32.get();
      epi.col_ld = col_ld;
      if (log_proba == 1) {
        // pass 1: online (max, sum exp) per row and column tile, nothing stored
        const int n_tiles = static_cast<int>(ceil_div(k, k >= 256 ? 256 : round_up(k, 16)));
        ws_lmax.reserve(static_cast<size_t>(rows) * n_tiles * 2);
        ws_lsum.reserve(static_cast<size_t>(rows) * n_tiles * 2);
        ws_neglse.reserve(rows);
        GemmEpilogue e1 = epi;
        e1.lse_max = ws_lmax.get();
        e1.lse_sum = ws_lsum.get();
        gemm_bf16x3(ctx, ws_x.view(), coef_split.view(), rows, k, d, e1);
        lse_combine(ctx, ws_lmax.get(), ws_lsum.get(), rows, n_tiles, ws_neglse.get());
        epi.row_add = ws_neglse.get();     // pass 2 recomputes the tile and stores z - lse
      }
      epi.out = dst;
      epi.ldo = ldo_dev;
      gemm_bf16x3(ctx, ws_x.view(), coef_split.view(), rows, k, d, epi);
    }
    if (log_proba == 2) {     // OvR sigmoid probabilities from the decision values
      proba_rows_kernel<<<static_cast<unsigned>(rows), 256, 0, ctx.stream>>>(dst, rows, k, ldo_dev, k == 2 ? 0 : 1);
      PB_CUDA(cudaGetLastError());
      ctx.count_launch();
    }
    if (out_loc == 0) {
      PB_CUDA(cudaMemcpy2DAsync(out + r0 * ldo, ldo * sizeof(float), dst, ldo_dev * sizeof(float), k * sizeof(float),
                                rows, cudaMemcpyDeviceToHost, ctx.stream));
    }
  }
  ctx.sync();
}

}  // namespace pb
