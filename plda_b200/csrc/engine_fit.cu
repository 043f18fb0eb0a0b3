// PldaEngine::fit -- stats pass, EM iterations in the jointly-diagonalising basis, GetOutput.
//
// Reference call stack replaced:
//   MPlda_fit (src/pldamodule.cpp:42-109)
//     -> PldaStats::AddSamples per speaker with weight 1/n_s (:94-98), Sort (:100)
//     -> PldaEstimator::Estimate (:102-106): EstimateOneIter x iters, GetOutput
#include <algorithm>
#include <chrono>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "engine.h"

namespace pb {
namespace {

__global__ void scale_vec_kernel(double* __restrict__ v, int n, const double* __restrict__ denom) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = v[i] / *denom;
}

__global__ void symmetrise_kernel(double* __restrict__ a, int d) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<long long>(d) * d) return;
  const int i = static_cast<int>(idx / d), j = static_cast<int>(idx % d);
  if (j < i) {
    const double v = 0.5 * (a[idx] + a[static_cast<long long>(j) * d + i]);
    a[idx] = v;
    a[static_cast<long long>(j) * d + i] = v;
  }
}

double ms_between(cudaEvent_t a, cudaEvent_t b) {
  float ms = 0.f;
  cudaEventElapsedTime(&ms, a, b);
  return static_cast<double>(ms);
}

}  // namespace

void PldaEngine::allreduce_parts(const std::vector<std::pair<double*, int64_t>>& parts) {
  if (reduce_fn == nullptr) return;
  int64_t total = 0;
  for (const auto& pr : parts) total += pr.second;
  PB_CHECK(total <= reduce_capacity, kInvalidArg, "all-reduce scratch buffer too small (need 2*d*d + d + 2 doubles)");
  int64_t off = 0;
  for (const auto& pr : parts) {
    PB_CUDA(cudaMemcpyAsync(reduce_scratch + off, pr.first, pr.second * sizeof(double), cudaMemcpyDeviceToDevice,
                            ctx.stream));
    off += pr.second;
  }
  PB_CHECK(reduce_fn(reduce_user, total) == 0, kInternal, "all-reduce callback failed");
  off = 0;
  for (const auto& pr : parts) {
    PB_CUDA(cudaMemcpyAsync(pr.first, reduce_scratch + off, pr.second * sizeof(double), cudaMemcpyDeviceToDevice,
                            ctx.stream));
    off += pr.second;
  }
}

void PldaEngine::em_mark(const char* what) {
  static const bool on = getenv("PLDA_B200_EM_PROFILE") != nullptr;
  if (!on) return;
  cudaEvent_t e;
  PB_CUDA(cudaEventCreate(&e));
  PB_CUDA(cudaEventRecord(e, ctx.stream));
  em_marks.emplace_back(what, e);
}

// phase = the work between the previous mark and this one
void PldaEngine::em_report() {
  if (em_marks.empty()) return;
  ctx.sync();
  std::vector<std::pair<const char*, std::vector<double>>> acc;
  for (size_t i = 1; i < em_marks.size(); ++i) {
    const double ms = ms_between(em_marks[i - 1].second, em_marks[i].second);
    auto it = std::find_if(acc.begin(), acc.end(), [&](const auto& a) { return strcmp(a.first, em_marks[i].first) == 0; });
    if (it == acc.end()) acc.push_back({em_marks[i].first, {ms}});
    else it->second.push_back(ms);
  }
  for (auto& a : acc) {
    std::vector<double>& v = a.second;
    double sum = 0.0;
    for (double x : v) sum += x;
    std::sort(v.begin(), v.end());
    fprintf(stderr, "plda_b200 em phase: %-28s avg %8.4f  median %8.4f  min %8.4f ms over %d\n", a.first,
            sum / v.size(), v[v.size() / 2], v[0], static_cast<int>(v.size()));
  }
  for (auto& m : em_marks) cudaEventDestroy(m.second);
  em_marks.clear();
}

// C = chol(W); T1 = C^-1; B' = T1 B T1^T; B' = U diag(psi) U^T; A = U^T T1; A^-1 = C U
// (PldaEstimator::GetOutput / ComputeNormalizingTransform).  Leaves A in em_a, A^-1 in em_ainv, psi in em_psi.
void PldaEngine::joint_diagonalise(int64_t d, bool warm, bool final_pass) {
  const size_t dd = static_cast<size_t>(d) * d;
  em_c.reserve(dd); em_t1.reserve(dd); em_bp.reserve(dd); em_u.reserve(dd); em_a.reserve(dd); em_ainv.reserve(dd);
  em_psi.reserve(d); em_tmp.reserve(dd); em_info.reserve(1);
  em_mark("(iteration start)");
  PB_CUDA(cudaMemcpyAsync(em_c.get(), model.within.get(), dd * sizeof(double), cudaMemcpyDeviceToDevice, ctx.stream));
  if (!cholesky_inverse_fused(ctx, em_c.get(), em_t1.get(), d, em_info.get())) {
    cholesky_lower(ctx, em_c.get(), d, em_info.get());
    tri_inverse_lower(ctx, em_c.get(), em_t1.get(), d);
  }
  em_mark("cholesky + inverse");
  // B' = T1 B T1^T
  gemm_f64(ctx, false, false, d, d, d, 1.0, em_t1.get(), d, model.between.get(), d, 0.0, em_tmp.get(), d);
  gemm_f64(ctx, false, true, d, d, d, 1.0, em_tmp.get(), d, em_t1.get(), d, 0.0, em_bp.get(), d);
  symmetrise_kernel<<<static_cast<unsigned>(ceil_div(d * d, 256)), 256, 0, ctx.stream>>>(em_bp.get(), static_cast<int>(d));
  ctx.count_launch();
  em_mark("B' = T1 B T1^T");
  // eigenvectors as ROWS of em_u (= U^T), eigenvalues descending, floored at 0
  int sweeps = 0;
  const bool dbg = getenv("PLDA_B200_DBG") != nullptr;
  // Inside the EM loop the solve ends after the first sweep whose largest rotation stayed below 3e-3: Jacobi converges
  // quadratically, so the couplings left are ~1e-5 relative.  The statistics of the iteration inherit that error
  // linearly (the E-step formulas are exact for an exactly diagonalising basis) -- two orders below the 1e-3 parity
  // tolerance -- and it does not accumulate: every iteration re-diagonalises the new (W, B).  GetOutput (the model the
  // caller sees) and the exact fp64 mode are solved to full fp64 accuracy.  (3e-4 costs 2-3 more sweeps per 10-iteration
  // fit at d = 200 for no measurable change of psi / W / B against the oracle.)
  static const char* eig_exact = getenv("PLDA_B200_EIG_EXACT");
  const double stop_rotation = (final_pass || precision == 1 || eig_exact != nullptr) ? 1e-7 : 3e-3;
  eig_sym_jacobi(ctx, em_bp.get(), d, (warm && em_have_basis) ? em_u.get() : nullptr, em_psi.get(), em_tmp.get(), eig,
                 dbg ? &sweeps : nullptr, stop_rotation);
  if (dbg) fprintf(stderr, "plda_b200: joint_diagonalise d=%lld warm=%d sweeps=%d\n", static_cast<long long>(d),
                   static_cast<int>(warm && em_have_basis), sweeps);
  em_mark("eigensolver");
  PB_CUDA(cudaMemcpyAsync(em_u.get(), em_tmp.get(), dd * sizeof(double), cudaMemcpyDeviceToDevice, ctx.stream));
  em_have_basis = true;
  gemm_f64(ctx, false, false, d, d, d, 1.0, em_u.get(), d, em_t1.get(), d, 0.0, em_a.get(), d);     // A = U^T T1
  gemm_f64(ctx, false, true, d, d, d, 1.0, em_c.get(), d, em_u.get(), d, 0.0, em_ainv.get(), d);    // A^-1 = C U
}

void PldaEngine::em_iteration(int64_t k, int64_t d, const double* scatter, const SplitBuf& mc_split,
                              const double* mc_f64, const int32_t* counts_dev, double w_count, double b_count,
                              bool warm) {
  const size_t dd = static_cast<size_t>(d) * d;
  joint_diagonalise(d, warm, /*final_pass=*/false);
  em_bs.reserve(dd); em_ws.reserve(dd); em_db.reserve(d); em_dw.reserve(d); em_tmp2.reserve(dd);
  if (precision == 1) {
    // exact mode: everything fp64
    ws_f64a.reserve(k * d);   // U
    ws_f64b.reserve(k * d);   // P
    ws_f64c.reserve(k * d);   // Q
    gemm_f64(ctx, false, true, k, d, d, 1.0, mc_f64, d, em_a.get(), d, 0.0, ws_f64a.get(), d);
    em_posterior_f64(ctx, ws_f64a.get(), k, d, counts_dev, em_psi.get(), ws_f64b.get(), ws_f64c.get(), em_db.get(),
                     em_dw.get());
    gemm_f64(ctx, true, false, d, d, k, 1.0, ws_f64b.get(), d, ws_f64b.get(), d, 0.0, em_bs.get(), d);   // P^T P
    gemm_f64(ctx, true, false, d, d, k, 1.0, ws_f64c.get(), d, ws_f64c.get(), d, 0.0, em_ws.get(), d);   // Q^T Q
  } else {
    // U = Mc A^T on tcgen05
    SplitBuf& a_split = ws_x;
    split_rows(ctx, em_a.get(), false, d, d, d, nullptr, nullptr, nullptr, a_split);
    const int64_t ldu = round_up(d, 4);
    ws_u.reserve(k * ldu);
    GemmEpilogue epi;
    epi.out = ws_u.get();
    epi.ldo = ldu;
    gemm_bf16x3(ctx, mc_split.view(), a_split.view(), k, d, d, epi);
    em_mark("U = Mc A^T");
    // P and Q stacked in one operand: ONE split-K SYRK ([P ; Q][P ; Q]^T, the diagonal blocks are the two statistics)
    // and one fp64 reduction that also symmetrises and adds the diagonal terms
    em_posterior_stacked(ctx, ws_u.get(), ldu, k, d, counts_dev, em_psi.get(), ws_pt, em_db.get(), em_dw.get());
    const int ks = choose_ksplit(ctx, 2 * d, 2 * d, k);
    const int eff = effective_ksplit(ctx, 2 * d, 2 * d, k, ks);
    const int64_t mpad = round_up(2 * d, 128), npad = round_up(2 * d, 4);
    ws_partial.reserve(static_cast<size_t>(eff) * mpad * npad);
    gemm_bf16x3_splitk(ctx, ws_pt.view(), ws_pt.view(), 2 * d, 2 * d, k, ks, ws_partial.get());
    em_stats_reduce(ctx, ws_partial.get(), eff, d, em_db.get(), em_dw.get(), em_bs.get(), em_ws.get());
    em_mark("posterior + class SYRK");
  }
  if (precision == 1) {
    // add the diagonal terms (no scaling yet)
    add_diag_scale(ctx, em_bs.get(), em_db.get(), d, 1.0, nullptr);
    add_diag_scale(ctx, em_ws.get(), em_dw.get(), d, 1.0, nullptr);
  }
  // sharded fit: every rank holds a shard of the classes and the same (A, psi); the two d x d statistics are
  // the only per-iteration exchange
  allreduce_parts({{em_bs.get(), static_cast<int64_t>(dd)}, {em_ws.get(), static_cast<int64_t>(dd)}});
  // back to the original basis:  X -> A^-1 X A^-T  (the two statistics side by side: two launches, not four)
  em_tmp.reserve(dd);
  gemm_f64_pair(ctx, false, false, d, d, d, 1.0, em_ainv.get(), em_bs.get(), em_tmp2.get(), em_ainv.get(), em_ws.get(),
                em_tmp.get(), d, d, 0.0, d);
  gemm_f64_pair(ctx, false, true, d, d, d, 1.0, em_tmp2.get(), em_ainv.get(), model.between.get(), em_tmp.get(),
                em_ainv.get(), model.within.get(), d, d, 0.0, d);
  // W = (S + .)/W_count ;  B = ./B_count      (EstimateFromStats), symmetrised
  em_finalize(ctx, model.between.get(), model.within.get(), scatter, 1.0 / b_count, 1.0 / w_count, d);
  em_mark("back-transform + finalize");
}

void PldaEngine::fit(const void* x, int64_t n, int64_t d, int64_t ldx, int dtype, int loc, const uint64_t* labels,
                     int iters, int labels_loc) {
  PB_CHECK(n > 0 && d > 0, kInvalidArg, "fit: empty input");
  PB_CHECK(d <= 1024, kInvalidArg, "fit: feature dimension above 1024 is not supported");
  PB_CHECK(labels != nullptr, kInvalidArg, "fit: labels are required");
  PB_CHECK(iters >= 0, kInvalidArg, "fit: iters must be >= 0");
  cudaEvent_t ev[4];
  for (auto& e : ev) PB_CUDA(cudaEventCreate(&e));
  const size_t dd = static_cast<size_t>(d) * d;
  PB_CUDA(cudaEventRecord(ev[0], ctx.stream));

  // PLDA_B200_TRACE=1: host-clock laps of the stats pass, each after a stream synchronisation
  const bool trace = getenv("PLDA_B200_TRACE") != nullptr;
  const auto t_begin = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!trace) return;
    cudaStreamSynchronize(ctx.stream);
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
    fprintf(stderr, "plda_b200 fit: %-32s +%.3f ms\n", what, ms);
  };

  // ---- stats pass (PldaStats::AddSamples for every speaker) ----
  Staged sx;
  stage(x, n, d, ldx, dtype, loc, sx);
  lap("rows staged");
  const uint64_t* lab_dev = labels;
  if (labels_loc == 0) {
    ws_labels.reserve(n);
    PB_CUDA(cudaMemcpyAsync(ws_labels.get(), labels, n * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx.stream));
    lab_dev = ws_labels.get();
  }
  lap("labels uploaded");
  build_segments(ctx, lab_dev, n, segs);                                    // K1
  lap("segments built");
  const int64_t k = segs.nseg;
  if (k < 2 && reduce_fn == nullptr) {
    for (auto& e : ev) cudaEventDestroy(e);
    throw Error(kValueError,
                "Number of speakers is 1. Aborting PLDA esimation, at least two speakers are required!");
  }
  // grow-only workspaces owned by the handle: a cudaMalloc / cudaFree pair per fit costs milliseconds at C4 sizes
  // (200 MB of class means) and cudaFree synchronises the device in the middle of the pass
  fit_means.reserve(static_cast<size_t>(k) * d);
  fit_counts.reserve(k);
  fit_scatter.reserve(dd);
  fit_scalars.reserve(2);
  fit_mc.reserve(static_cast<size_t>(k) * d);
  DevBuf<double>& means = fit_means;
  DevBuf<int32_t>& counts = fit_counts;
  DevBuf<double>& scatter = fit_scatter;
  static const char* stats_mode = getenv("PLDA_B200_STATS");    // "legacy": round-1 materialised-operand path (A/B)
  const bool fused = precision != 1 && d <= scatter_fused_max_dim() &&
                     !(stats_mode != nullptr && strcmp(stats_mode, "legacy") == 0);
  if (fused) {
    // K2 + K3 in one read of the rows: anchor-centred SYRK on tcgen05, class sums from the same registers
    scatter_fused(ctx, sx.ptr, sx.is_f32, d, sx.ld, segs, true, scatter.get(), means.get(), counts.get(), scat);
    lap("fused scatter + class means");
  } else {
    segment_sums(ctx, sx.ptr, sx.is_f32, d, sx.ld, segs, means.get());        // K2
    segment_finalize_means(ctx, means.get(), d, segs, counts.get());
    lap("class means");
    if (precision == 1) {
      ws_gram.reserve(static_cast<size_t>(n) * d);
      center_scale_f64(ctx, sx.ptr, sx.is_f32, d, sx.ld, segs, means.get(), true, ws_gram.get());
      gemm_f64(ctx, true, false, d, d, n, 1.0, ws_gram.get(), d, ws_gram.get(), d, 0.0, scatter.get(), d);
    } else {
      center_scale_split_t(ctx, sx.ptr, sx.is_f32, d, sx.ld, segs, means.get(), true, ws_xt);   // K3 operand
      lap("centred split operand");
      const int ks = choose_ksplit(ctx, d, d, n);
      const int eff = effective_ksplit(ctx, d, d, n, ks);
      const int64_t mpad = round_up(d, 128), npad = round_up(d, 4);
      ws_partial.reserve(static_cast<size_t>(eff) * mpad * npad);
      gemm_bf16x3_splitk(ctx, ws_xt.view(), ws_xt.view(), d, d, n, ks, ws_partial.get());   // K3: S = X~^T X~
      reduce_partials_f64(ctx, ws_partial.get(), eff, d, d, scatter.get(), d, 1.0, true);
    }
  }
  lap("scatter SYRK");
  // sum_ = sum_s w_s m_s, class_weight = sum_s w_s ; mu = sum_/class_weight
  double* class_weight = fit_scalars.get();
  model.d = d;
  model.mean.reserve(d);
  class_weighted_sum(ctx, means.get(), counts.get(), k, d, model.mean.get(), class_weight);
  // sharded fit (whole speakers per rank): S, sum_, class_weight and the class count are the only quantities
  // exchanged by the stats pass
  double* k_dev = fit_scalars.get() + 1;
  const double k_local = static_cast<double>(k);
  PB_CUDA(cudaMemcpyAsync(k_dev, &k_local, sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
  allreduce_parts({{scatter.get(), static_cast<int64_t>(dd)}, {model.mean.get(), d}, {class_weight, 1},
                   {k_dev, 1}});
  scale_vec_kernel<<<static_cast<unsigned>(ceil_div(d, 128)), 128, 0, ctx.stream>>>(model.mean.get(), static_cast<int>(d),
                                                                                  class_weight);
  ctx.count_launch();
  double h_cw = 0.0, h_k = 0.0;
  PB_CUDA(cudaMemcpyAsync(&h_cw, class_weight, sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
  PB_CUDA(cudaMemcpyAsync(&h_k, k_dev, sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
  // centred class means (constant across iterations)
  DevBuf<double>& mc = fit_mc;
  convert_to_f64(ctx, means.get(), false, k, d, d, mc.get(), d, model.mean.get());
  if (precision != 1) split_rows(ctx, mc.get(), false, k, d, d, nullptr, nullptr, nullptr, ws_mc);
  PB_CUDA(cudaEventRecord(ev[1], ctx.stream));
  ctx.sync();
  lap("stats pass done");
  // example_weight = sum w_s n_s = K;  W_count = (K - class_weight) + class_weight = K;  B_count = class_weight
  if (h_k < 2.0) {
    for (auto& e : ev) cudaEventDestroy(e);
    throw Error(kValueError,
                "Number of speakers is 1. Aborting PLDA esimation, at least two speakers are required!");
  }
  const double w_count = h_k;          // global number of classes
  const double b_count = h_cw;

  // ---- EM (InitParameters: W = B = I) ----
  model.within.reserve(dd);
  model.between.reserve(dd);
  set_identity(ctx, model.within.get(), d);
  set_identity(ctx, model.between.get(), d);
  em_have_basis = false;
  for (int it = 0; it < iters; ++it)
    em_iteration(k, d, scatter.get(), ws_mc, mc.get(), counts.get(), w_count, b_count, /*warm=*/it > 0);
  PB_CUDA(cudaEventRecord(ev[2], ctx.stream));

  // ---- GetOutput ----
  joint_diagonalise(d, iters > 0, /*final_pass=*/true);
  int h_info = 0;
  PB_CUDA(cudaMemcpyAsync(&h_info, em_info.get(), sizeof(int), cudaMemcpyDeviceToHost, ctx.stream));
  model.transform.reserve(dd);
  model.psi.reserve(d);
  PB_CUDA(cudaMemcpyAsync(model.transform.get(), em_a.get(), dd * sizeof(double), cudaMemcpyDeviceToDevice, ctx.stream));
  PB_CUDA(cudaMemcpyAsync(model.psi.get(), em_psi.get(), d * sizeof(double), cudaMemcpyDeviceToDevice, ctx.stream));
  PB_CUDA(cudaEventRecord(ev[3], ctx.stream));
  ctx.sync();
  fit_ms[0] = ms_between(ev[0], ev[1]);
  fit_ms[1] = ms_between(ev[1], ev[2]);
  fit_ms[2] = ms_between(ev[2], ev[3]);
  fit_ms[3] = ms_between(ev[0], ev[3]);
  fit_ms[4] = iters;
  em_report();
  for (auto& e : ev) cudaEventDestroy(e);
  PB_CHECK(h_info == 0, kInternal, "within-class covariance is not positive definite (Cholesky failed at column " +
                                      std::to_string(h_info) + ")");
  refresh_model_operands();
}

// test hook: the fused stats pass on its own (HOST rows and labels in; scatter [d*d], class means [k*d] (label order
// ascending) and k out)
void PldaEngine::test_scatter(const void* x, int64_t n, int64_t d, int dtype, const uint64_t* labels, int scale_by_count,
                              double* scatter_out, double* means_out, int64_t means_capacity, int64_t* k_out) {
  PB_CHECK(n > 0 && d > 0 && d <= scatter_fused_max_dim(), kInvalidArg, "test_scatter: bad shape");
  Staged sx;
  stage(x, n, d, d, dtype, 0, sx);
  DevBuf<uint64_t> lab(n);
  PB_CUDA(cudaMemcpyAsync(lab.get(), labels, n * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx.stream));
  build_segments(ctx, lab.get(), n, segs);
  const int64_t k = segs.nseg;
  DevBuf<double> sc(static_cast<size_t>(d) * d), means(static_cast<size_t>(k) * d);
  DevBuf<int32_t> counts(k);
  scatter_fused(ctx, sx.ptr, sx.is_f32, d, sx.ld, segs, scale_by_count != 0, sc.get(), means.get(), counts.get(), scat);
  PB_CUDA(cudaMemcpyAsync(scatter_out, sc.get(), d * d * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
  if (means_out != nullptr && means_capacity >= k * d)
    PB_CUDA(cudaMemcpyAsync(means_out, means.get(), k * d * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
  if (k_out) *k_out = k;
  ctx.sync();
}

void PldaEngine::test_linalg(int op, const double* a, int64_t d, double* out, double* out2) {
  const size_t dd = static_cast<size_t>(d) * d;
  DevBuf<double> da(dd), db(dd), dv(d);
  DevBuf<int> info(1);
  PB_CUDA(cudaMemcpyAsync(da.get(), a, dd * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
  if (op == 0) {
    cholesky_lower(ctx, da.get(), d, info.get());
    PB_CUDA(cudaMemcpyAsync(out, da.get(), dd * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
  } else if (op == 1) {
    tri_inverse_lower(ctx, da.get(), db.get(), d);
    PB_CUDA(cudaMemcpyAsync(out, db.get(), dd * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
  } else if (op == 2) {
    int sweeps = 0;
    eig_sym_jacobi(ctx, da.get(), d, nullptr, dv.get(), db.get(), eig, &sweeps);
    if (getenv("PLDA_B200_DBG")) fprintf(stderr, "plda_b200: jacobi d=%lld converged after %d sweeps\n",
                                         static_cast<long long>(d), sweeps);
    // return eigenvectors as columns
    std::vector<double> vt(dd);
    PB_CUDA(cudaMemcpyAsync(vt.data(), db.get(), dd * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
    PB_CUDA(cudaMemcpyAsync(out2, dv.get(), d * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
    ctx.sync();
    for (int64_t p = 0; p < d; ++p)
      for (int64_t i = 0; i < d; ++i) out[i * d + p] = vt[p * d + i];
  } else if (op == 3 || op == 4) {
    // fused cluster Cholesky + inverse: op 3 returns L, op 4 returns L^-1
    PB_CHECK(cholesky_inverse_fused(ctx, da.get(), db.get(), d, info.get()), kInvalidArg,
             "test_linalg: the fused Cholesky / inverse kernel is not available for this size");
    PB_CUDA(cudaMemcpyAsync(out, op == 3 ? da.get() : db.get(), dd * sizeof(double), cudaMemcpyDeviceToHost, ctx.stream));
  } else {
    throw Error(kInvalidArg, "test_linalg: unknown op");
  }
  ctx.sync();
}

}  // namespace pb
