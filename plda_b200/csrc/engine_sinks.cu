// PldaEngine: sinks of the score grid that never send the whole matrix to the caller.
//
//   score_trials   the trials of a LIST -- what scoring/scorePLDA.py:302-318 asks of MPlda_score (src/pldamodule.cpp:
//                  258-277) -- scored on the device: sparse lists straight from the transformed vectors (one warp per
//                  trial, fp64 accumulation), dense lists as grid slabs + a gather, so the host sees n_trials floats
//   score_hist     target / non-target score histograms for the EER (scoring/eer.py:68-73) of a grid that is too
//                  large to materialise (BASELINE configs[3]: 2e11 trials): fused into the GEMM epilogue; non-targets
//                  below `theta_lo` are only counted, so the bulk of the grid costs no atomic
#include <algorithm>

#include "engine.h"

namespace pb {

void PldaEngine::score_trials(const void* enrol, int64_t ne, int64_t ld_enrol, const int32_t* counts,
                              const uint64_t* ids, const void* test, int64_t nt, int64_t ld_test, int64_t dim,
                              int dtype, int loc, const int32_t* te, const int32_t* tt, int64_t n_trials, int idx_loc,
                              float* out, int out_loc, const double* zmean_in, const double* zstd_in, int z_loc,
                              int mode) {
  require_model();
  PB_CHECK(dim > 0 && dim <= model.d, kInvalidArg, "score_trials: vector dimension does not match the model");
  PB_CHECK(ne >= 0 && nt >= 0 && n_trials >= 0, kInvalidArg, "score_trials: negative size");
  PB_CHECK(ne < (1ll << 31) && nt < (1ll << 31), kInvalidArg, "score_trials: too many vectors");
  PB_CHECK(mode >= 0 && mode <= 2, kInvalidArg, "score_trials: mode must be 0 (auto), 1 (direct) or 2 (grid)");
  PB_CHECK(idx_loc == 0 || idx_loc == 1, kInvalidArg, "score_trials: bad index location");
  if (n_trials == 0) return;
  PB_CHECK(te != nullptr && tt != nullptr && out != nullptr, kInvalidArg, "score_trials: null pointer");
  PB_CHECK(ne > 0 && nt > 0, kInvalidArg, "score_trials: trials listed against an empty enrol or test set");
  const ScoreGroups g = prepare_groups(counts, ne, dim);
  Staged se, st;
  stage(enrol, ne, dim, ld_enrol, dtype, loc, se, &ws_stage[0]);
  stage(test, nt, dim, ld_test, dtype, loc, st, &ws_stage[1]);
  const float* zmean = nullptr;
  const float* zinv = nullptr;
  znorm_affine(ids, ne, zmean_in, zstd_in, z_loc, &zmean, &zinv);

  const int32_t* te_d = te;
  const int32_t* tt_d = tt;
  if (idx_loc == 0) {
    for (int64_t i = 0; i < n_trials; ++i)
      PB_CHECK(te[i] >= 0 && te[i] < ne && tt[i] >= 0 && tt[i] < nt, kInvalidArg, "score_trials: trial index out of range");
    ws_te.reserve(n_trials);
    ws_tt.reserve(n_trials);
    PB_CUDA(cudaMemcpyAsync(ws_te.get(), te, n_trials * sizeof(int32_t), cudaMemcpyHostToDevice, ctx.stream));
    PB_CUDA(cudaMemcpyAsync(ws_tt.get(), tt, n_trials * sizeof(int32_t), cudaMemcpyHostToDevice, ctx.stream));
    te_d = ws_te.get();
    tt_d = ws_tt.get();
  }
  float* out_d = out;
  if (out_loc == 0) {
    ws_trial_out.reserve(n_trials);
    out_d = ws_trial_out.get();
  }
  // direct: 2 rows of `dim` values gathered per trial; grid: 4 bytes written + read per grid cell at tensor rate.
  // The grid wins once the list holds more than ~0.5 / dim of the cells (0.25 % at d = 200).
  const double density = static_cast<double>(n_trials) / (static_cast<double>(ne) * static_cast<double>(nt));
  const bool direct = mode == 1 || precision == 1 || (mode == 0 && density * static_cast<double>(dim) < 0.5);
  if (direct) {
    score_trials_direct(ctx, se.ptr, se.ld, st.ptr, st.ld, se.is_f32, dim, g.tables, g.grp_dev, zmean, zinv, te_d, tt_d,
                        n_trials, out_d);
  } else {
    const int64_t col_ld = round_up(nt, 32);
    const int64_t ld_slab = round_up(nt, 4);
    produce_score_operands(se, ne, st, nt, dim, g, col_ld);
    int64_t chunk = std::max<int64_t>(128, ((1ll << 30) / 4 / ld_slab) / 128 * 128);     // 1 GB slab
    chunk = std::min(chunk, ne);
    ws_out[0].reserve(static_cast<size_t>(chunk) * ld_slab);
    for (int64_t r0 = 0; r0 < ne; r0 += chunk) {
      const int64_t rows = std::min(chunk, ne - r0);
      SplitOperand a = ws_l.view();
      a.hi += r0 * a.ld;
      a.lo += r0 * a.ld;
      a.rows = rows;
      GemmEpilogue epi;
      epi.out = ws_out[0].get();
      epi.ldo = ld_slab;
      epi.row_add = ws_row.get() + r0;
      bind_col_terms(epi, g, r0, col_ld);
      epi.zmean = zmean ? zmean + r0 : nullptr;
      epi.zinv = zinv ? zinv + r0 : nullptr;
      gemm_bf16x3(ctx, a, ws_r.view(), rows, nt, score_k, epi);
      gather_trials(ctx, ws_out[0].get(), ld_slab, r0, rows, te_d, tt_d, n_trials, out_d);
    }
  }
  if (out_loc == 0) {
    PB_CUDA(cudaMemcpyAsync(out, out_d, n_trials * sizeof(float), cudaMemcpyDeviceToHost, ctx.stream));
    ctx.sync();
  } else if (ctx.owns_stream) {
    ctx.sync();
  }
}

void PldaEngine::score_hist(const void* enrol, int64_t ne, int64_t ld_enrol, int32_t enrol_count, const void* test,
                            int64_t nt, int64_t ld_test, int64_t dim, int dtype, int loc, const int32_t* enrol_spk,
                            const int32_t* test_spk, int spk_loc, double lo, double hi, int nbins, double theta_lo,
                            const double* zmean_in, const double* zstd_in, int z_loc, uint64_t* hist_target,
                            uint64_t* hist_nontarget, uint64_t* below, int out_loc) {
  require_model();
  PB_CHECK(precision == 0, kInvalidArg, "score_hist runs on the tensor path (precision bf16x3)");
  PB_CHECK(dim > 0 && dim <= model.d, kInvalidArg, "score_hist: vector dimension does not match the model");
  PB_CHECK(ne > 0 && nt > 0, kInvalidArg, "score_hist: empty grid");
  PB_CHECK(enrol_count > 0, kInvalidArg, "score_hist: enrol count must be positive (uniform counts only)");
  PB_CHECK(nbins >= 2 && nbins <= (1 << 24) && hi > lo, kInvalidArg, "score_hist: bad histogram range");
  PB_CHECK(enrol_spk && test_spk && hist_target && hist_nontarget && below, kInvalidArg, "score_hist: null pointer");
  PB_CHECK((spk_loc == 0 || spk_loc == 1) && (out_loc == 0 || out_loc == 1), kInvalidArg, "score_hist: bad location");
  Staged se, st;
  stage(enrol, ne, dim, ld_enrol, dtype, loc, se, &ws_stage[0]);
  stage(test, nt, dim, ld_test, dtype, loc, st, &ws_stage[1]);
  const float* zmean = nullptr;
  const float* zinv = nullptr;
  znorm_affine(nullptr, ne, zmean_in, zstd_in, z_loc, &zmean, &zinv);
  const int32_t* rs = enrol_spk;
  const int32_t* cs = test_spk;
  if (spk_loc == 0) {
    ws_spk.reserve(ne + nt);
    PB_CUDA(cudaMemcpyAsync(ws_spk.get(), enrol_spk, ne * sizeof(int32_t), cudaMemcpyHostToDevice, ctx.stream));
    PB_CUDA(cudaMemcpyAsync(ws_spk.get() + ne, test_spk, nt * sizeof(int32_t), cudaMemcpyHostToDevice, ctx.stream));
    rs = ws_spk.get();
    cs = ws_spk.get() + ne;
  }
  const size_t words = 2 * static_cast<size_t>(nbins) + 1;
  ws_hist.reserve(words);
  PB_CUDA(cudaMemsetAsync(ws_hist.get(), 0, words * sizeof(unsigned long long), ctx.stream));
  const int64_t col_ld = round_up(nt, 32);
  ws_row.reserve(ne);
  ws_col.reserve(col_ld);
  score_prep_uniform(ctx, se.ptr, ne, se.ld, st.ptr, nt, st.ld, se.is_f32, dim, score_consts_for(enrol_count, dim), ws_l,
                     ws_r, ws_row.get(), ws_col.get(), col_ld);
  GemmEpilogue epi;
  epi.row_add = ws_row.get();
  epi.col_add = ws_col.get();
  epi.col_ld = col_ld;
  epi.zmean = zmean;
  epi.zinv = zinv;
  epi.row_spk = rs;
  epi.col_spk = cs;
  epi.hist_t = ws_hist.get();
  epi.hist_n = ws_hist.get() + nbins;
  epi.below = ws_hist.get() + 2 * nbins;
  epi.hist_lo = static_cast<float>(lo);
  epi.hist_scale = static_cast<float>(static_cast<double>(nbins) / (hi - lo));
  epi.theta_lo = static_cast<float>(theta_lo);
  epi.nbins = nbins;
  gemm_bf16x3(ctx, ws_l.view(), ws_r.view(), ne, nt, dim, epi);
  const cudaMemcpyKind kind = out_loc == 1 ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
  PB_CUDA(cudaMemcpyAsync(hist_target, ws_hist.get(), nbins * sizeof(uint64_t), kind, ctx.stream));
  PB_CUDA(cudaMemcpyAsync(hist_nontarget, ws_hist.get() + nbins, nbins * sizeof(uint64_t), kind, ctx.stream));
  PB_CUDA(cudaMemcpyAsync(below, ws_hist.get() + 2 * nbins, sizeof(uint64_t), kind, ctx.stream));
  if (out_loc == 0 || ctx.owns_stream) ctx.sync();
}

}  // namespace pb
