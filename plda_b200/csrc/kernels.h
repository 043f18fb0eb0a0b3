// Launch wrappers of the SIMT (HBM-bound / latency-bound) kernels of the PLDA + LDA path.
// Kernel numbering follows SURVEY.md section 8a (K1..K9).
#pragma once

#include "runtime.h"

namespace pb {

// ---- generic operand producers ------------------------------------------------ //
// out(hi,lo)[r, c] = split( (in[r, c] - sub[c]) * col_scale[c] * row_scale[r] ), zero padded to ld.
// in is fp64 (is_f32 = false) or fp32, row-major with pitch ld_in.  Null sub/col_scale/row_scale = identity.
void split_rows(Context& ctx, const void* in, bool is_f32, int64_t rows, int64_t cols, int64_t ld_in,
                const double* sub, const double* col_scale, const double* row_scale, SplitBuf& out);

// out[r,c] = (double)in[r,c] - sub[c]   (sub may be null)
void convert_to_f64(Context& ctx, const void* in, bool is_f32, int64_t rows, int64_t cols, int64_t ld_in, double* out,
                    int64_t ld_out, const double* sub = nullptr);

void convert_f64_to_f32(Context& ctx, const double* in, int64_t rows, int64_t cols, int64_t ld_in, float* out,
                        int64_t ld_out);

// ---- scoring (K6 operand prep; reference: Plda::LogLikelihoodRatio via src/pldamodule.cpp:235,266) ---- //
// Enrol side: L = E * (a/v) (per-row n_e), row_term[e] = c(n_e) - 1/2 sum a^2/v e^2.
// enrol is fp64/fp32 [ne x d]; counts int32 [ne] (device); psi fp64 [d] (device).
// Writes the split operand (if l_out), the fp64 operand (if l_f64) and row terms (fp32 and/or fp64).
void score_prep_enrol(Context& ctx, const void* enrol, bool is_f32, int64_t ne, int64_t d, int64_t ld,
                      const int32_t* counts, int const_count, const double* psi, SplitBuf* l_out, double* l_f64, float* row_term,
                      double* row_term_f64);
// Test side: R = T (split), col_term[g][t] = sum_i q_i(n_g) t_i^2 for each distinct enrol count n_g.
// group_counts int32 [g] (device).  col_term pitch col_ld (floats, zero padded).
void score_prep_test(Context& ctx, const void* test, bool is_f32, int64_t nt, int64_t d, int64_t ld,
                     const int32_t* group_counts, int ngroups, int const_count, const double* psi, SplitBuf* r_out, float* col_term,
                     int64_t col_ld, double* col_term_f64);
// Fused enrol + test operand producer for a single enrol count (one launch, per-column constants computed once per
// block, writes the zero padding of col_term itself).  enrol and test share dtype.
// `consts`: device table of the per-column LLR constants for one (model, enrol count), layout kScoreConsts*.
constexpr int kScoreConstsScale = 0, kScoreConstsEnrolSq = 1024, kScoreConstsTestSq = 2048, kScoreConstsLogdet = 3072,
              kScoreConstsSize = 3073;
void score_prep_uniform(Context& ctx, const void* enrol, int64_t ne, int64_t ld_e, const void* test, int64_t nt,
                        int64_t ld_t, bool is_f32, int64_t d, const double* consts, SplitBuf& l_out,
                        SplitBuf& r_out, float* row_term, float* col_term, int64_t col_ld);
// The same producer with the test-side outputs written to up to kMaxPeers destinations (operand planes with pitch
// ld_out and a column-term row each) at global row offset test_row0, and an optional completion signal: the last
// block stores `epoch` (release, system scope) to each `flag[w]`.  Sharded score grid: the
// destinations are every rank's operand buffer over NVLink peer memory -- producer and all-gather in one kernel.
// l_out == nullptr skips the enrol side; tdst.n == 0 skips the test side.
constexpr int kMaxPeers = 16;
struct PrepDst {
  __nv_bfloat16* hi[kMaxPeers];
  __nv_bfloat16* lo[kMaxPeers];
  float* term[kMaxPeers];
  int n = 0;
};
struct PrepSignal {
  unsigned* counter = nullptr;      // local block counter (zero between launches)
  unsigned* flag[kMaxPeers];        // ready flag of THIS source rank inside each destination region
  int n = 0;
  unsigned epoch = 0;
  int fence_per_thread = 0;         // A/B switch (env PLDA_B200_FENCE=thread): system fence in every thread
};
void score_prep_uniform_multi(Context& ctx, const void* enrol, int64_t ne, int64_t ld_e, SplitBuf* l_out,
                              float* row_term, const void* test, int64_t nt, int64_t ld_t, int64_t test_row0,
                              int64_t test_pad_end, const PrepDst& tdst, int64_t ld_out, bool is_f32, int64_t d,
                              const double* consts, const PrepSignal& sig);
// Ragged enrol counts: one launch for both sides with the per-column constants read from one kScoreConsts* table per
// distinct count (tables_dev: [ng][kScoreConstsSize]); grp_dev: group index per enrol row.  Writes one column-term row
// per group (pitch col_ld; the caller zeroes the padding).
void score_prep_grouped(Context& ctx, const void* enrol, int64_t ne, int64_t ld_e, const int32_t* grp_dev,
                        const void* test, int64_t nt, int64_t ld_t, bool is_f32, int64_t d, int ng,
                        const double* tables_dev, SplitBuf& l_out, SplitBuf& r_out, float* row_term, float* col_term,
                        int64_t col_ld);
// exact-mode epilogue applied in place on an fp64 grid: s = (s + row[m] + col[grp[m]][n] - zmean[m]) * zinv[m] -> fp32 out
void score_epilogue_f64(Context& ctx, const double* gram, int64_t ne, int64_t nt, const double* row_term,
                        const double* col_term, int64_t col_ld, const int32_t* grp, const float* zmean,
                        const float* zinv, float* out, int64_t ldo, double* rsum, double* rsq);
// z-norm moments -> mean / population std (src/pldamodule.cpp:240-253)
void znorm_finalize(Context& ctx, const double* rsum, const double* rsq, int64_t ne, int64_t m, float* zmean,
                    float* zinv, double* mean_out, double* std_out);

// ---- sinks of the score grid other than "whole fp32 matrix" (sinks.cu) ------------------------------- //
// z-norm moments from the MOM epilogue's per-slot partials (GemmEpilogue::mom, [ne][n_tiles][2] float4): Chan's
// pairwise update in fp64 -> mean and POPULATION std per enrol row (src/pldamodule.cpp:240-253); any output may be null
void moments_reduce(Context& ctx, const float4* mom, int64_t ne, int n_tiles, float* zmean, float* zinv,
                    double* mean_out, double* std_out);
// Listed trials, scored directly (one warp per trial, fp64 accumulation): out[i] = LLR(enrol[te[i]], test[tt[i]]) in
// the Gram form of SURVEY App. A.7 with the per-count constants of `tables` ([ng][kScoreConstsSize]; grp == null ->
// table 0), optional z-norm affine per enrol row.  Replaces the per-trial loop of scoring/scorePLDA.py:302-318 when
// the list is sparse (the grid + gather_trials wins above ~0.5/dim density).
void score_trials_direct(Context& ctx, const void* enrol, int64_t ld_e, const void* test, int64_t ld_t, bool is_f32,
                         int64_t dim, const double* tables, const int32_t* grp, const float* zmean, const float* zinv,
                         const int32_t* te, const int32_t* tt, int64_t n_trials, float* out);
// out[i] = slab[(te[i] - r0) * ld + tt[i]] for the trials with r0 <= te[i] < r0 + rows (others untouched)
void gather_trials(Context& ctx, const float* slab, int64_t ld, int64_t r0, int64_t rows, const int32_t* te,
                   const int32_t* tt, int64_t n_trials, float* out);
// Vectorised operand producer for ragged enrol counts (16-byte loads / stores, per-row constants table).
// embed: the operands get 2 * ng extra K columns that carry the column terms through the product itself (enrol
// rows: one-hot pair of their group; test rows: the group's term in four bf16 pieces) -> K = d + 2 * ng.
// shard_dst (sharded grid, embed only): the test rows are written at row offset test_row0 into every destination
// (pitch shard_ld) instead of r_out and `sig` raises the ready flags; col_term may then be null.
void score_prep_grouped_vec(Context& ctx, const void* enrol, int64_t ne, int64_t ld_e, const int32_t* grp_dev,
                            const void* test, int64_t nt, int64_t ld_t, bool is_f32, int64_t d, int ng,
                            const double* tables_dev, SplitBuf& l_out, SplitBuf& r_out, float* row_term,
                            float* col_term, int64_t col_ld, bool embed, const PrepDst* shard_dst = nullptr,
                            int64_t test_row0 = 0, int64_t shard_ld = 0, const PrepSignal* sig = nullptr);

// ---- label segmentation (K1) + segmented sums (K2/K4) --------------------------- //
struct Segments {
  DevBuf<int32_t> order;       // [n] row index of sorted position p
  DevBuf<int32_t> seg_of_pos;  // [n] segment id of sorted position p
  DevBuf<uint64_t> seg_label;  // [nseg] label value (ascending)
  DevBuf<int32_t> seg_start;   // [nseg+1] CSR offsets into the sorted order
  int64_t n = 0;
  int64_t nseg = 0;
  // scratch
  DevBuf<uint64_t> keys_in, keys_out;
  DevBuf<int32_t> vals_in, flags;
  DevBuf<uint8_t> cub_tmp;
};
// labels: uint64 [n] on device.  Synchronises once (to learn nseg).
void build_segments(Context& ctx, const uint64_t* labels_dev, int64_t n, Segments& seg);
// sums[seg, :] = sum of rows of that segment (fp64 accumulate); x fp64/fp32 [n x d]
void segment_sums(Context& ctx, const void* x, bool is_f32, int64_t d, int64_t ld, const Segments& seg, double* sums);
// means = sums / count  (in place), counts_out int32 [nseg]
void segment_finalize_means(Context& ctx, double* sums, int64_t d, const Segments& seg, int32_t* counts_out);

// ---- transform (K5 epilogue; reference: Plda::TransformIvector, src/pldamodule.cpp:171,224) ---- //
// y[r,:] *= sqrt(dim / sum_i y_i^2/(psi_i + 1/n_r)); y fp32 [rows x ld] from the tensor GEMM (or fp64 exact),
// writes fp64 (and/or fp32) normalised rows.  dim = number of columns used (targetdim semantics: first `dim` rows of A).
void length_normalise(Context& ctx, const void* y, bool y_is_f32, int64_t rows, int64_t dim, int64_t ld_y,
                      const double* psi, const int32_t* counts, int32_t const_count, double* out64, int64_t ld64,
                      float* out32, int64_t ld32);

// ---- fit: stats (K3 operand) ------------------------------------------------------ //
// Writes the centred, 1/sqrt(n_s)-scaled rows TRANSPOSED as a split operand [d x npad] (K-major in the
// row index, which is the SYRK reduction axis): xt[c, p] = (x[order[p], c] - mean[seg(p), c]) / sqrt(n_seg).
void center_scale_split_t(Context& ctx, const void* x, bool is_f32, int64_t d, int64_t ld, const Segments& seg,
                          const double* means, bool scale_by_count, SplitBuf& xt);
// Fused stats pass (scatter_tc.cu): within-class scatter + class means + counts in ONE read of the rows, no
// materialised operand.  scatter_out [d x d] fp64 symmetric = sum_p w_p (x_p - m_c)(x_p - m_c)^T with w_p = 1/n_c
// (scale_by_count; PldaStats::AddSamples with the reference's weights) or 1; means_out [nseg x d] fp64; counts_out
// [nseg].  Tensor path (split-bf16 x3), d <= scatter_fused_max_dim().
struct ScatterWork {
  DevBuf<int4> meta;
  DevBuf<uint8_t> csum;          // class sums in the rows' own type
  DevBuf<float> delta, partial;
};
int scatter_fused_max_dim();
void scatter_fused(Context& ctx, const void* x, bool is_f32, int64_t d, int64_t ld, const Segments& seg,
                   bool scale_by_count, double* scatter_out, double* means_out, int32_t* counts_out, ScatterWork& w);
// exact mode: same rows, fp64, not transposed [n x d]
void center_scale_f64(Context& ctx, const void* x, bool is_f32, int64_t d, int64_t ld, const Segments& seg,
                      const double* means, bool scale_by_count, double* out);
// sum_out[c] = sum_s w_s * means[s,c] with w_s = 1/n_s; class_weight = sum_s w_s  (PldaStats::AddSamples)
void class_weighted_sum(Context& ctx, const double* means, const int32_t* counts, int64_t k, int64_t d,
                        double* sum_out, double* class_weight_out);

// ---- EM (K7) ----------------------------------------------------------------------- //
// From U = (M - mu) A^T [k x d] build the two weighted SYRK operands (transposed, split) and diagonal terms:
//   r = psi/(1+n psi), g = n r;  P = sqrt(w) g u;  Q = sqrt(w n) (1-g) u;  db = sum_s w r;  dw = sum_s w n r
void em_posterior_t(Context& ctx, const float* u, int64_t ldu, int64_t k, int64_t d, const int32_t* counts,
                    const double* psi, SplitBuf& pt, SplitBuf& qt, double* db, double* dw);
// the same with P and Q stacked in ONE operand ([P ; Q], 2d rows): a single SYRK launch yields both statistics
void em_posterior_stacked(Context& ctx, const float* u, int64_t ldu, int64_t k, int64_t d, const int32_t* counts,
                          const double* psi, SplitBuf& pq, double* db, double* dw);
// bs / ws [d x d] from the split-K partials of the stacked SYRK (+ symmetrise + diagonal terms), one launch
void em_stats_reduce(Context& ctx, const float* partial, int ksplit, int64_t d, const double* db, const double* dw,
                     double* bs, double* ws);
// B = sym(B) / B_count ; W = (sym(W) + S) / W_count   (EstimateFromStats), one launch
void em_finalize(Context& ctx, double* between, double* within, const double* scatter, double inv_b, double inv_w,
                 int64_t d);
void em_posterior_f64(Context& ctx, const double* u, int64_t k, int64_t d, const int32_t* counts, const double* psi,
                      double* p, double* q, double* db, double* dw);
// out = (base? base:0) + scale * x ; adds diag (if diag) before scaling:  out = base + scale*(x + diag(dg))
void add_diag_scale(Context& ctx, double* x, const double* dg, int64_t d, double scale, const double* base);
void set_identity(Context& ctx, double* a, int64_t d);

// ---- d x d fp64 factorizations (K8) ------------------------------------------------- //
// In-place lower Cholesky of a (row-major, ld = d); upper triangle zeroed.  *info (device int) = 0 or failing column+1.
void cholesky_lower(Context& ctx, double* a, int64_t d, int* info_dev);
// a <- chol(a) (lower, upper zeroed) and inv <- a^-1 in ONE cluster launch (d <= 512); false: not available on this
// device / size (the caller uses cholesky_lower + tri_inverse_lower)
bool cholesky_inverse_fused(Context& ctx, double* a, double* inv, int64_t d, int* info_dev);
// inv = L^-1 (lower triangular), row-major
void tri_inverse_lower(Context& ctx, const double* l, double* inv, int64_t d);
// Symmetric eigendecomposition by one-sided (Hestenes) Jacobi, cooperative grid kernel.
//   in : b (d x d symmetric PSD-ish), v0_t = optional warm-start orthogonal basis (ROWS are the vectors), may be null
//   out: evals (descending, floored at 0), evecs_t (d x d row-major, ROW p = eigenvector of the p-th largest value)
struct EigWork {
  DevBuf<double> g, v, lam, tmp;
  DevBuf<int> flags;
  DevBuf<int32_t> perm;
  DevBuf<long long> dbg;     // PLDA_B200_DBG=1: SM cycles of CTA 1 per phase (load, Gram, tournament, apply, barrier, rounds)
};
// stop_rotation: the solve ends after the first sweep whose largest rotation |gamma| / sqrt(alpha beta) stayed below
// it; Jacobi converges quadratically, so the couplings left are ~stop_rotation^2 relative (1e-7 -> full fp64
// accuracy; the EM iterations in between use a looser value, see PldaEngine::joint_diagonalise)
void eig_sym_jacobi(Context& ctx, const double* b, int64_t d, const double* v0_t, double* evals, double* evecs_t,
                    EigWork& work, int* sweeps_out, double stop_rotation = 1e-7);

// ---- d-vector pooling (scoring/extractdvector.py:19-58) ------------------------------------ //
// frames [n_frames x d] (device), utterance u = frames [offsets[u], offsets[u+1]) (host CSR offsets);
// out[u, :] = mean / max / population variance (mode 0 / 1 / 2) over the utterance's frames, each frame divided by
// its L2 norm first when l2norm.  out is a device fp64 matrix.  Synchronises.
void dvector_pool(Context& ctx, const void* frames, bool is_f32, int64_t n_frames, int64_t d, int64_t ld,
                  const int64_t* offsets_host, int64_t n_utts, int mode, bool l2norm, double* out_dev, int64_t ldo);

// ---- LDA ----------------------------------------------------------------------------- //
// log-softmax finalisation: lse[m] from per-tile (max,sum) partials
void lse_combine(Context& ctx, const float* lmax, const float* lsum, int64_t m, int n_tiles, float* neg_lse);

}  // namespace pb
