// PldaEngine / LdaEngine: device-resident state behind the opaque C-ABI handles.
#pragma once

#include <mutex>
#include <unordered_map>
#include <vector>

#include "kernels.h"

namespace pb {

// Stages a caller matrix on the device: HOST -> copied into an owned buffer, DEVICE -> used in place.
struct Staged {
  const void* ptr = nullptr;
  int64_t ld = 0;
  bool is_f32 = false;
  DevBuf<uint8_t> own;
};

struct PldaModel {
  int64_t d = 0;
  bool ready = false;
  DevBuf<double> mean, transform, psi, within, between;
  SplitBuf a_split;              // rows of transform_ as split-bf16 (K-major) for the tensor GEMM
  std::vector<double> h_psi;     // host mirror for cheap validation
};

// Sum-all-reduce hook for the sharded fit (stats pass + EM iteration): the engine packs the fp64
// quantities to reduce into the caller's device scratch buffer and calls `fn(user, count)`; the callee all-reduces
// the first `count` doubles of the scratch in place, stream-ordered on the handle's stream (NCCL via torch).
typedef int (*AllReduceFn)(void* user, int64_t count);

// Sharded score grid over NVLink peer memory (SURVEY 8e "scoring grid"): every rank owns one REGION holding
// ready flags and two generations of the full test operand (split planes + column terms).  A rank's push kernel
// writes ITS test rows into every region (its own and, through CUDA-IPC mappings, the peers'); the score GEMM reads
// the local region, waiting per column tile for the owner's flag.  Two generations + monotonic epochs make the
// exchange safe without any host-side barrier per step (a rank can be at most one step ahead of a peer).
struct ShardSession {
  bool open = false;
  int world = 0, rank = 0;
  int64_t nt_total = 0, dim = 0, ldk = 0, col_ld = 0;
  std::vector<int64_t> bounds;          // [world+1] row partition of the test set
  uint8_t* region = nullptr;            // own region (cudaMalloc)
  size_t bytes = 0;
  size_t off_gen[2] = {0, 0}, off_lo = 0, off_col = 0;   // generation base; lo plane / column terms inside it
  std::vector<uint8_t*> peer;           // [world] region base of every rank as mapped in THIS process
  std::vector<bool> peer_ipc;           // mapped with cudaIpcOpenMemHandle (to be closed)
  unsigned epoch = 0;                   // number of pushes so far (same on every rank)
  int push_count = 0;                   // enrol count the column terms of the current generation were built for
  int max_groups = 0;                   // distinct enrol counts a ragged step may carry (room in the operand pitch)
};

// rows of the background set used by norm(numutts, seed): defined splitmix64 Fisher-Yates (engine_score.cu)
void norm_selection(int64_t m, int64_t numutts, uint64_t seed, int32_t* out);

class PldaEngine {
 public:
  explicit PldaEngine(int device) : ctx(device) {}
  AllReduceFn reduce_fn = nullptr;
  void* reduce_user = nullptr;
  double* reduce_scratch = nullptr;
  int64_t reduce_capacity = 0;
  void allreduce_parts(const std::vector<std::pair<double*, int64_t>>& parts);
  Context ctx;
  std::mutex mu;
  int precision = 0;
  PldaModel model;
  std::unordered_map<uint64_t, std::pair<double, double>> znorm;   // id -> (mean, std)  (meanz/stdvz, pldamodule.cpp:33)
  double fit_ms[5] = {0, 0, 0, 0, 0};

  void set_model(int64_t d, const double* mean, const double* transform, const double* psi);
  void get_model(double* mean, double* transform, double* psi);
  void get_covariances(double* within, double* between);
  void smooth(double factor);
  // labels: uint64 [n], HOST (labels_loc 0) or DEVICE (1)
  void fit(const void* x, int64_t n, int64_t d, int64_t ldx, int dtype, int loc, const uint64_t* labels, int iters,
           int labels_loc = 0);
  DevBuf<uint64_t> ws_labels;
  DevBuf<double> fit_means, fit_scatter, fit_scalars, fit_mc;
  DevBuf<int32_t> fit_counts;
  void transform_grouped(const void* x, int64_t n, int64_t d, int64_t ldx, int dtype, int loc, const uint64_t* labels,
                         int64_t targetdim, uint64_t* out_labels, int64_t* out_counts, double* out_vecs,
                         int64_t* n_out);
  void transform_rows(const void* x, int64_t n, int64_t d, int64_t ldx, int dtype, int loc, const int32_t* counts,
                      int32_t const_count, int64_t targetdim, void* out, int64_t ldo, int out_dtype, int out_loc);
  void score_pair(uint64_t id, int64_t n_enrol, const double* enrol, const double* test, int64_t dim, float* out);
  // zmean_in / zstd_in (fp64 [ne] at z_loc, both or neither): per-row z-norm given as arrays instead of the id table
  void score_grid(const void* enrol, int64_t ne, int64_t ld_enrol, const int32_t* counts, const uint64_t* ids,
                  const void* test, int64_t nt, int64_t ld_test, int64_t dim, int dtype, int loc, float* out,
                  int64_t ldo, int out_loc, const double* zmean_in = nullptr, const double* zstd_in = nullptr,
                  int z_loc = 0);
  // enrol_ids may be null when mean_out / std_out (fp64 [ne] at out_loc) receive the statistics instead of the table
  void norm(const void* bkg, int64_t m, int64_t d, int64_t ldb, int dtype, int loc, const uint64_t* enrol_ids,
            const void* enrol, int64_t ne, int64_t ld_enrol, int64_t dim, int enrol_dtype, int enrol_loc,
            int64_t numutts, uint64_t seed, double* mean_out = nullptr, double* std_out = nullptr, int out_loc = 0);
  // listed trials (engine_sinks.cu): out[i] = LLR(enrol[te[i]], test[tt[i]]); mode 0 auto, 1 direct, 2 grid + gather
  void score_trials(const void* enrol, int64_t ne, int64_t ld_enrol, const int32_t* counts, const uint64_t* ids,
                    const void* test, int64_t nt, int64_t ld_test, int64_t dim, int dtype, int loc, const int32_t* te,
                    const int32_t* tt, int64_t n_trials, int idx_loc, float* out, int out_loc, const double* zmean_in,
                    const double* zstd_in, int z_loc, int mode);
  // target / tail non-target histograms of the grid, nothing materialised (engine_sinks.cu)
  void score_hist(const void* enrol, int64_t ne, int64_t ld_enrol, int32_t enrol_count, const void* test, int64_t nt,
                  int64_t ld_test, int64_t dim, int dtype, int loc, const int32_t* enrol_spk, const int32_t* test_spk,
                  int spk_loc, double lo, double hi, int nbins, double theta_lo, const double* zmean_in,
                  const double* zstd_in, int z_loc, uint64_t* hist_target, uint64_t* hist_nontarget, uint64_t* below,
                  int out_loc);
  // sharded score grid (engine_shard.cu)
  ShardSession shard;
  void shard_open(int world, int rank, const int64_t* bounds, int64_t dim, unsigned char* ipc_handle_out,
                  void** region_out, int max_groups = 0);
  void shard_connect(int peer_rank, const unsigned char* ipc_handle, void* same_process_region);
  void shard_push(const void* test_shard, int64_t nt_local, int64_t ld, int dtype, int enrol_count);
  void shard_score(const void* enrol, int64_t ne, int64_t ld_enrol, int enrol_count, const uint64_t* ids, int dtype,
                   float* out, int64_t ldo);
  // push + score with the enrol-side producer fused into the push kernel (one launch less per step)
  void shard_step(const void* test_shard, int64_t nt_local, int64_t ld_test, const void* enrol, int64_t ne,
                  int64_t ld_enrol, int enrol_count, const uint64_t* ids, int dtype, float* out, int64_t ldo);
  // ragged enrol counts on the sharded grid: group_counts = the distinct counts of ALL ranks (ascending, the same list
  // on every rank, at most the session's max_groups); their column terms travel inside the pushed operand rows
  void shard_step_ragged(const void* test_shard, int64_t nt_local, int64_t ld_test, const void* enrol, int64_t ne,
                         int64_t ld_enrol, const int32_t* enrol_counts, const int32_t* group_counts, int n_groups,
                         const uint64_t* ids, int dtype, float* out, int64_t ldo);
  void shard_status(int64_t* epoch, int64_t* timeouts);
  void shard_close();
  ~PldaEngine();

 private:
  void shard_produce(const void* test_shard, int64_t nt_local, int64_t ld_test, const void* enrol, int64_t ne,
                     int64_t ld_enrol, int enrol_count, int dtype);
  void shard_gemm(int64_t ne, const uint64_t* ids, float* out, int64_t ldo);
  GemmShard shard_desc() const;
  void shard_targets(PrepDst& dst, PrepSignal& sig);

 public:
  void test_gemm(const double* a, const double* b, int64_t m, int64_t n, int64_t k, int ksplit, float* out);
  void test_linalg(int op, const double* a, int64_t d, double* out, double* out2);
  void test_scatter(const void* x, int64_t n, int64_t d, int dtype, const uint64_t* labels, int scale_by_count,
                    double* scatter_out, double* means_out, int64_t means_capacity, int64_t* k_out);

 private:
  void require_model() const { PB_CHECK(model.ready, kNotFitted, "PLDA model is not fitted (call fit or set_model)"); }
  void refresh_model_operands();
  // enrol counts -> groups of equal count + their constants tables (engine_score.cu)
  struct ScoreGroups {
    bool uniform = true;
    int32_t uniform_count = 1;
    int ng = 1;
    const int32_t* counts_dev = nullptr;   // ragged only
    const int32_t* grp_dev = nullptr;      // ragged only: table index per enrol row
    const int32_t* gcounts_dev = nullptr;  // ragged only: count of each group
    const double* tables = nullptr;        // [ng][kScoreConstsSize]
  };
  ScoreGroups prepare_groups(const int32_t* counts, int64_t ne, int64_t dim);
  void znorm_affine(const uint64_t* ids, int64_t ne, const double* zmean_in, const double* zstd_in, int z_loc,
                    const float** zmean, const float** zinv);
  void produce_score_operands(const Staged& se, int64_t ne, const Staged& st, int64_t nt, int64_t dim,
                              const ScoreGroups& g, int64_t col_ld);
  // set by produce_score_operands: reduction length of the score GEMM (dim, or dim + 2 * groups when the ragged
  // column terms ride inside the operands) and the epilogue's column-term binding
  int64_t score_k = 0;
  bool score_cols_embedded = false;
  void bind_col_terms(GemmEpilogue& epi, const ScoreGroups& g, int64_t r0, int64_t col_ld) const;
  std::vector<int32_t> ragged_key;   // distinct counts ws_tables was built for
  DevBuf<int32_t> rg_counts, rg_grp, rg_gcounts;   // device copies owned by the ragged-count cache
  std::vector<int32_t> last_counts;  // the ragged counts whose device copies (ws_counts / ws_grp) are current
  int64_t ragged_dim = 0;
  DevBuf<float4> ws_mom;
  DevBuf<unsigned long long> ws_hist;
  DevBuf<int32_t> ws_te, ws_tt, ws_spk;
  DevBuf<float> ws_trial_out;
  DevBuf<double> ws_bt, ws_zstat;
  DevBuf<uint8_t> ws_gather;
  // `keep`: grow-only staging buffer owned by the engine (hot host paths must not cudaMalloc/cudaFree per call)
  void stage(const void* p, int64_t rows, int64_t cols, int64_t ld, int dtype, int loc, Staged& s,
             DevBuf<uint8_t>* keep = nullptr);
  // device rows [n x d] (centred by the model mean) -> transformed + length-normalised rows
  void transform_device_rows(const void* x, bool is_f32, int64_t n, int64_t d, int64_t ld, const double* sub,
                             const int32_t* counts_dev, int32_t const_count, int64_t dim, double* out64, int64_t ld64,
                             float* out32, int64_t ld32);
  void em_iteration(int64_t k, int64_t d, const double* scatter, const SplitBuf& mc_split, const double* mc_f64,
                    const int32_t* counts_dev, double w_count, double b_count, bool warm);
  void joint_diagonalise(int64_t d, bool warm, bool final_pass);

  // persistent workspaces (grow-only)
  SplitBuf ws_l, ws_r, ws_x, ws_xt, ws_pt, ws_qt, ws_mc;
  DevBuf<float> ws_row, ws_col, ws_out[2], ws_y, ws_zmean, ws_zinv, ws_partial, ws_u;
  DevBuf<double> ws_rsum, ws_rsq, ws_f64a, ws_f64b, ws_f64c, ws_gram, ws_row64, ws_col64;
  DevBuf<int32_t> ws_counts, ws_grp, ws_gcounts;
  DevBuf<uint8_t> ws_stage[2];
  // per-column LLR constants (a/v, a^2/v, q, log-determinant term) for one (enrol count, dim): computed on the host
  // from the psi mirror once and kept on the device until the model changes
  struct ScoreConsts { int count = 0; int64_t dim = 0; DevBuf<double> dev; };
  std::vector<ScoreConsts> score_consts;
  const double* score_consts_for(int count, int64_t dim);
  static void fill_score_consts(const double* psi, int64_t dim, int count, double* out /* kScoreConstsSize */);
  DevBuf<double> ws_tables;   // ragged counts: [ng][kScoreConstsSize]
  // device->host drain of the score grid: second stream + events, created on first use
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_done[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
  void ensure_copy_stream();
  Segments segs;
  ScatterWork scat;
  EigWork eig;
  // EM state (d x d, fp64)
  DevBuf<double> em_c, em_t1, em_bp, em_u, em_a, em_ainv, em_psi, em_tmp, em_tmp2, em_bs, em_ws, em_db, em_dw;
  DevBuf<int> em_info;
  bool em_have_basis = false;
  // PLDA_B200_EM_PROFILE=1: CUDA events between the phases of the EM loop, averaged per phase when fit returns
  std::vector<std::pair<const char*, cudaEvent_t>> em_marks;
  void em_mark(const char* what);
  void em_report();
};

class LdaEngine {
 public:
  explicit LdaEngine(int device) : ctx(device) {}
  Context ctx;
  std::mutex mu;
  int precision = 0;
  int64_t k = 0, d = 0;
  bool ready = false;
  std::vector<double> h_coef, h_intercept;
  std::vector<int64_t> h_classes;
  DevBuf<double> coef, intercept;
  SplitBuf coef_split;
  DevBuf<float> intercept_f32;

  void fit_svd(const void* x, int64_t n, int64_t d, int64_t ldx, int dtype, int loc, const int64_t* labels,
               const double* priors, int64_t n_priors);
  // lda.py:223-251 (_solve_lsqr): coef = means cov^-1 with cov = sum_k priors_k cov_k
  void fit_lsqr(const void* x, int64_t n, int64_t d, int64_t ldx, int dtype, int loc, const int64_t* labels,
                const double* priors, int64_t n_priors);
  // lda.py:140-176 (_solve_eigen): generalised symmetric eigenproblem of (Sb, Sw), unit-norm eigenvectors
  void fit_eigen(const void* x, int64_t n, int64_t d, int64_t ldx, int dtype, int loc, const int64_t* labels,
                 const double* priors, int64_t n_priors);
  std::vector<double> h_evals;              // eigen solver: generalised eigenvalues, descending
  // lda.py:328-349 (transform, svd solver): (X - xbar) scalings[:, :n_components]
  void transform(const void* x, int64_t nt, int64_t d, int64_t ldx, int dtype, int loc, int64_t n_components,
                 float* out, int64_t ldo, int out_loc);
  std::vector<double> h_xbar, h_scalings;   // svd solver: xbar [d], scalings [d x rank]
  int64_t rank = 0;
  SplitBuf scal_split;                      // scalings^T as the B operand [rank x d]
  DevBuf<double> xbar_dev;
  void set_coef(int64_t k, int64_t d, const double* coef, const double* intercept);
  void predict(const void* x, int64_t nt, int64_t d, int64_t ldx, int dtype, int loc, int log_proba, float* out,
               int64_t ldo, int out_loc);
  // sharded fit: statistics of this rank's rows, then the solver on the merged statistics
  void local_class_stats(const void* x, int64_t n, int64_t d, int64_t ldx, int dtype, int loc, const int64_t* labels);
  void fit_from_stats(int solver, int64_t n, int64_t k, int64_t d, const double* sw, const double* means,
                      const int64_t* counts, const int64_t* classes, const double* priors, int64_t n_priors);

  // shared front end of the solvers: stage rows, segment by class, class means / counts, unscaled within scatter
  struct ClassStats {
    int64_t k = 0, n = 0, d = 0;
    std::vector<double> sw, means, priors;
    std::vector<int32_t> counts;
    std::vector<int64_t> classes;
  };
  ClassStats shard_stats;

 private:
  void refresh_operands();
  void class_stats(const void* x, int64_t n, int64_t d, int64_t ldx, int dtype, int loc, const int64_t* labels,
                   const double* priors, int64_t n_priors, ClassStats& out, bool allow_single_class);
  void finish_priors(ClassStats& out, const double* priors, int64_t n_priors);
  void solve_svd(const ClassStats& cs);
  void solve_lsqr(const ClassStats& cs);
  void solve_eigen(const ClassStats& cs);
  void refresh_transform_operands();
  SplitBuf ws_x;
  DevBuf<float> ws_out[2], ws_lmax, ws_lsum, ws_neglse;
  ScatterWork scat;
  DevBuf<double> ws_gram;
  DevBuf<uint8_t> ws_in;
};

}  // namespace pb
