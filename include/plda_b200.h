/* plda_b200 -- C ABI of the B200-native PLDA / LDA hot path.
 *
 * This header is the drop-in boundary: every entry point replaces one method of the
 * reference's native object `libplda.MPlda` (src/pldamodule.cpp) or one method of the
 * reference's pure-Python `LDA` class (python/liblda/lda.py).  Plain C types only
 * (pointers + sizes); no CUDA or torch types.  All functions return 0 on success and
 * a negative PLDA_E_* status otherwise; plda_last_error() returns the message of the
 * last failure on the calling thread.  There is NO CPU fallback: without a sm_100
 * device plda_create()/lda_create() fail with PLDA_E_CUDA.
 *
 * Conventions
 *   - matrices are row-major; `ld*` are row pitches in ELEMENTS
 *   - dtype: PLDA_F64 (reference contract: C-contiguous float64, kaldi-utils.hpp:99-111)
 *            or PLDA_F32 (superset)
 *   - `loc`: PLDA_HOST pointers are copied in/out by the call (the reference deep-copies
 *            its inputs too, src/pldamodule.cpp:72,120,202,264-265); PLDA_DEVICE pointers
 *            are used in place on the handle's device
 *   - labels are 64-bit unsigned (src/pldamodule.cpp:74,139), always HOST for *_fit /
 *     *_transform (they are 8 bytes per row)
 *   - calls on one handle are serialised internally; distinct handles are independent
 */
#ifndef PLDA_B200_H_
#define PLDA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PLDA_OK 0
#define PLDA_E_INVALID (-1)   /* bad argument                                             */
#define PLDA_E_CUDA (-2)      /* CUDA failure / no device                                  */
#define PLDA_E_NOTFITTED (-3) /* model not fitted / set                                    */
#define PLDA_E_VALUE (-4)     /* the reference raises ValueError here (e.g. one speaker)   */
#define PLDA_E_INTERNAL (-5)

#define PLDA_F64 0
#define PLDA_F32 1
#define PLDA_HOST 0
#define PLDA_DEVICE 1

/* precision of the dense contractions */
#define PLDA_PREC_BF16X3 0 /* tcgen05, split-bf16 x3, fp32 accumulate (default)            */
#define PLDA_PREC_FP64 1   /* exact mode: fp64 SIMT kernels                                */

typedef struct plda_handle_s* plda_handle_t;
typedef struct lda_handle_s* lda_handle_t;

const char* plda_last_error(void);
const char* plda_version(void);

/* ---- lifetime: replaces Plda_new / MPLDA_dealloc (src/pldamodule.cpp:297-316) ---------- */
int plda_create(int device, plda_handle_t* out);
int plda_destroy(plda_handle_t h);
int plda_set_precision(plda_handle_t h, int precision);
/* run on an externally owned cudaStream_t (e.g. torch's current stream); NULL restores the own stream */
int plda_set_stream(plda_handle_t h, void* cuda_stream);
int plda_synchronize(plda_handle_t h);
/* Order the handle's stream AFTER everything enqueued so far on `producer_stream` (a cudaStream_t; NULL = the legacy
 * default stream): callers that fill DEVICE operands with their own kernels / copies (e.g. torch's current stream) call
 * this before a plda_* call that reads them, instead of synchronising the host.  DEVICE outputs are complete when a
 * call returns on the handle's own stream; on a stream installed with plda_set_stream they are stream-ordered. */
int plda_stream_wait(plda_handle_t h, void* producer_stream);
/* number of plda_b200 kernels launched through this handle so far (bench "gpu_launches") */
int plda_launch_count(plda_handle_t h, int64_t* out);

/* per-launch CUDA-event timing of the tensor-core GEMM kernel on the launching stream (bench roofline):
 * enable/disable (resets the record); collect synchronises and returns the summed kernel ms + launch count */
int plda_profile_gemm(plda_handle_t h, int enable);
int plda_profile_collect(plda_handle_t h, double* total_ms, int64_t* count);

/* Sharded fit (one process per GPU, whole speakers per rank): the library calls `fn(user, count)` whenever the
 * first `count` fp64 values of `scratch_dev` (a device buffer of `capacity` >= 2*d*d + d + 2 doubles owned by the
 * caller) must be SUM-all-reduced in place across ranks, stream-ordered on the handle's stream (plda_set_stream).
 * Calls per fit: one after the stats pass (scatter, weighted mean sum, class weight, class count) and one per EM
 * iteration (the two d x d statistics).  fn == NULL disables it.  Returns 0 on success from fn. */
typedef int (*plda_allreduce_fn)(void* user, int64_t count);
int plda_set_allreduce(plda_handle_t h, plda_allreduce_fn fn, void* user, double* scratch_dev, int64_t capacity);

/* ---- fit: replaces MPlda_fit (src/pldamodule.cpp:42-109) ------------------------------- *
 * x: [n x d]; labels: [n] uint64 (any values; the reference requires dense 0..K-1, :88-92 --
 * dense labels give identical results).  iters = EM iterations (default 10 in the reference).
 * Errors: PLDA_E_VALUE if only one distinct label (:83-86).                                 */
int plda_fit(plda_handle_t h, const void* x, int64_t n, int64_t d, int64_t ldx, int dtype, int loc,
             const uint64_t* labels, int iters);
/* same with the labels at `labels_loc` (PLDA_DEVICE: uint64 [n] on the handle's device -- no 8 B/row upload) */
int plda_fit_labels(plda_handle_t h, const void* x, int64_t n, int64_t d, int64_t ldx, int dtype, int loc,
                    const uint64_t* labels, int labels_loc, int iters);
/* timing breakdown of the last fit in milliseconds: [0] ingest+stats pass, [1] EM iterations,
 * [2] GetOutput, [3] total; [4] EM iterations run */
int plda_fit_timings(plda_handle_t h, double out[5]);

/* model access (Kaldi Plda{mean_, transform_, psi_}); HOST fp64 buffers: mean[d], transform[d*d], psi[d] */
int plda_dim(plda_handle_t h, int64_t* d);
int plda_get_model(plda_handle_t h, double* mean, double* transform, double* psi);
int plda_set_model(plda_handle_t h, int64_t d, const double* mean, const double* transform, const double* psi);
/* final EM covariances (within, between) [d*d] each, HOST fp64 -- for parity checks */
int plda_get_covariances(plda_handle_t h, double* within, double* between);
/* Plda::SmoothWithinClassCovariance (src/pldamodule.cpp:158-160); mutates the model like the reference */
int plda_smooth(plda_handle_t h, double factor);

/* ---- transform: replaces Mplda_transform (src/pldamodule.cpp:111-194) ------------------- *
 * Groups rows by label (ascending label order = std::map order, :164), averages, applies
 * Plda::TransformIvector with num_examples = group size.  targetdim = 0 keeps d outputs;
 * targetdim = r keeps the r leading (largest-psi) directions (defined semantics; the
 * reference's own plumbing is broken).
 * Outputs (HOST, caller-allocated for the worst case of n groups):
 *   out_labels[n_out], out_counts[n_out], out_vecs[n_out x out_dim] fp64                    */
int plda_transform(plda_handle_t h, const void* x, int64_t n, int64_t d, int64_t ldx, int dtype, int loc,
                   const uint64_t* labels, int64_t targetdim, uint64_t* out_labels, int64_t* out_counts,
                   double* out_vecs, int64_t* n_out);
/* batched variant without grouping: row r is one vector averaged over counts[r] utterances
 * (counts HOST int32, NULL = all ones).  out: [n x out_dim] of out_dtype at out_loc.         */
int plda_transform_rows(plda_handle_t h, const void* x, int64_t n, int64_t d, int64_t ldx, int dtype, int loc,
                        const int32_t* counts, int32_t const_count, int64_t targetdim, void* out, int64_t ldo,
                        int out_dtype, int out_loc);

/* ---- score: replaces MPlda_score (src/pldamodule.cpp:258-277) --------------------------- *
 * One (enrol, test) pair, already transformed, HOST fp64 [dim]; z-normalised iff model_id was
 * seen by plda_norm (:269-273).  The reference returns the fp64 LLR (Py_BuildValue("f", ...) takes a C double, :276);
 * this entry point hands back fp32 (the grid dtype): <= 6e-8 relative, far inside the 1e-3 score tolerance. */
int plda_score_pair(plda_handle_t h, uint64_t model_id, int64_t n_enrol, const double* enrol, const double* test,
                    int64_t dim, float* out);
/* All-pairs grid: out[e, t] = LLR(enrol_e (n = counts[e]), test_t)  [ne x nt] fp32.
 * enrol [ne x dim], test [nt x dim] (dtype/loc as given); enrol_counts HOST int32 [ne];
 * enrol_ids HOST uint64 [ne] or NULL: if given, rows whose id was seen by plda_norm are z-normalised. */
int plda_score_grid(plda_handle_t h, const void* enrol, int64_t ne, int64_t ld_enrol, const int32_t* enrol_counts,
                    const uint64_t* enrol_ids, const void* test, int64_t nt, int64_t ld_test, int64_t dim, int dtype,
                    int loc, float* out, int64_t ldo, int out_loc);

/* plda_score_grid with the z-norm given as arrays (fp64 [ne] at z_loc, from plda_norm_rows; NULL = none) */
int plda_score_grid_z(plda_handle_t h, const void* enrol, int64_t ne, int64_t ld_enrol, const int32_t* enrol_counts,
                      const void* test, int64_t nt, int64_t ld_test, int64_t dim, int dtype, int loc, float* out,
                      int64_t ldo, int out_loc, const double* zmean, const double* zstd, int z_loc);
/* Listed trials -- what the reference's callers actually need (scoring/scorePLDA.py:302-318 calls MPlda_score once per
 * line of the trial list): out[i] = LLR(enrol[trial_enrol[i]], test[trial_test[i]]), z-normalised like plda_score_grid
 * (ids seen by plda_norm) or by the zmean / zstd arrays.  trial_* : int32 [n_trials] at idx_loc (HOST indices are range
 * checked; DEVICE indices must be valid); out: fp32 [n_trials] at out_loc.  mode 0 picks the cheaper of 1 = direct
 * (one warp per trial, fp64 accumulation, wins below ~0.5/dim list density) and 2 = grid slabs + gather on the device. */
int plda_score_trials(plda_handle_t h, const void* enrol, int64_t ne, int64_t ld_enrol, const int32_t* enrol_counts,
                      const uint64_t* enrol_ids, const void* test, int64_t nt, int64_t ld_test, int64_t dim, int dtype,
                      int loc, const int32_t* trial_enrol, const int32_t* trial_test, int64_t n_trials, int idx_loc,
                      float* out, int out_loc, const double* zmean, const double* zstd, int z_loc, int mode);
/* Histogram sink for the EER (scoring/eer.py:68-73) of a grid that is never materialised: trial (e, t) is a TARGET iff
 * enrol_spk[e] == test_spk[t] (int32 at spk_loc).  bin = clamp(floor((score - lo) * nbins / (hi - lo)), 0, nbins-1);
 * every target goes into hist_target, a non-target into hist_nontarget only if score >= theta_lo and otherwise into the
 * single counter *below (theta_lo = -inf histograms everything; choose it below the EER threshold, e.g. a low quantile of
 * the target scores, and the bulk of the grid costs no atomic).  One enrol count for all rows.  Outputs: uint64
 * [nbins], [nbins], [1] at out_loc, overwritten (the caller sums slabs / ranks). */
int plda_score_hist(plda_handle_t h, const void* enrol, int64_t ne, int64_t ld_enrol, int32_t enrol_count,
                    const void* test, int64_t nt, int64_t ld_test, int64_t dim, int dtype, int loc,
                    const int32_t* enrol_spk, const int32_t* test_spk, int spk_loc, double lo, double hi, int nbins,
                    double theta_lo, const double* zmean, const double* zstd, int z_loc, uint64_t* hist_target,
                    uint64_t* hist_nontarget, uint64_t* below, int out_loc);

/* ---- norm: replaces MPlda_norm (src/pldamodule.cpp:196-256) ------------------------------ *
 * bkg: [m x d] RAW (untransformed) background vectors; each selected row is transformed with
 * num_examples = m (sic, :224) and scored as LLR(bkg, n=1, enrol_k) against every enrol vector
 * (:235); mean and population std per enrol id are stored (first insert wins, :245,250).
 * numutts = 0 -> all rows; otherwise a seeded random subset of that size (:204-213).          */
int plda_norm(plda_handle_t h, const void* bkg, int64_t m, int64_t d, int64_t ldb, int dtype, int loc,
              const uint64_t* enrol_ids, const void* enrol, int64_t ne, int64_t ld_enrol, int64_t dim, int enrol_dtype,
              int enrol_loc, int64_t numutts, uint64_t seed);
/* Array form of plda_norm for large enrol sets (same statistics, src/pldamodule.cpp:196-256): nothing is inserted into
 * the id table; mean_out / std_out (fp64 [ne], HOST or DEVICE) receive the per-row mean and population std and go
 * back into plda_score_grid_z / plda_score_trials / plda_score_hist as `zmean` / `zstd`. */
int plda_norm_rows(plda_handle_t h, const void* bkg, int64_t m, int64_t d, int64_t ldb, int dtype, int loc,
                   const void* enrol, int64_t ne, int64_t ld_enrol, int64_t dim, int enrol_dtype, int enrol_loc,
                   int64_t numutts, uint64_t seed, double* mean_out, double* std_out, int out_loc);
/* The background rows plda_norm / plda_norm_rows use for (m, numutts, seed): the first numutts entries of a
 * Fisher-Yates shuffle of 0..m-1 driven by splitmix64(seed) (the reference shuffles with an unseeded
 * std::random_shuffle, :204-213); numutts = 0 -> all m rows in order.  rows_out: HOST int32 [numutts or m]. */
int plda_norm_selection(int64_t m, int64_t numutts, uint64_t seed, int32_t* rows_out);
int plda_znorm_size(plda_handle_t h, int64_t* n);
int plda_znorm_get(plda_handle_t h, uint64_t* ids, double* mean, double* stdv, int64_t capacity, int64_t* n);
int plda_znorm_clear(plda_handle_t h);
/* restore tables saved with plda_znorm_get (first insert wins, like plda_norm) */
int plda_znorm_set(plda_handle_t h, const uint64_t* ids, const double* mean, const double* stdv, int64_t n);

/* ---- sharded score grid over NVLink peer memory (multi-GPU form of plda_score_grid) --------- *
 * Replaces, for one process per GPU, the all-pairs loop over MPlda_score (src/pldamodule.cpp:258-277)
 * when the grid is sharded by ENROL BLOCK and the test vectors by row block (`bounds`, [world+1]).
 * Instead of an all-gather collective, plda_shard_push writes this rank's test rows (split operand +
 * column terms for `enrol_count`) into the operand buffer of EVERY rank through peer mappings and
 * raises a ready flag; plda_shard_score runs the grid of this rank's enrol block against all test
 * rows, waiting per column tile for the rank that owns the rows.  Uniform enrol count only (ragged
 * counts: all-gather + plda_score_grid).  All pointers are DEVICE pointers; both calls are
 * stream-ordered on the handle's stream and do NOT synchronise (plda_synchronize does).
 *   open     allocates the region; ipc_handle_out (64 bytes, cudaIpcMemHandle_t) is what the peers
 *            need, region_out the raw pointer for peers living in the same process
 *   connect  once per peer, with that peer's IPC handle (other process) or region pointer (same
 *            process); all ranks must have connected (barrier) before the first push
 *   push / score must be called the same number of times, in the same order, on every rank
 *   status   epoch = pushes so far; timeouts = waits that gave up (2 s; sticky) -- results are invalid if > 0      */
#define PLDA_IPC_HANDLE_BYTES 64
int plda_shard_open(plda_handle_t h, int world, int rank, const int64_t* bounds, int64_t dim,
                    unsigned char* ipc_handle_out, void** region_out);
int plda_shard_connect(plda_handle_t h, int peer_rank, const unsigned char* ipc_handle, void* same_process_region);
int plda_shard_push(plda_handle_t h, const void* test_shard, int64_t nt_local, int64_t ld_test, int dtype,
                    int enrol_count);
int plda_shard_score(plda_handle_t h, const void* enrol, int64_t ne, int64_t ld_enrol, int enrol_count,
                     const uint64_t* enrol_ids, int dtype, float* out, int64_t ldo);
/* push + score in one call: the enrol-side operand producer shares the push kernel's launch (test blocks first,
 * so their peer stores are in flight while the enrol rows are processed) */
int plda_shard_step(plda_handle_t h, const void* test_shard, int64_t nt_local, int64_t ld_test, const void* enrol,
                    int64_t ne, int64_t ld_enrol, int enrol_count, const uint64_t* enrol_ids, int dtype, float* out,
                    int64_t ldo);
/* Ragged enrol counts on the sharded grid (scoring/scorePLDA.py enrols speakers with differing numbers of utterances;
 * the reference scores them pair by pair, src/pldamodule.cpp:258-277).  open_ragged = open with room in the operand
 * rows for `max_groups` (<= 8) distinct enrol counts; step_ragged = plda_shard_step with a count per enrol row
 * (host int32[ne]) and `group_counts`: the distinct counts over ALL ranks, strictly ascending, the SAME list on
 * every rank (n_groups <= max_groups).  The column terms of every group travel inside the pushed operand rows, so
 * the grid runs the same kernel as for one uniform count.  Uniform steps remain valid on such a session.           */
int plda_shard_open_ragged(plda_handle_t h, int world, int rank, const int64_t* bounds, int64_t dim, int max_groups,
                           unsigned char* ipc_handle_out, void** region_out);
int plda_shard_step_ragged(plda_handle_t h, const void* test_shard, int64_t nt_local, int64_t ld_test, const void* enrol,
                           int64_t ne, int64_t ld_enrol, const int32_t* enrol_counts, const int32_t* group_counts,
                           int n_groups, const uint64_t* enrol_ids, int dtype, float* out, int64_t ldo);
int plda_shard_status(plda_handle_t h, int64_t* epoch, int64_t* timeouts);
int plda_shard_close(plda_handle_t h);

/* ---- d-vector pooling: replaces scoring/extractdvector.py:19-58 for a batch of utterances ------------- *
 * frames [n_frames x d]: frame-level activations of ALL utterances, utterance u = rows [offsets[u], offsets[u+1])
 * (offsets: HOST int64 [n_utts+1], ascending; every utterance needs >= 1 frame).  l2norm != 0: each frame is divided
 * by its L2 norm first (getnormalizedvector, :19-29).  mode: mean (extractdvectormean :37-39), max
 * (extractdvectormax :32-34), population variance (extractdvectorvar :42-46); l2norm = 0 gives the *_nol2 variants
 * (:49-58).  out: fp64 [n_utts x d].  Needs no fitted model (any plda handle provides the device and stream). */
#define PLDA_POOL_MEAN 0
#define PLDA_POOL_MAX 1
#define PLDA_POOL_VAR 2
int plda_dvector_pool(plda_handle_t h, const void* frames, int64_t n_frames, int64_t d, int64_t ld, int dtype, int loc,
                      const int64_t* offsets, int64_t n_utts, int mode, int l2norm, double* out, int64_t ldo,
                      int out_loc);

/* ---- LDA: replaces python/liblda/lda.py (LDA.fit svd :178-221, decision_function :253-279,
 *      predict_log_proba :306-325) ---------------------------------------------------------- */
int lda_create(int device, lda_handle_t* out);
int lda_destroy(lda_handle_t h);
int lda_set_precision(lda_handle_t h, int precision);
int lda_launch_count(lda_handle_t h, int64_t* out);
int lda_synchronize(lda_handle_t h);
int lda_stream_wait(lda_handle_t h, void* producer_stream);   /* see plda_stream_wait */
/* labels: HOST int64 [n] (any values; classes = sorted unique, lda.py:118); priors NULL = empirical */
int lda_fit_svd(lda_handle_t h, const void* x, int64_t n, int64_t d, int64_t ldx, int dtype, int loc,
                const int64_t* labels, const double* priors, int64_t n_priors);
/* lsqr solver (lda.py:223-251): coef = means cov^-1, cov = pooled class covariance; empirical priors only */
int lda_fit_lsqr(lda_handle_t h, const void* x, int64_t n, int64_t d, int64_t ldx, int dtype, int loc,
                 const int64_t* labels, const double* priors, int64_t n_priors);
/* eigen solver (lda.py:140-176): generalised symmetric eigenproblem Sb v = lambda Sw v (scipy eigh(Sb, Sw)),
 * eigenvectors by descending eigenvalue rescaled to unit 2-norm, coef = means V V^T; empirical priors only.
 * lda_get_svd then returns rank = d, xbar = 0 and scalings = V (transform is X V, lda.py:345-346);
 * lda_get_eigenvalues the generalised eigenvalues (descending, floored at 0). */
int lda_fit_eigen(lda_handle_t h, const void* x, int64_t n, int64_t d, int64_t ldx, int dtype, int loc,
                  const int64_t* labels, const double* priors, int64_t n_priors);
int lda_get_eigenvalues(lda_handle_t h, double* evals, int64_t capacity, int64_t* n);
/* Sharded fit (SURVEY 8e "LDA fit": rows sharded by class, one d x d all-reduce + one all-gather of class rows).
 * lda_class_stats: statistics of THIS rank's rows -- k local classes, their means / counts / labels (ascending) and the
 * unscaled within-class scatter sum_i (x_i - m_class(i))(x_i - m_class(i))^T; read them with lda_get_class_stats
 * (sw[d*d], means[k*d], counts[k], classes[k]; any pointer may be NULL).  After the caller has summed `sw` over ranks
 * and concatenated the per-class rows (classes strictly increasing, every class on exactly one rank),
 * lda_fit_from_stats runs the solver (0 = svd, lda.py:178-221; 1 = lsqr, lda.py:223-251) on every rank. */
int lda_class_stats(lda_handle_t h, const void* x, int64_t n, int64_t d, int64_t ldx, int dtype, int loc,
                    const int64_t* labels, int64_t* k);
int lda_get_class_stats(lda_handle_t h, double* sw, double* means, int64_t* counts, int64_t* classes);
int lda_fit_from_stats(lda_handle_t h, int solver /* 0 svd, 1 lsqr, 2 eigen */, int64_t n, int64_t k, int64_t d, const double* sw,
                       const double* means, const int64_t* counts, const int64_t* classes, const double* priors,
                       int64_t n_priors);
/* svd solver state: rank (0 for other solvers), xbar[d], scalings[d*rank] row-major (any pointer may be NULL) */
int lda_get_svd(lda_handle_t h, int64_t* rank, double* xbar, double* scalings);
/* LDA.transform for the svd solver (lda.py:328-349, evident intent): out[nt x n_components] = (X - xbar) scalings */
int lda_transform(lda_handle_t h, const void* x, int64_t nt, int64_t d, int64_t ldx, int dtype, int loc,
                  int64_t n_components, float* out, int64_t ldo, int out_loc);
int lda_num_classes(lda_handle_t h, int64_t* k, int64_t* d);
int lda_get_coef(lda_handle_t h, double* coef /* [k*d] */, double* intercept /* [k] */, int64_t* classes /* [k] */);
int lda_set_coef(lda_handle_t h, int64_t k, int64_t d, const double* coef, const double* intercept);
/* out[nt x k] fp32: decision values (log_proba = 0), row log-softmax of them (1: predict_log_proba) or
 * OvR sigmoid probabilities (2: predict_proba, lda.py:281-304; not row-normalised when k == 2) */
int lda_predict(lda_handle_t h, const void* x, int64_t nt, int64_t d, int64_t ldx, int dtype, int loc, int log_proba,
                float* out, int64_t ldo, int out_loc);

/* ---- memory helpers for callers that want resident data without torch --------------------- */
int plda_device_malloc(int device, size_t bytes, void** out);
int plda_device_free(int device, void* p);
int plda_host_malloc_pinned(size_t bytes, void** out);
int plda_host_free_pinned(void* p);
int plda_memcpy(void* dst, const void* src, size_t bytes, int kind /* 0 h2d, 1 d2h, 2 d2d */);

/* ---- kernel-level entry used by the parity tests: C[m x n] = A[m x k] * B[n x k]^T (HOST fp64 in,
 *      HOST fp32 out) through the tensor-core kernel (ksplit <= 1: fused-store path; > 1: split-K) */
int plda_test_gemm(plda_handle_t h, const double* a, const double* b, int64_t m, int64_t n, int64_t k, int ksplit,
                   float* out);
/* stall counters (clock cycles) of CTA 0 and 1 of the last tensor GEMM launch; needs env PLDA_B200_DBG=1.
 * per CTA 16 slots: 0 producer wait-empty, 1 producer total, 2 mma wait-full, 3 mma wait-tmem-empty, 4 mma total,
 * 5 tiles, 6/7/8 epilogue warp 0 wait-tmem-full / wait-store / total, 9/10/11 same for epilogue warp 7 */
int plda_debug_counters(plda_handle_t h, int64_t* out, int n);
/* the fused stats pass on its own (csrc/scatter_tc.cu): HOST rows [n x d] (contiguous) and labels in; scatter [d*d] =
 * sum_p w_p (x_p - m_c)(x_p - m_c)^T with w_p = 1/n_c (scale_by_count != 0, PldaStats::AddSamples with the reference's
 * weights, src/pldamodule.cpp:97) or 1; class means [k*d] in ascending label order if means_capacity >= k*d; d <= 512 */
int plda_test_scatter(plda_handle_t h, const void* x, int64_t n, int64_t d, int dtype, const uint64_t* labels,
                      int scale_by_count, double* scatter_out, double* means_out, int64_t means_capacity,
                      int64_t* k_out);
/* fp64 d x d helpers exposed for tests: op 0 = cholesky (lower), 1 = lower-triangular inverse,
 * 2 = symmetric eig (out = eigenvectors as columns, out2 = eigenvalues descending),
 * 3 / 4 = the fused single-launch Cholesky + inverse (out = L / out = L^-1; d <= 512) */
int plda_test_linalg(plda_handle_t h, int op, const double* a, int64_t d, double* out, double* out2);

#ifdef __cplusplus
}
#endif
#endif /* PLDA_B200_H_ */
