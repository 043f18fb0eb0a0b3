// extern "C" boundary (include/plda_b200.h): converts C++ exceptions into status codes and a
// thread-local last-error string, serialises calls per handle.
#include "../../include/plda_b200.h"

#include <string.h>

#include <algorithm>

#include "engine.h"

struct plda_handle_s {
  pb::PldaEngine eng;
  explicit plda_handle_s(int dev) : eng(dev) {}
};
struct lda_handle_s {
  pb::LdaEngine eng;
  explicit lda_handle_s(int dev) : eng(dev) {}
};

namespace {

thread_local std::string g_last_error;

template <typename F>
int guarded(F&& f) {
  try {
    f();
    return PLDA_OK;
  } catch (const pb::Error& e) {
    g_last_error = e.what();
    // a sticky CUDA error (e.g. a trapped kernel) must not be silently reused
    cudaGetLastError();
    return e.code;
  } catch (const std::bad_alloc&) {
    g_last_error = "out of host memory";
    return PLDA_E_INTERNAL;
  } catch (const std::exception& e) {
    g_last_error = e.what();
    return PLDA_E_INTERNAL;
  } catch (...) {
    g_last_error = "unknown error";
    return PLDA_E_INTERNAL;
  }
}

template <typename H, typename F>
int with_handle(H h, F&& f) {
  if (h == nullptr) {
    g_last_error = "null handle";
    return PLDA_E_INVALID;
  }
  return guarded([&] {
    std::lock_guard<std::mutex> lock(h->eng.mu);
    PB_CUDA(cudaSetDevice(h->eng.ctx.device));
    f(h->eng);
  });
}

}  // namespace

extern "C" {

const char* plda_last_error(void) { return g_last_error.c_str(); }
const char* plda_version(void) { return "plda_b200 0.1.0 (sm_100a)"; }

int plda_create(int device, plda_handle_t* out) {
  if (out == nullptr) { g_last_error = "null output pointer"; return PLDA_E_INVALID; }
  *out = nullptr;
  return guarded([&] { *out = new plda_handle_s(device); });
}
int plda_destroy(plda_handle_t h) {
  if (h == nullptr) return PLDA_OK;
  return guarded([&] { cudaSetDevice(h->eng.ctx.device); delete h; });
}
int plda_set_precision(plda_handle_t h, int precision) {
  return with_handle(h, [&](pb::PldaEngine& e) {
    PB_CHECK(precision == PLDA_PREC_BF16X3 || precision == PLDA_PREC_FP64, pb::kInvalidArg, "unknown precision");
    e.precision = precision;
  });
}
int plda_set_stream(plda_handle_t h, void* cuda_stream) {
  return with_handle(h, [&](pb::PldaEngine& e) {
    e.ctx.sync();
    if (cuda_stream == nullptr) {
      if (!e.ctx.owns_stream) {
        PB_CUDA(cudaStreamCreateWithFlags(&e.ctx.stream, cudaStreamNonBlocking));
        e.ctx.owns_stream = true;
      }
    } else {
      if (e.ctx.owns_stream && e.ctx.stream) cudaStreamDestroy(e.ctx.stream);
      e.ctx.stream = static_cast<cudaStream_t>(cuda_stream);
      e.ctx.owns_stream = false;
    }
  });
}
int plda_stream_wait(plda_handle_t h, void* producer_stream) {
  return with_handle(h, [&](pb::PldaEngine& e) {
    cudaStream_t ps = static_cast<cudaStream_t>(producer_stream);
    if (ps == e.ctx.stream) return;
    cudaEvent_t ev;
    PB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    cudaError_t err = cudaEventRecord(ev, ps);
    if (err == cudaSuccess) err = cudaStreamWaitEvent(e.ctx.stream, ev, 0);
    cudaEventDestroy(ev);      // released once the wait has been satisfied
    PB_CUDA(err);
  });
}
int plda_synchronize(plda_handle_t h) { return with_handle(h, [&](pb::PldaEngine& e) { e.ctx.sync(); }); }
int plda_launch_count(plda_handle_t h, int64_t* out) {
  return with_handle(h, [&](pb::PldaEngine& e) { *out = e.ctx.launches.load(); });
}

int plda_profile_gemm(plda_handle_t h, int enable) {
  return with_handle(h, [&](pb::PldaEngine& e) {
    e.ctx.profile_reset();
    e.ctx.profile_gemm = enable != 0;
  });
}
int plda_profile_collect(plda_handle_t h, double* total_ms, int64_t* count) {
  return with_handle(h, [&](pb::PldaEngine& e) { e.ctx.profile_collect(total_ms, count); });
}

int plda_set_allreduce(plda_handle_t h, plda_allreduce_fn fn, void* user, double* scratch_dev, int64_t capacity) {
  return with_handle(h, [&](pb::PldaEngine& e) {
    PB_CHECK(fn == nullptr || (scratch_dev != nullptr && capacity > 0), pb::kInvalidArg,
             "set_allreduce: a device scratch buffer is required");
    e.reduce_fn = fn;
    e.reduce_user = user;
    e.reduce_scratch = scratch_dev;
    e.reduce_capacity = capacity;
  });
}

int plda_fit(plda_handle_t h, const void* x, int64_t n, int64_t d, int64_t ldx, int dtype, int loc,
             const uint64_t* labels, int iters) {
  return with_handle(h, [&](pb::PldaEngine& e) { e.fit(x, n, d, ldx, dtype, loc, labels, iters); });
}
int plda_fit_labels(plda_handle_t h, const void* x, int64_t n, int64_t d, int64_t ldx, int dtype, int loc,
                    const uint64_t* labels, int labels_loc, int iters) {
  return with_handle(h, [&](pb::PldaEngine& e) {
    PB_CHECK(labels_loc == PLDA_HOST || labels_loc == PLDA_DEVICE, pb::kInvalidArg, "fit: bad label location");
    e.fit(x, n, d, ldx, dtype, loc, labels, iters, labels_loc);
  });
}
int plda_fit_timings(plda_handle_t h, double out[5]) {
  return with_handle(h, [&](pb::PldaEngine& e) { memcpy(out, e.fit_ms, sizeof(e.fit_ms)); });
}
int plda_dim(plda_handle_t h, int64_t* d) {
  return with_handle(h, [&](pb::PldaEngine& e) { *d = e.model.ready ? e.model.d : 0; });
}
int plda_get_model(plda_handle_t h, double* mean, double* transform, double* psi) {
  return with_handle(h, [&](pb::PldaEngine& e) { e.get_model(mean, transform, psi); });
}
int plda_set_model(plda_handle_t h, int64_t d, const double* mean, const double* transform, const double* psi) {
  return with_handle(h, [&](pb::PldaEngine& e) { e.set_model(d, mean, transform, psi); });
}
int plda_get_covariances(plda_handle_t h, double* within, double* between) {
  return with_handle(h, [&](pb::PldaEngine& e) { e.get_covariances(within, between); });
}
int plda_smooth(plda_handle_t h, double factor) {
  return with_handle(h, [&](pb::PldaEngine& e) { e.smooth(factor); });
}

int plda_transform(plda_handle_t h, const void* x, int64_t n, int64_t d, int64_t ldx, int dtype, int loc,
                   const uint64_t* labels, int64_t targetdim, uint64_t* out_labels, int64_t* out_counts,
                   double* out_vecs, int64_t* n_out) {
  return with_handle(h, [&](pb::PldaEngine& e) {
    PB_CHECK(out_labels && out_counts && out_vecs && n_out, pb::kInvalidArg, "transform: null output");
    e.transform_grouped(x, n, d, ldx, dtype, loc, labels, targetdim, out_labels, out_counts, out_vecs, n_out);
  });
}
int plda_transform_rows(plda_handle_t h, const void* x, int64_t n, int64_t d, int64_t ldx, int dtype, int loc,
                        const int32_t* counts, int32_t const_count, int64_t targetdim, void* out, int64_t ldo,
                        int out_dtype, int out_loc) {
  return with_handle(h, [&](pb::PldaEngine& e) {
    PB_CHECK(out != nullptr || n == 0, pb::kInvalidArg, "transform_rows: null output");
    e.transform_rows(x, n, d, ldx, dtype, loc, counts, const_count, targetdim, out, ldo, out_dtype, out_loc);
  });
}

int plda_score_pair(plda_handle_t h, uint64_t model_id, int64_t n_enrol, const double* enrol, const double* test,
                    int64_t dim, float* out) {
  return with_handle(h, [&](pb::PldaEngine& e) {
    PB_CHECK(enrol && test && out, pb::kInvalidArg, "score: null pointer");
    e.score_pair(model_id, n_enrol, enrol, test, dim, out);
  });
}
int plda_score_grid(plda_handle_t h, const void* enrol, int64_t ne, int64_t ld_enrol, const int32_t* enrol_counts,
                    const uint64_t* enrol_ids, const void* test, int64_t nt, int64_t ld_test, int64_t dim, int dtype,
                    int loc, float* out, int64_t ldo, int out_loc) {
  return with_handle(h, [&](pb::PldaEngine& e) {
    e.score_grid(enrol, ne, ld_enrol, enrol_counts, enrol_ids, test, nt, ld_test, dim, dtype, loc, out, ldo, out_loc);
  });
}
int plda_norm(plda_handle_t h, const void* bkg, int64_t m, int64_t d, int64_t ldb, int dtype, int loc,
              const uint64_t* enrol_ids, const void* enrol, int64_t ne, int64_t ld_enrol, int64_t dim, int enrol_dtype,
              int enrol_loc, int64_t numutts, uint64_t seed) {
  return with_handle(h, [&](pb::PldaEngine& e) {
    e.norm(bkg, m, d, ldb, dtype, loc, enrol_ids, enrol, ne, ld_enrol, dim, enrol_dtype, enrol_loc, numutts, seed);
  });
}
int plda_norm_rows(plda_handle_t h, const void* bkg, int64_t m, int64_t d, int64_t ldb, int dtype, int loc,
                   const void* enrol, int64_t ne, int64_t ld_enrol, int64_t dim, int enrol_dtype, int enrol_loc,
                   int64_t numutts, uint64_t seed, double* mean_out, double* std_out, int out_loc) {
  return with_handle(h, [&](pb::PldaEngine& e) {
    PB_CHECK(ne == 0 || (mean_out != nullptr && std_out != nullptr), pb::kInvalidArg, "norm_rows: null output");
    PB_CHECK(out_loc == PLDA_HOST || out_loc == PLDA_DEVICE, pb::kInvalidArg, "norm_rows: bad output location");
    e.norm(bkg, m, d, ldb, dtype, loc, nullptr, enrol, ne, ld_enrol, dim, enrol_dtype, enrol_loc, numutts, seed, mean_out,
           std_out, out_loc);
  });
}
int plda_norm_selection(int64_t m, int64_t numutts, uint64_t seed, int32_t* rows_out) {
  return guarded([&] {
    PB_CHECK(m > 0 && m < (1ll << 31) && numutts >= 0 && numutts <= m && rows_out != nullptr, pb::kInvalidArg,
             "norm_selection: bad arguments");
    pb::norm_selection(m, numutts == 0 ? m : numutts, seed, rows_out);
  });
}
int plda_score_grid_z(plda_handle_t h, const void* enrol, int64_t ne, int64_t ld_enrol, const int32_t* enrol_counts,
                      const void* test, int64_t nt, int64_t ld_test, int64_t dim, int dtype, int loc, float* out,
                      int64_t ldo, int out_loc, const double* zmean, const double* zstd, int z_loc) {
  return with_handle(h, [&](pb::PldaEngine& e) {
    PB_CHECK((zmean == nullptr) == (zstd == nullptr), pb::kInvalidArg, "score_grid_z: zmean and zstd go together");
    e.score_grid(enrol, ne, ld_enrol, enrol_counts, nullptr, test, nt, ld_test, dim, dtype, loc, out, ldo, out_loc, zmean,
                 zstd, z_loc);
  });
}
int plda_score_trials(plda_handle_t h, const void* enrol, int64_t ne, int64_t ld_enrol, const int32_t* enrol_counts,
                      const uint64_t* enrol_ids, const void* test, int64_t nt, int64_t ld_test, int64_t dim, int dtype,
                      int loc, const int32_t* trial_enrol, const int32_t* trial_test, int64_t n_trials, int idx_loc,
                      float* out, int out_loc, const double* zmean, const double* zstd, int z_loc, int mode) {
  return with_handle(h, [&](pb::PldaEngine& e) {
    PB_CHECK((zmean == nullptr) == (zstd == nullptr), pb::kInvalidArg, "score_trials: zmean and zstd go together");
    e.score_trials(enrol, ne, ld_enrol, enrol_counts, enrol_ids, test, nt, ld_test, dim, dtype, loc, trial_enrol,
                   trial_test, n_trials, idx_loc, out, out_loc, zmean, zstd, z_loc, mode);
  });
}
int plda_score_hist(plda_handle_t h, const void* enrol, int64_t ne, int64_t ld_enrol, int32_t enrol_count,
                    const void* test, int64_t nt, int64_t ld_test, int64_t dim, int dtype, int loc,
                    const int32_t* enrol_spk, const int32_t* test_spk, int spk_loc, double lo, double hi, int nbins,
                    double theta_lo, const double* zmean, const double* zstd, int z_loc, uint64_t* hist_target,
                    uint64_t* hist_nontarget, uint64_t* below, int out_loc) {
  return with_handle(h, [&](pb::PldaEngine& e) {
    PB_CHECK((zmean == nullptr) == (zstd == nullptr), pb::kInvalidArg, "score_hist: zmean and zstd go together");
    e.score_hist(enrol, ne, ld_enrol, enrol_count, test, nt, ld_test, dim, dtype, loc, enrol_spk, test_spk, spk_loc, lo, hi,
                 nbins, theta_lo, zmean, zstd, z_loc, hist_target, hist_nontarget, below, out_loc);
  });
}
int plda_znorm_size(plda_handle_t h, int64_t* n) {
  return with_handle(h, [&](pb::PldaEngine& e) { *n = static_cast<int64_t>(e.znorm.size()); });
}
int plda_znorm_get(plda_handle_t h, uint64_t* ids, double* mean, double* stdv, int64_t capacity, int64_t* n) {
  return with_handle(h, [&](pb::PldaEngine& e) {
    int64_t i = 0;
    for (const auto& kv : e.znorm) {
      if (i >= capacity) break;
      ids[i] = kv.first;
      mean[i] = kv.second.first;
      stdv[i] = kv.second.second;
      ++i;
    }
    *n = i;
  });
}
int plda_znorm_set(plda_handle_t h, const uint64_t* ids, const double* mean, const double* stdv, int64_t n) {
  return with_handle(h, [&](pb::PldaEngine& e) {
    PB_CHECK(n == 0 || (ids && mean && stdv), pb::kInvalidArg, "znorm_set: null pointer");
    for (int64_t i = 0; i < n; ++i) e.znorm.emplace(ids[i], std::make_pair(mean[i], stdv[i]));
  });
}
int plda_znorm_clear(plda_handle_t h) { return with_handle(h, [&](pb::PldaEngine& e) { e.znorm.clear(); }); }

int plda_shard_open(plda_handle_t h, int world, int rank, const int64_t* bounds, int64_t dim,
                    unsigned char* ipc_handle_out, void** region_out) {
  return with_handle(h, [&](pb::PldaEngine& e) { e.shard_open(world, rank, bounds, dim, ipc_handle_out, region_out); });
}
int plda_shard_open_ragged(plda_handle_t h, int world, int rank, const int64_t* bounds, int64_t dim, int max_groups,
                           unsigned char* ipc_handle_out, void** region_out) {
  return with_handle(h, [&](pb::PldaEngine& e) {
    e.shard_open(world, rank, bounds, dim, ipc_handle_out, region_out, max_groups);
  });
}
int plda_shard_step_ragged(plda_handle_t h, const void* test_shard, int64_t nt_local, int64_t ld_test, const void* enrol,
                           int64_t ne, int64_t ld_enrol, const int32_t* enrol_counts, const int32_t* group_counts,
                           int n_groups, const uint64_t* enrol_ids, int dtype, float* out, int64_t ldo) {
  return with_handle(h, [&](pb::PldaEngine& e) {
    e.shard_step_ragged(test_shard, nt_local, ld_test, enrol, ne, ld_enrol, enrol_counts, group_counts, n_groups,
                        enrol_ids, dtype, out, ldo);
  });
}
int plda_shard_connect(plda_handle_t h, int peer_rank, const unsigned char* ipc_handle, void* same_process_region) {
  return with_handle(h, [&](pb::PldaEngine& e) { e.shard_connect(peer_rank, ipc_handle, same_process_region); });
}
int plda_shard_push(plda_handle_t h, const void* test_shard, int64_t nt_local, int64_t ld_test, int dtype,
                    int enrol_count) {
  return with_handle(h, [&](pb::PldaEngine& e) { e.shard_push(test_shard, nt_local, ld_test, dtype, enrol_count); });
}
int plda_shard_score(plda_handle_t h, const void* enrol, int64_t ne, int64_t ld_enrol, int enrol_count,
                     const uint64_t* enrol_ids, int dtype, float* out, int64_t ldo) {
  return with_handle(h, [&](pb::PldaEngine& e) {
    e.shard_score(enrol, ne, ld_enrol, enrol_count, enrol_ids, dtype, out, ldo);
  });
}
int plda_shard_step(plda_handle_t h, const void* test_shard, int64_t nt_local, int64_t ld_test, const void* enrol,
                    int64_t ne, int64_t ld_enrol, int enrol_count, const uint64_t* enrol_ids, int dtype, float* out,
                    int64_t ldo) {
  return with_handle(h, [&](pb::PldaEngine& e) {
    e.shard_step(test_shard, nt_local, ld_test, enrol, ne, ld_enrol, enrol_count, enrol_ids, dtype, out, ldo);
  });
}
int plda_shard_status(plda_handle_t h, int64_t* epoch, int64_t* timeouts) {
  return with_handle(h, [&](pb::PldaEngine& e) { e.shard_status(epoch, timeouts); });
}
int plda_shard_close(plda_handle_t h) {
  return with_handle(h, [&](pb::PldaEngine& e) { e.shard_close(); });
}

int plda_dvector_pool(plda_handle_t h, const void* frames, int64_t n_frames, int64_t d, int64_t ld, int dtype, int loc,
                      const int64_t* offsets, int64_t n_utts, int mode, int l2norm, double* out, int64_t ldo,
                      int out_loc) {
  return with_handle(h, [&](pb::PldaEngine& e) {
    PB_CHECK(dtype == PLDA_F64 || dtype == PLDA_F32, pb::kInvalidArg, "dtype must be PLDA_F64 or PLDA_F32");
    PB_CHECK((loc == 0 || loc == 1) && (out_loc == 0 || out_loc == 1), pb::kInvalidArg, "bad loc");
    PB_CHECK(n_frames >= 0 && (frames != nullptr || n_frames == 0) && (out != nullptr || n_utts == 0), pb::kInvalidArg,
             "dvector_pool: null pointer");
    PB_CHECK(ld >= d && ldo >= d, pb::kInvalidArg, "dvector_pool: pitch smaller than the dimension");
    if (n_utts == 0) return;
    const size_t es = dtype == PLDA_F32 ? 4 : 8;
    pb::DevBuf<uint8_t> stage;
    const void* fd = frames;
    int64_t ldf = ld;
    if (loc == PLDA_HOST) {
      stage.alloc(static_cast<size_t>(n_frames > 0 ? n_frames : 1) * d * es);
      if (n_frames > 0)
        PB_CUDA(cudaMemcpy2DAsync(stage.get(), d * es, frames, ld * es, d * es, n_frames, cudaMemcpyHostToDevice, e.ctx.stream));
      fd = stage.get();
      ldf = d;
    }
    if (out_loc == PLDA_DEVICE) {
      pb::dvector_pool(e.ctx, fd, dtype == PLDA_F32, n_frames, d, ldf, offsets, n_utts, mode, l2norm != 0, out, ldo);
      return;
    }
    pb::DevBuf<double> dout(static_cast<size_t>(n_utts) * d);
    pb::dvector_pool(e.ctx, fd, dtype == PLDA_F32, n_frames, d, ldf, offsets, n_utts, mode, l2norm != 0, dout.get(), d);
    PB_CUDA(cudaMemcpy2DAsync(out, ldo * sizeof(double), dout.get(), d * sizeof(double), d * sizeof(double), n_utts,
                              cudaMemcpyDeviceToHost, e.ctx.stream));
    e.ctx.sync();
  });
}

int lda_create(int device, lda_handle_t* out) {
  if (out == nullptr) { g_last_error = "null output pointer"; return PLDA_E_INVALID; }
  *out = nullptr;
  return guarded([&] { *out = new lda_handle_s(device); });
}
int lda_destroy(lda_handle_t h) {
  if (h == nullptr) return PLDA_OK;
  return guarded([&] { cudaSetDevice(h->eng.ctx.device); delete h; });
}
int lda_set_precision(lda_handle_t h, int precision) {
  return with_handle(h, [&](pb::LdaEngine& e) {
    PB_CHECK(precision == PLDA_PREC_BF16X3 || precision == PLDA_PREC_FP64, pb::kInvalidArg, "unknown precision");
    e.precision = precision;
  });
}
int lda_launch_count(lda_handle_t h, int64_t* out) {
  return with_handle(h, [&](pb::LdaEngine& e) { *out = e.ctx.launches.load(); });
}
int lda_stream_wait(lda_handle_t h, void* producer_stream) {
  return with_handle(h, [&](pb::LdaEngine& e) {
    cudaStream_t ps = static_cast<cudaStream_t>(producer_stream);
    if (ps == e.ctx.stream) return;
    cudaEvent_t ev;
    PB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    cudaError_t err = cudaEventRecord(ev, ps);
    if (err == cudaSuccess) err = cudaStreamWaitEvent(e.ctx.stream, ev, 0);
    cudaEventDestroy(ev);
    PB_CUDA(err);
  });
}
int lda_synchronize(lda_handle_t h) { return with_handle(h, [&](pb::LdaEngine& e) { e.ctx.sync(); }); }
int lda_fit_svd(lda_handle_t h, const void* x, int64_t n, int64_t d, int64_t ldx, int dtype, int loc,
                const int64_t* labels, const double* priors, int64_t n_priors) {
  return with_handle(h, [&](pb::LdaEngine& e) { e.fit_svd(x, n, d, ldx, dtype, loc, labels, priors, n_priors); });
}
int lda_fit_lsqr(lda_handle_t h, const void* x, int64_t n, int64_t d, int64_t ldx, int dtype, int loc,
                 const int64_t* labels, const double* priors, int64_t n_priors) {
  return with_handle(h, [&](pb::LdaEngine& e) { e.fit_lsqr(x, n, d, ldx, dtype, loc, labels, priors, n_priors); });
}
int lda_fit_eigen(lda_handle_t h, const void* x, int64_t n, int64_t d, int64_t ldx, int dtype, int loc,
                  const int64_t* labels, const double* priors, int64_t n_priors) {
  return with_handle(h, [&](pb::LdaEngine& e) { e.fit_eigen(x, n, d, ldx, dtype, loc, labels, priors, n_priors); });
}
int lda_get_eigenvalues(lda_handle_t h, double* evals, int64_t capacity, int64_t* n) {
  return with_handle(h, [&](pb::LdaEngine& e) {
    PB_CHECK(e.ready, pb::kNotFitted, "This LDA instance is not fitted yet");
    const int64_t m = std::min<int64_t>(capacity, static_cast<int64_t>(e.h_evals.size()));
    if (evals && m > 0) memcpy(evals, e.h_evals.data(), m * sizeof(double));
    if (n) *n = static_cast<int64_t>(e.h_evals.size());
  });
}
int lda_class_stats(lda_handle_t h, const void* x, int64_t n, int64_t d, int64_t ldx, int dtype, int loc,
                    const int64_t* labels, int64_t* k) {
  return with_handle(h, [&](pb::LdaEngine& e) {
    e.local_class_stats(x, n, d, ldx, dtype, loc, labels);
    if (k) *k = e.shard_stats.k;
  });
}
int lda_get_class_stats(lda_handle_t h, double* sw, double* means, int64_t* counts, int64_t* classes) {
  return with_handle(h, [&](pb::LdaEngine& e) {
    const auto& cs = e.shard_stats;
    PB_CHECK(cs.k > 0, pb::kNotFitted, "lda_get_class_stats: call lda_class_stats first");
    if (sw) memcpy(sw, cs.sw.data(), cs.sw.size() * sizeof(double));
    if (means) memcpy(means, cs.means.data(), cs.means.size() * sizeof(double));
    if (counts) for (int64_t c = 0; c < cs.k; ++c) counts[c] = cs.counts[c];
    if (classes) memcpy(classes, cs.classes.data(), cs.k * sizeof(int64_t));
  });
}
int lda_fit_from_stats(lda_handle_t h, int solver, int64_t n, int64_t k, int64_t d, const double* sw,
                       const double* means, const int64_t* counts, const int64_t* classes, const double* priors,
                       int64_t n_priors) {
  return with_handle(h, [&](pb::LdaEngine& e) {
    PB_CHECK(solver >= 0 && solver <= 2, pb::kInvalidArg, "lda_fit_from_stats: solver must be 0 (svd), 1 (lsqr) or 2 (eigen)");
    e.fit_from_stats(solver, n, k, d, sw, means, counts, classes, priors, n_priors);
  });
}
int lda_get_svd(lda_handle_t h, int64_t* rank, double* xbar, double* scalings) {
  return with_handle(h, [&](pb::LdaEngine& e) {
    PB_CHECK(e.ready, pb::kNotFitted, "This LDA instance is not fitted yet");
    if (rank) *rank = e.rank;
    if (xbar && e.rank > 0) memcpy(xbar, e.h_xbar.data(), e.h_xbar.size() * sizeof(double));
    if (scalings && e.rank > 0) memcpy(scalings, e.h_scalings.data(), e.h_scalings.size() * sizeof(double));
  });
}
int lda_transform(lda_handle_t h, const void* x, int64_t nt, int64_t d, int64_t ldx, int dtype, int loc,
                  int64_t n_components, float* out, int64_t ldo, int out_loc) {
  return with_handle(h, [&](pb::LdaEngine& e) { e.transform(x, nt, d, ldx, dtype, loc, n_components, out, ldo, out_loc); });
}
int lda_num_classes(lda_handle_t h, int64_t* k, int64_t* d) {
  return with_handle(h, [&](pb::LdaEngine& e) { *k = e.ready ? e.k : 0; *d = e.ready ? e.d : 0; });
}
int lda_get_coef(lda_handle_t h, double* coef, double* intercept, int64_t* classes) {
  return with_handle(h, [&](pb::LdaEngine& e) {
    PB_CHECK(e.ready, pb::kNotFitted, "This LDA instance is not fitted yet");
    if (coef) memcpy(coef, e.h_coef.data(), e.h_coef.size() * sizeof(double));
    if (intercept) memcpy(intercept, e.h_intercept.data(), e.h_intercept.size() * sizeof(double));
    if (classes) memcpy(classes, e.h_classes.data(), e.h_classes.size() * sizeof(int64_t));
  });
}
int lda_set_coef(lda_handle_t h, int64_t k, int64_t d, const double* coef, const double* intercept) {
  return with_handle(h, [&](pb::LdaEngine& e) { e.set_coef(k, d, coef, intercept); });
}
int lda_predict(lda_handle_t h, const void* x, int64_t nt, int64_t d, int64_t ldx, int dtype, int loc, int log_proba,
                float* out, int64_t ldo, int out_loc) {
  return with_handle(h, [&](pb::LdaEngine& e) { e.predict(x, nt, d, ldx, dtype, loc, log_proba, out, ldo, out_loc); });
}

int plda_device_malloc(int device, size_t bytes, void** out) {
  return guarded([&] {
    PB_CUDA(cudaSetDevice(device));
    PB_CUDA(cudaMalloc(out, bytes == 0 ? 1 : bytes));
  });
}
int plda_device_free(int device, void* p) {
  return guarded([&] {
    PB_CUDA(cudaSetDevice(device));
    PB_CUDA(cudaFree(p));
  });
}
int plda_host_malloc_pinned(size_t bytes, void** out) {
  return guarded([&] { PB_CUDA(cudaMallocHost(out, bytes == 0 ? 1 : bytes)); });
}
int plda_host_free_pinned(void* p) { return guarded([&] { PB_CUDA(cudaFreeHost(p)); }); }
int plda_memcpy(void* dst, const void* src, size_t bytes, int kind) {
  return guarded([&] {
    const cudaMemcpyKind k = kind == 0 ? cudaMemcpyHostToDevice : kind == 1 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    PB_CUDA(cudaMemcpy(dst, src, bytes, k));
  });
}

int plda_test_gemm(plda_handle_t h, const double* a, const double* b, int64_t m, int64_t n, int64_t k, int ksplit,
                   float* out) {
  return with_handle(h, [&](pb::PldaEngine& e) { e.test_gemm(a, b, m, n, k, ksplit, out); });
}
int plda_debug_counters(plda_handle_t h, int64_t* out, int n) {
  return with_handle(h, [&](pb::PldaEngine& e) {
    e.ctx.sync();
    PB_CHECK(e.ctx.gemm_dbg.size() >= 32 && n <= 32, pb::kInvalidArg, "debug counters need PLDA_B200_DBG=1");
    PB_CUDA(cudaMemcpy(out, e.ctx.gemm_dbg.get(), n * sizeof(long long), cudaMemcpyDeviceToHost));
  });
}
int plda_test_scatter(plda_handle_t h, const void* x, int64_t n, int64_t d, int dtype, const uint64_t* labels,
                      int scale_by_count, double* scatter_out, double* means_out, int64_t means_capacity,
                      int64_t* k_out) {
  return with_handle(h, [&](pb::PldaEngine& e) {
    PB_CHECK(x && labels && scatter_out, pb::kInvalidArg, "test_scatter: null pointer");
    e.test_scatter(x, n, d, dtype, labels, scale_by_count, scatter_out, means_out, means_capacity, k_out);
  });
}
int plda_test_linalg(plda_handle_t h, int op, const double* a, int64_t d, double* out, double* out2) {
  return with_handle(h, [&](pb::PldaEngine& e) { e.test_linalg(op, a, d, out, out2); });
}

}  // extern "C"
