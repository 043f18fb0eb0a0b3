// Host-side runtime shared by all translation units: device buffers, the per-handle
// context (device, stream, SM count, launch counter), TMA descriptor encoding and the
// launch wrappers of every kernel family.  No torch types anywhere: the boundary above
// this is the C ABI in include/plda_b200.h.
#pragma once

#include <stdint.h>

#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

namespace pb {

// ------------------------------------------------------------------------- //
// RAII device buffer
// ------------------------------------------------------------------------- //
template <typename T>
class DevBuf {
 public:
  DevBuf() = default;
  explicit DevBuf(size_t n) { alloc(n); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p_(o.p_), n_(o.n_) { o.p_ = nullptr; o.n_ = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { release(); p_ = o.p_; n_ = o.n_; o.p_ = nullptr; o.n_ = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  void alloc(size_t n) {
    release();
    if (n == 0) n = 1;
    PB_CUDA(cudaMalloc(reinterpret_cast<void**>(&p_), n * sizeof(T)));
    n_ = n;
  }
  // grow-only (contents not preserved)
  void reserve(size_t n) { if (n > n_) alloc(n); }
  void release() { if (p_) { cudaFree(p_); p_ = nullptr; n_ = 0; } }
  T* get() const { return p_; }
  size_t size() const { return n_; }
  void zero(cudaStream_t s) { PB_CUDA(cudaMemsetAsync(p_, 0, n_ * sizeof(T), s)); }
 private:
  T* p_ = nullptr;
  size_t n_ = 0;
};

// ------------------------------------------------------------------------- //
// Context
// ------------------------------------------------------------------------- //
struct Context {
  int device = 0;
  int num_sms = 148;
  cudaStream_t stream = nullptr;     // launching stream (owned unless external)
  bool owns_stream = true;
  std::atomic<long long> launches{0};   // number of OUR kernels launched (bench "gpu_launches")
  int epi_mode = 0;                  // debug (env PLDA_B200_EPI): 0 default (TMA store), 1 "direct" register stores, 2 "skip", 3 "lsu"
  bool epi_hybrid = false;           // env PLDA_B200_EPI=hybrid: every other chunk of the score-grid epilogue as register sector stores
  bool epi_sector = false;           // env PLDA_B200_EPI=sector: register-direct sector stores in the score-grid epilogue
  int dbg_skip_a = 0;                // debug (env PLDA_B200_DBGSKIPA=1): timing experiment, the TMA producer loads the A
                                     // tiles only for the first item of a CTA (WRONG results) -> upper bound of an
                                     // A-resident schedule
  bool k_tail_boxes = true;          // env PLDA_B200_KTAIL=0 disables the narrow K-tail boxes
  DevBuf<long long> gemm_dbg;        // env PLDA_B200_DBG=1: stall counters written by CTA 0/1 of the last GEMM launch
  bool gemm_two_cta = true;          // env PLDA_B200_GEMM=1cta forces the single-CTA (cta_group::1) kernel
  bool gemm_ts = false;              // score grid with the enrol operand in TMEM (gemm_ts.cu); env PLDA_B200_TS=0/1
  // optional per-launch timing of the tensor-core GEMM (CUDA events on the launching stream); bench roofline
  // programmatic dependent launch: set by a producer kernel that executes griddepcontrol.launch_dependents; the
  // next tensor GEMM launch then carries the programmatic-serialization attribute, so its prologue (barrier init,
  // TMEM allocation, descriptor prefetch) overlaps the producer's tail.  env PLDA_B200_PDL=0 disables it.
  bool pdl_enabled = true;
  bool pdl_pending = false;
  bool profile_gemm = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> gemm_events;
  void profile_reset();
  // synchronises; returns total ms and number of launches recorded since the last reset
  void profile_collect(double* total_ms, int64_t* count);
  explicit Context(int dev);
  ~Context();
  void sync() { PB_CUDA(cudaStreamSynchronize(stream)); }
  void count_launch(int n = 1) { launches.fetch_add(n, std::memory_order_relaxed); }
};

// Encode a 2-D row-major tensor map.  inner = contiguous dimension (elements).
// swizzle_bytes (128 / 64 / 32) must equal the box inner extent in bytes.
enum class TmaType { BF16, F32 };
void encode_tmap_2d(CUtensorMap* out, TmaType type, const void* base, uint64_t inner, uint64_t outer,
                    uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer, uint32_t swizzle_bytes);

// ------------------------------------------------------------------------- //
// Split-bf16 operand: a logical [rows x k] matrix stored K-major as two bf16
// planes (hi, lo), each [rows x ld] with ld % 8 == 0 and columns k..ld-1 zero.
// ------------------------------------------------------------------------- //
struct SplitOperand {
  const __nv_bfloat16* hi = nullptr;
  const __nv_bfloat16* lo = nullptr;
  int64_t rows = 0;
  int64_t k = 0;      // logical reduction length
  int64_t ld = 0;     // elements per row in memory (>= round_up(k,16))
};

struct SplitBuf {           // owning version
  DevBuf<__nv_bfloat16> hi, lo;
  int64_t rows = 0, k = 0, ld = 0;
  void reserve(int64_t r, int64_t kk, int64_t pitch = 0) {
    rows = r; k = kk; ld = pitch > 0 ? pitch : round_up(kk, 16);
    size_t n = static_cast<size_t>(r > 0 ? r : 1) * ld;
    hi.reserve(n); lo.reserve(n);
  }
  SplitOperand view() const { return SplitOperand{hi.get(), lo.get(), rows, k, ld}; }
};

// Epilogue of the tensor-core GEMM:  v = acc[m,n]
//   + row_add[m]                      (if row_add)
//   + col_add[grp[m]*col_ld + n]      (if col_add; grp==nullptr -> group 0)
//   v = (v - zmean[m]) * zinv[m]      (if zmean)
// then either stored to out[m*ldo+n] (fp32) and/or reduced per row into
// rsum[m] += v, rsq[m] += v*v (fp64 atomics; for z-norm moments / log-sum-exp passes).
struct GemmEpilogue {
  float* out = nullptr;
  int64_t ldo = 0;
  const float* row_add = nullptr;
  const float* col_add = nullptr;
  const int32_t* grp = nullptr;
  int64_t col_ld = 0;
  const float* zmean = nullptr;
  const float* zinv = nullptr;
  double* rsum = nullptr;
  double* rsq = nullptr;
  // online log-sum-exp per row over the columns (LDA): rmax/rsumexp are [M x n_tiles]
  float* lse_max = nullptr;
  float* lse_sum = nullptr;
  // moments sink (z-norm, nothing stored): per (row, column tile, epilogue half) the fp32 partials
  // {sum(v - p), sum((v - p)^2), p, count} with the pivot p = the first score of that slot -- combined in fp64 by
  // moments_reduce (Chan's update); [M][n_tiles][2] float4
  float4* mom = nullptr;
  // tail-histogram sink (EER at sizes where the grid is not materialised, nothing stored): a trial (m, n) is a
  // TARGET iff row_spk[m] == col_spk[n].  Targets are always binned into hist_t; non-targets are binned into hist_n
  // only when v >= theta_lo and otherwise just counted in *below (the bulk of the non-target mass never touches an
  // atomic).  bin = clamp(floor((v - hist_lo) * hist_scale), 0, nbins - 1)
  const int32_t* row_spk = nullptr;
  const int32_t* col_spk = nullptr;
  unsigned long long* hist_t = nullptr;
  unsigned long long* hist_n = nullptr;
  unsigned long long* below = nullptr;
  float hist_lo = 0.f, hist_scale = 0.f, theta_lo = 0.f;
  int nbins = 0;
};

// Sharded B operand (scoring grid): rows [bounds[r], bounds[r+1]) of B (and of the column-term row)
// are produced by rank r and pushed into this GPU's buffer over NVLink peer memory.  The GEMM does not wait for
// the whole exchange: its TMA producer polls `flags[r] >= epoch` (acquire, system scope) only before the first
// tile that touches rows of rank r, and the column tiles are visited starting with this rank's own rows.
struct GemmShard {
  const unsigned* flags = nullptr;   // [world] ready epochs in the LOCAL region (written by the peers)
  unsigned* err = nullptr;           // timeouts are counted here instead of hanging the GPU
  unsigned epoch = 0;
  int world = 0;
  int rank = 0;
  int bounds[17] = {0};
};

// waits (bounded, like the GEMM's own polls) until every rank's flag has reached shard.epoch -- for a rank that has
// no GEMM to launch in a step
void shard_wait_all(Context& ctx, const GemmShard& shard);

// C[M,N] = A[M,K] * B[N,K]^T with bf16x3 split operands on tcgen05 (fp32 TMEM accumulate).
// ksplit > 1: reduction axis split into ksplit chunks, partial tiles written to
// `partial` ([ksplit][Mpad][Npad] fp32, Mpad = round_up(M,128), Npad = round_up(N,4)) and
// the caller reduces them (reduce_partials_f64).
void gemm_bf16x3(Context& ctx, const SplitOperand& a, const SplitOperand& b, int64_t m, int64_t n, int64_t k,
                 const GemmEpilogue& epi, const GemmShard* shard = nullptr);
// Score-grid form with the A operand in tensor memory (gemm_ts.cu): k <= 256, plain store epilogue with row / uniform
// column / z-norm terms.  Returns false when the problem does not qualify (nothing launched).
bool gemm_ts_score(Context& ctx, const SplitOperand& a, const SplitOperand& b, int64_t m, int64_t n, int64_t k,
                   const GemmEpilogue& epi, const GemmShard* shard);
void gemm_bf16x3_splitk(Context& ctx, const SplitOperand& a, const SplitOperand& b, int64_t m, int64_t n, int64_t k,
                        int ksplit, float* partial);
int choose_ksplit(const Context& ctx, int64_t m, int64_t n, int64_t k);
// number of column tiles gemm_bf16x3 uses for n columns (slot count of the per-tile epilogue outputs)
int gemm_n_tiles(int64_t n);
// number of partial planes gemm_bf16x3_splitk actually writes for a requested ksplit
int effective_ksplit(const Context& ctx, int64_t m, int64_t n, int64_t k, int ksplit);
// out[m*ldo+n] = alpha * sum_s partial[s][m][n]  (+ beta * out) in fp64; optional symmetrise
void reduce_partials_f64(Context& ctx, const float* partial, int ksplit, int64_t m, int64_t n, double* out,
                         int64_t ldo, double alpha, bool symmetrise);

// Exact-mode / cross-check fp64 SIMT GEMM:  C = alpha * op(A) * op(B) + beta * C, row-major.
//   ta: A is [K x M] (read transposed) else [M x K];  tb: B is [N x K] (i.e. C = A*B^T) else [K x N].
void gemm_f64(Context& ctx, bool ta, bool tb, int64_t m, int64_t n, int64_t k, double alpha, const double* a,
              int64_t lda, const double* b, int64_t ldb, double beta, double* c, int64_t ldc);

// two independent products of the same shape / transposition / pitches in ONE launch (d x d algebra of the EM
// iteration: halves the launch count on a latency-bound chain)
void gemm_f64_pair(Context& ctx, bool ta, bool tb, int64_t m, int64_t n, int64_t k, double alpha, const double* a0,
                   const double* b0, double* c0, const double* a1, const double* b1, double* c1, int64_t lda,
                   int64_t ldb, double beta, int64_t ldc);

}  // namespace pb
