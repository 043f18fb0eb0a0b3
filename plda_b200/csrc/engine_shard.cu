// Sharded enrol x test score grid over NVLink peer memory (SURVEY 8e "scoring grid").
//
// The reference scores one pair per call (Plda::LogLikelihoodRatio via src/pldamodule.cpp:258-277); the grid
// shards by ENROL BLOCK across GPUs and needs the transformed test vectors of every rank.  Instead of an NCCL
// all-gather followed by the operand producer, the producer kernel itself writes each rank's test rows (split
// bf16 planes + column terms) into the operand buffer of EVERY rank through peer mappings and raises a ready flag;
// the tcgen05 GEMM consumes the local buffer and waits per column tile for the flag of the rank that owns the
// rows (gemm_tc.cu, GemmShard), starting with its own rows.  No NCCL, no host synchronisation per step.
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "engine.h"

namespace pb {
namespace {

constexpr size_t kFlagsOff = 0;       // unsigned[64]: ready epoch per source rank
constexpr size_t kErrOff = 256;       // unsigned: number of timed-out waits
constexpr size_t kCounterOff = 512;   // unsigned: block counter of the local push kernel
constexpr size_t kHeaderBytes = 1024;

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

PldaEngine::~PldaEngine() {
  try {
    cudaSetDevice(ctx.device);
    shard_close();
    if (copy_stream) {
      cudaStreamDestroy(copy_stream);
      for (int i = 0; i < 2; ++i) { cudaEventDestroy(ev_done[i]); cudaEventDestroy(ev_free[i]); }
    }
  } catch (...) {
  }
}

void PldaEngine::shard_open(int world, int rank, const int64_t* bounds, int64_t dim, unsigned char* ipc_handle_out,
                            void** region_out, int max_groups) {
  require_model();
  PB_CHECK(!shard.open, kInvalidArg, "shard_open: a session is already open on this handle");
  PB_CHECK(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, kInvalidArg, "shard_open: bad world/rank");
  PB_CHECK(bounds != nullptr && bounds[0] == 0, kInvalidArg, "shard_open: bounds must start at 0");
  for (int r = 0; r < world; ++r) PB_CHECK(bounds[r + 1] >= bounds[r], kInvalidArg, "shard_open: bounds must ascend");
  PB_CHECK(bounds[world] > 0 && bounds[world] < (1ll << 31), kInvalidArg, "shard_open: bad number of test rows");
  PB_CHECK(dim > 0 && dim <= model.d && dim <= 1024, kInvalidArg, "shard_open: dimension does not match the model");
  PB_CHECK(precision == 0, kInvalidArg, "shard_open: the sharded grid runs on the tensor path (precision bf16x3)");
  ShardSession s;
  s.world = world;
  s.rank = rank;
  s.bounds.assign(bounds, bounds + world + 1);
  s.nt_total = bounds[world];
  s.dim = dim;
  PB_CHECK(max_groups >= 0 && max_groups <= 8, kInvalidArg, "shard_open: at most 8 distinct enrol counts per session");
  s.max_groups = max_groups;
  // ragged steps carry two extra K columns per distinct enrol count inside the operand rows (prep.cu)
  s.ldk = round_up(dim + 2 * max_groups, 16);
  s.col_ld = round_up(s.nt_total, 32);
  const size_t plane = align_up(static_cast<size_t>(s.nt_total) * s.ldk * sizeof(__nv_bfloat16), 1024);
  const size_t colb = align_up(static_cast<size_t>(s.col_ld) * sizeof(float), 1024);
  s.off_lo = plane;
  s.off_col = 2 * plane;
  const size_t gen = 2 * plane + colb;
  s.off_gen[0] = kHeaderBytes;
  s.off_gen[1] = kHeaderBytes + gen;
  s.bytes = kHeaderBytes + 2 * gen;
  PB_CUDA(cudaMalloc(reinterpret_cast<void**>(&s.region), s.bytes));
  cudaError_t e = cudaMemsetAsync(s.region, 0, s.bytes, ctx.stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx.stream);
  if (e == cudaSuccess && ipc_handle_out != nullptr) {
    cudaIpcMemHandle_t hnd;
    static_assert(sizeof(hnd) == 64, "cudaIpcMemHandle_t is 64 bytes");
    e = cudaIpcGetMemHandle(&hnd, s.region);
    if (e == cudaSuccess) memcpy(ipc_handle_out, &hnd, sizeof(hnd));
  }
  if (e != cudaSuccess) {
    cudaFree(s.region);
    throw Error(kCudaError, std::string("shard_open failed: ") + cudaGetErrorString(e));
  }
  s.peer.assign(world, nullptr);
  s.peer_ipc.assign(world, false);
  s.peer[rank] = s.region;
  s.open = true;
  shard = std::move(s);
  if (region_out) *region_out = shard.region;
}

void PldaEngine::shard_connect(int peer_rank, const unsigned char* ipc_handle, void* same_process_region) {
  PB_CHECK(shard.open, kInvalidArg, "shard_connect: no open session");
  PB_CHECK(peer_rank >= 0 && peer_rank < shard.world, kInvalidArg, "shard_connect: bad peer rank");
  if (peer_rank == shard.rank) return;
  PB_CHECK(shard.peer[peer_rank] == nullptr, kInvalidArg, "shard_connect: peer already connected");
  if (same_process_region != nullptr) {
    // a peer handle of the same process (tests: several ranks on one GPU; multi-GPU single process)
    shard.peer[peer_rank] = static_cast<uint8_t*>(same_process_region);
    return;
  }
  PB_CHECK(ipc_handle != nullptr, kInvalidArg, "shard_connect: an IPC handle or a region pointer is required");
  cudaIpcMemHandle_t hnd;
  memcpy(&hnd, ipc_handle, sizeof(hnd));
  void* p = nullptr;
  PB_CUDA(cudaIpcOpenMemHandle(&p, hnd, cudaIpcMemLazyEnablePeerAccess));
  shard.peer[peer_rank] = static_cast<uint8_t*>(p);
  shard.peer_ipc[peer_rank] = true;
}

// Producer launch: this rank's test rows -> every region (+ ready flags); optionally the enrol-side operand in the
// same launch.  Advances the epoch.
void PldaEngine::shard_produce(const void* test_shard, int64_t nt_local, int64_t ld_test, const void* enrol, int64_t ne,
                               int64_t ld_enrol, int enrol_count, int dtype) {
  require_model();
  PB_CHECK(shard.open, kInvalidArg, "shard: no open session");
  for (int r = 0; r < shard.world; ++r) PB_CHECK(shard.peer[r] != nullptr, kInvalidArg, "shard: a peer is not connected");
  PB_CHECK(dtype == 0 || dtype == 1, kInvalidArg, "dtype must be PLDA_F64 or PLDA_F32");
  const int64_t row0 = shard.bounds[shard.rank];
  PB_CHECK(nt_local == shard.bounds[shard.rank + 1] - row0, kInvalidArg, "shard: shard size does not match the bounds");
  PB_CHECK(nt_local == 0 || (test_shard != nullptr && ld_test >= shard.dim), kInvalidArg, "shard: bad test rows");
  PB_CHECK(enrol_count > 0, kInvalidArg, "shard: enrol count must be positive");
  PB_CHECK(ne == 0 || (enrol != nullptr && ld_enrol >= shard.dim), kInvalidArg, "shard: bad enrol rows");
  const double* consts = score_consts_for(enrol_count, shard.dim);
  shard.epoch += 1;
  shard.push_count = enrol_count;
  const size_t gen = shard.off_gen[shard.epoch & 1u];
  (void)gen;
  PrepDst dst;
  PrepSignal sig;
  shard_targets(dst, sig);
  if (ne > 0) ws_row.reserve(ne);
  score_prep_uniform_multi(ctx, enrol, ne, ld_enrol, ne > 0 ? &ws_l : nullptr, ne > 0 ? ws_row.get() : nullptr,
                           test_shard, nt_local, ld_test, row0, row0 + nt_local, dst, shard.ldk, dtype == 1, shard.dim,
                           consts, sig);
}

// Destinations of the current generation in every region (0 = this rank's own) and the ready flags of this source rank.
void PldaEngine::shard_targets(PrepDst& dst, PrepSignal& sig) {
  const size_t gen = shard.off_gen[shard.epoch & 1u];
  dst.n = sig.n = shard.world;
  for (int i = 0; i < shard.world; ++i) {
    const int r = (shard.rank + i) % shard.world;
    uint8_t* base = shard.peer[r];
    dst.hi[i] = reinterpret_cast<__nv_bfloat16*>(base + gen);
    dst.lo[i] = reinterpret_cast<__nv_bfloat16*>(base + gen + shard.off_lo);
    dst.term[i] = reinterpret_cast<float*>(base + gen + shard.off_col);
    sig.flag[i] = reinterpret_cast<unsigned*>(base + kFlagsOff) + shard.rank;
  }
  sig.counter = reinterpret_cast<unsigned*>(shard.region + kCounterOff);
  sig.epoch = shard.epoch;
  static const char* fence_mode = getenv("PLDA_B200_FENCE");
  sig.fence_per_thread = (fence_mode != nullptr && strcmp(fence_mode, "thread") == 0) ? 1 : 0;
}

// One sharded step with RAGGED enrol counts: the producer writes this rank's test rows, with the column terms of
// every group of `group_counts` in their extra K columns, into every region; the enrol rows get the one-hot pair of
// their group; the GEMM is the uniform-count kernel over K = dim + 2 * n_groups.
void PldaEngine::shard_step_ragged(const void* test_shard, int64_t nt_local, int64_t ld_test, const void* enrol,
                                   int64_t ne, int64_t ld_enrol, const int32_t* enrol_counts,
                                   const int32_t* group_counts, int n_groups, const uint64_t* ids, int dtype,
                                   float* out, int64_t ldo) {
  require_model();
  PB_CHECK(shard.open, kInvalidArg, "shard: no open session");
  for (int r = 0; r < shard.world; ++r) PB_CHECK(shard.peer[r] != nullptr, kInvalidArg, "shard: a peer is not connected");
  PB_CHECK(dtype == 0 || dtype == 1, kInvalidArg, "dtype must be PLDA_F64 or PLDA_F32");
  const int64_t row0 = shard.bounds[shard.rank];
  PB_CHECK(nt_local == shard.bounds[shard.rank + 1] - row0, kInvalidArg, "shard: shard size does not match the bounds");
  PB_CHECK(nt_local == 0 || (test_shard != nullptr && ld_test >= shard.dim), kInvalidArg, "shard: bad test rows");
  PB_CHECK(ne >= 0 && (ne == 0 || (enrol != nullptr && enrol_counts != nullptr && out != nullptr && ld_enrol >= shard.dim)),
           kInvalidArg, "shard: bad enrol rows");
  PB_CHECK(ne == 0 || ldo >= shard.nt_total, kInvalidArg, "shard_step: output pitch too small");
  PB_CHECK(group_counts != nullptr && n_groups >= 1 && n_groups <= shard.max_groups, kInvalidArg,
           "shard: the session was opened for fewer distinct enrol counts (plda_shard_open_ragged max_groups)");
  for (int i = 0; i < n_groups; ++i)
    PB_CHECK(group_counts[i] > 0 && (i == 0 || group_counts[i] > group_counts[i - 1]), kInvalidArg,
             "shard: group counts must be positive and strictly ascending");
  // group of every local enrol row
  std::vector<int32_t> grp(static_cast<size_t>(ne));
  for (int64_t i = 0; i < ne; ++i) {
    const int32_t* it = std::lower_bound(group_counts, group_counts + n_groups, enrol_counts[i]);
    PB_CHECK(it != group_counts + n_groups && *it == enrol_counts[i], kInvalidArg,
             "shard: an enrol count is missing from the group list");
    grp[i] = static_cast<int32_t>(it - group_counts);
  }
  // the ragged caches of score_grid describe other rows / tables from here on
  ragged_key.clear();
  last_counts.clear();
  rg_grp.reserve(std::max<int64_t>(ne, 1));
  if (ne > 0)
    PB_CUDA(cudaMemcpyAsync(rg_grp.get(), grp.data(), ne * sizeof(int32_t), cudaMemcpyHostToDevice, ctx.stream));
  std::vector<double> tabs(static_cast<size_t>(n_groups) * kScoreConstsSize, 0.0);
  for (int i = 0; i < n_groups; ++i)
    fill_score_consts(model.h_psi.data(), shard.dim, group_counts[i], tabs.data() + static_cast<size_t>(i) * kScoreConstsSize);
  ws_tables.reserve(tabs.size());
  PB_CUDA(cudaMemcpyAsync(ws_tables.get(), tabs.data(), tabs.size() * sizeof(double), cudaMemcpyHostToDevice, ctx.stream));
  shard.epoch += 1;
  shard.push_count = -n_groups;
  PrepDst dst;
  PrepSignal sig;
  shard_targets(dst, sig);
  ws_row.reserve(std::max<int64_t>(ne, 1));
  score_prep_grouped_vec(ctx, enrol, ne, ld_enrol, rg_grp.get(), test_shard, nt_local, ld_test, dtype == 1, shard.dim,
                         n_groups, ws_tables.get(), ws_l, ws_r, ws_row.get(), nullptr, 0, /*embed=*/true, &dst, row0,
                         shard.ldk, &sig);
  if (ne == 0) {
    shard_wait_all(ctx, shard_desc());     // an empty enrol block must not run ahead of its peers
    return;
  }
  const float* zmean = nullptr;
  const float* zinv = nullptr;
  znorm_affine(ids, ne, nullptr, nullptr, 0, &zmean, &zinv);
  const size_t gen = shard.off_gen[shard.epoch & 1u];
  SplitOperand b;
  b.hi = reinterpret_cast<const __nv_bfloat16*>(shard.region + gen);
  b.lo = reinterpret_cast<const __nv_bfloat16*>(shard.region + gen + shard.off_lo);
  b.rows = shard.nt_total;
  b.k = shard.dim + 2 * n_groups;
  b.ld = shard.ldk;
  GemmEpilogue epi;
  epi.out = out;
  epi.ldo = ldo;
  epi.row_add = ws_row.get();
  epi.zmean = zmean;
  epi.zinv = zinv;
  const GemmShard gs = shard_desc();
  gemm_bf16x3(ctx, ws_l.view(), b, ne, shard.nt_total, shard.dim + 2 * n_groups, epi, &gs);
}

// The grid of this rank's enrol block (operand already in ws_l / ws_row) against the current generation.
GemmShard PldaEngine::shard_desc() const {
  GemmShard gs;
  gs.flags = reinterpret_cast<const unsigned*>(shard.region + kFlagsOff);
  gs.err = reinterpret_cast<unsigned*>(shard.region + kErrOff);
  gs.epoch = shard.epoch;
  gs.world = shard.world;
  gs.rank = shard.rank;
  for (int r = 0; r <= shard.world; ++r) gs.bounds[r] = static_cast<int>(shard.bounds[r]);
  return gs;
}

void PldaEngine::shard_gemm(int64_t ne, const uint64_t* ids, float* out, int64_t ldo) {
  const float* zmean = nullptr;
  const float* zinv = nullptr;
  znorm_affine(ids, ne, nullptr, nullptr, 0, &zmean, &zinv);
  const size_t gen = shard.off_gen[shard.epoch & 1u];
  SplitOperand b;
  b.hi = reinterpret_cast<const __nv_bfloat16*>(shard.region + gen);
  b.lo = reinterpret_cast<const __nv_bfloat16*>(shard.region + gen + shard.off_lo);
  b.rows = shard.nt_total;
  b.k = shard.dim;
  b.ld = shard.ldk;
  GemmEpilogue epi;
  epi.out = out;
  epi.ldo = ldo;
  epi.row_add = ws_row.get();
  epi.col_add = reinterpret_cast<const float*>(shard.region + gen + shard.off_col);
  epi.col_ld = shard.col_ld;
  epi.zmean = zmean;
  epi.zinv = zinv;
  const GemmShard gs = shard_desc();
  gemm_bf16x3(ctx, ws_l.view(), b, ne, shard.nt_total, shard.dim, epi, &gs);
}

void PldaEngine::shard_push(const void* test_shard, int64_t nt_local, int64_t ld, int dtype, int enrol_count) {
  shard_produce(test_shard, nt_local, ld, nullptr, 0, 0, enrol_count, dtype);
}

void PldaEngine::shard_score(const void* enrol, int64_t ne, int64_t ld_enrol, int enrol_count, const uint64_t* ids,
                             int dtype, float* out, int64_t ldo) {
  require_model();
  PB_CHECK(shard.open, kInvalidArg, "shard_score: no open session");
  PB_CHECK(shard.epoch > 0, kInvalidArg, "shard_score: nothing pushed yet");
  PB_CHECK(enrol_count == shard.push_count, kInvalidArg,
           "shard_score: the column terms were pushed for a different enrol count");
  PB_CHECK(dtype == 0 || dtype == 1, kInvalidArg, "dtype must be PLDA_F64 or PLDA_F32");
  PB_CHECK(ne >= 0 && (ne == 0 || (enrol != nullptr && out != nullptr)), kInvalidArg, "shard_score: null pointer");
  PB_CHECK(ne == 0 || (ld_enrol >= shard.dim && ldo >= shard.nt_total), kInvalidArg, "shard_score: pitch too small");
  if (ne == 0) {
    // no grid to compute, but the step's back-pressure still applies: see every peer's flag before the next push
    shard_wait_all(ctx, shard_desc());
    return;
  }
  ws_row.reserve(ne);
  PrepDst none;
  score_prep_uniform_multi(ctx, enrol, ne, ld_enrol, &ws_l, ws_row.get(), nullptr, 0, 0, 0, 0, none, shard.ldk,
                           dtype == 1, shard.dim, score_consts_for(enrol_count, shard.dim), PrepSignal{});
  shard_gemm(ne, ids, out, ldo);
}

void PldaEngine::shard_step(const void* test_shard, int64_t nt_local, int64_t ld_test, const void* enrol, int64_t ne,
                            int64_t ld_enrol, int enrol_count, const uint64_t* ids, int dtype, float* out,
                            int64_t ldo) {
  PB_CHECK(ne >= 0 && (ne == 0 || out != nullptr), kInvalidArg, "shard_step: null output");
  PB_CHECK(!shard.open || ldo >= shard.nt_total, kInvalidArg, "shard_step: output pitch too small");
  shard_produce(test_shard, nt_local, ld_test, enrol, ne, ld_enrol, enrol_count, dtype);
  if (ne > 0) shard_gemm(ne, ids, out, ldo);
  else shard_wait_all(ctx, shard_desc());     // an empty enrol block must not run ahead of its peers
}

void PldaEngine::shard_status(int64_t* epoch, int64_t* timeouts) {
  PB_CHECK(shard.open, kInvalidArg, "shard_status: no open session");
  unsigned err = 0;
  PB_CUDA(cudaMemcpyAsync(&err, shard.region + kErrOff, sizeof(err), cudaMemcpyDeviceToHost, ctx.stream));
  ctx.sync();
  if (epoch) *epoch = shard.epoch;
  if (timeouts) *timeouts = err;
}

void PldaEngine::shard_close() {
  if (!shard.open) return;
  cudaStreamSynchronize(ctx.stream);
  for (int r = 0; r < shard.world; ++r)
    if (shard.peer_ipc[r] && shard.peer[r] != nullptr) cudaIpcCloseMemHandle(shard.peer[r]);
  cudaFree(shard.region);
  shard = ShardSession{};
}

}  // namespace pb
