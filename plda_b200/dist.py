"""Multi-GPU sharding of the scoring grid.

One process per GPU (``torch.distributed``, NCCL on GPUs, gloo in the CPU tests).  The grid
shards by ENROL BLOCK: rank g owns a contiguous block of enrol models (their operand rows, row
terms and z-norm statistics) and the ``block x Nt`` slab of the score grid.  Test vectors are
transformed shard-wise and every rank needs all of them -- the only exchange on the data path:

* ``PeerShardedScorer`` (GPUs with peer access, uniform enrol counts): no collective at all.  The
  operand-producer kernel of every rank writes its test rows into the operand buffer of every
  other rank over NVLink peer memory (CUDA-IPC mapped regions) and the tcgen05 GEMM waits per
  column tile for the owner's ready flag (C ABI ``plda_shard_*``, ``csrc/engine_shard.cu``).
* ``ShardedScorer``: ONE NCCL all-gather of the transformed test vectors, then ``score_grid``.

The host-side plumbing lives here so that it can be exercised with world_size-2 gloo tests on CPU
(``tests/test_dist_gloo.py``: partition, ragged all-gather, slab assembly with a stand-in scorer,
collective failure agreement of the peer scorer); the peer-memory protocol itself is tested on one
GPU with several handles (``tests/test_gpu_shard.py``) and across processes
(``scripts/dist_shard_check.py``).
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np


def block_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block partition of ``n`` items: sizes differ by at most one."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_rows(local, n_total: int, group=None):
    """All-gather row blocks laid out by ``block_bounds`` into one ``[n_total, d]`` tensor.
    Ragged blocks are padded to the largest block for the collective and trimmed afterwards."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if world == 1:
        return local
    sizes = [block_bounds(n_total, world, r) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    d = local.shape[1]
    pad = torch.zeros((mx, d), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    gathered = torch.empty((world * mx, d), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered, pad, group=group)
    if all(hi - lo == mx for lo, hi in sizes):
        return gathered
    parts = [gathered[r * mx: r * mx + (hi - lo)] for r, (lo, hi) in enumerate(sizes)]
    return torch.cat(parts, dim=0)


class ShardedScorer:
    """Enrol-block sharded all-pairs scoring.

    ``score_fn(enrol_block, counts_block, test_all, ids_block) -> [block, Nt]`` is
    ``PLDA.score_grid`` on GPUs.  Every rank passes ITS enrol block and ITS shard of the test
    vectors (``block_bounds`` partition of the global arrays); ``score`` returns this rank's slab.
    """

    def __init__(self, score_fn: Callable, group=None):
        self.score_fn = score_fn
        self.group = group

    def score(self, enrol_block, counts_block, test_shard, n_test_total: int, ids_block: Optional[np.ndarray] = None):
        test_all = all_gather_rows(test_shard, n_test_total, self.group)
        return self.score_fn(enrol_block, counts_block, test_all, ids_block)

    def gather_slabs_to_rank0(self, slab, n_enrol_total: int):
        """Debug / small-problem helper: assemble the full grid on rank 0 (not used on the hot path)."""
        import torch
        import torch.distributed as dist
        world = dist.get_world_size(self.group)
        rank = dist.get_rank(self.group)
        if world == 1:
            return slab
        sizes = [block_bounds(n_enrol_total, world, r) for r in range(world)]
        mx = max(hi - lo for lo, hi in sizes)
        nt = slab.shape[1]
        pad = torch.zeros((mx, nt), dtype=slab.dtype, device=slab.device)
        pad[: slab.shape[0]] = slab
        out = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
        dist.gather(pad, out, dst=0, group=self.group)
        if rank != 0:
            return None
        return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(out, sizes)], dim=0)


def agree_group_counts(enrol_counts, max_groups: int, group=None) -> np.ndarray:
    """Collective: the distinct (positive) enrol counts over all ranks of ``group``, ascending, as int32 -- the list
    ``plda_shard_step_ragged`` needs to be the same everywhere.  Raises ``ValueError`` on EVERY rank when there are more
    than ``max_groups`` of them."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    mine = np.unique(np.asarray(enrol_counts, dtype=np.int64).reshape(-1))
    if mine.size and mine[0] <= 0:
        raise ValueError("enrol counts must be positive")
    width = int(max_groups) + 1                       # one more than fits, so that an overflow is seen by every rank
    pad = np.zeros(width, dtype=np.int64)
    pad[: min(mine.size, width)] = mine[:width]
    dev = _coll_device(group)
    allc = torch.empty(world * width, dtype=torch.int64, device=dev)      # flat: gloo accepts no other shape
    dist.all_gather_into_tensor(allc, torch.from_numpy(pad).to(dev), group=group)
    allc = np.unique(allc.cpu().numpy())
    allc = allc[allc > 0]
    if allc.size > max_groups:
        raise ValueError("%d or more distinct enrol counts over all ranks, the session has room for %d"
                         % (allc.size, max_groups))
    return allc.astype(np.int32)


class PeerShardedScorer:
    """Enrol-block sharded scoring WITHOUT a collective: the operand producer of every rank writes its test rows
    straight into the operand buffer of every other rank over NVLink peer memory and the tcgen05 GEMM waits, per
    column tile, for the rank that owns the rows (C ABI ``plda_shard_*``, ``csrc/engine_shard.cu``).

    One instance per process (rank); ``torch.distributed`` is used once, at construction, to exchange the 64-byte
    CUDA-IPC handles of the regions -- no collective runs on the data path.  ``score`` must be called the same
    number of times on every rank.  ``score`` takes one enrol count for the block, ``score_ragged`` a count per
    row (scorer opened with ``max_groups``).

    ``peers``: same-process alternative to ``torch.distributed`` (tests, several handles in one process): a list of
    all ranks' ``PeerShardedScorer`` objects is connected with ``PeerShardedScorer.connect_local``.
    """

    def __init__(self, plda, n_test_total: int, dim: int, group=None, world: Optional[int] = None,
                 rank: Optional[int] = None, max_groups: int = 0):
        import ctypes as C
        from . import _ffi
        self._ffi, self._C = _ffi, C
        self._lib = _ffi.lib()
        self.plda = plda
        self.group = group
        self._local_only = world is not None
        if world is None:
            import torch.distributed as dist
            world, rank = dist.get_world_size(group), dist.get_rank(group)
        self.world, self.rank = int(world), int(rank)
        self.n_test_total, self.dim = int(n_test_total), int(dim)
        self.max_groups = int(max_groups)     # > 0: room for that many distinct enrol counts (score_ragged)
        self.bounds = np.array([block_bounds(self.n_test_total, self.world, r)[0] for r in range(self.world)]
                               + [self.n_test_total], dtype=np.int64)
        self._handle = (C.c_ubyte * 64)()
        self._region = C.c_void_p()
        self._open = False
        err = None
        try:
            if self.max_groups > 0:
                _ffi.check(self._lib.plda_shard_open_ragged(plda._h, self.world, self.rank, _ffi.ptr(self.bounds),
                                                            self.dim, self.max_groups,
                                                            C.cast(self._handle, C.c_void_p), C.byref(self._region)))
            else:
                _ffi.check(self._lib.plda_shard_open(plda._h, self.world, self.rank, _ffi.ptr(self.bounds), self.dim,
                                                     C.cast(self._handle, C.c_void_p), C.byref(self._region)))
            self._open = True
        except Exception as e:          # decided collectively below: a rank must not leave the others in a collective
            err = e
        if self._local_only:
            if err is not None:
                raise err
            return
        self._agree(err, "allocating the exchange region")
        self._exchange()

    # -- wiring ---------------------------------------------------------------------------------
    def _agree(self, err, what):
        """Collective: every rank learns whether ANY rank failed; then all of them release their region and raise
        (so that a caller can fall back to the all-gather path on every rank), or none does."""
        import torch
        import torch.distributed as dist
        flag = torch.tensor([0 if err is None else 1], dtype=torch.int32, device=_coll_device(self.group))
        dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=self.group)
        if int(flag.item()) == 0:
            return
        self.close(barrier=False)
        raise RuntimeError("peer-memory scorer unavailable (%s failed on %s): %r"
                           % (what, "this rank" if err is not None else "another rank", err))

    def _exchange(self):
        import torch
        import torch.distributed as dist
        dev = _coll_device(self.group)
        mine = torch.tensor(list(bytes(self._handle)), dtype=torch.uint8, device=dev)
        allh = torch.empty((self.world, 64), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(allh, mine, group=self.group)
        allh = allh.cpu().numpy()
        err = None
        try:
            for r in range(self.world):
                if r == self.rank:
                    continue
                buf = (self._C.c_ubyte * 64)(*allh[r].tolist())
                self._ffi.check(self._lib.plda_shard_connect(self.plda._h, r, self._C.cast(buf, self._C.c_void_p),
                                                             None))
        except Exception as e:
            err = e
        # doubles as the barrier: every region is mapped everywhere before the first push
        self._agree(err, "mapping the peers' regions (CUDA IPC)")

    @staticmethod
    def connect_local(scorers):
        """Wire scorers that live in ONE process (their regions are plain device pointers to each other)."""
        for a in scorers:
            for b in scorers:
                if a is not b:
                    a._ffi.check(a._lib.plda_shard_connect(a.plda._h, b.rank, None, b._region))

    # -- data path --------------------------------------------------------------------------------
    def push(self, test_shard, enrol_count: int):
        """Producer + exchange: this rank's transformed test rows (CUDA tensor ``[bounds[rank+1]-bounds[rank], dim]``,
        fp32 or fp64) -> every rank's operand buffer.  Stream-ordered, does not synchronise."""
        tt, dtype = _cuda_matrix(test_shard)
        self.plda._after_torch(tt)
        self._ffi.check(self._lib.plda_shard_push(self.plda._h, self._C.c_void_p(tt.data_ptr()), tt.shape[0],
                                                  tt.stride(0) if tt.shape[0] else self.dim, dtype, int(enrol_count)))

    def grid(self, enrol_block, enrol_count: int, out=None, enrol_ids=None):
        """This rank's ``[block, n_test_total]`` slab of the grid (fp32 CUDA tensor); stream-ordered."""
        import torch
        et, dtype = _cuda_matrix(enrol_block)
        ne = et.shape[0]
        if out is None:
            ldo = (self.n_test_total + 3) // 4 * 4
            out = torch.empty((ne, ldo), dtype=torch.float32, device=et.device)[:, : self.n_test_total]
        ids = None if enrol_ids is None else np.ascontiguousarray(enrol_ids, dtype=np.uint64).reshape(-1)
        self.plda._after_torch(et)
        self._ffi.check(self._lib.plda_shard_score(self.plda._h, self._C.c_void_p(et.data_ptr()), ne,
                                                   et.stride(0) if ne else self.dim, int(enrol_count),
                                                   self._ffi.ptr(ids), dtype, self._C.c_void_p(out.data_ptr()),
                                                   out.stride(0)))
        return out

    def score(self, enrol_block, enrol_count: int, test_shard, out=None, enrol_ids=None, sync: bool = True):
        """One sharded scoring step (``plda_shard_step``): ONE producer launch (this rank's test rows -> every rank,
        plus the enrol operand) and the grid GEMM (+ a stream synchronisation unless ``sync=False``)."""
        import torch
        tt, dtype = _cuda_matrix(test_shard)
        et, dtype_e = _cuda_matrix(enrol_block)
        if dtype != dtype_e:
            raise ValueError("enrol and test dtypes differ")
        ne = et.shape[0]
        if out is None:
            ldo = (self.n_test_total + 3) // 4 * 4
            out = torch.empty((ne, ldo), dtype=torch.float32, device=et.device)[:, : self.n_test_total]
        ids = None if enrol_ids is None else np.ascontiguousarray(enrol_ids, dtype=np.uint64).reshape(-1)
        C = self._C
        self.plda._after_torch(tt)
        self._ffi.check(self._lib.plda_shard_step(
            self.plda._h, C.c_void_p(tt.data_ptr()), tt.shape[0], tt.stride(0) if tt.shape[0] else self.dim,
            C.c_void_p(et.data_ptr()), ne, et.stride(0) if ne else self.dim, int(enrol_count), self._ffi.ptr(ids), dtype,
            C.c_void_p(out.data_ptr()), out.stride(0)))
        if sync:
            self.check()
        return out

    def group_counts(self, enrol_counts):
        """The distinct enrol counts over ALL ranks, ascending (one small all-gather; the list every rank must pass to
        ``score_ragged``).  Same-process scorers (tests) pass the list themselves."""
        return agree_group_counts(enrol_counts, self.max_groups, self.group)

    def score_ragged(self, enrol_block, enrol_counts, test_shard, group_counts=None, out=None, enrol_ids=None,
                     sync: bool = True):
        """One sharded scoring step with a count PER ENROL ROW (``plda_shard_step_ragged``): the column terms of
        every distinct count travel inside the pushed operand rows, the grid is the uniform-count kernel.  Needs a
        scorer opened with ``max_groups`` >= the number of distinct counts over all ranks."""
        import torch
        tt, dtype = _cuda_matrix(test_shard)
        et, dtype_e = _cuda_matrix(enrol_block)
        if dtype != dtype_e:
            raise ValueError("enrol and test dtypes differ")
        ne = et.shape[0]
        cnt = np.ascontiguousarray(enrol_counts, dtype=np.int32).reshape(-1)
        if cnt.shape[0] != ne:
            raise ValueError("score_ragged: enrol_counts length mismatch")
        if group_counts is None:
            group_counts = self.group_counts(cnt)
        groups = np.ascontiguousarray(group_counts, dtype=np.int32).reshape(-1)
        if out is None:
            ldo = (self.n_test_total + 3) // 4 * 4
            out = torch.empty((ne, ldo), dtype=torch.float32, device=et.device)[:, : self.n_test_total]
        ids = None if enrol_ids is None else np.ascontiguousarray(enrol_ids, dtype=np.uint64).reshape(-1)
        C = self._C
        self.plda._after_torch(tt)
        self._ffi.check(self._lib.plda_shard_step_ragged(
            self.plda._h, C.c_void_p(tt.data_ptr()), tt.shape[0], tt.stride(0) if tt.shape[0] else self.dim,
            C.c_void_p(et.data_ptr()), ne, et.stride(0) if ne else self.dim, self._ffi.ptr(cnt), self._ffi.ptr(groups),
            int(groups.shape[0]), self._ffi.ptr(ids), dtype, C.c_void_p(out.data_ptr()), out.stride(0)))
        if sync:
            self.check()
        return out

    def check(self):
        """Synchronise and raise if any wait on a peer's flag has timed out (a stalled or dead peer): the slab was then
        computed on stale operands and must not be used."""
        self._ffi.check(self._lib.plda_synchronize(self.plda._h))
        epoch, timeouts = self.status()
        if timeouts:
            raise RuntimeError("peer-memory exchange: %d wait(s) on a peer's ready flag timed out by step %d; the "
                               "scores of this rank are invalid" % (timeouts, epoch))

    def status(self):
        """``(pushes so far, waits that timed out)`` -- a non-zero second value invalidates the results."""
        e, t = self._C.c_int64(), self._C.c_int64()
        self._ffi.check(self._lib.plda_shard_status(self.plda._h, self._C.byref(e), self._C.byref(t)))
        return int(e.value), int(t.value)

    def close(self, barrier: bool = True):
        if not self._open:
            return
        self._open = False
        timeouts = 0
        if barrier:
            try:
                self._ffi.check(self._lib.plda_synchronize(self.plda._h))
                timeouts = self.status()[1]
            except Exception:
                timeouts = -1
        if barrier and not self._local_only:
            import torch.distributed as dist
            dist.barrier(group=self.group)     # nobody writes into a region that is about to be freed
        self._ffi.check(self._lib.plda_shard_close(self.plda._h))
        if timeouts:
            raise RuntimeError("peer-memory exchange: waits on a peer's ready flag timed out during this session; "
                               "slabs computed after the first timeout are invalid")


def _cuda_matrix(x):
    import torch
    if not (isinstance(x, torch.Tensor) and x.is_cuda):
        raise ValueError("the sharded scorer works on CUDA tensors (resident operands)")
    if x.dim() != 2 or x.dtype not in (torch.float32, torch.float64):
        raise ValueError("expected a 2-D float32/float64 CUDA tensor")
    if x.shape[0] and x.stride(1) != 1:
        x = x.contiguous()
    return x, (1 if x.dtype == torch.float32 else 0)


def _coll_device(group=None):
    """NCCL moves CUDA tensors only; gloo (CPU tests) moves host tensors."""
    import torch
    import torch.distributed as dist
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")


def merge_class_stats(sw, means, counts, classes, group=None):
    """Merge per-rank LDA class statistics (SURVEY 8e "LDA fit"): ONE sum-all-reduce of the ``d x d`` within scatter
    and ONE all-gather of the packed per-class rows ``[class, count, mean...]`` (ragged -> padded to the largest
    rank).  Returns ``(sw, means, counts, classes)`` with classes ascending; a class seen on two ranks is an error
    (its scatter would be centred on two different means)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    sw = np.ascontiguousarray(sw, dtype=np.float64)
    means = np.ascontiguousarray(means, dtype=np.float64)
    k, d = means.shape
    if world > 1:
        dev = _coll_device(group)
        t_sw = torch.from_numpy(sw.copy()).to(dev)
        dist.all_reduce(t_sw, group=group)
        sw = t_sw.cpu().numpy()
        t_k = torch.tensor([k], dtype=torch.int64, device=dev)
        ks = [torch.zeros_like(t_k) for _ in range(world)]
        dist.all_gather(ks, t_k, group=group)
        ks = [int(v.item()) for v in ks]
        mx = max(ks)
        # class ids and counts travel as int64 next to the fp64 means (bit-exact for any label value)
        ids = torch.zeros((mx, 2), dtype=torch.int64, device=dev)
        ids[:k, 0] = torch.from_numpy(np.asarray(classes, dtype=np.int64)).to(dev)
        ids[:k, 1] = torch.from_numpy(np.asarray(counts, dtype=np.int64)).to(dev)
        mm = torch.zeros((mx, d), dtype=torch.float64, device=dev)
        mm[:k] = torch.from_numpy(means).to(dev)
        g_ids = torch.empty((world * mx, 2), dtype=torch.int64, device=dev)
        g_mm = torch.empty((world * mx, d), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(g_ids, ids, group=group)
        dist.all_gather_into_tensor(g_mm, mm, group=group)
        keep = np.concatenate([np.arange(r * mx, r * mx + kr) for r, kr in enumerate(ks)])
        g_ids = g_ids.cpu().numpy()[keep]
        means = g_mm.cpu().numpy()[keep]
        classes, counts = g_ids[:, 0], g_ids[:, 1]
    classes = np.asarray(classes, dtype=np.int64)
    counts = np.asarray(counts, dtype=np.int64)
    order = np.argsort(classes, kind="stable")
    classes, counts, means = classes[order], counts[order], means[order]
    if np.any(classes[1:] == classes[:-1]):
        raise ValueError("sharded LDA fit: every class must live on exactly one rank")
    return sw, np.ascontiguousarray(means), counts, classes


def broadcast_lda(lda, src: int = 0, group=None):
    """Replicate a fitted LDA from rank ``src`` (SURVEY 8e "LDA predict": test rows are sharded, the model is
    broadcast once after the fit): ``coef`` / ``intercept`` / the class label of every row, so ``predict`` returns the
    same labels on every rank.  (``transform`` needs the fitted projection and stays with the fitting rank /
    ``fit_distributed``, which fits replicas everywhere.)"""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank(group)
    dev = _coll_device(group)
    shape = torch.zeros(2, dtype=torch.int64, device=dev)
    if rank == src:
        shape[0], shape[1] = lda._coef.shape
    dist.broadcast(shape, src=src, group=group)
    k, d = int(shape[0].item()), int(shape[1].item())
    buf = torch.empty((k, d + 1), dtype=torch.float64, device=dev)
    cls = torch.empty(k, dtype=torch.int64, device=dev)
    if rank == src:
        buf[:, :d] = torch.from_numpy(np.asarray(lda._coef, dtype=np.float64)).to(dev)
        buf[:, d] = torch.from_numpy(np.asarray(lda._intercept, dtype=np.float64)).to(dev)
        cls.copy_(torch.from_numpy(np.asarray(lda._classes, dtype=np.int64)).to(dev))
    dist.broadcast(buf, src=src, group=group)
    dist.broadcast(cls, src=src, group=group)
    if rank != src:
        h = buf.cpu().numpy()
        lda.set_coef(np.ascontiguousarray(h[:, :d]), np.ascontiguousarray(h[:, d]), classes=cls.cpu().numpy())
    return lda


def sharded_norm(plda, cohort_shard, n_cohort_total: int, enrol_block: dict, numutts: int = 0, seed: int = 0,
                 group=None):
    """z-norm with enrol-block ownership (SURVEY 8e "z-norm"): the cohort rows are all-gathered (one small
    collective), the moments of this rank's enrol models stay local to it."""
    import torch
    t = torch.as_tensor(np.ascontiguousarray(cohort_shard, dtype=np.float64)).to(_coll_device(group))
    cohort = all_gather_rows(t, n_cohort_total, group).cpu().numpy()
    return plda.norm(cohort, enrol_block, numutts, seed)
