"""world_size-2 (and 3) gloo tests of the multi-GPU host logic on CPU: enrol-block partition + the single
all-gather of the test vectors (plda_b200/dist.py).  The scorer is a stand-in (the oracle's Gram-form grid)
because the CUDA path needs a GPU; on GPUs the same class drives PLDA.score_grid (bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from plda_b200.dist import ShardedScorer, all_gather_rows, block_bounds


def test_block_bounds_partition():
    for n in (0, 1, 7, 10, 1000, 10001):
        for world in (1, 2, 3, 8):
            spans = [block_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        block_bounds(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ne, nt, d, result_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import kaldi_plda as kp
        rng = np.random.RandomState(0)           # same global problem on every rank
        q, _ = np.linalg.qr(rng.randn(d, d))
        model = kp.Plda()
        model.mean = rng.randn(d)
        model.transform = q
        model.psi = np.sort(rng.rand(d))[::-1] + 0.01
        model.compute_derived_vars()
        enrol = rng.randn(ne, d)
        counts = rng.randint(1, 4, size=ne)
        test = rng.randn(nt, d)

        def score_fn(e_blk, c_blk, t_all, ids):
            return torch.from_numpy(kp.score_grid(model, e_blk.numpy(), c_blk, t_all.numpy()))

        elo, ehi = block_bounds(ne, world, rank)
        tlo, thi = block_bounds(nt, world, rank)
        scorer = ShardedScorer(score_fn)
        gathered = all_gather_rows(torch.from_numpy(test[tlo:thi]), nt)
        assert torch.equal(gathered, torch.from_numpy(test))
        slab = scorer.score(torch.from_numpy(enrol[elo:ehi]), counts[elo:ehi], torch.from_numpy(test[tlo:thi]), nt)
        assert slab.shape == (ehi - elo, nt)
        full = scorer.gather_slabs_to_rank0(slab, ne)
        if rank == 0:
            want = kp.score_grid(model, enrol, counts, test)
            assert np.allclose(full.numpy(), want, atol=1e-12)
            open(os.path.join(result_dir, "ok_%d" % world), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,ne,nt", [(2, 11, 9), (3, 10, 7)])
def test_sharded_scoring_gloo(tmp_path, world, ne, nt):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, ne, nt, 6, str(tmp_path)), nprocs=world, join=True)
    assert os.path.exists(os.path.join(str(tmp_path), "ok_%d" % world))


class _FakeLda:
    """Stand-in with the attributes broadcast_lda touches (the real LDA needs a GPU)."""
    def __init__(self):
        self._coef = self._intercept = self._classes = None

    def set_coef(self, coef, intercept, classes=None):
        self._coef, self._intercept = coef, intercept
        self._classes = np.arange(coef.shape[0]) if classes is None else np.asarray(classes)


class _FakePlda:
    def norm(self, cohort, enrol, numutts, seed):
        return cohort.copy(), sorted(enrol), numutts, seed


def _lda_worker(rank, world, port, result_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.lda_port import LDAOracle
        from plda_b200.dist import broadcast_lda, merge_class_stats, sharded_norm
        rng = np.random.RandomState(3)           # same global problem on every rank
        k, d, n = 7, 5, 210
        y = rng.randint(0, k, n) * 3 - 4         # sparse, partly negative labels
        x = rng.randn(n, d) + y[:, None] * 0.1
        classes = np.unique(y)
        # whole classes per rank, ragged (4 + 3 for world 2), interleaved so the merge has to sort
        mine = classes[rank::world]
        rows = np.isin(y, mine)
        xl, yl = x[rows], y[rows]
        means = np.stack([xl[yl == c].mean(0) for c in mine])
        counts = np.array([(yl == c).sum() for c in mine])
        sw = sum((xl[yl == c] - xl[yl == c].mean(0)).T @ (xl[yl == c] - xl[yl == c].mean(0)) for c in mine)
        g_sw, g_means, g_counts, g_classes = merge_class_stats(sw, means, counts, mine)
        assert np.array_equal(g_classes, classes)
        assert np.array_equal(g_counts, np.array([(y == c).sum() for c in classes]))
        assert np.allclose(g_means, np.stack([x[y == c].mean(0) for c in classes]), atol=1e-13)
        want_sw = sum((x[y == c] - x[y == c].mean(0)).T @ (x[y == c] - x[y == c].mean(0)) for c in classes)
        assert np.allclose(g_sw, want_sw, atol=1e-10)
        # a class present on two ranks is refused
        with pytest.raises(ValueError):
            merge_class_stats(sw, means[:1], counts[:1], classes[:1])
        # coefficient broadcast
        lda = _FakeLda()
        if rank == 0:
            o = LDAOracle("svd")
            o.fit(x, y)
            lda.set_coef(np.asarray(o._coef), np.asarray(o._intercept), classes=classes)
        broadcast_lda(lda, src=0)
        o = LDAOracle("svd")
        o.fit(x, y)
        assert np.array_equal(lda._coef, o._coef) and np.array_equal(lda._intercept, o._intercept)
        # the label VALUES travel too (sparse, partly negative here): predict() agrees on every rank
        assert np.array_equal(lda._classes, classes)
        # cohort all-gather in front of norm
        cohort = rng.randn(9, d)
        lo, hi = block_bounds(9, world, rank)
        got = sharded_norm(_FakePlda(), cohort[lo:hi], 9, {5: None, 2: None}, numutts=4, seed=11)
        assert np.array_equal(got[0], cohort) and got[1:] == ([2, 5], 4, 11)
        open(os.path.join(result_dir, "lda_ok_%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_lda_and_norm_plumbing_gloo(tmp_path, world):
    port = _free_port()
    mp.spawn(_lda_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(os.path.join(str(tmp_path), "lda_ok_%d" % r)) for r in range(world))


def _peer_fail_worker(rank, world, port, result_dir):
    """No GPU here: plda_shard_open fails on every rank (null handle).  The point is the COLLECTIVE agreement of
    PeerShardedScorer: all ranks raise together (so a caller can fall back to the all-gather path on every rank)
    instead of one rank walking away from a collective the others are blocked in."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import ctypes as C
        from plda_b200.dist import PeerShardedScorer

        class FakePlda:
            _h = C.c_void_p(None)

        outcome = "no error"
        try:
            PeerShardedScorer(FakePlda(), 100, 8)
        except RuntimeError as e:
            outcome = "RuntimeError" if "peer-memory scorer unavailable" in str(e) else "other: %r" % (e,)
        dist.barrier()                          # both ranks are still in step
        with open(os.path.join(result_dir, "peer_%d.txt" % rank), "w") as f:
            f.write(outcome)
    finally:
        dist.destroy_process_group()


def test_peer_scorer_failure_is_collective(tmp_path):
    world = 2
    mp.spawn(_peer_fail_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert (tmp_path / ("peer_%d.txt" % r)).read_text() == "RuntimeError"


def _group_worker(rank, world, port, result_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from plda_b200.dist import agree_group_counts
        mine = [[1, 2, 2, 1], [5, 2, 5], [7]][rank]
        got = agree_group_counts(np.array(mine, dtype=np.int32), 4)
        over = "no error"
        try:
            agree_group_counts(np.array(mine, dtype=np.int32), 3)       # 4 distinct counts, room for 3: every rank raises
        except ValueError:
            over = "ValueError"
        empty = agree_group_counts(np.zeros(0, dtype=np.int32) if rank == 1 else np.array([3]), 2)   # a rank without rows
        dist.barrier()
        with open(os.path.join(result_dir, "groups_%d.txt" % rank), "w") as f:
            f.write("%s|%s|%s|%s" % (got.tolist(), got.dtype, over, empty.tolist()))
    finally:
        dist.destroy_process_group()


def test_ragged_group_list_is_agreed_collectively(tmp_path):
    """The sharded ragged grid needs the SAME ascending list of distinct enrol counts on every rank
    (plda_shard_step_ragged); too many counts must fail on every rank, not on one."""
    world = 3
    mp.spawn(_group_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert (tmp_path / ("groups_%d.txt" % r)).read_text() == "[1, 2, 5, 7]|int32|ValueError|[3]"
