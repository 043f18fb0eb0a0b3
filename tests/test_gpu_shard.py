"""GPU tests of the sharded score grid over peer memory (C ABI plda_shard_*, csrc/engine_shard.cu).

Several ranks are emulated in ONE process on one GPU: every rank is its own PLDA handle (own stream, own region),
wired with PeerShardedScorer.connect_local, so the push kernel of rank r really writes into the regions of the other
ranks and the GEMM of every rank really polls the flags.  The cross-process (CUDA IPC / NVLink) wiring is covered by
tests/test_gpu_multi.py on >= 2 GPUs.  The checker is the CPU oracle (oracle/kaldi_plda.py) plus bit-equality with
the single-GPU plda_score_grid (same kernels, same operand values, different tile order).
"""
import time

import numpy as np
import pytest

from oracle import kaldi_plda as kp

pytestmark = pytest.mark.gpu


def _model(d, seed=5):
    rs = np.random.RandomState(seed)
    q, _ = np.linalg.qr(rs.randn(d, d))
    psi = 2.0 * np.exp(-np.arange(d) / (0.15 * d))
    return np.full(d, 0.5), q, psi


def _ranks(world, d, nt_total, model, max_groups=0):
    from plda_b200 import PLDA
    from plda_b200.dist import PeerShardedScorer
    hs, scorers = [], []
    for r in range(world):
        p = PLDA()
        p.set_model(*model)
        hs.append(p)
        scorers.append(PeerShardedScorer(p, nt_total, d, world=world, rank=r, max_groups=max_groups))
    PeerShardedScorer.connect_local(scorers)
    return hs, scorers


def _oracle_grid(psi, enrol, count, test):
    ref = kp.Plda()
    ref.psi = np.asarray(psi, dtype=np.float64)
    return kp.score_grid(ref, np.asarray(enrol, dtype=np.float64), np.full(enrol.shape[0], count),
                         np.asarray(test, dtype=np.float64))


@pytest.mark.parametrize("world,d,nt_total,count", [(1, 40, 300, 1), (2, 200, 1000, 3), (3, 200, 1001, 2),
                                                     (4, 150, 77, 5), (8, 64, 5, 1)])
def test_peer_sharded_grid_matches_single_and_oracle(world, d, nt_total, count):
    import torch
    from plda_b200 import PLDA
    from plda_b200.dist import block_bounds
    model = _model(d)
    hs, scorers = _ranks(world, d, nt_total, model)
    single = PLDA()
    single.set_model(*model)
    rng = np.random.RandomState(11)
    ne = [130 + 37 * r for r in range(world)]
    try:
        for step in range(3):                      # three steps: both operand generations get reused
            test = rng.randn(nt_total, d).astype(np.float32)
            enrol = [rng.randn(n, d).astype(np.float32) for n in ne]
            t_dev = torch.from_numpy(test).cuda()
            outs = []
            for r, sc in enumerate(scorers):       # all pushes first: one GPU cannot co-schedule a full-size GEMM
                lo, hi = block_bounds(nt_total, world, r)
                sc.push(t_dev[lo:hi], count)
            for r, sc in enumerate(scorers):
                outs.append(sc.grid(torch.from_numpy(enrol[r]).cuda(), count))
            torch.cuda.synchronize()
            for r, sc in enumerate(scorers):
                epoch, timeouts = sc.status()
                assert epoch == step + 1 and timeouts == 0
                got = outs[r].cpu().numpy()
                want = single.score_grid(torch.from_numpy(enrol[r]).cuda(), np.full(ne[r], count, np.int32), t_dev)
                torch.cuda.synchronize()
                assert np.array_equal(got, want.cpu().numpy())
                ref = _oracle_grid(model[2], enrol[r].astype(np.float64), count, test.astype(np.float64))
                assert np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1.0)) <= 1e-3
    finally:
        for sc in scorers:
            sc.close()


def test_grid_waits_for_a_late_push():
    """The GEMM of rank 0 is launched BEFORE rank 1 has pushed: its TMA producer has to spin on rank 1's flag.
    Small problem: the waiting GEMM occupies a handful of SMs, so the late push kernel can be scheduled."""
    import torch
    from plda_b200.dist import block_bounds
    d, nt_total, count = 96, 600, 2
    model = _model(d)
    hs, scorers = _ranks(2, d, nt_total, model)
    rng = np.random.RandomState(3)
    test = rng.randn(nt_total, d)
    enrol = rng.randn(200, d)
    t_dev = torch.from_numpy(test).cuda()
    e_dev = torch.from_numpy(enrol).cuda()
    try:
        lo, hi = block_bounds(nt_total, 2, 0)
        scorers[0].push(t_dev[lo:hi], count)
        out0 = scorers[0].grid(e_dev, count)           # in flight, waiting for rank 1
        time.sleep(0.05)
        lo1, hi1 = block_bounds(nt_total, 2, 1)
        scorers[1].push(t_dev[lo1:hi1], count)
        out1 = scorers[1].grid(e_dev, count)
        torch.cuda.synchronize()
        assert scorers[0].status() == (1, 0) and scorers[1].status() == (1, 0)
        ref = _oracle_grid(model[2], enrol, count, test)
        for out in (out0, out1):
            got = out.cpu().numpy()
            assert np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1.0)) <= 1e-3
        assert torch.equal(out0, out1)
    finally:
        for sc in scorers:
            sc.close()


@pytest.mark.parametrize("world", [2, 3])
def test_fused_step_equals_push_then_grid(world):
    """plda_shard_step (one producer launch for test push + enrol operand) == plda_shard_push + plda_shard_score.
    Ranks are stepped one after the other without synchronisation: every GEMM but the last one launched waits on
    flags of ranks that have not pushed yet (small grids, so the late producers find free SMs)."""
    import torch
    from plda_b200.dist import block_bounds
    d, nt_total, count = 200, 700, 3
    model = _model(d)
    hs, scorers = _ranks(world, d, nt_total, model)
    rng = np.random.RandomState(8)
    try:
        for step in range(2):
            test = rng.randn(nt_total, d).astype(np.float32)
            enrol = [rng.randn(150 + 20 * r, d).astype(np.float32) for r in range(world)]
            t_dev = torch.from_numpy(test).cuda()
            e_dev = [torch.from_numpy(e).cuda() for e in enrol]
            torch.cuda.synchronize()
            fused = []
            for r, sc in enumerate(scorers):
                lo, hi = block_bounds(nt_total, world, r)
                fused.append(sc.score(e_dev[r], count, t_dev[lo:hi], sync=False))
            torch.cuda.synchronize()
            for r, sc in enumerate(scorers):
                lo, hi = block_bounds(nt_total, world, r)
                sc.push(t_dev[lo:hi], count)
            split = [sc.grid(e_dev[r], count) for r, sc in enumerate(scorers)]
            torch.cuda.synchronize()
            for r, sc in enumerate(scorers):
                assert sc.status() == (2 * step + 2, 0)
                assert torch.equal(fused[r], split[r])
                ref = _oracle_grid(model[2], enrol[r], count, test)
                got = fused[r].cpu().numpy()
                assert np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1.0)) <= 1e-3
    finally:
        for sc in scorers:
            sc.close()


@pytest.mark.parametrize("world,d,nt_total", [(2, 200, 700), (3, 64, 333)])
def test_ragged_counts_on_the_sharded_grid(world, d, nt_total):
    """plda_shard_step_ragged: a count per enrol row, different count sets on different ranks; the column terms of
    every group travel inside the pushed operand rows.  Checked against the oracle's per-pair LLR and the single-GPU
    ragged grid; a uniform step on the same (ragged-capable) session still equals the plain sharded grid."""
    import torch
    from plda_b200 import PLDA
    from plda_b200.dist import block_bounds
    model = _model(d)
    hs, scorers = _ranks(world, d, nt_total, model, max_groups=4)
    single = PLDA()
    single.set_model(*model)
    rng = np.random.RandomState(21)
    ref = kp.Plda()
    ref.psi = np.asarray(model[2], dtype=np.float64)
    rank_counts = [[1, 2], [2, 5], [1, 7]][:world]
    groups = np.unique(np.concatenate(rank_counts)).astype(np.int32)
    try:
        for step in range(2):
            test = rng.randn(nt_total, d).astype(np.float32)
            enrol = [rng.randn(150 + 20 * r, d).astype(np.float32) for r in range(world)]
            counts = [rng.choice(rank_counts[r], size=enrol[r].shape[0]).astype(np.int32) for r in range(world)]
            t_dev = torch.from_numpy(test).cuda()
            e_dev = [torch.from_numpy(e).cuda() for e in enrol]
            torch.cuda.synchronize()
            outs = []
            for r, sc in enumerate(scorers):
                lo, hi = block_bounds(nt_total, world, r)
                outs.append(sc.score_ragged(e_dev[r], counts[r], t_dev[lo:hi], group_counts=groups, sync=False))
            torch.cuda.synchronize()
            for r, sc in enumerate(scorers):
                assert sc.status() == (2 * step + 1, 0)
                got = outs[r].cpu().numpy()
                want = kp.score_grid(ref, enrol[r].astype(np.float64), counts[r], test.astype(np.float64))
                assert np.max(np.abs(got - want) / np.maximum(np.abs(want), 1.0)) <= 1e-3
                one = single.score_grid(e_dev[r], counts[r], t_dev).cpu().numpy()
                assert np.max(np.abs(got - one)) <= 1e-4 * max(1.0, np.abs(one).max())
            # a uniform step on the ragged-capable session (wider operand pitch, same kernels)
            uni = []
            for r, sc in enumerate(scorers):
                lo, hi = block_bounds(nt_total, world, r)
                uni.append(sc.score(e_dev[r], 3, t_dev[lo:hi], sync=False))
            torch.cuda.synchronize()
            for r, sc in enumerate(scorers):
                assert sc.status() == (2 * step + 2, 0)
                want = single.score_grid(e_dev[r], np.full(enrol[r].shape[0], 3, np.int32), t_dev)
                assert torch.equal(uni[r], want)
        with pytest.raises(ValueError):              # a count outside the agreed group list
            scorers[0].score_ragged(e_dev[0], np.full(enrol[0].shape[0], 9, np.int32), t_dev[:block_bounds(nt_total, world, 0)[1]],
                                    group_counts=groups, sync=False)
        with pytest.raises(ValueError):              # more groups than the session has room for
            scorers[0].score_ragged(e_dev[0], counts[0], t_dev[:block_bounds(nt_total, world, 0)[1]],
                                    group_counts=np.arange(1, 7, dtype=np.int32), sync=False)
    finally:
        for sc in scorers:
            sc.close()


def test_sharded_znorm_and_errors():
    import torch
    from plda_b200.dist import block_bounds
    d, nt_total = 40, 64
    model = _model(d)
    hs, scorers = _ranks(2, d, nt_total, model)
    rng = np.random.RandomState(4)
    test = torch.from_numpy(rng.randn(nt_total, d)).cuda()
    enrol = torch.from_numpy(rng.randn(10, d)).cuda()
    try:
        with pytest.raises(ValueError):
            scorers[0].grid(enrol, 1)                  # nothing pushed yet
        ids = np.arange(10, dtype=np.uint64)
        for h in hs:
            h.norm(rng.randn(32, d) + 0.5, {int(i): (1, enrol[i].cpu().numpy()) for i in ids})
        for r, sc in enumerate(scorers):
            lo, hi = block_bounds(nt_total, 2, r)
            sc.push(test[lo:hi], 1)
        with pytest.raises(ValueError):
            scorers[0].grid(enrol, 2)                  # column terms were pushed for count 1
        with pytest.raises(ValueError):
            scorers[0].push(test[:5], 1)               # wrong shard size
        got = scorers[0].grid(enrol, 1, enrol_ids=ids)
        torch.cuda.synchronize()
        want = hs[0].score_grid(enrol, np.ones(10, np.int32), test, enrol_ids=ids)
        torch.cuda.synchronize()
        assert torch.equal(got, want)
    finally:
        for sc in scorers:
            sc.close()


@pytest.mark.parametrize("d", [40, 150, 200, 203])
@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_vectorised_prep_equals_scalar_prep(d, dtype):
    """The operand producer takes 16-byte loads when the rows allow it and scalar loads otherwise (odd pitch /
    unaligned base): both must give the same grid, and match the oracle."""
    import torch
    from plda_b200 import PLDA
    model = _model(d)
    p = PLDA()
    p.set_model(*model)
    rng = np.random.RandomState(d)
    ne, nt, count = 70, 90, 2
    tdt = getattr(torch, dtype)
    e_al = torch.from_numpy(rng.randn(ne, d)).to(tdt).cuda()
    t_al = torch.from_numpy(rng.randn(nt, d)).to(tdt).cuda()
    # same values at an odd pitch and an element-shifted base
    e_buf = torch.zeros((ne, d + 3), dtype=tdt, device="cuda")
    t_buf = torch.zeros((nt, d + 3), dtype=tdt, device="cuda")
    e_un, t_un = e_buf[:, 1:d + 1], t_buf[:, 1:d + 1]
    e_un.copy_(e_al)
    t_un.copy_(t_al)
    torch.cuda.synchronize()                         # the handle launches on its own stream
    cnt = np.full(ne, count, np.int32)
    a = p.score_grid(e_al, cnt, t_al)
    b = p.score_grid(e_un, cnt, t_un)
    torch.cuda.synchronize()
    assert torch.equal(a, b)
    ref = _oracle_grid(model[2], e_al.double().cpu().numpy(), count, t_al.double().cpu().numpy())
    got = a.cpu().numpy()
    assert np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1.0)) <= 1e-3
