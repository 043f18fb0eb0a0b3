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
