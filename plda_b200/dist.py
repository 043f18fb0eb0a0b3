"""Multi-GPU sharding of the scoring grid.

One process per GPU (``torch.distributed``, NCCL on GPUs, gloo in the CPU tests).  The grid
shards by ENROL BLOCK: rank g owns a contiguous block of enrol models (their operand rows, row
terms and z-norm statistics) and the ``block x Nt`` slab of the score grid.  Test vectors are
transformed shard-wise and exchanged with ONE all-gather; there is no other data-path collective.

The collective plumbing lives here so that it can be exercised with world_size-2 gloo tests on CPU
(``tests/test_dist_gloo.py``) with a stand-in scorer; on GPUs the scorer is ``PLDA.score_grid``.
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
