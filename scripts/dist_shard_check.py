"""N-GPU check of the peer-memory sharded score grid (PeerShardedScorer: CUDA-IPC regions, NVLink pushes, flag waits
inside the GEMM) against the NCCL all-gather path (ShardedScorer) and the single-GPU grid.
Launch: torchrun --nproc-per-node N scripts/dist_shard_check.py"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plda_b200 import PLDA
from plda_b200.dist import PeerShardedScorer, all_gather_rows, block_bounds

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
d, nt_total, count = 200, 4001, 3
rs = np.random.RandomState(5)
q, _ = np.linalg.qr(rs.randn(d, d))
p = PLDA(device=local)
p.set_model(np.full(d, 0.5), q, 2.0 * np.exp(-np.arange(d) / (0.15 * d)))
ne = 700 + 50 * rank
ok = True
peer = PeerShardedScorer(p, nt_total, d)
lo, hi = block_bounds(nt_total, world, rank)
worst = 0.0
for step in range(4):
    rng = np.random.RandomState(100 + step)                       # same test set on every rank
    test = rng.randn(nt_total, d).astype(np.float32)
    enrol = np.random.RandomState(1000 * rank + step).randn(ne, d).astype(np.float32)
    t_shard = torch.from_numpy(test[lo:hi]).cuda()
    e_dev = torch.from_numpy(enrol).cuda()
    torch.cuda.synchronize()
    got = peer.score(e_dev, count, t_shard)
    test_all = all_gather_rows(t_shard, nt_total)                 # NCCL, torch's stream
    torch.cuda.synchronize()                                      # the handle launches on its own stream
    want = p.score_grid(e_dev, np.full(ne, count, np.int32), test_all)
    torch.cuda.synchronize()
    same = bool(torch.equal(got, want))
    worst = max(worst, float((got - want).abs().max().item()))
    ok = ok and same
epoch, timeouts = peer.status()
ok = ok and epoch == 4 and timeouts == 0
# back-to-back steps without host synchronisation in between (the bench pattern), checked at the end
outs = [torch.empty((ne, (nt_total + 3) // 4 * 4), dtype=torch.float32, device="cuda") for _ in range(2)]
for i in range(20):
    peer.score(e_dev, count, t_shard, out=outs[i & 1][:, :nt_total], sync=False)
torch.cuda.synchronize()
ok = ok and bool(torch.equal(outs[1][:, :nt_total], want)) and peer.status() == (24, 0)
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("peer-sharded grid == all-gather grid on every rank: %s (max abs diff %.3g, timeouts %d)"
          % (bool(flag.item()), worst, timeouts))
peer.close()
# ragged enrol counts: a count per row, different count sets per rank; the group list is agreed with one all-gather
rag = PeerShardedScorer(p, nt_total, d, max_groups=6)
counts = np.random.RandomState(7 + rank).choice([1 + rank, 2, 5], size=ne).astype(np.int32)
ok_r, worst_r = True, 0.0
for step in range(3):
    got = rag.score_ragged(e_dev, counts, t_shard)
    torch.cuda.synchronize()
    want_r = p.score_grid(e_dev, counts, test_all)
    torch.cuda.synchronize()
    diff = float((got - want_r).abs().max().item())
    worst_r = max(worst_r, diff)
    ok_r = ok_r and diff <= 1e-4 * max(1.0, float(want_r.abs().max().item()))
ok_r = ok_r and rag.status() == (3, 0)
flag_r = torch.tensor([1 if ok_r else 0], device="cuda")
dist.all_reduce(flag_r, op=dist.ReduceOp.MIN)
if rank == 0:
    print("ragged peer-sharded grid == single-GPU ragged grid on every rank: %s (max abs diff %.3g)"
          % (bool(flag_r.item()), worst_r))
rag.close()
flag = torch.minimum(flag, flag_r)
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)
