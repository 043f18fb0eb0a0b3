"""Concurrent pinned D2H of the e2e payload (400 MB fp32 per rank) on N GPUs of one host -- the bound of bench.py's
`e2e` matrix sink at N > 1.  Launch: torchrun --nproc-per-node N scripts/pcie_probe_ranks.py
Prints per-rank and aggregate GB/s with all ranks copying at once, and each rank alone for comparison."""
import os, time
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 100_000_000
d_buf = torch.empty(n, dtype=torch.float32, device="cuda")
h_buf = torch.empty(n, dtype=torch.float32).pin_memory()


def timed(reps=5):
    for _ in range(2):
        h_buf.copy_(d_buf, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        h_buf.copy_(d_buf, non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


dist.barrier()
dt_all = timed()                                   # every rank at once
t = torch.tensor([dt_all], device="cuda")
allt = [torch.zeros_like(t) for _ in range(world)]
dist.all_gather(allt, t)
solo = []
for r in range(world):                             # one rank at a time
    dist.barrier()
    if r == rank:
        solo_dt = timed()
    dist.barrier()
ts = torch.tensor([solo_dt], device="cuda")
alls = [torch.zeros_like(ts) for _ in range(world)]
dist.all_gather(alls, ts)
if rank == 0:
    con = [0.4 / float(x.item()) for x in allt]
    alone = [0.4 / float(x.item()) for x in alls]
    print("pinned D2H, 400 MB per rank, %d ranks" % world)
    print("  alone      GB/s per rank: " + " ".join("%.1f" % v for v in alone))
    print("  concurrent GB/s per rank: " + " ".join("%.1f" % v for v in con))
    print("  concurrent aggregate: %.1f GB/s (sum of alone: %.1f); slowest rank needs %.2f ms per 400 MB -> e2e matrix-sink "
          "bound %.2e trials/s for all ranks" % (sum(con), sum(alone), 400.0 / min(con), world * 1e8 / (0.4 / min(con))))
dist.destroy_process_group()
