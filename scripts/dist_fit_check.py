"""2-GPU check of the sharded fit: rank r fits its speakers' rows with PLDA.fit_distributed; the model must equal
a single-GPU fit of all rows (psi, mean, scores).  Launch: torchrun --nproc-per-node 2 scripts/dist_fit_check.py"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plda_b200 import PLDA
from plda_b200.dist import block_bounds

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
d, k = 96, 240
rng = np.random.RandomState(0)
q, _ = np.linalg.qr(rng.randn(d, d))
a_b = q * np.sqrt(2.0 * np.exp(-np.arange(d) / (0.15 * d)))[None, :]
counts = rng.randint(2, 15, size=k)
z = rng.randn(k, d)
labels = np.repeat(np.arange(k), counts)
x = 0.5 + (z @ a_b.T)[labels] + rng.randn(labels.shape[0], d)
lo, hi = block_bounds(k, world, rank)                      # whole speakers per rank
sel = (labels >= lo) & (labels < hi)
p = PLDA(device=local)
p.fit_distributed(x[sel], (labels[sel] - lo).astype(np.uint64), 6)
mean, tr, psi = p.get_model()
ok = True
if rank == 0:
    ref = PLDA(device=local)
    ref.fit(x, labels.astype(np.uint64), 6)
    m2, t2, psi2 = ref.get_model()
    e1 = np.max(np.abs(psi - psi2) / np.maximum(psi2, 1e-12))
    e2 = np.max(np.abs(mean - m2))
    xt = 0.5 + rng.randn(50, d)
    s1 = p.score_grid(p.transform_batch(xt[:20]), np.ones(20, np.int32), p.transform_batch(xt[20:]))
    s2 = ref.score_grid(ref.transform_batch(xt[:20]), np.ones(20, np.int32), ref.transform_batch(xt[20:]))
    e3 = np.max(np.abs(s1 - s2) / np.maximum(np.abs(s2), 1.0))
    print("dist fit vs single fit: psi rel %.2e  mean abs %.2e  score tol-err %.2e" % (e1, e2, e3))
    ok = e1 < 1e-3 and e2 < 1e-9 and e3 < 1e-3
psi_t = torch.from_numpy(psi).cuda()
gathered = [torch.empty_like(psi_t) for _ in range(world)]
dist.all_gather(gathered, psi_t)
same = all(torch.equal(gathered[0], g) for g in gathered)     # replicas stay bit-identical
if rank == 0:
    print("replica psi bit-identical across ranks:", same)
dist.destroy_process_group()
sys.exit(0 if (ok and same) else 1)
