"""2-GPU check of the sharded fit: rank r fits its speakers' rows with PLDA.fit_distributed; the model must equal
a single-GPU fit of all rows (psi, mean, scores).  Launch: torchrun --nproc-per-node 2 scripts/dist_fit_check.py"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plda_b200 import LDA, PLDA
from plda_b200.dist import block_bounds, broadcast_lda

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

# ---- LDA: rows sharded by class (fit), test rows sharded + coefficients broadcast (predict) ----
kc, dl = 40, 24
rng = np.random.RandomState(1)                            # fresh stream: rank 0 drew extra numbers above
yl = rng.randint(0, kc, 4000)
xl = rng.randn(4000, dl) + rng.randn(kc, dl)[yl]
tl = rng.randn(300, dl)
clo, chi = block_bounds(kc, world, rank)
mine = (yl >= clo) & (yl < chi)
ld = LDA(device=local, precision="fp64")
ld.fit_distributed(xl[mine], yl[mine])
lda_ok = True
coef_t = torch.from_numpy(ld._coef).cuda()
gath = [torch.empty_like(coef_t) for _ in range(world)]
dist.all_gather(gath, coef_t)
lda_same = all(torch.equal(gath[0], g) for g in gath)
one = LDA(device=local, precision="fp64")
if rank == 0:
    one.fit(xl, yl)
    e4 = np.max(np.abs(ld._coef - one._coef)) + np.max(np.abs(ld._intercept - one._intercept))
    print("sharded LDA fit vs single fit: coef+intercept abs %.2e  identical across ranks: %s" % (e4, lda_same))
    lda_ok = e4 < 1e-8
broadcast_lda(one, src=0)
rlo, rhi = block_bounds(300, world, rank)
part = torch.from_numpy(np.asarray(one.predict_log_proba(tl[rlo:rhi]), dtype=np.float64)).cuda()
want = torch.from_numpy(np.asarray(ld.predict_log_proba(tl[rlo:rhi]), dtype=np.float64)).cuda()
lda_ok = lda_ok and bool(torch.allclose(part, want, atol=1e-4))
dist.destroy_process_group()
sys.exit(0 if (ok and same and lda_ok and lda_same) else 1)
