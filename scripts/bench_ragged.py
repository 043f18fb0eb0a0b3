"""Score grid with RAGGED enrol counts (several distinct n): step time on resident inputs.
usage: [PLDA_B200_RAGGED=old] python scripts/bench_ragged.py [ne nt d groups reps]"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plda_b200 import PLDA

ne, nt, d, groups, reps = (int(a) for a in (sys.argv[1:6] + ["10000", "10000", "200", "5", "20"][len(sys.argv) - 1:]))
rng = np.random.RandomState(0)
q, _ = np.linalg.qr(rng.randn(d, d))
p = PLDA()
p.set_model(np.full(d, 0.5), q, 2.0 * np.exp(-np.arange(d) / (0.15 * d)))
e = torch.randn(ne, d, device="cuda")
t = torch.randn(nt, d, device="cuda")
cnt = rng.randint(1, groups + 1, size=ne).astype(np.int32)
out = torch.empty((ne, (nt + 3) // 4 * 4), device="cuda")
for _ in range(3):
    p.score_grid(e, cnt, t, out=out[:, :nt])
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(reps):
    p.score_grid(e, cnt, t, out=out[:, :nt])
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / reps
uni = np.full(ne, 3, np.int32)
for _ in range(3):
    p.score_grid(e, uni, t, out=out[:, :nt])
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(reps):
    p.score_grid(e, uni, t, out=out[:, :nt])
torch.cuda.synchronize()
du = (time.perf_counter() - t0) / reps
print("mode=%s ne=%d nt=%d d=%d groups=%d: ragged %.4f ms/step (%.3e trials/s) | uniform %.4f ms/step (host-synchronous calls)"
      % (os.environ.get("PLDA_B200_RAGGED", "tables"), ne, nt, d, groups, dt * 1e3, ne * nt / dt, du * 1e3))
