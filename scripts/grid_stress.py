"""Repeats the C2 score grid on the same inputs and compares every run bit-for-bit with the first one and with a
float64 Gram-form grid computed by torch on the device (race / nondeterminism hunt; not a parity test)."""
import sys

import numpy as np
import torch

import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plda_b200 import PLDA  # noqa: E402

ne = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000
nt = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
d = int(sys.argv[3]) if len(sys.argv) > 3 else 200
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 30
rng = np.random.RandomState(0)
g = PLDA()
q, _ = np.linalg.qr(rng.randn(d, d))
g.set_model(rng.randn(d) * 0.1, q * (1 + rng.rand(d))[:, None], np.sort(rng.rand(d) * 5 + 0.05)[::-1].copy())
e = g.transform_batch(rng.randn(ne, d), counts=3)
t = g.transform_batch(rng.randn(nt, d), counts=1)
n = np.full(ne, 3, dtype=np.int32)
ed = torch.as_tensor(e, device="cuda", dtype=torch.float32)
td = torch.as_tensor(t, device="cuda", dtype=torch.float32)
first = None
bad = 0
for r in range(reps):
    out = g.score_grid(ed, n, td)
    torch.cuda.synchronize()
    if first is None:
        first = out.clone()
        psi = torch.as_tensor(g.get_model()[2], device="cuda")
        n3 = 3.0
        a = psi / (n3 * psi + 1.0)
        v = 1.0 + a
        e64, t64 = ed.double(), td.double()
        # Gram form of Plda::LogLikelihoodRatio with uniform n
        m = n3 * a
        row = (-0.5 * (torch.log(v).sum() + ((m * e64) ** 2 / v).sum(1)))
        col = (-0.5 * (t64 ** 2 / v).sum(1)) + 0.5 * (torch.log(psi + 1).sum() + (t64 ** 2 / (psi + 1)).sum(1))
        want = row[:, None] + col[None, :] + ((m / v) * e64) @ t64.T
        err = ((first.double() - want).abs() / want.abs().clamp(min=1)).max().item()
        print("vs fp64 torch grid: max tol-err %.3e" % err)
    else:
        diff = (out != first)
        nb = int(diff.sum().item())
        if nb:
            bad += 1
            idx = diff.nonzero()
            print("run %d: %d differing scores; rows %d..%d cols %d..%d; max abs diff %.3e" % (
                r, nb, idx[:, 0].min(), idx[:, 0].max(), idx[:, 1].min(), idx[:, 1].max(),
                (out - first).abs().max().item()))
print("runs with differences: %d / %d" % (bad, reps - 1))
