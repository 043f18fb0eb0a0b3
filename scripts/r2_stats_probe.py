"""Stats pass / EM probe: fused vs legacy stats pass timing (PLDA_B200_STATS=legacy), EM ms/iter, on device rows.
usage: python scripts/r2_stats_probe.py n d k iters [f32|f64]"""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plda_b200 import PLDA
n, d, k, iters = (int(a) for a in sys.argv[1:5])
dt = torch.float32 if (len(sys.argv) < 6 or sys.argv[5] == "f32") else torch.float64
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1)
z = torch.randn(k, d, device=dev, generator=g, dtype=dt)
x = (0.5 + z).repeat_interleave(n // k, dim=0) + torch.randn(n, d, device=dev, generator=g, dtype=dt)
labels = torch.arange(k, device=dev).repeat_interleave(n // k)
p = PLDA()
for _ in range(2):
    p.fit(x, labels, iters)
ft = p.fit_timings()
es = 4 if dt == torch.float32 else 8
print(json.dumps({"n": n, "d": d, "k": k, "dtype": str(dt), "mode": os.environ.get("PLDA_B200_STATS", "fused"),
                  "stats_ms": ft["stats"], "em_ms_per_iter": ft["em"] / max(1, ft["iters"]), "output_ms": ft["output"],
                  "stats_algo_gbs": n * d * es / (ft["stats"] * 1e-3) / 1e9}))
_, _, psi = p.get_model()
print("psi[:4]", psi[:4])
