"""EM phase profile on the bench's own C2 data (two_cov spectrum, fp64 rows resident):
PLDA_B200_EM_PROFILE=1 PLDA_B200_DBG=1 python scripts/em_bench_probe.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from plda_b200 import PLDA
a_b = bench.two_cov(bench.D)
x, labels, _ = bench.speakers(a_b, bench.K_TRAIN, bench.N_TRAIN // bench.K_TRAIN, 1234)
dev = torch.device("cuda", 0)
xd = torch.from_numpy(x).to(dev)
ld = torch.from_numpy(labels.astype(np.int64)).to(dev)
p = PLDA()
for _ in range(2):
    p.fit(xd, ld, bench.EM_ITERS)
print(p.fit_timings())
