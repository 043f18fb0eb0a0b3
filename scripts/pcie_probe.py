"""Raw pinned-memory PCIe rates next to the e2e scoring call (explains bench.py's e2e number)."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plda_b200 import PLDA

dev = torch.device("cuda", 0)
n = 100_000_000
d_buf = torch.empty(n, dtype=torch.float32, device=dev)
for label, h_buf in (("torch pinned", torch.empty(n, dtype=torch.float32).pin_memory()),):
    for _ in range(2):
        h_buf.copy_(d_buf, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        h_buf.copy_(d_buf, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print("D2H 400 MB %s: %.2f ms  %.1f GB/s" % (label, dt * 1e3, 0.4 / dt))
    t0 = time.perf_counter()
    for _ in range(5):
        d_buf.copy_(h_buf, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print("H2D 400 MB %s: %.2f ms  %.1f GB/s" % (label, dt * 1e3, 0.4 / dt))
    # split in 4 chunks on 2 streams (is one DMA stream rate-limited?)
    s = [torch.cuda.Stream() for _ in range(2)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        for i in range(4):
            with torch.cuda.stream(s[i & 1]):
                h_buf[i * n // 4:(i + 1) * n // 4].copy_(d_buf[i * n // 4:(i + 1) * n // 4], non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print("D2H 400 MB in 4 chunks on 2 streams: %.2f ms  %.1f GB/s" % (dt * 1e3, 0.4 / dt))

D = 200
rs = np.random.RandomState(5)
q, _ = np.linalg.qr(rs.randn(D, D))
p = PLDA()
p.set_model(np.full(D, 0.5), q, 2.0 * np.exp(-np.arange(D) / (0.15 * D)))
e = torch.empty((10000, D), dtype=torch.float64).pin_memory().numpy()
t = torch.empty((10000, D), dtype=torch.float64).pin_memory().numpy()
o = torch.empty((10000, 10000), dtype=torch.float32).pin_memory().numpy()
e[:] = rs.randn(10000, D); t[:] = rs.randn(10000, D)
cnt = np.full(10000, 3, np.int32)
p.score_grid(e, cnt, t, out=o)
for rep in range(3):
    t0 = time.perf_counter()
    for _ in range(5):
        p.score_grid(e, cnt, t, out=o)
    dt = (time.perf_counter() - t0) / 5
    print("e2e score_grid host->host: %.2f ms/step (%.2e trials/s)" % (dt * 1e3, 1e8 / dt))
