"""Sinks probe for ncu captures / timings: usage  python scripts/r2_sink_probe.py [znorm|hist|trials|ragged] """
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plda_b200 import PLDA
mode = sys.argv[1] if len(sys.argv) > 1 else "znorm"
d = 160
rs = np.random.RandomState(5)
q, _ = np.linalg.qr(rs.randn(d, d))
p = PLDA()
p.set_model(np.full(d, 0.5), q, 2.0 * np.exp(-np.arange(d) / (0.15 * d)))
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1)
def wall(fn, n=10):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
if mode == "znorm":
    ne, m = 50000, 10000
    e = torch.randn(ne, d, device=dev, generator=g); cohort = 0.5 + torch.randn(m, d, device=dev, generator=g)
    print("znorm %d x %d: %.3f ms/call" % (ne, m, wall(lambda: p.norm_rows(cohort, e))))
elif mode == "hist":
    ne, nt = 20000, 20000
    e = torch.randn(ne, d, device=dev, generator=g); t = torch.randn(nt, d, device=dev, generator=g)
    es = torch.arange(ne, device=dev, dtype=torch.int32); ts = torch.randint(0, ne, (nt,), device=dev, generator=g).to(torch.int32)
    print("hist full %.3f ms" % wall(lambda: p.score_hist(e, 3, t, es, ts, -200.0, 200.0, 1 << 16), 5))
    print("hist tail %.3f ms" % wall(lambda: p.score_hist(e, 3, t, es, ts, 0.0, 200.0, 1 << 16, theta_lo=0.0), 5))
    print("grid      %.3f ms" % wall(lambda: p.score_grid(e, 3, t), 5))
elif mode == "trials":
    ne, nt, n = 10000, 10000, 1000000
    e = torch.randn(ne, d, device=dev, generator=g); t = torch.randn(nt, d, device=dev, generator=g)
    te = torch.sort(torch.randint(0, ne, (n,), device=dev, generator=g))[0].to(torch.int32)
    tt = torch.randint(0, nt, (n,), device=dev, generator=g).to(torch.int32)
    for md in ("direct", "grid"):
        print("trials %s: %.3f ms / 1e6 trials" % (md, wall(lambda: p.score_trials(e, 3, t, te, tt, mode=md))))
elif mode == "ragged":
    ne, nt = 10000, 10000
    e = torch.randn(ne, d, device=dev, generator=g); t = torch.randn(nt, d, device=dev, generator=g)
    cnt = rs.randint(1, 6, size=ne).astype(np.int32)
    out = torch.empty((ne, nt), device=dev)
    print("ragged %.3f ms | uniform %.3f ms" % (wall(lambda: p.score_grid(e, cnt, t, out=out), 20), wall(lambda: p.score_grid(e, 3, t, out=out), 20)))
