"""Host-side cost of one scoring step through the Python layer (the GPU must never wait for the host inside a timed
step): enqueue N steps without synchronising and divide the host wall clock."""
import os, sys, time, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plda_b200 import PLDA, _ffi
from plda_b200.dist import PeerShardedScorer
d, ne, nt = 200, 10000, 10000
rs = np.random.RandomState(5)
q, _ = np.linalg.qr(rs.randn(d, d))
p = PLDA()
p.set_model(np.full(d, 0.5), q, 2.0 * np.exp(-np.arange(d) / (0.15 * d)))
dev = torch.device("cuda", 0)
e = torch.randn(ne, d, device=dev); t = torch.randn(nt, d, device=dev)
out = torch.empty((ne, nt), device=dev)
cnt = np.full(ne, 3, np.int32)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
_ffi.check(p._lib.plda_set_stream(p._h, C.c_void_p(stream.cuda_stream)))
def host_time(fn, n=300):
    for _ in range(20): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    return (t1 - t0) / n * 1e6, (t2 - t0) / n * 1e6
print("score_grid  host us/call %.1f  wall us/call %.1f" % host_time(lambda: p.score_grid(e, cnt, t, out=out)))
print("score_grid scalar count  host us/call %.1f  wall us/call %.1f" % host_time(lambda: p.score_grid(e, 3, t, out=out)))
peer = PeerShardedScorer(p, nt, d, world=1, rank=0)
print("peer.score  host us/call %.1f  wall us/call %.1f" % host_time(lambda: peer.score(e, 3, t, out=out, sync=False)))
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
flush = torch.empty(64 * 1024 * 1024, device=dev)
def timed(fn, n=50):
    tot = 0.0
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    torch.cuda.synchronize()
    for a, b in evs:
        flush.zero_(); a.record(stream); fn(); b.record(stream)
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in evs) / n * 1e3
print("event-timed step: score_grid %.1f us, peer.score %.1f us" % (timed(lambda: p.score_grid(e, cnt, t, out=out)), timed(lambda: peer.score(e, 3, t, out=out, sync=False))))
peer.close()
