"""Run a BASELINE.json config end to end on one GPU with device-generated synthetic data and report timings.
usage: python scripts/run_config.py c3|c2|c4shard"""
import os, sys, time, json
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plda_b200 import PLDA

cfg = sys.argv[1] if len(sys.argv) > 1 else "c3"
P = {"c2": dict(n=100_000, d=200, k=1_000, r=0, ne=10_000, nt=10_000, m=0),
     "c3": dict(n=1_000_000, d=256, k=10_000, r=150, ne=50_000, nt=100_000, m=10_000),
     # one GPU's share of C4: full fit (5M x 512, 50k speakers) + a 25k x 1M enrol-block slab of the 200k x 1M grid
     "c4shard": dict(n=5_000_000, d=512, k=50_000, r=0, ne=25_000, nt=1_000_000, m=0)}[cfg]
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1234)
d, k, n = P["d"], P["k"], P["n"]
spec = 2.0 * torch.exp(-torch.arange(d, device=dev, dtype=torch.float64) / (0.15 * d))
q, _ = torch.linalg.qr(torch.randn(d, d, device=dev, dtype=torch.float64, generator=g))
a_b = q * spec.sqrt()[None, :]
def speakers(k, per, dtype=torch.float32):
    z = torch.randn(k, d, device=dev, dtype=torch.float64, generator=g)
    x = (0.5 + (z @ a_b.T)).repeat_interleave(per, dim=0)
    x += torch.randn(k * per, d, device=dev, dtype=torch.float64, generator=g)
    return x.to(dtype), z
out = {"config": cfg, **P}
x, _ = speakers(k, n // k, torch.float32)
labels = np.repeat(np.arange(k), n // k).astype(np.uint64)
p = PLDA()
p.fit(x, labels, 1)                       # first call: workspace allocations (10 GB operand buffer at C4)
torch.cuda.synchronize(); t0 = time.perf_counter()
p.fit(x, labels, 10)
torch.cuda.synchronize(); out["fit_s"] = time.perf_counter() - t0
out["fit_timings_ms"] = p.fit_timings()
del x
r = P["r"]
xe, ze = speakers(P["ne"], 3)
enrol = p.transform_batch(xe.view(P["ne"], 3, d).mean(dim=1), counts=3, targetdim=r, out_dtype=np.float32)
del xe
zt = torch.randn(P["nt"], d, device=dev, dtype=torch.float64, generator=g)
xt = (0.5 + zt @ a_b.T + torch.randn(P["nt"], d, device=dev, dtype=torch.float64, generator=g)).float()
torch.cuda.synchronize(); t0 = time.perf_counter()
test = p.transform_batch(xt, counts=1, targetdim=r, out_dtype=np.float32)
torch.cuda.synchronize(); out["transform_test_s"] = time.perf_counter() - t0
del xt, zt
ids = None
if P["m"]:
    bkg, _ = speakers(P["m"], 1)
    # z-norm through the C ABI with device enrol vectors
    import ctypes as C
    from plda_b200 import _ffi
    ids = np.arange(P["ne"], dtype=np.uint64)
    _ffi.check(p._lib.plda_norm(p._h, C.c_void_p(bkg.data_ptr()), P["m"], d, d, _ffi.F32, _ffi.DEVICE, _ffi.ptr(ids),
                                C.c_void_p(enrol.data_ptr()), P["ne"], enrol.stride(0), enrol.shape[1], _ffi.F32,
                                _ffi.DEVICE, 0, 0))          # first call: allocations
    _ffi.check(p._lib.plda_znorm_clear(p._h))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    _ffi.check(p._lib.plda_norm(p._h, C.c_void_p(bkg.data_ptr()), P["m"], d, d, _ffi.F32, _ffi.DEVICE, _ffi.ptr(ids),
                                C.c_void_p(enrol.data_ptr()), P["ne"], enrol.stride(0), enrol.shape[1], _ffi.F32,
                                _ffi.DEVICE, 0, 0))
    torch.cuda.synchronize(); out["znorm_s"] = time.perf_counter() - t0
    out["znorm_trials_per_s"] = P["m"] * P["ne"] / out["znorm_s"]
cnt = np.full(P["ne"], 3, dtype=np.int32)
grid = p.score_grid(enrol, cnt, test, enrol_ids=ids)
torch.cuda.synchronize(); t0 = time.perf_counter()
grid = p.score_grid(enrol, cnt, test, enrol_ids=ids, out=grid)
torch.cuda.synchronize(); out["grid_s"] = time.perf_counter() - t0
out["grid_trials_per_s"] = P["ne"] * P["nt"] / out["grid_s"]
out["grid_gb"] = grid.numel() * 4 / 1e9
out["grid_finite"] = bool(torch.isfinite(grid[:1000]).all().item())
out["mem_gb"] = torch.cuda.max_memory_allocated() / 1e9
print(json.dumps(out))
