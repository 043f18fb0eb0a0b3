"""TS (A-in-TMEM) Gram kernel vs the SS kernel: parity of the two paths + kernel times.
usage: python scripts/ts_probe.py [ne nt d]"""
import os, sys, subprocess, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import ctypes as C, torch
    from plda_b200 import PLDA, _ffi
    ne, nt, d = (int(a) for a in sys.argv[2:5])
    rs = np.random.RandomState(5)
    q, _ = np.linalg.qr(rs.randn(d, d))
    p = PLDA()
    p.set_model(np.full(d, 0.5), q, 2.0 * np.exp(-np.arange(d) / (0.15 * d)))
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    e = torch.randn(ne, d, device="cuda", generator=g); t = torch.randn(nt, d, device="cuda", generator=g)
    out = torch.empty((ne, (nt + 3) // 4 * 4), device="cuda")
    zm = torch.randn(ne, device="cuda", generator=g).double(); zs = (1.0 + torch.rand(ne, device="cuda", generator=g)).double()
    lib = _ffi.lib()
    res = p.score_grid(e, 3, t, out=out[:, :nt]).clone()
    resz = p.score_grid(e, 3, t, znorm=(zm, zs)).clone()
    for _ in range(3): p.score_grid(e, 3, t, out=out[:, :nt])
    _ffi.check(lib.plda_profile_gemm(p._h, 1))
    for _ in range(20): p.score_grid(e, 3, t, out=out[:, :nt])
    ms, n = C.c_double(), C.c_int64(); _ffi.check(lib.plda_profile_collect(p._h, C.byref(ms), C.byref(n))); _ffi.check(lib.plda_profile_gemm(p._h, 0))
    torch.save({"res": res.cpu(), "resz": resz.cpu()}, sys.argv[5])
    print(json.dumps({"kernel_ms": ms.value / n.value, "launches": n.value}))
    sys.exit(0)
ne, nt, d = (int(a) for a in (sys.argv[1:4] + ["10000", "10000", "200"][len(sys.argv) - 1:]))
import torch
outs = {}
for mode in ("0", "1"):
    env = dict(os.environ, PLDA_B200_TS=mode)
    f = "/tmp/ts_probe_%s.pt" % mode
    r = subprocess.run([sys.executable, __file__, "child", str(ne), str(nt), str(d), f], env=env, capture_output=True, text=True, timeout=600)
    print("TS=%s" % mode, r.stdout.strip()[-200:], r.stderr.strip()[-600:])
    if r.returncode == 0: outs[mode] = torch.load(f)
if len(outs) == 2:
    for k in ("res", "resz"):
        a, b = outs["0"][k].double(), outs["1"][k].double()
        print(k, "max |TS - SS| =", float((a - b).abs().max()), " max |SS| =", float(a.abs().max()), " mismatching > 1e-3:", int(((a - b).abs() > 1e-3 * a.abs().clamp(min=1)).sum()))
