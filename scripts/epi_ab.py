"""A/B of score-grid epilogue variants in one process: bitwise comparison on an awkward shape, kernel time on the
bench shape.  usage: python scripts/epi_ab.py [variant ...]   (values of PLDA_B200_EPI; default: hybrid)"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plda_b200 import PLDA, _ffi  # noqa: E402

variants = sys.argv[1:] or ["hybrid"]
dev = torch.device("cuda", 0)
lib = _ffi.lib()


def handle(d, mode):
    if mode is None:
        os.environ.pop("PLDA_B200_EPI", None)
    else:
        os.environ["PLDA_B200_EPI"] = mode
    rng = np.random.RandomState(0)
    p = PLDA()
    q, _ = np.linalg.qr(rng.randn(d, d))
    p.set_model(rng.randn(d), q, np.sort(2.0 * np.exp(-np.arange(d) / (0.15 * d)))[::-1].copy())
    return p


def kernel_ms(p, e, cnt, t, out, reps=20):
    for _ in range(3):
        p.score_grid(e, cnt, t, out=out)
    torch.cuda.synchronize()
    _ffi.check(lib.plda_profile_gemm(p._h, 1))
    for _ in range(reps):
        p.score_grid(e, cnt, t, out=out)
    torch.cuda.synchronize()
    ms, n = C.c_double(), C.c_int64()
    _ffi.check(lib.plda_profile_collect(p._h, C.byref(ms), C.byref(n)))
    _ffi.check(lib.plda_profile_gemm(p._h, 0))
    return ms.value / n.value


for d in (200, 512):
    base = handle(d, None)
    g = torch.Generator(device=dev); g.manual_seed(1)
    e_s = torch.randn(1000, d, device=dev, generator=g); t_s = torch.randn(1003, d, device=dev, generator=g)
    cnt_s = np.full(1000, 3, dtype=np.int32)
    zm, zs = np.random.RandomState(2).randn(1000), 0.5 + np.random.RandomState(3).rand(1000)
    ref = base.score_grid(e_s, cnt_s, t_s, znorm=(zm, zs)).clone()
    ne = nt = 10000 if d == 200 else 20000
    e = torch.randn(ne, d, device=dev, generator=g); t = torch.randn(nt, d, device=dev, generator=g)
    cnt = np.full(ne, 3, dtype=np.int32)
    out = torch.empty((ne, nt), device=dev)
    print("d=%d default  %.4f ms" % (d, kernel_ms(base, e, cnt, t, out)))
    for v in variants:
        h = handle(d, v)
        got = h.score_grid(e_s, cnt_s, t_s, znorm=(zm, zs))
        same = bool(torch.equal(got, ref))
        print("d=%d %-8s %.4f ms  bit-identical to default: %s (max diff %.3e)" %
              (d, v, kernel_ms(h, e, cnt, t, out), same, float((got - ref).abs().max())))
    print("d=%d default  %.4f ms (again)" % (d, kernel_ms(base, e, cnt, t, out)))
