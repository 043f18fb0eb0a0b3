"""Micro-benchmark of the score-grid GEMM kernel (CUDA-event timed inside the library).
usage: python scripts/bench_gemm.py [ne nt d reps]   (env PLDA_B200_EPI = tma|direct|skip)"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plda_b200 import PLDA, _ffi  # noqa: E402

ne, nt, d, reps = (int(a) for a in (sys.argv[1:5] + ["10000", "10000", "200", "20"][len(sys.argv) - 1:]))
rng = np.random.RandomState(0)
p = PLDA()
q, _ = np.linalg.qr(rng.randn(d, d))
p.set_model(rng.randn(d), q, np.sort(2.0 * np.exp(-np.arange(d) / (0.15 * d)))[::-1].copy())
dev = torch.device("cuda", 0)
e = torch.randn(ne, d, device=dev)
t = torch.randn(nt, d, device=dev)
cnt = np.full(ne, 3, dtype=np.int32)
ldo = (nt + 3) // 4 * 4
out = torch.empty((ne, ldo), device=dev)
lib = _ffi.lib()
for _ in range(3):
    p.score_grid(e, cnt, t, out=out[:, :nt])
torch.cuda.synchronize()
_ffi.check(lib.plda_profile_gemm(p._h, 1))
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
_st = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(_st)
_ffi.check(lib.plda_set_stream(p._h, C.c_void_p(_st.cuda_stream)))
ev0.record()
for _ in range(reps):
    p.score_grid(e, cnt, t, out=out[:, :nt])
ev1.record()
torch.cuda.synchronize()
ms, n = C.c_double(), C.c_int64()
_ffi.check(lib.plda_profile_collect(p._h, C.byref(ms), C.byref(n)))
k_ms = ms.value / n.value
k16 = (d + 15) // 16 * 16
print("mode=%s ne=%d nt=%d d=%d  gemm %.4f ms  step %.4f ms  %.3e trials/s (kernel)  issued %.0f TFLOP/s  write %.0f GB/s"
      % (os.environ.get("PLDA_B200_EPI", "default") + "/" + os.environ.get("PLDA_B200_GEMM", "2cta") + "/ts" + os.environ.get("PLDA_B200_TS", "-"), ne, nt, d, k_ms, ev0.elapsed_time(ev1) / reps, ne * nt / k_ms * 1e3,
         3 * 2 * k16 * ne * nt / k_ms / 1e9, 4.0 * ne * nt / k_ms / 1e6))

if os.environ.get("PLDA_B200_DBG") == "1":
    c = np.zeros(32, dtype=np.int64)
    _ffi.check(lib.plda_debug_counters(p._h, _ffi.ptr(c), 32))
    names = ["prod_wait_empty", "prod_total", "mma_wait_full", "mma_wait_tempty", "mma_total", "tiles",
             "epi0_wait_tfull", "epi0_wait_store", "epi0_total", "epi7_wait_tfull", "epi7_wait_store", "epi7_total", "epi0_tmem_load",
             "mma_wait_a"]
    for b in range(2):
        print("  cta%d: " % b + "  ".join("%s=%d" % (n, c[b * 16 + i]) for i, n in enumerate(names)))

if os.environ.get("PLDA_B200_CUBLAS", "1") == "1":
    # Library yardstick at EQUAL ISSUED WORK: one cuBLAS bf16 GEMM with K = 3 * K16 (the three split products laid end
    # to end), no fused score terms.  bf16 output halves its store traffic; fp32 output where torch exposes it.
    kk = 3 * k16
    a = torch.randn(ne, kk, device=dev, dtype=torch.bfloat16)
    b = torch.randn(nt, kk, device=dev, dtype=torch.bfloat16)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timed(fn):
        for _ in range(3):
            fn()
        tot = 0.0
        for _ in range(reps):
            flush.zero_()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            fn()
            s1.record()
            torch.cuda.synchronize()
            tot += s0.elapsed_time(s1)
        return tot / reps
    ms16 = timed(lambda: torch.mm(a, b.t()))
    print("cuBLAS bf16 [%d x %d x %d] -> bf16 out: %.4f ms  %.0f TFLOP/s" % (ne, nt, kk, ms16, 2.0 * kk * ne * nt / ms16 / 1e9))
    try:
        o32 = torch.empty((ne, nt), device=dev, dtype=torch.float32)
        ms32 = timed(lambda: torch.mm(a, b.t(), out_dtype=torch.float32, out=o32))
        print("cuBLAS bf16 [%d x %d x %d] -> fp32 out: %.4f ms  %.0f TFLOP/s" % (ne, nt, kk, ms32, 2.0 * kk * ne * nt / ms32 / 1e9))
    except Exception as exc:  # noqa: BLE001
        print("cuBLAS fp32-out variant not available in this torch:", type(exc).__name__, str(exc)[:100])
    # fp32 (TF32 off) and TF32 GEMMs on the unsplit operands, for the precision / speed trade-off table
    a32, b32 = torch.randn(ne, d, device=dev), torch.randn(nt, d, device=dev)
    torch.backends.cuda.matmul.allow_tf32 = False
    ms_f32 = timed(lambda: torch.mm(a32, b32.t()))
    torch.backends.cuda.matmul.allow_tf32 = True
    ms_tf32 = timed(lambda: torch.mm(a32, b32.t()))
    print("cuBLAS fp32 (no TF32) [K=%d] -> fp32: %.4f ms ; TF32 -> fp32: %.4f ms" % (d, ms_f32, ms_tf32))
