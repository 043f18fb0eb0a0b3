#!/bin/bash
mkdir -p gpurun_out
echo "== host overhead"; timeout 300 python scripts/host_overhead_probe.py 2>&1 | tail -n 5
echo "== sink probes"; for m in znorm hist trials ragged; do timeout 300 python scripts/r2_sink_probe.py $m 2>&1 | tail -n 3; done
echo "== ragged kernel time"; PLDA_B200_DBG=0 timeout 300 python - <<'PY'
import sys, os, ctypes as C
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from plda_b200 import PLDA, _ffi
d=200; rs=np.random.RandomState(0); q,_=np.linalg.qr(rs.randn(d,d))
p=PLDA(); p.set_model(np.full(d,0.5), q, 2.0*np.exp(-np.arange(d)/(0.15*d)))
e=torch.randn(10000,d,device="cuda"); t=torch.randn(10000,d,device="cuda"); out=torch.empty((10000,10000),device="cuda")
cnt=rs.randint(1,6,size=10000).astype(np.int32)
lib=_ffi.lib()
for name,c in (("ragged",cnt),("uniform",np.full(10000,3,np.int32))):
    for _ in range(3): p.score_grid(e,c,t,out=out)
    _ffi.check(lib.plda_profile_gemm(p._h,1))
    for _ in range(10): p.score_grid(e,c,t,out=out)
    ms,n=C.c_double(),C.c_int64(); _ffi.check(lib.plda_profile_collect(p._h,C.byref(ms),C.byref(n))); _ffi.check(lib.plda_profile_gemm(p._h,0))
    print(name,"gemm kernel ms",ms.value/n.value)
PY
echo "== full gpu tests"; timeout 1800 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "exit=$?"; tail -n 6 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -n 2
echo "== probes"; for cfg in "100000 200 1000 10 f32" "5000000 512 50000 5 f32"; do timeout 300 python scripts/r2_stats_probe.py $cfg 2>&1 | grep stats_ms; done
echo "== lda c5"; timeout 300 python scripts/run_lda_config.py 2>&1 | tail -n 2
