#!/bin/bash
# round 2: register-tiled tile products in the Cholesky cluster kernel, approximate angle math in the Jacobi sweep
mkdir -p gpurun_out
export PLDA_B200_CUBLAS=0
echo "== kernel + fit tests"; timeout 1500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_plda.py tests/test_gpu_scale.py -q -x --timeout 900 2>&1 | tail -n 6
echo "== EM phases C2"; PLDA_B200_EM_PROFILE=1 timeout 300 python scripts/r2_stats_probe.py 100000 200 1000 10 f32 2>&1 | grep -E "em phase|stats_ms" | tail -n 9
echo "== EM phases C4"; PLDA_B200_EM_PROFILE=1 timeout 300 python scripts/r2_stats_probe.py 5000000 512 50000 5 f32 2>&1 | grep -E "em phase|stats_ms" | tail -n 9
echo "== C3"; timeout 300 python scripts/r2_stats_probe.py 1000000 256 10000 5 f32 2>&1 | grep stats_ms
