#!/bin/bash
export PLDA_B200_CUBLAS=0
echo "== kernel + fit tests"; timeout 1500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_plda.py tests/test_gpu_scale.py tests/test_gpu_lda.py -q -x --timeout 900 2>&1 | tail -n 6
echo "== jacobi phases (bench data)"; PLDA_B200_DBG=1 timeout 300 python scripts/em_bench_probe.py 2>&1 | grep -E "jacobi" | tail -n 2
echo "== EM (bench data)"; timeout 300 python scripts/em_bench_probe.py 2>&1 | tail -n 1
echo "== EM (bench data), one CTA per pair"; PLDA_B200_JACOBI_TEAM=1 timeout 300 python scripts/em_bench_probe.py 2>&1 | tail -n 1
echo "== EM phases C2 probe data"; PLDA_B200_EM_PROFILE=1 timeout 300 python scripts/r2_stats_probe.py 100000 200 1000 10 f32 2>&1 | grep -E "eigensolver|stats_ms" | tail -n 2
echo "== EM C3"; timeout 300 python scripts/r2_stats_probe.py 1000000 256 10000 10 f32 2>&1 | grep stats_ms
echo "== EM C4"; timeout 300 python scripts/r2_stats_probe.py 5000000 512 50000 10 f32 2>&1 | grep stats_ms
