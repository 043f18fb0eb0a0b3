#!/bin/bash
echo "== fit + kernel tests"; timeout 900 python -m pytest tests/test_gpu_plda.py tests/test_gpu_scale.py tests/test_gpu_lda.py tests/test_gpu_kernels.py -q -x --timeout 600 2>&1 | tail -n 3
echo "== sweeps"; PLDA_B200_DBG=1 timeout 300 python scripts/em_bench_probe.py 2>&1 | grep -E "sweeps" | tail -n 11 | tr '\n' ';'
echo; PLDA_B200_DBG=1 timeout 300 python scripts/em_bench_probe.py 2>&1 | grep -E "cholesky" | tail -n 1
echo "== EM (bench data)"; timeout 300 python scripts/em_bench_probe.py 2>&1 | tail -n 1
