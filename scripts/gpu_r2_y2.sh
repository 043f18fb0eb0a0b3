#!/bin/bash
for v in 0 1; do
echo "== jacobi phases (bench data, CTA 2) variant $v"; PLDA_B200_JACOBI=$v PLDA_B200_DBG=1 timeout 300 python scripts/em_bench_probe.py 2>&1 | grep -E "jacobi" | tail -n 2
PLDA_B200_JACOBI=$v timeout 300 python scripts/em_bench_probe.py 2>&1 | tail -n 1
done
echo "== tests with variant 1"; PLDA_B200_JACOBI=1 timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_plda.py -q -x --timeout 600 2>&1 | tail -n 3
echo "== eig tests default"; timeout 900 python -m pytest tests/test_gpu_kernels.py -q -x --timeout 600 2>&1 | tail -n 3
