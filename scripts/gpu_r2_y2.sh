#!/bin/bash
for v in 0 1 2 3; do
echo "== jacobi phases (bench data, CTA 1) variant $v"; PLDA_B200_JACOBI=$v PLDA_B200_DBG=1 timeout 300 python scripts/em_bench_probe.py 2>&1 | grep -E "jacobi" | tail -n 1
PLDA_B200_JACOBI=$v timeout 300 python scripts/em_bench_probe.py 2>&1 | tail -n 1
done
echo "== tests with variant 3"; PLDA_B200_JACOBI=3 timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_plda.py -q -x --timeout 600 2>&1 | tail -n 3
