#!/bin/bash
# round 2, call A: new kernels' parity + first measurements
mkdir -p gpurun_out
echo "== new tests"; timeout 900 python -m pytest tests/test_gpu_sinks.py -x -q --timeout 300 2>&1 | tail -n 25
echo "== stats probe (fused)"; for cfg in "100000 200 1000 10 f32" "100000 200 1000 10 f64" "1000000 256 10000 3 f32" "5000000 512 50000 2 f32"; do timeout 300 python scripts/r2_stats_probe.py $cfg 2>&1 | tail -n 2; done
echo "== stats probe (legacy)"; for cfg in "100000 200 1000 10 f32" "5000000 512 50000 2 f32"; do PLDA_B200_STATS=legacy timeout 300 python scripts/r2_stats_probe.py $cfg 2>&1 | tail -n 2; done
echo "== jacobi sweeps"; PLDA_B200_DBG=1 timeout 300 python scripts/fit_once.py 200 1000 100 10 2>&1 | grep -E "sweeps|stats" | tail -n 14
echo "== ragged"; timeout 300 python scripts/bench_ragged.py 2>&1 | tail -n 1; PLDA_B200_RAGGED=scalar timeout 300 python scripts/bench_ragged.py 2>&1 | tail -n 1
echo "== full gpu tests"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/pytest_gpu.log 2>&1; echo "exit=$?"; tail -n 15 gpurun_out/pytest_gpu.log
