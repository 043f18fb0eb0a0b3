#!/bin/bash
mkdir -p gpurun_out
echo "== kernel + fit tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_sinks.py tests/test_gpu_plda.py -x -q --timeout 300 2>&1 | tail -n 8
echo "== probes"; for cfg in "100000 200 1000 10 f32" "100000 200 1000 10 f64" "1000000 256 10000 5 f32" "5000000 512 50000 5 f32"; do timeout 300 python scripts/r2_stats_probe.py $cfg 2>&1 | grep stats_ms; done
echo "== legacy chol A/B"; PLDA_B200_CHOL=legacy timeout 300 python scripts/r2_stats_probe.py 100000 200 1000 10 f32 2>&1 | grep stats_ms
echo "== sweeps"; PLDA_B200_DBG=1 timeout 300 python scripts/fit_once.py 200 1000 100 10 2>&1 | grep -E "sweeps" | tail -n 11 | sed 's/plda_b200: joint_diagonalise//' | tr '\n' ';'; echo
PLDA_B200_DBG=1 timeout 300 python scripts/fit_once.py 512 2000 50 6 2>&1 | grep -E "sweeps" | tail -n 7 | sed 's/plda_b200: joint_diagonalise//' | tr '\n' ';'; echo
echo "== trace C4"; PLDA_B200_TRACE=1 timeout 300 python scripts/r2_stats_probe.py 5000000 512 50000 1 f32 2>&1 | grep -E "plda_b200 fit" | tail -n 6
echo "== ncu launch list fit C2"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_fit_launches2.csv python scripts/fit_once.py 200 1000 100 10 > gpurun_out/ncu_fit.log 2>&1; echo "exit=$?"
echo "== full gpu tests"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "exit=$?"; tail -n 6 gpurun_out/pytest_gpu.log
