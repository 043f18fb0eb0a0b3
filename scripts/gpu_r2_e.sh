#!/bin/bash
mkdir -p gpurun_out
echo "== kernel + fit tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_plda.py tests/test_gpu_scale.py -x -q --timeout 600 2>&1 | tail -n 8
echo "== probes"; for cfg in "100000 200 1000 10 f32" "100000 200 1000 10 f64" "1000000 256 10000 5 f32" "5000000 512 50000 5 f32"; do timeout 300 python scripts/r2_stats_probe.py $cfg 2>&1 | grep stats_ms; done
echo "== ncu launch list fit C2"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_fit_launches3.csv python scripts/fit_once.py 200 1000 100 10 > gpurun_out/ncu_fit.log 2>&1; echo "exit=$?"
