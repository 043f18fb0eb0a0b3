#!/bin/bash
mkdir -p gpurun_out
echo "== scatter + fit tests"; timeout 900 python -m pytest tests/test_gpu_sinks.py tests/test_gpu_plda.py -x -q --timeout 300 2>&1 | tail -n 8
echo "== probes"; for cfg in "100000 200 1000 10 f32" "100000 200 1000 10 f64" "1000000 256 10000 5 f32" "5000000 512 50000 5 f32"; do timeout 300 python scripts/r2_stats_probe.py $cfg 2>&1 | grep stats_ms; done
echo "== EIG grid barrier A/B"; PLDA_B200_EIG=grid timeout 300 python scripts/r2_stats_probe.py 100000 200 1000 10 f32 2>&1 | grep stats_ms
echo "== EIG exact A/B"; PLDA_B200_EIG_EXACT=1 timeout 300 python scripts/r2_stats_probe.py 100000 200 1000 10 f32 2>&1 | grep stats_ms
echo "== sweeps"; PLDA_B200_DBG=1 timeout 300 python scripts/fit_once.py 200 1000 100 10 2>&1 | grep -E "sweeps" | tail -n 11 | tr '\n' ';'; echo
PLDA_B200_DBG=1 timeout 300 python scripts/fit_once.py 512 2000 50 6 2>&1 | grep -E "sweeps" | tail -n 7 | tr '\n' ';'; echo
echo "== ncu full scatter"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:scatter_syrk -s 2 -c 1 -o gpurun_out/r02_prof_scatter2 -f python scripts/r2_stats_probe.py 2000000 512 20000 1 f32 > gpurun_out/ncu_scatter.log 2>&1; echo "exit=$?"
echo "== ncu full jacobi"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:block_jacobi -s 8 -c 1 -o gpurun_out/r02_prof_jacobi -f python scripts/fit_once.py 200 1000 100 10 > gpurun_out/ncu_jacobi.log 2>&1; echo "exit=$?"
echo "== ncu launch list fit C2"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_fit_launches.csv python scripts/fit_once.py 200 1000 100 10 > gpurun_out/ncu_fit.log 2>&1; echo "exit=$?"
