#!/bin/bash
# round 2: ragged FAST epilogue, sorted-label fast path of the segment builder, Jacobi on 1024 threads -- tests + timings
mkdir -p gpurun_out
export PLDA_B200_CUBLAS=0
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -n 8
echo "== probes"
for cfg in "100000 200 1000 10 f32" "100000 200 1000 10 f64" "1000000 256 10000 5 f32" "5000000 512 50000 5 f32"; do timeout 600 python scripts/r2_stats_probe.py $cfg 2>&1 | grep stats_ms; done
echo "== ragged"; timeout 300 python scripts/r2_sink_probe.py ragged 2>&1 | tail -n 2
echo "== ragged generic epilogue"; PLDA_B200_RAGGED_EPI=generic timeout 300 python scripts/r2_sink_probe.py ragged 2>&1 | tail -n 2
echo "== ragged launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_ragged_launches.csv python scripts/r2_sink_probe.py ragged > gpurun_out/ncu_ragged.log 2>&1; echo "exit=$?"
echo "== fit launch lists"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_fit_launches.csv python scripts/fit_once.py 200 1000 100 10 > gpurun_out/ncu_fit.log 2>&1; echo "exit=$?"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_stats_launches.csv python scripts/r2_stats_probe.py 2000000 512 20000 1 f32 > gpurun_out/ncu_stats.log 2>&1; echo "exit=$?"
