#!/bin/bash
# round 2: ragged column terms inside the operands, all-warp Jacobi apply phase, EM phase profile
mkdir -p gpurun_out
export PLDA_B200_CUBLAS=0
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -n 8
echo "== ragged"; timeout 300 python scripts/r2_sink_probe.py ragged 2>&1 | tail -n 2
echo "== ragged epilogue variant"; PLDA_B200_RAGGED=epilogue timeout 300 python scripts/r2_sink_probe.py ragged 2>&1 | tail -n 2
echo "== EM phases C2"; PLDA_B200_EM_PROFILE=1 timeout 300 python scripts/r2_stats_probe.py 100000 200 1000 10 f32 2>&1 | grep -E "em phase|stats_ms" | tail -n 12
echo "== EM phases C4"; PLDA_B200_EM_PROFILE=1 timeout 300 python scripts/r2_stats_probe.py 5000000 512 50000 5 f32 2>&1 | grep -E "em phase|stats_ms" | tail -n 12
echo "== ragged launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_ragged_launches.csv python scripts/r2_sink_probe.py ragged > gpurun_out/ncu_ragged.log 2>&1; echo "exit=$?"
