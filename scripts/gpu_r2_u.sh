#!/bin/bash
export PLDA_B200_CUBLAS=0
echo "== shard + sinks tests"; timeout 900 python -m pytest tests/test_gpu_shard.py tests/test_gpu_sinks.py tests/test_gpu_plda.py -q -x --timeout 600 2>&1 | tail -n 8
echo "== ragged"; timeout 300 python scripts/r2_sink_probe.py ragged 2>&1 | tail -n 2
