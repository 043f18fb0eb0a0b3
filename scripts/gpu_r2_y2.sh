#!/bin/bash
PLDA_B200_DBG=1 timeout 300 python scripts/em_bench_probe.py 2>&1 | grep -E "cholesky" | tail -n 3
PLDA_B200_DBG=1 timeout 300 python scripts/r2_stats_probe.py 500000 512 5000 3 f32 2>&1 | grep -E "cholesky" | tail -n 2
