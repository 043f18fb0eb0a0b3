#!/bin/bash
PLDA_B200_EM_PROFILE=1 PLDA_B200_DBG=1 timeout 300 python scripts/em_bench_probe.py 2>&1 | grep -E "sweeps|em phase|stats" | tail -n 24
