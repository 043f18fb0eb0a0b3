#!/bin/bash
export PLDA_B200_CUBLAS=0
for ts in 0 1; do
  PLDA_B200_TS=$ts PLDA_B200_DBG=1 timeout 300 python scripts/bench_gemm.py 10000 10000 200 20 2>&1 | tail -n 3
  PLDA_B200_TS=$ts PLDA_B200_DBG=1 PLDA_B200_EPI=skip timeout 300 python scripts/bench_gemm.py 10000 10000 200 20 2>&1 | tail -n 3
done
