#!/bin/bash
export PLDA_B200_CUBLAS=0
for m in default hybrid sector skip; do
  echo "== $m"
  if [ $m = default ]; then unset PLDA_B200_EPI; else export PLDA_B200_EPI=$m; fi
  PLDA_B200_DBG=1 timeout 300 python scripts/bench_gemm.py 10000 10000 200 20 2>&1 | tail -n 3
done
