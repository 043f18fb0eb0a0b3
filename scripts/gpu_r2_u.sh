#!/bin/bash
export PLDA_B200_CUBLAS=0
echo "== shard tests"; timeout 900 python -m pytest tests/test_gpu_shard.py -q -x --timeout 600 2>&1 | tail -n 3
for i in 1 2 3; do
echo "== headline (N=1)"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --skip-em --headline-only 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(j['value'], j['ms_per_step'], j['roofline']['kernel_ms'])"
done
timeout 300 python scripts/bench_gemm.py 10000 10000 200 20 2>&1 | tail -n 1
