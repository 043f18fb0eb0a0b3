#!/bin/bash
for i in 1 2 3; do
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --skip-em --headline-only 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(j['value'], j['ms_per_step'], j['ms_per_step_median'], j['ms_per_step_max'], j['roofline']['kernel_ms'])"
done
