#!/bin/bash
# one headline-only scaling point: bash scripts/gpu_r2_h3.sh N
mkdir -p gpurun_out
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 50 --warmup 5 --headline-only --no-cpu > gpurun_out/bench_g${N}_head.json 2> gpurun_out/bench_g${N}_head.err; echo "exit=$?"; python -c "
import json
j=json.loads(open('gpurun_out/bench_g${N}_head.json').read().strip().splitlines()[-1])
print({k:j[k] for k in ('value','ms_per_step','ms_per_step_median','ms_per_step_max','n_gpus') if k in j}, 'gemm kernel ms', j['roofline']['kernel_ms'])
"
