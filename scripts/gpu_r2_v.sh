#!/bin/bash
# round 2 final multi-GPU pass: bash scripts/gpu_r2_v.sh N   (multi-GPU tests at N = 2; bench with the sharded records)
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi -L | head -n 8
if [ "$N" = "2" ]; then
  echo "== multi tests"; timeout 900 python -m pytest tests/test_gpu_multi.py -q --timeout 600 2>&1 | tail -n 5
fi
timeout 1800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_g${N}.json 2> gpurun_out/bench_g${N}.err; echo "exit=$?"; tail -n 6 gpurun_out/bench_g${N}.err | cut -c 1-300; python -c "
import json
j=json.loads(open('gpurun_out/bench_g${N}.json').read().strip().splitlines()[-1])
print({k:j[k] for k in ('value','ms_per_step','ms_per_step_median','ms_per_step_max','n_gpus') if k in j})
print('gemm kernel ms', j['roofline']['kernel_ms'], 'e2e', j['e2e']['value'], j['e2e']['ms_per_step'], 'e2e_trials', j['e2e_trials']['value'], j['e2e_trials']['ms_per_step'])
print(json.dumps(j.get('sharded'), indent=1)[:7000])
"
