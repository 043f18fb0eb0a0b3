#!/bin/bash
# 2-GPU validation: multi-GPU tests + bench at N=2 (sharded side records)
mkdir -p gpurun_out
nvidia-smi -L
echo "== multi tests"; timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_kernels.py -q --timeout 600 2>&1 | tail -n 5
echo "== bench N=2"; timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_g2.json 2> gpurun_out/bench_g2.err; echo "exit=$?"; tail -n 8 gpurun_out/bench_g2.err; python -c "
import json
j=json.loads(open('gpurun_out/bench_g2.json').read().strip().splitlines()[-1])
print({k:j[k] for k in ('value','ms_per_step','n_gpus','parallelism') if k in j})
print('e2e', j['e2e']['value'], 'e2e_trials', j['e2e_trials']['value'])
print(json.dumps(j.get('sharded'), indent=1)[:5000])
"
echo "== bench N=2 reference"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 5 --warmup 2 | cut -c 1-300
