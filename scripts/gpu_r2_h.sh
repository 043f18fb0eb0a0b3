#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
run() { tag=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 5 --headline-only --no-cpu "$@" > gpurun_out/bench_g${N}_$tag.json 2> gpurun_out/bench_g${N}_$tag.err; echo "$tag exit=$?"; python -c "
import json,sys
j=json.loads(open('gpurun_out/bench_g${N}_$tag.json').read().strip().splitlines()[-1])
print('$tag', 'value %.4g' % j['value'], 'ms/step %.4f' % j['ms_per_step'], 'gemm kernel ms %.4f' % j['roofline']['kernel_ms'], 'e2e %.4g' % j['e2e']['value'], 'e2e_trials %.4g' % j['e2e_trials']['value'], j['parallelism'][:40])
"; }
nvidia-smi topo -m | head -n 12
run peer
PLDA_B200_FENCE=thread run peer_fence_thread
run nccl --nccl-allgather
