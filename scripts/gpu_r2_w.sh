#!/bin/bash
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 scripts/pcie_probe_ranks.py 2>&1 | grep -v "OMP_NUM_THREADS\|\*\*\*\*\|^$" | tail -n 6
nvidia-smi topo -m 2>/dev/null | head -n 14
