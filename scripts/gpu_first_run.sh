#!/bin/bash
# First-contact GPU run: every stage in its own process under a timeout so a trapped or hung kernel
# cannot take the rest (or the box) with it.  Output -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { # name, timeout, cmd...
  local name=$1; shift; local t=$1; shift
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $t "$@" > gpurun_out/$name.log 2>&1
  echo "exit=$? ($name)" | tee -a gpurun_out/summary.txt
  tail -n 15 gpurun_out/$name.log | tee -a gpurun_out/summary.txt
}
run gemm_tma 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "gemm" --timeout 120
PLDA_B200_EPI=direct run gemm_direct 300 env PLDA_B200_EPI=direct python -m pytest tests/test_gpu_kernels.py -q -k "gemm" --timeout 120
run linalg 300 python -m pytest tests/test_gpu_kernels.py -q -k "not gemm" --timeout 120
run plda 600 python -m pytest tests/test_gpu_plda.py -q --timeout 200
run lda 300 python -m pytest tests/test_gpu_lda.py -q --timeout 120
