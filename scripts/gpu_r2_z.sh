#!/bin/bash
mkdir -p gpurun_out
echo "== memcheck: ragged grids (single GPU + sharded)"
timeout 1200 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_shard.py tests/test_gpu_sinks.py -q -x -k "ragged" --timeout 1000 > gpurun_out/sanitizer_mem_ragged.log 2>&1; echo "exit=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_mem_ragged.log | tail -n 3
echo "== memcheck: eig forms"
timeout 1200 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_kernels.py -q -x -k "jacobi or cholesky" --timeout 1000 > gpurun_out/sanitizer_mem_eig.log 2>&1; echo "exit=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitizer_mem_eig.log | tail -n 3
