#!/bin/bash
# closing tree: re-capture the kernels that changed after the last full profile pass (Gram kernel: watcher warp;
# Cholesky: hoisted tile load) and the fit launch list
mkdir -p gpurun_out
export PLDA_B200_CUBLAS=0
NCU="ncu --set full --clock-control none --import-source on -f"
cap() { name=$1; regex=$2; skip=$3; shift 3; timeout 600 $NCU -k regex:$regex -s $skip -c 1 -o gpurun_out/r02_prof_$name "$@" > gpurun_out/ncu_$name.log 2>&1; echo "$name exit=$?"; }
cap gemm gemm_bf16x3 4 python scripts/bench_gemm.py 10000 10000 200 3
cap chol chol_inverse_cluster 6 python scripts/fit_once.py 200 1000 100 10
cap jacobi block_jacobi 8 python scripts/fit_once.py 200 1000 100 10
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_fit_launches.csv python scripts/fit_once.py 200 1000 100 10 > gpurun_out/ncu_fit.log 2>&1; echo "exit=$?"
