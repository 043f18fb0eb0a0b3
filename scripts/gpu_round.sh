#!/bin/bash
# Round-end style GPU pass: parity tests, smoke, bench (both arms), ncu launch lists + full captures.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "exit=$?"; tail -n 4 gpurun_out/pytest_gpu.log
echo "== smoke" ; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -n 2
echo "== bench ours" ; timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "exit=$?"; cut -c 1-600 gpurun_out/bench.json; tail -n 3 gpurun_out/bench.err
echo "== bench reference" ; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "exit=$?"; cut -c 1-300 gpurun_out/bench_ref.json
echo "== ncu launch list (bench)" ; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --skip-em > gpurun_out/ncu_launch.log 2>&1; echo "exit=$?"
echo "== ncu launch list (fit)" ; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/fit_launches.csv python scripts/fit_once.py 200 1000 100 5 > gpurun_out/ncu_fit.log 2>&1; echo "exit=$?"
echo "== ncu full gemm d=200" ; timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16x3 -s 4 -c 1 -o gpurun_out/prof_gemm -f python scripts/bench_gemm.py 10000 10000 200 3 > gpurun_out/ncu_full.log 2>&1; echo "exit=$?"
echo "== ncu full producer kernel" ; timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_prep -s 4 -c 1 -o gpurun_out/prof_prep -f python bench.py --steps 3 --warmup 3 --no-cpu --skip-em > gpurun_out/ncu_prep.log 2>&1; echo "exit=$?"
echo "== ncu full gemm d=512" ; timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16x3 -s 4 -c 1 -o gpurun_out/prof_gemm_d512 -f python scripts/bench_gemm.py 20000 20000 512 3 > gpurun_out/ncu_full2.log 2>&1; echo "exit=$?"
ls gpurun_out | tr '\n' ' '
