#!/bin/bash
# round 2 final single-GPU pass: tests, smoke, ncu captures of every hot kernel, launch lists, bench both arms
mkdir -p gpurun_out
export PLDA_B200_CUBLAS=0
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -n 6
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
NCU="ncu --set full --clock-control none --import-source on -f"
cap() { name=$1; regex=$2; skip=$3; shift 3; timeout 600 $NCU -k regex:$regex -s $skip -c 1 -o gpurun_out/r02_prof_$name "$@" > gpurun_out/ncu_$name.log 2>&1; echo "$name exit=$?"; }
cap gemm gemm_bf16x3 4 python scripts/bench_gemm.py 10000 10000 200 3
cap gemm_d512 gemm_bf16x3 4 python scripts/bench_gemm.py 20000 20000 512 3
cap prep score_prep_uniform 4 python scripts/bench_gemm.py 10000 10000 200 3
cap gemm_ragged gemm_bf16x3 2 python scripts/r2_sink_probe.py ragged
cap prep_ragged score_prep_grouped_vec 2 python scripts/r2_sink_probe.py ragged
cap gemm_mom gemm_bf16x3 2 python scripts/r2_sink_probe.py znorm
cap trials score_trials_kernel 2 python scripts/r2_sink_probe.py trials
cap chol chol_inverse_cluster 6 python scripts/fit_once.py 200 1000 100 10
cap jacobi block_jacobi 8 python scripts/fit_once.py 200 1000 100 10
cap scatter_c2 scatter_syrk 2 python scripts/fit_once.py 200 1000 100 2
cap scatter scatter_syrk 2 python scripts/r2_stats_probe.py 2000000 512 20000 1 f32
echo "== launch lists"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --skip-em --headline-only > gpurun_out/ncu_launch.log 2>&1; echo "exit=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_fit_launches.csv python scripts/fit_once.py 200 1000 100 10 > gpurun_out/ncu_fit.log 2>&1; echo "exit=$?"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_stats_launches.csv python scripts/r2_stats_probe.py 2000000 512 20000 1 f32 > gpurun_out/ncu_stats.log 2>&1; echo "exit=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_ragged_launches.csv python scripts/r2_sink_probe.py ragged > gpurun_out/ncu_ragged.log 2>&1; echo "exit=$?"
echo "== bench ours"; timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "exit=$?"; tail -n 3 gpurun_out/bench.err; cut -c 1-400 gpurun_out/bench.json
echo "== bench reference"; timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "exit=$?"; cut -c 1-300 gpurun_out/bench_ref.json
echo "== sanitizer (small fit: Cholesky / Jacobi cluster kernels, fused stats kernel)"
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis python scripts/fit_once.py 72 60 10 2 > gpurun_out/sanitizer_race_fit.log 2>&1; echo "racecheck exit=$?"; grep -E "RACECHECK SUMMARY|ERROR SUMMARY|hazard" gpurun_out/sanitizer_race_fit.log | tail -n 5
timeout 900 compute-sanitizer --tool memcheck python scripts/fit_once.py 72 60 10 2 > gpurun_out/sanitizer_mem_fit.log 2>&1; echo "memcheck exit=$?"; grep -E "ERROR SUMMARY" gpurun_out/sanitizer_mem_fit.log | tail -n 2
