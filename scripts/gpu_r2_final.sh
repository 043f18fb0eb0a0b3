#!/bin/bash
# round 2 closing pass on one GPU: full GPU test suite, smoke, both bench arms
mkdir -p gpurun_out
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -n 6
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
echo "== bench ours"; timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "exit=$?"; tail -n 3 gpurun_out/bench.err; cut -c 1-300 gpurun_out/bench.json
echo "== bench reference"; timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "exit=$?"; cut -c 1-200 gpurun_out/bench_ref.json
echo "== default bench invocation (no flags)"; timeout 1200 python bench.py 2>/dev/null | cut -c 1-200
