#!/bin/bash
# round 2: validate the 1024-thread Jacobi Gram phase and the tightened tolerances; EM and ragged timings
mkdir -p gpurun_out
export PLDA_B200_CUBLAS=0
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -n 12
echo "== probes"
for a in "100000 200 1000 float32" "100000 200 1000 float64" "1000000 256 10000 float32" "5000000 512 50000 float32"; do timeout 600 python scripts/r2_stats_probe.py $a 2>&1 | tail -n 1; done
echo "== ragged"; timeout 300 python scripts/r2_sink_probe.py ragged 2>&1 | tail -n 6
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 3
