#!/bin/bash
mkdir -p gpurun_out
export PLDA_B200_CUBLAS=0
echo "== epilogue A/B"; timeout 600 python scripts/epi_ab.py hybrid sector 2>&1 | tail -n 12
echo "== LDA tolerance probe"; PLDA_LDA_TOL=1e-3 timeout 600 python -m pytest tests/test_gpu_lda.py -q --timeout 300 2>&1 | tail -n 6
