#!/bin/bash
# ncu full captures (source-level) of the kernels below the 0.60 bar: strain_k, the four fftb_k passes, gauss2_k, correc_k
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
T=r2o
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'strain_k|fftb_k|gauss2_k' --launch-skip 14 -c 12 -o gpurun_out/${T}_full python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-phases > gpurun_out/${T}_ncu_full.log 2>&1
ls -la gpurun_out/${T}_*
