#!/bin/bash
# first GPU call of round 2: parity of both library variants, kernel tables of both, bench line, BASELINE-size parity
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2a_gpu.txt
nproc >> gpurun_out/r2a_gpu.txt
python -m pytest tests -m "gpu and not fullsize" -x -q > gpurun_out/r2a_gtest.log 2>&1; echo "gtest rc=$?" >> gpurun_out/r2a_gtest.log
tail -3 gpurun_out/r2a_gtest.log
CALES_B200_ARITH=strict python tools/kbench.py > gpurun_out/r2a_kbench_strict.txt 2>&1
CALES_B200_ARITH=fma python tools/kbench.py > gpurun_out/r2a_kbench_fma.txt 2>&1
paste gpurun_out/r2a_kbench_strict.txt gpurun_out/r2a_kbench_fma.txt | cut -c1-200
python bench.py --steps 20 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r2a_bench.err
python bench.py --steps 20 --arith strict --no-cpu-baseline --no-e2e > gpurun_out/r2a_bench_strict.json 2>> gpurun_out/r2a_bench.err
python -m pytest tests -m "gpu and fullsize" -q -n 3 > gpurun_out/r2a_fullsize.log 2>&1; echo "fullsize rc=$?" >> gpurun_out/r2a_fullsize.log
tail -15 gpurun_out/r2a_fullsize.log
