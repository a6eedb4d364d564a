#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
python -m pytest tests -m "gpu and not fullsize" -q > gpurun_out/r2b_gtest.log 2>&1; echo "gtest rc=$?" >> gpurun_out/r2b_gtest.log
tail -5 gpurun_out/r2b_gtest.log
python bench.py --steps 20 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"; tail -c 800 gpurun_out/r2b_bench.err
CALES_B200_FUSED=0 python bench.py --steps 20 --no-cpu-baseline --no-e2e --no-phases > gpurun_out/r2b_bench_unfused.json 2>> gpurun_out/r2b_bench.err
CALES_B200_GRAPH=0 python bench.py --steps 20 --no-cpu-baseline --no-e2e --no-phases > gpurun_out/r2b_bench_nograph.json 2>> gpurun_out/r2b_bench.err
python bench.py --workload channel1 --steps 100 --no-e2e > gpurun_out/r2b_bench_channel1.json 2>> gpurun_out/r2b_bench.err
CALES_B200_GRAPH=0 python bench.py --workload channel1 --steps 100 --no-e2e --no-phases > gpurun_out/r2b_bench_channel1_nograph.json 2>> gpurun_out/r2b_bench.err
CALES_B200_FUSED=0 python bench.py --workload channel1 --steps 100 --no-e2e --no-phases > gpurun_out/r2b_bench_channel1_unfused.json 2>> gpurun_out/r2b_bench.err
for f in gpurun_out/r2b_bench*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], d["ms_per_step"], d["value"], d["gpu_launches"], d["sanity"]["ok"], d["kernels"].get("rk(mom+update+forcing)",{}).get("ms"))
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
