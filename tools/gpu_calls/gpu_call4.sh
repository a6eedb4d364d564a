#!/bin/bash
# 2-GPU call: multi-GPU parity tests, bench lines (tgv256 weak, channel3 strong) with parity check + nvlink + phases
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
N=${1:-2}
python -m pytest tests/test_gpu_multi.py -q -k "two_gpus" > gpurun_out/r2d_gtest_n$N.log 2>&1; echo "gtest rc=$?" >> gpurun_out/r2d_gtest_n$N.log
tail -5 gpurun_out/r2d_gtest_n$N.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
$TR bench.py --gpus $N --steps 20 > gpurun_out/r2d_bench_n$N.json 2> gpurun_out/r2d_bench_n$N.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r2d_bench_n$N.err
$TR bench.py --gpus $N --steps 10 --workload channel3 --no-e2e --no-parity-check > gpurun_out/r2d_bench_channel3_n$N.json 2>> gpurun_out/r2d_bench_n$N.err; echo "bench rc=$?"
python bench.py --steps 10 --workload channel3 --no-e2e > gpurun_out/r2d_bench_channel3_n1.json 2>> gpurun_out/r2d_bench_n$N.err
for f in gpurun_out/r2d_bench*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[1], "ms/step", d["ms_per_step"], "Mcell/s", d["value"], "poisson", d["poisson_ms"], "sanity", d["sanity"]["ok"], "parity", d.get("parity_check",{}).get("ok"))
    print("   phases", {k: round(v,3) for k,v in d["phases_ms"].items()})
    if "nvlink" in d: print("   nvlink", d["nvlink"]["ms"], d["nvlink"]["gbs_per_direction"], d["nvlink"]["frac_of_900"])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
