#!/bin/bash
# N-GPU call: multi-GPU parity tests + bench lines, A/B of the pipelined solver exchange and the fused halo kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
N=${1:-2}; TAG=${2:-r2f}
python -m pytest tests/test_gpu_multi.py -q -k "two_gpus" > gpurun_out/${TAG}_gtest_n$N.log 2>&1; echo "gtest rc=$?" >> gpurun_out/${TAG}_gtest_n$N.log
tail -5 gpurun_out/${TAG}_gtest_n$N.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
$TR bench.py --gpus $N --steps 20 --no-e2e > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err; echo "bench rc=$?"; grep -v "^$\|\*\*\*\|OMP_NUM" gpurun_out/${TAG}_bench_n$N.err | tail -5
CALES_SOLVER_PIPE=0 $TR bench.py --gpus $N --steps 20 --no-e2e --no-parity-check > gpurun_out/${TAG}_bench_nopipe_n$N.json 2>> gpurun_out/${TAG}_bench_n$N.err
CALES_HALO_FUSED=0 $TR bench.py --gpus $N --steps 20 --no-e2e --no-parity-check > gpurun_out/${TAG}_bench_nohalofuse_n$N.json 2>> gpurun_out/${TAG}_bench_n$N.err
CALES_SOLVER_CHUNKS=2 $TR bench.py --gpus $N --steps 20 --no-e2e --no-parity-check --no-phases > gpurun_out/${TAG}_bench_chunks2_n$N.json 2>> gpurun_out/${TAG}_bench_n$N.err
$TR bench.py --gpus $N --steps 10 --workload channel3 --no-e2e --no-parity-check > gpurun_out/${TAG}_bench_channel3_n$N.json 2>> gpurun_out/${TAG}_bench_n$N.err
for f in gpurun_out/${TAG}_bench*.json; do python - "$f" <<'PY'
import json,sys
try:
    l=[x for x in open(sys.argv[1]) if x.startswith("{")][-1]
    d=json.loads(l); print(sys.argv[1], "ms/step", round(d["ms_per_step"],3), "Mcell/s", round(d["value"]), "poisson", round(d["poisson_ms"],3), "sanity", d["sanity"]["ok"], "parity", d.get("parity_check",{}).get("ok"))
    print("   phases", {k: round(v,3) for k,v in d["phases_ms"].items()})
    if "nvlink" in d: print("   nvlink", d["nvlink"]["ms"], d["nvlink"]["gbs_per_direction"], d["nvlink"]["frac_of_900"])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
