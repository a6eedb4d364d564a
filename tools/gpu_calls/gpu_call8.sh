#!/bin/bash
# 8-GPU call: 4- and 8-GPU parity tests, weak-scaled bench (A/B of the exchange paths), BASELINE config 5 (1024x512x512) and config 4
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
TAG=r2i
timeout 600 python -m pytest tests/test_gpu_multi.py -q -k "four_gpus or eight_gpus" > gpurun_out/${TAG}_gtest_n8.log 2>&1; echo "gtest rc=$?" >> gpurun_out/${TAG}_gtest_n8.log
tail -4 gpurun_out/${TAG}_gtest_n8.log
run() { # name nprocs env... -- args
  local name=$1 np=$2; shift 2
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29540 bench.py --gpus $np "$@" > gpurun_out/${TAG}_${name}.json 2> gpurun_out/${TAG}_${name}.err
  echo "$name rc=$?"
}
run tgv_n8 8 X=1 -- --steps 20
run tgv_n8_pipe0 8 CALES_SOLVER_PIPE=0 -- --steps 20 --no-e2e --no-parity-check
run tgv_n8_pipe1c4 8 CALES_SOLVER_PIPE=1 CALES_SOLVER_CHUNKS=4 -- --steps 20 --no-e2e --no-parity-check --no-phases
run tgv_n8_pipe1c1 8 CALES_SOLVER_PIPE=1 CALES_SOLVER_CHUNKS=1 -- --steps 20 --no-e2e --no-parity-check --no-phases
run tgv_n4 4 X=1 -- --steps 20 --no-e2e --no-parity-check
run ch5_dsmag_n8 8 X=1 -- --workload channel5 --sgs dsmag --steps 10 --no-e2e --no-parity-check
run ch5_dsmag_n8_pipe1 8 CALES_SOLVER_PIPE=1 -- --workload channel5 --sgs dsmag --steps 10 --no-e2e --no-parity-check --no-phases
run ch5_smag_n8 8 X=1 -- --workload channel5 --sgs smag --steps 10 --no-e2e --no-parity-check
run ch5_smag_n8_pipe0 8 CALES_SOLVER_PIPE=0 -- --workload channel5 --sgs smag --steps 10 --no-e2e --no-parity-check --no-phases
run ch5_smag_n4 4 X=1 -- --workload channel5 --sgs smag --steps 10 --no-e2e --no-parity-check --no-phases
run ch5_smag_n2 2 X=1 -- --workload channel5 --sgs smag --steps 10 --no-e2e --no-parity-check --no-phases
run duct4_n4 4 X=1 -- --workload duct4 --steps 10 --no-e2e
run cavity4_n4 4 X=1 -- --workload cavity4 --steps 10 --no-e2e --no-parity-check
for f in gpurun_out/${TAG}_*.json; do python - "$f" <<'PY'
import json,sys
try:
    l=[x for x in open(sys.argv[1]) if x.startswith("{")][-1]
    d=json.loads(l); print(sys.argv[1].split("/")[-1], "dims", d["config"]["dims"], "ms/step", round(d["ms_per_step"],3), "Mcell/s", round(d["value"]), "poisson", round(d["poisson_ms"],3), "sanity", d["sanity"]["ok"], "parity", d.get("parity_check",{}).get("ok"))
    if d["phases_ms"]: print("   phases", {k: round(v,3) for k,v in d["phases_ms"].items()})
    if "nvlink" in d: print("   nvlink ms", round(d["nvlink"]["ms"],4), "GB/s", round(d["nvlink"]["gbs_per_direction"]), "frac", round(d["nvlink"]["frac_of_900"],3))
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
