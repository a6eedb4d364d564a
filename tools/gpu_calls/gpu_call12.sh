#!/bin/bash
# round 2, distributed z solve on 2 GPUs: parity tests of the paths it touches, then the weak-scaling bench line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
T=r2n
timeout 600 python -m pytest tests/test_gpu_multi.py -q -x -k "test_two_gpus[tgv_smag-1-2] or test_two_gpus[channel_dsmag-1-2] or test_two_gpus[duct_smag-1-2] or (exchange_variants and strict) or (implicit and 3d)" 2>&1 | tail -30 > gpurun_out/${T}_gtest_n2.log
tail -5 gpurun_out/${T}_gtest_n2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --no-e2e --no-cpu-baseline > gpurun_out/${T}_bench_n2.json 2> gpurun_out/${T}_bench_n2.err
echo "bench rc=$?"; tail -3 gpurun_out/${T}_bench_n2.err | cut -c1-600
python - <<'PY'
import json
l=[x for x in open("gpurun_out/r2n_bench_n2.json") if x.startswith("{")]
if l:
    d=json.loads(l[-1]); print("N=2 ms/step", d.get("ms_per_step"), "poisson", d.get("poisson_ms"), d.get("nvlink",{}).get("solver_exchange"), "parity", d.get("parity_check",{}).get("ok"))
    print(json.dumps(d.get("phases")))
PY
CALES_ZDIST=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 2 --no-e2e --no-cpu-baseline --no-parity-check --no-phases 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ZDIST=0 N=2 ms/step', d['ms_per_step'], 'poisson', d['poisson_ms'])"
