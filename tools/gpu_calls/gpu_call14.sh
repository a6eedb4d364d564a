#!/bin/bash
# round 2, 8 GPUs: distributed z solve at N = 8 (weak 256^3 per GPU: the driver's SCALE workload), config 5 smag on the 4 x 2 grid,
# and the 4 x 2 wall-model parity case that failed earlier in the round
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
T=r2p
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $R --master-port 29571 bench.py --gpus 8 --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${T}_tgv_n8.json 2> gpurun_out/${T}_tgv_n8.err; echo "tgv rc=$?"
timeout 300 $R --master-port 29572 bench.py --gpus 8 --workload channel5 --sgs smag --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-phases > gpurun_out/${T}_ch5_smag_n8.json 2> gpurun_out/${T}_ch5_smag_n8.err; echo "ch5 rc=$?"
timeout 200 $R --master-port 29573 tests/mgpu_worker.py channel_wm_smag 4 2 5 > gpurun_out/${T}_wm42.log 2>&1; echo "wm42 rc=$?"
grep "^{" gpurun_out/${T}_wm42.log | cut -c1-900
python - <<'PY'
import json
for f in ("r2p_tgv_n8", "r2p_ch5_smag_n8"):
    try:
        l=[x for x in open("gpurun_out/%s.json"%f) if x.startswith("{")]
        d=json.loads(l[-1]); pc=d.get("parity_check",{})
        print(f, "ms/step", d.get("ms_per_step"), "poisson", d.get("poisson_ms"), d.get("nvlink",{}).get("solver_exchange","")[:30], "parity", pc.get("ok"), "zdist_off" if "zdist_disabled_after_failed_check" in pc else "", d.get("error",""))
    except Exception as e:
        print(f, "no line", e)
PY
tail -3 gpurun_out/${T}_tgv_n8.err | cut -c1-400
