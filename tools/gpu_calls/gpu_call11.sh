#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
T=r2m
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mom_k -c 6 -o gpurun_out/${T}_mom_full python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-phases > gpurun_out/${T}_ncu_full.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-phases > gpurun_out/${T}_ncu_list.log 2>&1
ls -la gpurun_out/${T}_*
timeout 900 python bench.py --workload channel5 --sgs dsmag --steps 3 --no-e2e --no-phases > gpurun_out/${T}_ch5_dsmag_n1.json 2> gpurun_out/${T}_ch5_dsmag_n1.err; echo "ch5 dsmag rc=$?"; tail -3 gpurun_out/${T}_ch5_dsmag_n1.err
python - <<'PY'
import json
l=[x for x in open("gpurun_out/r2m_ch5_dsmag_n1.json") if x.startswith("{")]
if l:
    d=json.loads(l[-1]); print("ch5 dsmag N=1 ms/step", d["ms_per_step"], "Mcell/s", d["value"], "poisson", d["poisson_ms"])
PY
