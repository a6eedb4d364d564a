#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
T=r2k
python -m pytest tests -m "gpu and not fullsize" -q -x --deselect tests/test_gpu_multi.py > gpurun_out/${T}_gtest.log 2>&1; echo "gtest rc=$?" >> gpurun_out/${T}_gtest.log
grep -E "^FAILED|passed|failed|^E  " gpurun_out/${T}_gtest.log | cut -c1-300 | tail -12
timeout 300 python tools/kbench.py --only gaussel,solver,cmpt_sgs,step > gpurun_out/${T}_kbench_256.txt 2>&1
timeout 300 python tools/kbench.py --ng 512 256 192 --deck channel --wall-model --only gaussel,solver,cmpt_sgs,step > gpurun_out/${T}_kbench_channel3.txt 2>&1
tail -n +1 gpurun_out/${T}_kbench_*.txt | cut -c1-110
python bench.py --steps 20 --no-e2e --no-cpu-baseline --no-phases > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"; tail -c 500 gpurun_out/${T}_bench.err
python - <<'PY'
import json
d=json.loads([x for x in open("gpurun_out/r2k_bench.json") if x.startswith("{")][-1])
print(d["ms_per_step"], d["value"], d["roofline"]["kernel"], d["roofline"]["frac"], d["gpu_launches"])
for k,v in d["kernels"].items(): print("  %-26s %.4f ms  %.3f" % (k, v["ms"], v["frac"]))
PY
