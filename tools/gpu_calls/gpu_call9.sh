#!/bin/bash
# 1-GPU call: config 5 on one GPU (the strong-scaling denominator), dsmag and implicit-diffusion timings, cuFFT comparison, dsmag launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
mkdir -p gpurun_out
T=r2j
timeout 900 python bench.py --workload channel5 --sgs smag --steps 5 --no-e2e > gpurun_out/${T}_ch5_smag_n1.json 2> gpurun_out/${T}_ch5_smag_n1.err; echo "ch5 smag rc=$?"
timeout 300 python tools/kbench.py --only fft_PP,cufft,gaussel,solver > gpurun_out/${T}_kbench_cufft_256.txt 2>&1
timeout 300 python tools/kbench.py --ng 512 256 192 --deck channel --wall-model --only fft_PP,fft_NN,cufft,gaussel,solver,cmpt_sgs,mom,rk,step > gpurun_out/${T}_kbench_channel3.txt 2>&1
timeout 300 python tools/kbench.py --sgs dsmag --deck channel --only cmpt_sgs,step,solver > gpurun_out/${T}_kbench_dsmag_256.txt 2>&1
timeout 300 python tools/kbench.py --deck channel --impdiff 3d --only solver,substep,step,rk > gpurun_out/${T}_kbench_impdiff3d_256.txt 2>&1
timeout 300 python tools/kbench.py --deck channel --impdiff 1d --only solver,substep,step,rk > gpurun_out/${T}_kbench_impdiff1d_256.txt 2>&1
timeout 300 python tools/kbench.py --ng 192 192 192 --only fft_PP,cufft > gpurun_out/${T}_kbench_fft_192.txt 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_dsmag_launches.csv python tools/kbench.py --sgs dsmag --deck channel --only cmpt_sgs --iters 1 --warm 1 > /dev/null 2>&1
tail -n +1 gpurun_out/${T}_kbench_*.txt | cut -c1-110
python - <<'PY'
import json
l=[x for x in open("gpurun_out/r2j_ch5_smag_n1.json") if x.startswith("{")]
if l:
    d=json.loads(l[-1]); print("ch5 smag N=1 ms/step", d["ms_per_step"], "Mcell/s", d["value"], "poisson", d["poisson_ms"], d["phases_ms"])
PY
