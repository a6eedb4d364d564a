# round-end measurement recipe (1x B200): parity tests, bench line (both arms), ncu launch list, ncu full capture of the top kernels,
# kernel tables.  Usage (from the repo root, under gpurun):  bash tools/final_run.sh r3a
R=${1:-r3a}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/${R}_gtest.log; tail -2 gpurun_out/${R}_gtest.log
python bench.py 2>&1 | tail -1 > gpurun_out/${R}_bench.json; cut -c1-300 gpurun_out/${R}_bench.json
python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/${R}_bench_reference.json; cut -c1-300 gpurun_out/${R}_bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-phases > /dev/null 2>&1
python tools/launches.py gpurun_out/${R}_launches.csv > gpurun_out/${R}_launches_summary.txt 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:'mom_k|strain2_k|strain_k|gauss2_k|gauss_tma_k|fftb_k' --launch-skip 40 -c 12 -o gpurun_out/${R}_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-phases > /dev/null 2>&1
python tools/kbench.py --iters 20 2>&1 | tail -28 > gpurun_out/${R}_kbench_tgv_256.txt
python tools/kbench.py --deck channel --ng 512 256 192 --iters 10 2>&1 | tail -28 > gpurun_out/${R}_kbench_channel_512x256x192.txt
