# round-end measurement recipe (1x B200): parity tests, bench line, ncu launch list, ncu full capture of the top kernels
R=${1:-r1h}
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py 2>&1 | tail -1 > gpurun_out/${R}_bench.json; cut -c1-300 gpurun_out/${R}_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'mom_k|rk_update_k|strain_k|gauss_tma_k|fftb_k' --launch-skip 40 -c 12 -o gpurun_out/${R}_full python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python tools/kbench.py --deck channel --ng 512 256 192 --iters 10 2>&1 | tail -28 > gpurun_out/${R}_kbench_channel_512x256x192.txt; tail -3 gpurun_out/${R}_kbench_channel_512x256x192.txt
python tools/kbench.py --iters 20 2>&1 | tail -28 > gpurun_out/${R}_kbench_tgv_256.txt
python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | cut -c1-400
