#!/bin/bash
# verify + bench c2 + ncu of the DP kernel + first c5 (frameshift) bench line
bash tools/gpu_verify.sh
DMND_LANES=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:swipe16_kernel -c 3 -f -o gpurun_out/swipe16_r2d python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_full_dp.log 2>&1; echo "ncu dp rc $?"
timeout 900 python bench.py --config c5 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r2_c5.json 2> gpurun_out/bench_r2_c5.err; echo "bench c5 rc $?"; tail -c 1500 gpurun_out/bench_r2_c5.json; tail -5 gpurun_out/bench_r2_c5.err
