#!/bin/bash
# verify + bench + ncu --set full of the DP kernels (lane-replicated score table)
bash tools/gpu_verify.sh
DMND_LANES=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:swipe16_kernel -c 3 -f -o gpurun_out/swipe16_r2c python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_full_dp.log 2>&1; echo "ncu dp rc $?"
ls -la gpurun_out/*.ncu-rep
