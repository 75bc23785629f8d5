#!/bin/bash
# ncu --set full of the frameshift DP kernels on a small c5-like step
mkdir -p gpurun_out
DMND_LANES=1 timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:fs_swipe_kernel|fs_walk_kernel' -c 4 -f -o gpurun_out/fs_r2 python bench.py --config c5 --queries 2000 --db 200000 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_fs.log 2>&1; echo "ncu fs rc $?"
ls -la gpurun_out/fs_r2.ncu-rep; tail -3 gpurun_out/ncu_fs.log | cut -c1-300
