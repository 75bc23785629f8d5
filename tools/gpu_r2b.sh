#!/bin/bash
# One GPU call: verify (tests, smoke, bench c2), serialised launch list, and ncu --set full captures of the DP and seed kernels.
bash tools/gpu_verify.sh
bash tools/gpu_diag.sh 2>&1 | head -70
DMND_LANES=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:swipe16_kernel -c 3 -f -o gpurun_out/swipe16_r2b python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_full_dp.log 2>&1; echo "ncu dp rc $?"
DMND_LANES=1 timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:probe_kernel|stage12_kernel|xdrop_kernel|chain_pair_kernel|walk16_kernel' -c 12 -f -o gpurun_out/seed_r2b python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_full_seed.log 2>&1; echo "ncu seed rc $?"
ls -la gpurun_out/*.ncu-rep
