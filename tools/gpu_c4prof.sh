#!/bin/bash
# where does the seed stage of --sensitive (c4) spend its time?  serialised launch list of one step on a tenth of the queries
mkdir -p gpurun_out
DMND_LANES=1 timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/launches_r2_c4.csv python bench.py --config c4 --queries 100000 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/bench_c4_small.json 2> gpurun_out/bench_c4_small.err; echo "ncu c4 rc $?"
python tools/ncu_summary.py launches gpurun_out/launches_r2_c4.csv > gpurun_out/launch_shares_r2_c4.txt 2>&1; head -30 gpurun_out/launch_shares_r2_c4.txt
python -c "
import json; d=json.load(open('gpurun_out/bench_c4_small.json')); print(d['work']); print(d['breakdown_ms_per_step'])"
