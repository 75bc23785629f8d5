#!/bin/bash
# verify (bulk-copy probe tile, async bias) + bench c2, c3, c5 + serialised launch list
bash tools/gpu_verify.sh
timeout 600 python bench.py --config c3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_c3.json 2> gpurun_out/bench_r2_c3.err; echo "bench c3 rc $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_r2_c3.json')); print('c3', round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), d['breakdown_ms_per_step'])"
timeout 1200 python bench.py --config c5 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r2_c5.json 2> gpurun_out/bench_r2_c5.err; echo "bench c5 rc $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_r2_c5.json')); print('c5', round(d['value'],2), round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],2), d['breakdown_ms_per_step'], d['work'])"; tail -3 gpurun_out/bench_r2_c5.err | cut -c1-300
bash tools/gpu_diag.sh 2>&1 | head -45
