#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rA -p no:cacheprovider > gpurun_out/pytest_gpu_r2.txt 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/pytest_gpu_r2.txt
timeout 1500 python bench.py --config c4 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r2_c4.json 2> gpurun_out/bench_r2_c4.err; echo "bench c4 rc $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_r2_c4.json')); print('c4', round(d['value'],2), round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],2), d['breakdown_ms_per_step'], 'seed kernel ms', round(d['roofline_seed']['kernel_ms_per_step'],1), d['work']['seed_counters'])"; tail -3 gpurun_out/bench_r2_c4.err | cut -c1-300
