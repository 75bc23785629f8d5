#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_sens1_gpu.py tests/test_zzz_blastx_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -3
for c in c4 c5; do
timeout 1500 python bench.py --config $c --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r2_$c.json 2> gpurun_out/bench_r2_$c.err; echo "bench $c rc $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_r2_$c.json')); print('$c', round(d['value'],2), round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],2), d['breakdown_ms_per_step'], 'seed kernel ms', round(d['roofline_seed']['kernel_ms_per_step'],1), 'dp kernel ms', round(d['roofline']['kernel_ms_per_step'],1), d['work']['seed_counters'])"; tail -2 gpurun_out/bench_r2_$c.err | cut -c1-300
done
