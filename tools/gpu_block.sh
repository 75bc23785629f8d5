#!/bin/bash
# block mode (default) against the per-lane pipeline on the same box
bash tools/gpu_verify.sh
DMND_PIPELINE_LANES=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_lanes.json 2> gpurun_out/bench_lanes.err; python -c "
import json; d=json.load(open('gpurun_out/bench_lanes.json')); print('per-lane pipeline:', round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],2), d['step_ms'])"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_block.json 2> gpurun_out/bench_block.err; python -c "
import json; d=json.load(open('gpurun_out/bench_block.json')); print('block mode again:', round(d['value'],1), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],2), d['step_ms'])"
