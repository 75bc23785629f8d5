#!/bin/bash
# lane sweep on the c2 step (two interleaved rounds: boxes are noisy)
mkdir -p gpurun_out
for round in 1 2; do for L in 2 3 4 5 6; do
DMND_LANES=$L timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/lanes_$L.json 2> gpurun_out/lanes_$L.err; python -c "
import json; d=json.load(open('gpurun_out/lanes_$L.json')); r=sorted(d['step_ms']['resident']); e=sorted(d['step_ms']['e2e']); print('lanes $L round $round: resident median', r[len(r)//2], 'min', r[0], '| e2e median', e[len(e)//2], 'min', e[0])"
done; done
