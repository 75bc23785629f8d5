#!/bin/bash
# one multi-GPU bench line (argument: number of GPUs) + the 2-rank broadcast test when N >= 2
N=${1:-2}
mkdir -p gpurun_out
nproc; cat /sys/fs/cgroup/cpu.max 2>/dev/null
if [ "$N" = "2" ]; then timeout 600 python -m pytest tests/test_zzz_broadcast_gpu.py -m gpu -q -rA -p no:cacheprovider 2>&1 | tail -5; fi
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_${N}gpu.json 2> gpurun_out/bench_r2_${N}gpu.err; echo "bench N=$N rc $?"
python - <<PY
import json
try:
    d = json.load(open('gpurun_out/bench_r2_${N}gpu.json'))
    print('N=${N} value', round(d['value'], 1), round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 1), round(d['e2e']['ms_per_step'], 2), d['breakdown_ms_per_step'], d['step_ms'], d.get('host_pinning_rank0'), d['config'].get('reference_threads'))
except Exception as e:
    print('parse failed', e); print(open('gpurun_out/bench_r2_${N}gpu.err').read()[-2000:])
PY
