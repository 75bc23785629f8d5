#!/bin/bash
# One GPU call: device parity tests, smoke, the three bench configurations.  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rA -p no:cacheprovider > gpurun_out/pytest_gpu_r2.txt 2>&1; echo "pytest rc $?"; tail -4 gpurun_out/pytest_gpu_r2.txt
grep -E "FAILED|ERROR" gpurun_out/pytest_gpu_r2.txt | head -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2_c2.json 2> gpurun_out/bench_r2_c2.err; echo "bench c2 rc $?"
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/bench_r2_c2.json'))
    r = d['roofline']
    print('c2 value', round(d['value'], 1), round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 1), round(d['e2e']['ms_per_step'], 2),
          'kernel_ms', round(r['kernel_ms_per_step'], 2), 'frac', round(r['frac'], 3), round(r['frac_s32_peak'], 3),
          'seed_ms', round(d['roofline_seed']['kernel_ms_per_step'], 2), d['breakdown_ms_per_step'], d['step_ms'], d.get('cpu_baseline'))
except Exception as e:
    print('c2 parse failed', e); print(open('gpurun_out/bench_r2_c2.err').read()[-1500:])
PY
