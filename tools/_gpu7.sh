for L in 3 4 6 8; do echo "== lanes $L"; DMND_LANES=$L timeout 200 python bench.py --steps 6 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],2), d['breakdown_ms_per_step'])"; done
echo "== host threads 32 lanes 4"; DMND_HOST_THREADS=32 DMND_LANES=4 timeout 200 python bench.py --steps 6 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],2), d['breakdown_ms_per_step'])"
DMND_LANES=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"stage12_kernel|probe_kernel|walk_kernel|chain_pair_kernel" -c 8 -o gpurun_out/seedk_r2 -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --queries 200000 > gpurun_out/ncu_seedk.log 2>&1; echo "ncu rc $?"
