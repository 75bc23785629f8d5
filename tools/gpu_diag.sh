#!/bin/bash
# Diagnostics: serialised launch list (ncu) of one single-lane step and a host timeline of the default lanes.
mkdir -p gpurun_out
DMND_LANES=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r2_lane1.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_lane1.log 2>&1; echo "ncu rc $?"
python - <<'PY'
import csv, collections, re
rows = [r for r in csv.reader(l for l in open('gpurun_out/launches_r2_lane1.csv') if not l.startswith('=='))]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
tot = collections.defaultdict(lambda: [0.0, 0])
seq = []
for r in rows[1:]:
    if len(r) <= vi: continue
    v = float(r[vi].replace(',', '')); u = r[ui]
    ms = v / 1e6 if u in ('ns', 'nsecond') else v / 1e3 if u in ('us', 'usecond') else v
    name = re.sub(r'\(.*', '', r[ki])[:90]
    tot[name][0] += ms; tot[name][1] += 1; seq.append((name, ms))
allms = sum(v[0] for v in tot.values())
with open('gpurun_out/launch_shares_r2_lane1.txt', 'w') as f:
    f.write(f"# total kernel time {allms:.1f} ms over {sum(v[1] for v in tot.values())} launches (DMND_LANES=1 bench.py --steps 1 --warmup 1)\n")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][0]):
        f.write(f"{v[0]:10.3f} {100*v[0]/allms:6.1f}% {v[1]:6d}  {k}\n")
print(open('gpurun_out/launch_shares_r2_lane1.txt').read()[:3500])
PY
DMND_PROFILE=1 timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/prof_c2.json 2> gpurun_out/prof_c2.err; grep "dmnd profile" gpurun_out/prof_c2.err | tail -150 > gpurun_out/prof_c2_tail.txt; tail -75 gpurun_out/prof_c2_tail.txt
