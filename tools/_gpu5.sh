mkdir -p gpurun_out
timeout 300 python bench.py > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; echo "c2 rc $?"
timeout 200 python bench.py --emit-cells 16,96 > gpurun_out/emit_cells.log 2>&1; echo "emit rc $?"
timeout 300 python bench.py --config c3 --steps 5 --warmup 2 > gpurun_out/bench_r2_c3.json 2> gpurun_out/bench_r2_c3.err; echo "c3 rc $?"
timeout 120 python bench.py --config c3 --emit-cells 16,96 >> gpurun_out/emit_cells.log 2>&1
timeout 240 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2_ref.json 2> gpurun_out/bench_r2_ref.err; echo "ref rc $?"
DMND_LANES=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --queries 200000 > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc $?"
DMND_LANES=1 timeout 500 ncu --set full --clock-control none --import-source on -k regex:swipe16_kernel -c 3 -o gpurun_out/swipe16_r2 -f python bench.py --steps 1 --warmup 0 --no-cpu-baseline --queries 200000 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc $?"
timeout 700 python bench.py --config c4 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_r2_c4.json 2> gpurun_out/bench_r2_c4.err; echo "c4 rc $?"
tail -c 600 gpurun_out/bench_r2c.err; cat gpurun_out/bench_r2c.json | cut -c1-3000
