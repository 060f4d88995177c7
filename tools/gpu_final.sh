# round-end check on one B200: GPU tests, smoke, the default bench line (with CPU baseline), the reference arm, launch list
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout -s KILL 600 python bench.py > gpurun_out/bench_final.log 2>&1; tail -1 gpurun_out/bench_final.log | cut -c1-300
timeout -s KILL 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final_ref.log 2>&1; tail -1 gpurun_out/bench_final_ref.log | cut -c1-400
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"moe_block|rows_bulk|reduce_bulk|gate_kernel" -c 400 --csv --log-file gpurun_out/launches_final.csv \
    python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_list_final.log 2>&1
ls -la gpurun_out | tail -4
