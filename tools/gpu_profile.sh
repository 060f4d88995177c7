# One ncu --set full capture of the hot kernels + the launch list of a short bench run (see B200_PROFILING.md).
# usage (on the GPU box): bash tools/gpu_profile.sh <tag> [kernel-regex] [extra bench args / env via ENVV]
tag=${1:-rXX}; rx=${2:-"moe_block|rows_bulk|reduce_bulk"}
mkdir -p gpurun_out
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:"$rx" -s 10 -c 2 \
    -o gpurun_out/prof_$tag python bench.py --steps 1 --warmup 1 --resident-layers 2 --no-graph --no-cpu-baseline > gpurun_out/ncu_full_$tag.log 2>&1
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"moe_block|rows_bulk|reduce_bulk|rows_pipe|reduce_pipe|rows_kernel|reduce_kernel|gate_kernel|mla_|ep_" -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/ncu_list_$tag.log 2>&1
ls -la gpurun_out | tail -5
