# Round-2 ncu evidence (see /opt/skills/guides/B200_PROFILING.md): one --set full capture per hot kernel + the launch list of a
# short bench run.  usage (GPU box): bash tools/gpu_profile_r02.sh
mkdir -p gpurun_out
export KTB200_BLK_COOP=0
# 1. persistent MoE-block kernel (bench, eager, 2 resident layer sets)
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"moe_block_kernel" -s 6 -c 1 -o gpurun_out/prof_r02_block \
    python bench.py --steps 1 --warmup 1 --resident-layers 2 --no-graph --no-cpu-baseline --no-full-step > gpurun_out/ncu_block.log 2>&1
# 2. MLA decode on tcgen05 (ctx 32768, bs 1) and the dense segment-ring linear (o_proj shape)
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"mla_decode_tc_kernel|dense_q4k_kernel" -s 4 -c 2 -o gpurun_out/prof_r02_mla_dense \
    python tools/ncu_targets.py > gpurun_out/ncu_mla.log 2>&1
# 3. expert-parallel block kernel (world = 1 loop-back: the same code path, one GPU)
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"moe_ep_block_kernel" -s 4 -c 1 -o gpurun_out/prof_r02_ep \
    env REPRO_GRAPH=0 REPRO_N=6 REPRO_E=256 python tools/ep_pdl_repro.py > gpurun_out/ncu_ep.log 2>&1
# 4. launch list of the bench (MoE path + the eager passes of the whole-step leg)
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_r02.csv \
    python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline --resident-layers 2 > gpurun_out/ncu_list.log 2>&1
ls -la gpurun_out | tail -8
