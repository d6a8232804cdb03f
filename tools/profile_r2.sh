#!/bin/bash
# Round-2 profiling pass (run under gpurun, ONE GPU): launch list of the bench command + full ncu captures of the hot kernels.
mkdir -p gpurun_out
# (1) per-launch durations of the timed bench command (eager launches so every kernel is a plain launch)
ORL_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2_launches_bench.log 2>&1
# (2) full captures (one launch each, after warm-up launches)
timeout 300 ncu --set full --import-source on --clock-control none -k regex:ppo_fwdbwd_tc -s 4 -c 1 -f -o gpurun_out/r2_tc_update python tools/prof_update.py > gpurun_out/r2_ncu_a.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:rollout_ -s 1 -c 1 -f -o gpurun_out/r2_tc_rollout python tools/prof_update.py > gpurun_out/r2_ncu_b.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:critic_values_tc -s 1 -c 1 -f -o gpurun_out/r2_tc_critic python tools/prof_update.py > gpurun_out/r2_ncu_c.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:gae -s 3 -c 1 -f -o gpurun_out/r2_gae1g python tools/gae_1gb.py > gpurun_out/r2_ncu_d.log 2>&1
tail -2 gpurun_out/r2_ncu_*.log; wc -l gpurun_out/r2_launches.csv
