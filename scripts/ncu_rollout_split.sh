#!/bin/bash
# ncu --set full of the headline launch (gw_rollout_split_kernel, 4 096 envs, FeAR on, 20 steps per launch), only after the same
# command has exited 0 without the profiler
CMD="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-scale-points --no-train --no-e2e"
$CMD > gpurun_out/r2r_ncu_plain.log 2>&1 || exit 1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gw_rollout_split -c 3 -o gpurun_out/r2r_rollout_split -f $CMD > gpurun_out/r2r_rollout_split_ncu.log 2>&1
tail -3 gpurun_out/r2r_rollout_split_ncu.log
