#!/bin/bash
# Round-end evidence run (one gpurun call): tests, full bench line, ncu launch list of the bench command, full ncu
# captures of the step kernel at the bench workload (4096 envs) and at 1M envs and of the actor kernel, resident-kernel
# phase trace, host-driven step breakdown, rollout and training-loop throughput.  Outputs land in gpurun_out/.
R=${1:-r1}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${R}_gpu_tests.log 2>&1; tail -3 gpurun_out/${R}_gpu_tests.log
python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err; tail -c 400 gpurun_out/${R}_bench.json
python bench.py --impl reference --steps 300 --warmup 20 > gpurun_out/${R}_bench_reference.json 2>> gpurun_out/${R}_bench.err
timeout 200 python scripts/trace_server.py 4096 > gpurun_out/${R}_server_trace.txt 2>&1
for E in 4096 8192 16384; do echo "== $E envs"; timeout 100 python scripts/e2e_breakdown.py $E 2>&1; done > gpurun_out/${R}_e2e_breakdown.txt
timeout 300 python scripts/bench_rollout.py > gpurun_out/${R}_rollout.log 2>&1
timeout 300 python -m marl_responsible_nav_b200.train --config custom_fear_10 --envs 4096 --steps 1200 --report 300 2>&1 | grep env_steps > gpurun_out/${R}_train_1gpu.log
timeout 100 python scripts/trainer_breakdown.py > gpurun_out/${R}_trainer_breakdown.log 2>&1
timeout 100 python scripts/bench_sampler.py > gpurun_out/${R}_sampler.log 2>&1
CMD="python bench.py --steps 128 --warmup 64 --no-cpu-baseline --no-scale-points"
$CMD > gpurun_out/${R}_bench_short_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/${R}_bench_launches.csv $CMD > gpurun_out/${R}_ncu_launchlist.log 2>&1
for spec in "e4096_fear 4096 1" "e1m_fear 1048576 1" "e1m_nofear 1048576 0"; do set -- $spec
  C2="python bench.py --envs $2 --fear $3 --steps 64 --warmup 64 --no-cpu-baseline --no-e2e --no-scale-points"
  $C2 > gpurun_out/${R}_$1_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:gw_step -s 70 -c 2 -o gpurun_out/${R}_$1 $C2 > gpurun_out/${R}_$1_ncu.log 2>&1
done
CA="python scripts/bench_actor_only.py"
$CA > gpurun_out/${R}_actor_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:actor_forward -s 3 -c 1 -o gpurun_out/${R}_actor $CA > gpurun_out/${R}_actor_ncu.log 2>&1
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/${R}_gpu.txt; nproc >> gpurun_out/${R}_gpu.txt
ls gpurun_out | grep ${R}_ | head -60
