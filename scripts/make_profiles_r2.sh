#!/bin/bash
# Round-2 evidence run (one gpurun call, one GPU): GPU test suite, smoke, both bench arms at the driver's flags, the ncu launch
# list of the bench command (only after it has exited 0 without ncu), update-kernel timing / phase trace, training lines at
# both learn cadences.  Outputs land in gpurun_out/<tag>_*.
R=${1:-r2m}
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/${R}_gpu_tests.log 2>&1; tail -2 gpurun_out/${R}_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_smoke.log 2>&1; tail -1 gpurun_out/${R}_smoke.log | cut -c1-200
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${R}_bench_reference.json 2> gpurun_out/${R}_bench_reference.err
python bench.py --steps 20 --warmup 5 > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err; tail -c 300 gpurun_out/${R}_bench.json
python bench.py --steps 2000 --warmup 200 --no-cpu-baseline --no-scale-points --no-train > gpurun_out/${R}_bench_k2000.json 2>> gpurun_out/${R}_bench.err
CMD="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-scale-points --no-train"
$CMD > gpurun_out/${R}_bench_short_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${R}_bench_launches.csv $CMD > gpurun_out/${R}_ncu_launchlist.log 2>&1
timeout 120 python scripts/bench_learner.py 4096 cluster phase --torch > gpurun_out/${R}_learner_bench.log 2>&1; head -4 gpurun_out/${R}_learner_bench.log
timeout 120 python scripts/trace_learner.py 4096 cluster > gpurun_out/${R}_learner_trace.log 2>&1
timeout 120 python scripts/trace_learner.py 4096 phase >> gpurun_out/${R}_learner_trace.log 2>&1
timeout 300 python -m marl_responsible_nav_b200.train --config custom_fear_10 --envs 4096 --steps 60 --report 20 2>&1 | grep env_steps | cut -c1-400 > gpurun_out/${R}_train_1gpu_reference_cadence.log
timeout 300 python -m marl_responsible_nav_b200.train --config custom_fear_10 --envs 4096 --steps 1200 --report 300 --learn-cadence batched 2>&1 | grep env_steps | cut -c1-400 > gpurun_out/${R}_train_1gpu_batched.log
timeout 300 python scripts/bench_rollout.py > gpurun_out/${R}_rollout.log 2>&1
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/${R}_gpu.txt; nproc >> gpurun_out/${R}_gpu.txt
ls gpurun_out | grep ${R}_ | wc -l
# actor kernel alone (CUDA events) and under ncu --set full at 1 M envs x 2 learners (after the plain run has exited 0)
timeout 200 python scripts/time_actor.py > gpurun_out/${R}_time_actor.log 2>&1
CMD="python scripts/bench_actor_only.py"
$CMD > gpurun_out/${R}_actor_plain.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:actor_forward -s 3 -c 1 -f -o gpurun_out/${R}_actor $CMD > gpurun_out/${R}_actor_ncu.log 2>&1
ls gpurun_out | grep ${R}_ | wc -l
