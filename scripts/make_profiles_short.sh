#!/bin/bash
# Short evidence run (one gpurun call, ~10 min): GPU tests, full bench line + reference arm, ncu launch list of the
# bench command, full ncu capture of the actor kernel, rollout and 1-GPU training-loop throughput.
# The step kernels' full captures are in scripts/make_profiles.sh (unchanged kernels keep their earlier captures).
R=${1:-r1}
S=$(date +%s)
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${R}_gpu_tests.log 2>&1; tail -3 gpurun_out/${R}_gpu_tests.log
echo "tests done $(( $(date +%s) - S )) s"
python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err; tail -c 300 gpurun_out/${R}_bench.json
echo "bench done $(( $(date +%s) - S )) s"
python bench.py --impl reference --steps 300 --warmup 20 > gpurun_out/${R}_bench_reference.json 2>> gpurun_out/${R}_bench.err
CMD="python bench.py --steps 128 --warmup 64 --no-cpu-baseline --no-scale-points"
$CMD > gpurun_out/${R}_bench_short_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/${R}_bench_launches.csv $CMD > gpurun_out/${R}_ncu_launchlist.log 2>&1
echo "launch list done $(( $(date +%s) - S )) s"
timeout 200 python scripts/bench_rollout.py > gpurun_out/${R}_rollout.log 2>&1
timeout 200 python -m marl_responsible_nav_b200.train --config custom_fear_10 --envs 4096 --steps 1200 --report 300 2>&1 | grep env_steps > gpurun_out/${R}_train_1gpu.log
CA="python scripts/bench_actor_only.py"
$CA > gpurun_out/${R}_actor_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:actor_forward -s 3 -c 1 -o gpurun_out/${R}_actor $CA > gpurun_out/${R}_actor_ncu.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_smoke.log 2>&1; tail -1 gpurun_out/${R}_smoke.log
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/${R}_gpu.txt; nproc >> gpurun_out/${R}_gpu.txt
echo "all done $(( $(date +%s) - S )) s"
