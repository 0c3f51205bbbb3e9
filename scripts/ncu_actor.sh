CMD="python scripts/bench_actor_only.py"
$CMD > gpurun_out/actor_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:actor_forward -s 3 -c 1 -o gpurun_out/r1_actor $CMD > gpurun_out/actor_ncu.log 2>&1; tail -2 gpurun_out/actor_ncu.log
