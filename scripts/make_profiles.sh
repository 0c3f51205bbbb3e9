#!/bin/bash
# Round-end evidence run (one gpurun call): tests, full bench line, ncu launch list of the bench command, full ncu
# captures of the step kernel at the bench workload (4096 envs) and at 1M envs.  Outputs land in gpurun_out/.
R=${1:-r1}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${R}_gpu_tests.log 2>&1; tail -3 gpurun_out/${R}_gpu_tests.log
python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err; tail -c 400 gpurun_out/${R}_bench.json
python bench.py --impl reference --steps 300 --warmup 20 > gpurun_out/${R}_bench_reference.json 2>> gpurun_out/${R}_bench.err
CMD="python bench.py --steps 128 --warmup 64 --no-cpu-baseline --no-e2e --no-scale-points"
$CMD > gpurun_out/${R}_bench_short_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${R}_bench_launches.csv $CMD > gpurun_out/${R}_ncu_launchlist.log 2>&1
for spec in "e4096_fear 4096 1" "e1m_fear 1048576 1" "e1m_nofear 1048576 0"; do set -- $spec
  C2="python bench.py --envs $2 --fear $3 --steps 64 --warmup 64 --no-cpu-baseline --no-e2e --no-scale-points"
  $C2 > gpurun_out/${R}_$1_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:gw_step -s 70 -c 2 -o gpurun_out/${R}_$1 $C2 > gpurun_out/${R}_$1_ncu.log 2>&1
done
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/${R}_gpu.txt; nproc >> gpurun_out/${R}_gpu.txt
ls gpurun_out | head -40
