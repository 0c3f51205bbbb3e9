#!/bin/bash
# usage: scripts/ncu_step.sh <tag> <envs> <fear>   -> gpurun_out/<tag>_launches.csv and gpurun_out/<tag>.ncu-rep
set -e
TAG=$1; ENVS=$2; FEAR=$3
CMD="python bench.py --envs $ENVS --fear $FEAR --steps 40 --warmup 10 --no-cpu-baseline --no-e2e --no-scale-points"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gw_step -s 10 -c 40 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
$CMD > gpurun_out/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gw_step -s 20 -c 2 -o gpurun_out/${TAG} $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
tail -2 gpurun_out/${TAG}_ncu2.log
