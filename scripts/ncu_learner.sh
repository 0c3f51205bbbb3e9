#!/bin/bash
# full ncu capture of the update kernel (one launch of 8 updates), only after the plain run has exited 0.  Under the
# profiler the cooperative launch of all 33 co-resident clusters fails (LaunchFailed), so the capture runs without the
# cooperative attribute (GW_LEARN_NO_COOP=1) and without the helper clusters (GW_LEARN_HELPERS=0: the 16 clusters that own row blocks).
R=${1:-r2}
export GW_LEARN_HELPERS=${GW_LEARN_HELPERS:-0} GW_LEARN_NO_COOP=1
for K in cluster; do
  CMD="python scripts/ncu_learner_target.py $K"
  timeout 120 $CMD > gpurun_out/${R}_learner_${K}_plain.log 2>&1 &&
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gw_learn -s 2 -c 1 -o gpurun_out/${R}_learner_${K} $CMD > gpurun_out/${R}_learner_${K}_ncu.log 2>&1
  grep -i "error\|Report" gpurun_out/${R}_learner_${K}_ncu.log | tail -3
done
