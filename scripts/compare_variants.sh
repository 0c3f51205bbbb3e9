#!/bin/bash
# dev: the same rollout timing for two builds of the library (GW_LIB selects the .so), alternating, three rounds
for round in 1 2; do
  for v in "" "_varA" "_varC"; do
    lib=$PWD/marl_responsible_nav_b200/csrc/libgridworld_b200$v.so
    echo "--- round $round lib${v:-_current}"
    GW_LIB=$lib timeout 200 python scripts/bench_rollout_kernel.py 2>&1 | grep "E=4096 T=64 fear=True\|E=4096 T=256\|E=4096 T=8 "
  done
done
