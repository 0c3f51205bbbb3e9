#!/bin/bash
# Longer learning curves at the reference's learn cadence (409 updates per vector step at 4096 envs): 600 vector steps = 245 k
# updates = 2.4 M env steps per run, then evaluate.py on 16 384 episodes.
R=${1:-r2n}
for spec in "single custom_fear_5" "single custom" "multi custom_fear_5"; do set -- $spec
  CK=gpurun_out/${R}_trained_$1_$2.pt
  timeout 900 python -m marl_responsible_nav_b200.train --config $2 --env-kind $1 --envs 4096 --steps 600 --report 100 --save $CK 2>&1 | grep env_steps | cut -c1-330 > gpurun_out/${R}_train_$1_$2.log
  python - <<PY
import json
for l in open("gpurun_out/${R}_train_$1_$2.log"):
    d=json.loads(l[:l.rfind(",")]+"}") if not l.rstrip().endswith("}") else json.loads(l)
    print("$1 $2", d["env_steps"], "return %.2f len %.1f crashes/ep %.3f apples/ep %.3f updates %d (%.0f/s)" % (d["mean_return"], d["mean_len"], d["crashes_per_episode"], d["apples_per_episode"], d["updates"], d["updates_per_s"]))
PY
  timeout 300 python -m marl_responsible_nav_b200.evaluate --checkpoint $CK --episodes 16384 --envs 2048 --fear > gpurun_out/${R}_eval_$1_$2.log 2>&1
  tail -1 gpurun_out/${R}_eval_$1_$2.log | cut -c1-330
  rm -f $CK
done
