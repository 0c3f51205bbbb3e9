#!/bin/bash
# Learning-quality evidence: train at the reference's learn cadence through the update kernel,
# save in the reference's checkpoint format, evaluate with evaluate.py (the batched customeval.py).
R=${1:-r2}
for spec in "multi custom_fear_5 120" "multi custom 120" "single custom_fear_5 120"; do set -- $spec
  CK=gpurun_out/${R}_trained_$1_$2.pt
  timeout 600 python -m marl_responsible_nav_b200.train --config $2 --env-kind $1 --envs 4096 --steps $3 --report 40 --save $CK 2>&1 | grep env_steps | cut -c1-330 > gpurun_out/${R}_train_$1_$2.log
  tail -1 gpurun_out/${R}_train_$1_$2.log
  timeout 300 python -m marl_responsible_nav_b200.evaluate --checkpoint $CK --episodes 16384 --envs 2048 --fear > gpurun_out/${R}_eval_$1_$2.log 2>&1
  tail -1 gpurun_out/${R}_eval_$1_$2.log | cut -c1-420
  rm -f $CK
done
