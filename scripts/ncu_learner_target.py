"""ncu target: a few gw_learner_update launches of 8 updates each on a filled replay ring (custom_fear_10, 4096 envs)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from marl_responsible_nav_b200 import maddpg  # noqa: E402

hp = maddpg.preset("custom_fear_10")
env = maddpg.make_env(hp, 4096)
tr = maddpg.BatchedTrainer(env, hp=hp, seed=0)
tr.train(60, learn=False)
if len(sys.argv) > 1:
    tr.learner.set_kernel(sys.argv[1])
for _ in range(4):
    tr.learner.learn_from_ring(tr.ring, 8, tr.sample_seed)
torch.cuda.synchronize()
print("ok", tr.learner.kernel)
