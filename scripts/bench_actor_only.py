import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from marl_responsible_nav_b200 import BatchedGridWorld, FusedActor, maddpg
E = 1 << 20
env = BatchedGridWorld("Level 3", num_envs=E, fear=False, auto_reset=True, seed=1)
agent = maddpg.BatchedMADDPG(2, 160, 9, device="cuda", seed=0)
fused = FusedActor(env, agent.actors, seed=1)
out = env.reset()
for _ in range(6):
    fused.forward(out.obs_code, out.action_mask, training=True)
torch.cuda.synchronize()
print("ok")
