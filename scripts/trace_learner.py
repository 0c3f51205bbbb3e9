"""Per-phase time of the last update of a gw_learner_update launch (clock64 of CTA 0 at every phase start)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from marl_responsible_nav_b200 import maddpg  # noqa: E402

hp = maddpg.preset("custom_fear_10")
env = maddpg.make_env(hp, int(sys.argv[1]) if len(sys.argv) > 1 else 4096)
tr = maddpg.BatchedTrainer(env, hp=hp, seed=0)
tr.train(60, learn=False)
if len(sys.argv) > 2:
    tr.learner.set_kernel(sys.argv[2])
for _ in range(3):
    tr.learner.learn_from_ring(tr.ring, 16, tr.sample_seed)
torch.cuda.synchronize()
t = tr.learner.debug_tensor("trace").view(torch.int64).cpu().tolist()
names = tr.learner.phase_names if tr.learner.kernel == "phase" else ("A: critic gradients", "Adam C", "B: actor gradients", "Adam A")
n = len(names)
mhz = 1965.0
tot = (t[n] - t[0]) / mhz
print(f"update: {tot:.1f} us over {n} phases (clock64 at {mhz:.0f} MHz)")
for k in range(n):
    print(f"  phase {k:2d} {names[k]:20s} {(t[k + 1] - t[k]) / mhz:7.2f} us")

if tr.learner.kernel == "cluster" and t[8]:                # built with GW_NVCC_FLAGS=-DGW_LEARN_TRACE: checkpoints inside phases A and B
    pts = [x for x in t[8:108] if x]
    print("checkpoints (us since phase A start):", " ".join(f"{(x - t[0]) / mhz:.1f}" for x in pts))
