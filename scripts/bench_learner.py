"""GPU time of the fused MADDPG update (gw_learner_update, csrc/gw_maddpg.cu): us per update for U updates per launch,
batches drawn from a filled replay ring inside the kernel; next to it round 1's CUDA graph of PyTorch / library kernels."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from marl_responsible_nav_b200 import maddpg  # noqa: E402

hp = maddpg.preset("custom_fear_10")
E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
env = maddpg.make_env(hp, E)
tr = maddpg.BatchedTrainer(env, hp=hp, seed=0)
tr.train(60, learn=False)                                # fill the ring
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
kinds = [k for k in sys.argv[2:] if k in ("cluster", "phase")] or [tr.learner.kernel]
for kind, U in [(k, U) for k in kinds for U in (1, 8, 64, 409)]:
    tr.learner.set_kernel(kind)
    tr.learner.learn_from_ring(tr.ring, U, tr.sample_seed)          # warm-up
    torch.cuda.synchronize()
    reps = max(2, 512 // U)
    a.record()
    for _ in range(reps):
        tr.learner.learn_from_ring(tr.ring, U, tr.sample_seed)
    b.record()
    torch.cuda.synchronize()
    us = a.elapsed_time(b) * 1e3 / (reps * U)
    print(json.dumps({"kernel": kind, "updates_per_launch": U, "us_per_update": round(us, 2), "updates_per_s": round(1e6 / us),
                      "gflops": round(0.23e9 / us / 1e3, 1)}), flush=True)
losses = tr.learner.learn_from_ring(tr.ring, 4, tr.sample_seed)
print("losses of the last updates (actor | critic per agent):", losses.cpu().tolist()[-1], flush=True)
if "--torch" in sys.argv:
    tr2 = maddpg.BatchedTrainer(env, hp=hp, seed=0, fused_learner=False, learn_cadence="batched")
    tr2.train(200)
    g = tr2.agent._graph[0]
    torch.cuda.synchronize()
    a.record()
    for _ in range(100):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    print(json.dumps({"round1_update_graph_us": round(a.elapsed_time(b) * 10, 1)}), flush=True)
