"""Time gw_actor_forward alone (CUDA events, 50 graph-replayed launches after warm-up) at several batch sizes.
GW_ACTOR_GROUPS=1|2 forces a regime (read once per process)."""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from marl_responsible_nav_b200 import BatchedGridWorld, FusedActor, maddpg

for E in [int(x) for x in (sys.argv[1:] or ["4096", "65536", "1048576"])]:
    env = BatchedGridWorld("Level 3", num_envs=E, fear=False, auto_reset=True, seed=1)
    agent = maddpg.BatchedMADDPG(2, 160, 9, device="cuda", seed=0)
    fused = FusedActor(env, agent.actors, seed=1)
    out = env.reset()
    for training in (True, False):
        for _ in range(5):
            fused.forward(out.obs_code, out.action_mask, training=training)
        torch.cuda.synchronize()
        # 50 launches replayed from a CUDA graph: at small batches the eager call rate (~9 us per Python / ctypes call) would
        # be what is measured
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(50):
                fused.forward(out.obs_code, out.action_mask, training=training)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 50 * 1e3
        rows = E * 2
        print(json.dumps({"envs": E, "training": training, "actor_kernel_us": us, "tflops": rows * 76032 / us / 1e6}), flush=True)
    del env, fused, agent
