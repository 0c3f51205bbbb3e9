"""Rollout loop on the device: fused actor forward (K5) -> gw_step (K1-K4, K6), obs into the replay ring."""
import json, sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from marl_responsible_nav_b200 import BatchedGridWorld, FusedActor, ReplayRing, maddpg

def run(E, fear, steps=128, torch_actor=False):
    env = BatchedGridWorld("Level 3", num_envs=E, fear=bool(fear), fear_weight=-5.0, auto_reset=True, seed=1)
    agent = maddpg.BatchedMADDPG(2, 160, 9, device="cuda", seed=0)
    fused = FusedActor(env, agent.actors, seed=1)
    slots = max(3, min(64, (6 << 30) // (E * 2 * 160 * 4)))
    ring = torch.empty((slots, E, 2, 160), device="cuda")
    out = env.reset(obs_out=ring[0])
    def loop(n, t0):
        nonlocal out
        for t in range(t0, t0 + n):
            if torch_actor:
                cont, ids = agent.get_action(ring[t % slots], out.action_mask, training=True)
            else:
                cont, ids = fused.forward(out.obs_code, out.action_mask, training=True)
            out = env.step(ids, obs_out=ring[(t + 1) % slots])
    loop(16, 0)
    torch.cuda.synchronize()
    g = None
    if not torch_actor:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            loop(64, 16)
        g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if g is None:
        loop(steps, 16)
    else:
        for _ in range(steps // 64): g.replay()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    # actor kernel alone
    e0.record()
    for _ in range(50): fused.forward(out.obs_code, out.action_mask, training=True)
    e1.record(); torch.cuda.synchronize()
    return {"envs": E, "fear": bool(fear), "actor": "torch fp32" if torch_actor else "fused tcgen05", "ms_per_rollout_step": ms,
            "agent_steps_per_s": E * 2 / ms * 1e3, "actor_kernel_us": e0.elapsed_time(e1) / 50 * 1e3}

if __name__ == "__main__":
    for E in (4096, 65536, 1 << 20):
        for ta in (False, True):
            print(json.dumps(run(E, 1, torch_actor=ta)))
