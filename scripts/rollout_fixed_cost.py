"""Per-launch fixed cost of gw_rollout: the same 256 env steps as launches of T steps each (one CUDA graph per T),
time = launches * a + steps * b.  usage: python scripts/rollout_fixed_cost.py [envs]"""
import sys
import torch
sys.path.insert(0, ".")
from marl_responsible_nav_b200 import BatchedGridWorld

E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
STEPS = 256
env = BatchedGridWorld("Level 3", num_envs=E, fear=True, fear_weight=-5.0, auto_reset=True, max_steps=150, seed=42)
L = env.n_learners
obs_bytes = E * L * env.obs_len * 4
slots = max(STEPS + 1, -(-(320 << 20) // obs_bytes))
rings = env.new_rings(slots, fields=("obs", "reward", "shaped_reward", "fear", "terminated", "truncated", "ended", "info", "positions", "obs_code", "action_mask"))
gen = torch.Generator(device="cuda").manual_seed(1)
acts = torch.randint(0, 9, (64, E, L), generator=gen, device="cuda", dtype=torch.int8)
env.reset(obs_out=rings.obs[0])
res = []
for T in (1, 2, 4, 8, 16, 20, 32, 64):
    def run():
        t = 0
        while t < STEPS:
            env.rollout(acts, T, rings, first_slot=t % slots, first_action=t % 64)
            t += T
    run(); env.sync()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        run()
    g.replay(); env.sync()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(5):
        g.replay()
    ev1.record()
    torch.cuda.synchronize()
    us = ev0.elapsed_time(ev1) * 1e3 / 5
    n_l = -(-STEPS // T)
    res.append((T, n_l, us))
    print(f"E={E} T={T}: {n_l} launches, {us:.1f} us for {STEPS} steps = {us / STEPS:.3f} us/step", flush=True)
(t1, n1, u1), (t2, n2, u2) = res[0], res[-1]
a = (u1 - u2) / (n1 - n2)
b = (u2 - n2 * a) / STEPS
print(f"fit: a = {a:.2f} us per launch, b = {b:.3f} us per step")
