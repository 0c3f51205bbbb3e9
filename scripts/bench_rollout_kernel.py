"""Time gw_rollout (T steps per launch) against one gw_step launch per step replayed from a CUDA graph."""
import sys
import torch
sys.path.insert(0, ".")
from marl_responsible_nav_b200 import BatchedGridWorld

def run(E, T, fear=True, reps=20):
    env = BatchedGridWorld("Level 3", num_envs=E, fear=fear, fear_weight=-5.0, auto_reset=True, max_steps=150, seed=42)
    L = env.n_learners
    obs_bytes = E * L * env.obs_len * 4
    slots = max(T + 1, -(-(320 << 20) // obs_bytes))
    rings = env.new_rings(slots, fields=("obs", "reward", "shaped_reward", "fear", "terminated", "truncated", "ended", "info", "positions", "obs_code", "action_mask"))
    gen = torch.Generator(device="cuda").manual_seed(1)
    acts = torch.randint(0, 9, (64, E, L), generator=gen, device="cuda", dtype=torch.int8)
    env.reset(obs_out=rings.obs[0])
    t = 0
    for _ in range(3):
        env.rollout(acts, T, rings, first_slot=t % slots, first_action=t % 64); t += T
    env.sync()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(reps):
        env.rollout(acts, T, rings, first_slot=t % slots, first_action=t % 64); t += T
    ev1.record()
    torch.cuda.synchronize()
    us = ev0.elapsed_time(ev1) * 1e3 / (reps * T)
    print(f"E={E} T={T} fear={fear}: gw_rollout {us:.2f} us/step, {E * L / us / 1e3:.3f} G agent-steps/s, "
          f"{E * L * 668 / us / 1e3:.0f} GB/s algorithmic", flush=True)
    env.close()

for E in (4096, 16384, 65536):
    for T in (8, 64, 256):
        run(E, T)
run(4096, 64, fear=False)
