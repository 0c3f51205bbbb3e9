"""Replay sampling: gw_replay_sample (one kernel) vs the PyTorch indexing formulation, CUDA-event timed."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from marl_responsible_nav_b200 import BatchedGridWorld  # noqa: E402
from marl_responsible_nav_b200.replay import ReplayRing  # noqa: E402

E = 4096
env = BatchedGridWorld("Level 3", num_envs=E, fear=True, fear_weight=-10.0, auto_reset=True, seed=1)
ring = ReplayRing(E, env.n_learners, env.obs_len, capacity=200_000)
gen = torch.Generator(device="cuda").manual_seed(0)
env.reset(obs_out=ring.obs_slot(0))
for t in range(60):
    acts = torch.randint(0, 9, (E, 2), generator=gen, device="cuda", dtype=torch.int8)
    env.step(acts, obs_out=ring.obs_slot(t + 1), final_obs_out=ring.final_slot(t), buffers=ring.buffers_slot(t))
    ring.advance()


def timed(fn, n=300, warm=30):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


for B in (128, 1024, 16384):
    out = ring.new_batch(B)
    us_f = timed(lambda: ring.sample_fused(env, B, seed=1, out=out))
    us_t = timed(lambda: ring.sample(B, gen))
    print(json.dumps({"batch": B, "gw_replay_sample_us": us_f, "torch_indexing_us": us_t,
                      "bytes_gathered": B * (2 * 2 * 160 * 4 + 2 * 9 * 4 + 16)}))
