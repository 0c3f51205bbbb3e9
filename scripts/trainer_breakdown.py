"""Where the training loop's time goes (custom_fear_10, 4096 envs, one GPU): GPU time of the captured update, host time
of one learn() call, host and GPU time of the rollout steps between two updates."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from marl_responsible_nav_b200 import maddpg  # noqa: E402

hp = maddpg.preset("custom_fear_10")
env = maddpg.make_env(hp, 4096)
res = {}
for name, kw in (("default", {}), ("fused_linear_bwd", {"fused_linear_bwd": True}), ("torch_ops", {"fused_ops": False})):
    tr = maddpg.BatchedTrainer(env, hp=hp, seed=0, learn_cadence="batched", **kw)
    tr.train(200)                                        # graph captured, ring filled
    ag = tr.agent
    torch.cuda.synchronize()
    g = ag._graph[0]
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(100):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    upd_gpu = a.elapsed_time(b) * 10                     # us per replay (back to back: includes the launch rate of the host)
    t0 = time.perf_counter()
    for _ in range(100):
        ag.learn(tr._sample(hp["BATCH_SIZE"]))
    host_learn = (time.perf_counter() - t0) * 1e4        # us per call, no synchronisation (queue depth grows)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    tr.train(1000, learn=False)
    host_roll = (time.perf_counter() - t0) * 1e3         # us per env step incl. the final stats synchronisation
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    tr.train(1000)
    loop = (time.perf_counter() - t0) * 1e3
    res[name] = {"update_graph_us": upd_gpu, "learn_call_host_us": host_learn, "rollout_only_us_per_env_step": host_roll,
                 "loop_us_per_env_step": loop, "graph_nodes": None}
    print(json.dumps({name: res[name]}), flush=True)
