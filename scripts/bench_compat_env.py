"""Dev: per-step latency of the single-env drop-in classes (CustomMAEnv / CustomEnv, E = 1)."""
import os, sys, time, random
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from marl_responsible_nav_b200 import CustomMAEnv, CustomEnv

def run(make, act, n=3000):
    random.seed(0); np.random.seed(0)
    env = make()
    obs, info = env.reset()
    rng = np.random.default_rng(0)
    t0 = time.perf_counter(); steps = 0
    while steps < n:
        out = env.step(act(rng))
        steps += 1
        done = (all(out[3].values()) or all(out[2].values())) if isinstance(out[3], dict) else (out[2][0] or out[3])
        if done or steps % 150 == 0:
            env.reset()
    return (time.perf_counter() - t0) / n * 1e6

for fear in (False, True):
    print(f"CustomMAEnv fear={fear}: %.1f us per step" % run(lambda: CustomMAEnv(fear=fear, seed=1), lambda r: tuple(int(x) for x in r.integers(0, 9, 2))))
    print(f"CustomEnv   fear={fear}: %.1f us per step" % run(lambda: CustomEnv(fear=fear), lambda r: [int(r.integers(0, 9))]))
