"""ncu target: a few steps of the general layout's step kernel (20 x 28 x 7 scenario, 65 536 envs)."""
import os
import sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from marl_responsible_nav_b200 import BatchedGridWorld, load_scenario_json
fear = int(sys.argv[1]) if len(sys.argv) > 1 else 1
E = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
sc = load_scenario_json(os.path.join(ROOT, "tests", "golden", "wide_scenarios.json"), "Wide 20x28", walls="enforce")
env = BatchedGridWorld(sc, num_envs=E, fear=bool(fear), fear_weight=-5.0, seed=42)
env.reset()
acts = torch.randint(0, 9, (8, E, 2), dtype=torch.int8, device="cuda")
for t in range(12):
    env.step(acts[t % 8])
env.sync()
print("ok", env.stats()["episodes"])
