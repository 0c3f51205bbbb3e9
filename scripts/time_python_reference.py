"""Time the UNMODIFIED Python reference env (build container only: needs /root/reference).  One process, one core.
Output is committed under profiles/ as the record of what the reference itself does on a CPU."""
import os, sys, time, random
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden"))
from _ref_loader import load_reference
import contextlib, io
REF = load_reference()
for fear in (False, True):
    with contextlib.redirect_stdout(io.StringIO()):
        env = REF.ma_customenv.CustomMAEnv(render=False, fear=fear, seed=42)
    random.seed(0); np.random.seed(0)
    rng = np.random.default_rng(1)
    env.reset()
    budget = 10.0 if fear else 6.0
    steps = episodes = t_ep = 0
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < budget:
        _, _, term, trunc, _ = env.step(tuple(int(a) for a in rng.integers(0, 9, size=2)))
        steps += 1; t_ep += 1
        if all(term.values()) or all(trunc.values()) or t_ep >= 150:
            env.reset(); episodes += 1; t_ep = 0
    el = time.perf_counter() - t0
    print(f"reference CustomMAEnv fear={fear}: {steps} env-steps in {el:.1f} s -> {steps/el:.1f} env-steps/s = {2*steps/el:.1f} agent-steps/s "
          f"(1 core, {episodes} episodes, mean length {steps/max(1,episodes):.1f})")
