"""Time the UNMODIFIED Python reference env on this machine's host cores (test / bench infrastructure, not product).

The reference is pure Python (custom/ma_customenv.py:217-334 and what it calls).  `stage()` -- run by
`__graft_entry__.build()` in the build container, where /root/reference is mounted -- copies the hot-path files as they
are into the git-ignored `baseline/_ref/` so that they travel to the GPU box with the tree (like the built .so files);
nothing of them is tracked.  `python oracle/reference_python.py --seconds S --procs P` then drives P independent
`CustomMAEnv` instances (one process per core, the three absent third-party modules stubbed as in
tests/golden/_ref_loader.py) with uniform learner actions, episodes ended on all-terminated / all-truncated or 150
steps, with FeAR off and on, and prints one JSON object: per-core and aggregate agent-steps/s and the core count.
Only bench.py's cpu_baseline leg and scripts call this; the product never does."""
import argparse
import json
import multiprocessing as mp
import os
import shutil
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED = os.path.join(ROOT, "baseline", "_ref")
FILES = ("custom/grid_world.py", "custom/custom_agent.py", "custom/Responsibility.py", "custom/ma_customenv.py",
         "custom/customenv.py", "custom/Scenarios.json")


def stage(src_root="/root/reference", dst_root=STAGED):
    """Copy the reference's hot-path files, unmodified, to baseline/_ref (git-ignored).  Returns True when staged."""
    if not os.path.isfile(os.path.join(src_root, FILES[0])):
        return False
    for f in FILES:
        dst = os.path.join(dst_root, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(src_root, f), dst)
    return True


def reference_root():
    for r in ("/root/reference", STAGED):
        if os.path.isfile(os.path.join(r, FILES[0])):
            return r
    return None


def _worker(args):
    root, fear, seconds, seed = args
    os.environ["GW_REFERENCE_ROOT"] = root
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import contextlib
    import io
    import random
    import numpy as np
    import _ref_loader
    _ref_loader.REF_ROOT = root
    ref = _ref_loader.load_reference()
    with contextlib.redirect_stdout(io.StringIO()):
        env = ref.ma_customenv.CustomMAEnv(render=False, fear=fear, seed=seed)
        random.seed(seed)
        np.random.seed(seed)
        rng = np.random.default_rng(seed + 1)
        env.reset()
        steps = episodes = t_ep = 0
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < seconds:
            _, _, term, trunc, _ = env.step(tuple(int(a) for a in rng.integers(0, 9, size=2)))
            steps += 1
            t_ep += 1
            if all(term.values()) or all(trunc.values()) or t_ep >= 150:
                env.reset()
                episodes += 1
                t_ep = 0
        el = time.perf_counter() - t0
    return steps, el, episodes


def measure(seconds=8.0, procs=None):
    root = reference_root()
    if root is None:
        return {"kind": "reference-python", "unavailable": "reference files not staged (baseline/_ref is made by build() "
                                                           "where /root/reference is mounted)"}
    procs = procs or os.cpu_count() or 1
    out = {"kind": "reference-python", "cores": procs, "unit": "agent-steps/s", "root": root,
           "what": "unmodified custom/ma_customenv.py CustomMAEnv (stub pettingzoo/gymnasium/pygame), one process per core, "
                   f"uniform learner actions, {seconds:.0f} s per setting"}
    ctx = mp.get_context("fork")
    for fear in (False, True):
        with ctx.Pool(procs) as pool:
            res = pool.map(_worker, [(root, fear, seconds, 42 + i) for i in range(procs)])
        rates = [2.0 * s / el for s, el, _ in res]
        out["fear_on" if fear else "fear_off"] = {
            "value": sum(rates), "per_core": sum(rates) / len(rates), "env_steps": sum(s for s, _, _ in res),
            "episodes": sum(e for _, _, e in res)}
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=8.0)
    ap.add_argument("--procs", type=int, default=0)
    ap.add_argument("--stage", action="store_true")
    a = ap.parse_args()
    if a.stage:
        print(json.dumps({"staged": stage()}))
    else:
        print(json.dumps(measure(a.seconds, a.procs or None)))
