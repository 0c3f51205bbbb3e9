#!/bin/bash
# compute-sanitizer memcheck on a small end-to-end case (one tool per gpurun call, see B200_PROFILING.md)
cat > /tmp/san_case.py <<'PY'
import sys; sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle")
import numpy as np, torch
from marl_responsible_nav_b200 import BatchedGridWorld
for kw in (dict(fear=True), dict(fear=True, env_kind="single"), dict(fear=False, obs_dtype=torch.bfloat16), dict(fear=True, scenario="Level 5")):
    sc = kw.pop("scenario", "Level 3")
    for E in (77, 1000):
        env = BatchedGridWorld(sc, num_envs=E, auto_reset=True, max_steps=20, seed=1, **kw)
        fin = torch.zeros_like(env.buf.obs)
        out = env.reset()
        g = torch.Generator(device="cuda").manual_seed(0)
        for t in range(30):
            a = torch.randint(0, 9, (E, env.n_learners), generator=g, device="cuda", dtype=torch.int8)
            out = env.step(a, final_obs_out=fin)
        m = torch.zeros(E, dtype=torch.uint8, device="cuda"); m[::3] = 1
        env.reset(mask=m)
        env.sync(); print(sc, kw, E, env.stats()["episodes"])
print("SANITIZE CASE OK")
PY
python /tmp/san_case.py > gpurun_out/san_plain.log 2>&1 && timeout 800 compute-sanitizer --tool memcheck --error-exitcode 7 python /tmp/san_case.py > gpurun_out/san_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -5 gpurun_out/san_memcheck.log
