"""Dev: where a step of the resident kernel (gw_step_host mode 2) spends its time, per CTA (GW_TRACE=1, -DGW_ENABLE_TRACE)."""
import os, sys, time, ctypes as C, subprocess
os.environ["GW_TRACE"] = "1"
_root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "marl_responsible_nav_b200", "csrc")
_lib = os.path.join(_root, "libgridworld_b200_trace.so")
if not os.path.exists(_lib) or any(os.path.getmtime(os.path.join(_root, f)) > os.path.getmtime(_lib) for f in ("gw_kernels.cu", "gw_device.cuh", "gw_actor.cu")):
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
                    "-DGW_ENABLE_TRACE", "-o", _lib, "gw_kernels.cu", "gw_actor.cu"], cwd=_root, check=True)
os.environ["GW_LIB"] = _lib
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from marl_responsible_nav_b200 import BatchedGridWorld
E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
K = 2000
env = BatchedGridWorld("Level 3", num_envs=E, fear=True, fear_weight=-5.0, auto_reset=True, seed=1)
env.reset()
ha = torch.randint(0, 9, (8, E, 2), dtype=torch.int8).pin_memory()
hr = torch.empty((E, 2), dtype=torch.float32).pin_memory()
he = torch.empty((E,), dtype=torch.uint8).pin_memory()
for i in range(8): env.step_host(ha[i], hr, he, resident=True)
env.sync()   # the traced launch starts here
t0 = time.perf_counter()
for i in range(K): env.step_host(ha[i % 8], hr, he, resident=True)
el = (time.perf_counter() - t0) / K * 1e6
env.sync()
n = (E + 31) // 32
buf = np.zeros((n, 16), np.uint64)
env.lib.gw_debug_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
assert env.lib.gw_debug_trace(env._h, buf.ctypes.data, n) == 0
b = buf.astype(np.float64)
rounds = b[:, 5]
print(f"E={E}: {el:.2f} us per step_host (host clock), {int(rounds[0])} steps traced on {n} CTAs")
for k, nm in enumerate(["wait for the doorbell", "table entry -> smem", "step (tiles)", "bulk-store drain + barrier", "fence + arrive (+ flag)"]):
    v = b[:, k] / rounds / 1e3
    print(f"  {nm:28s} CTA0 {v[0]:7.2f} us   median {np.median(v):7.2f}   max {v.max():7.2f}")
