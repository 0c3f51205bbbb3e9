"""Dev: phase timeline of gw_step at a given batch (GW_TRACE=1).  Prints per-phase medians over CTAs in microseconds."""
import os, sys, ctypes as C
os.environ["GW_TRACE"] = "1"
# the stamps are compiled in only with -DGW_ENABLE_TRACE: build that variant next to the product library
import subprocess
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
fear = int(sys.argv[2]) if len(sys.argv) > 2 else 1
env = BatchedGridWorld("Level 3", num_envs=E, fear=bool(fear), auto_reset=True, seed=1)
env.reset()
a = torch.randint(0, 9, (E, 2), device="cuda", dtype=torch.int8)
for _ in range(20): env.step(a)
torch.cuda.synchronize()
n = min(4096, (E + 31) // 32 if E <= 32768 else 592)
buf = np.zeros((n, 16), np.uint64)
env.lib.gw_debug_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
rc = env.lib.gw_debug_trace(env._h, buf.ctypes.data, n)
assert rc == 0
b = buf.astype(np.int64)
b = b[b[:, 0] > 0]
t0 = b[:, 0].min()
names = ["start", "tables+P1a done", "P1b done", "P2 barrier", "P2 done", "P3 done", "P4 barrier", "tile end",
         "P1b: sampled", "P1b: geometry", "P1b: world_update", "P1b: fear tasks", "P1b: rewards", "P1b: outputs", "P1b: spawn", "-"]
if os.environ.get("GW_SMALL", "1") != "0" and E <= 6144:
    names = ["start", "tables+P1a done", "step done (warp 0)", "cp.async issued", "state arrived", "philox done", "render done (warp 0)", "tile end (warp 0)",
             "traj + effw", "collisions", "final cells", "rewards+spawn", "outputs", "specials+masks", "tables landed + barrier", "-"]
    order = [0, 3, 4, 5, 14, 1, 8, 9, 10, 11, 12, 13, 2, 6, 7]
print(f"E={E} fear={fear} CTAs={len(b)}  kernel span = {(b[:, 7].max() - t0) / 1e3:.2f} us (first start -> last end)")
order = locals().get('order', [0, 1, 8, 9, 10, 11, 12, 13, 14, 2, 3, 4, 5, 6, 7])
for i in order:
    nm = names[i]
    if not (b[:, i] > 0).all(): continue
    rel = (b[:, i] - t0) / 1e3
    print(f"  {nm:18s} median {np.median(rel):7.2f}  min {rel.min():7.2f}  max {rel.max():7.2f} us")
