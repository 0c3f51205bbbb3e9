"""Dev: where the time of the host-driven step goes (Python marshalling vs library call vs kernel)."""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from marl_responsible_nav_b200 import BatchedGridWorld
from marl_responsible_nav_b200 import _native as N
E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
env = BatchedGridWorld("Level 3", num_envs=E, fear=True, fear_weight=-5.0, auto_reset=True, seed=1)
env.reset()
L = env.n_learners
ha = torch.randint(0, 9, (64, E, L), dtype=torch.int8).pin_memory()
hr = torch.empty((E, L), dtype=torch.float32).pin_memory()
he = torch.empty((E,), dtype=torch.uint8).pin_memory()
K = 2000
def timeit(f, k=K):
    for i in range(50): f(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(k): f(i)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / k * 1e6
print("step_host memcpy   : %.2f us" % timeit(lambda i: env.step_host(ha[i % 64], hr, he, zero_copy=False)))
print("step_host zero-copy: %.2f us" % timeit(lambda i: env.step_host(ha[i % 64], hr, he, zero_copy=True, resident=False)))
for i in range(64): env.step_host(ha[i], hr, he, resident=True)
print("step_host resident : %.2f us" % timeit(lambda i: env.step_host(ha[i % 64], hr, he, resident=True)))
print("  server:", env.server_info())
env.sync()
# raw library call with prebuilt arguments
io = env._io(env.buf.obs, None, env._host_act_dev, None, None)
ioref = C.byref(io)
pa = [C.c_void_p(ha[i].data_ptr()) for i in range(64)]
pr, pe = C.c_void_p(hr.data_ptr()), C.c_void_p(he.data_ptr())
st = env._stream()
lib, h = env.lib, env._h
print("raw gw_step_host resident : %.2f us" % timeit(lambda i: lib.gw_step_host(h, ioref, pa[i % 64], pr, None, pe, 2, st)))
env.sync()
print("raw gw_step_host zero-copy: %.2f us" % timeit(lambda i: lib.gw_step_host(h, ioref, pa[i % 64], pr, None, pe, 1, st)))
print("raw gw_step_host memcpy   : %.2f us" % timeit(lambda i: lib.gw_step_host(h, ioref, pa[i % 64], pr, None, pe, 0, st)))
da = torch.randint(0, 9, (E, L), dtype=torch.int8, device="cuda")
io2 = env._io(env.buf.obs, None, da, None, None)
io2ref = C.byref(io2)
print("raw gw_step (async, queue deep): %.2f us" % timeit(lambda i: lib.gw_step(h, io2ref, st)))
def f(i):
    lib.gw_step(h, io2ref, st); lib.gw_sync(h, st)
print("raw gw_step + gw_sync          : %.2f us" % timeit(f))
print("python step() + sync()         : %.2f us" % timeit(lambda i: (env.step(da), env.sync())))
print("python step() async            : %.2f us" % timeit(lambda i: env.step(da)))
