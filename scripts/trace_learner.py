"""Per-phase time of the last update of a gw_learner_update launch (clock64 of CTA 0 at every phase start)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from marl_responsible_nav_b200 import maddpg  # noqa: E402

hp = maddpg.preset("custom_fear_10")
world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
if world > 1:                                            # torchrun: the in-kernel gradient exchange over peer memory
    import torch.distributed as dist
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
env = maddpg.make_env(hp, E, device=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))), env_id_base=rank * E)
tr = maddpg.BatchedTrainer(env, hp=hp, seed=0, global_envs=E * world)
if world > 1:
    print("gradient exchange:", tr.connect(0))
tr.train(60, learn=False)
if len(sys.argv) > 2:
    tr.learner.set_kernel(sys.argv[2])
for _ in range(3):
    tr.learner.learn_from_ring(tr.ring, 16, tr.sample_seed)
torch.cuda.synchronize()
t = tr.learner.debug_tensor("trace").view(torch.int64).cpu().tolist()
names = tr.learner.phase_names if tr.learner.kernel == "phase" else ("A: critic gradients", "Adam C", "B: actor gradients", "Adam A")
n = len(names)
mhz = 1965.0
tot = (t[n] - t[0]) / mhz
if rank != 0:
    sys.exit(0)
print(f"update: {tot:.1f} us over {n} phases (clock64 at {mhz:.0f} MHz)")
for k in range(n):
    print(f"  phase {k:2d} {names[k]:20s} {(t[k + 1] - t[k]) / mhz:7.2f} us")

if tr.learner.kernel == "cluster" and t[8]:                # built with GW_NVCC_FLAGS=-DGW_LEARN_TRACE: checkpoints inside phases A and B
    pts = [x for x in t[8:108] if x]
    print("checkpoints (us since phase A start):", " ".join(f"{(x - t[0]) / mhz:.1f}" for x in pts))
