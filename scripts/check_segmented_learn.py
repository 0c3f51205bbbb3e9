"""torchrun --nproc-per-node 2 scripts/check_segmented_learn.py : the multi-rank update (chain of CUDA graphs cut at the
gradient all-reduces) keeps the ranks' parameters identical, reduces the critic loss on a fixed batch, and its cost."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch, torch.distributed as dist
from marl_responsible_nav_b200 import maddpg

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
ag = maddpg.BatchedMADDPG(2, 160, 9, device=dev, seed=7)
ag.broadcast_parameters(0)
g = torch.Generator(device=dev).manual_seed(100 + rank)            # every rank its own data
def batch():
    return {"state": torch.randn(128, 2, 160, device=dev, generator=g), "next_state": torch.randn(128, 2, 160, device=dev, generator=g),
            "action": torch.rand(128, 2, 9, device=dev, generator=g), "reward": torch.randn(128, 2, device=dev, generator=g),
            "done": torch.zeros(128, 2, device=dev)}
fixed = batch()
first = None
for i in range(40):
    st = ag.learn(fixed)
    if i == 0: first = float(st.critic_loss.sum())
last = float(st.critic_loss.sum())
assert ag._segments is not None, "segmented path not taken"
for i in range(20):
    ag.learn(batch())
chk = torch.stack([p.detach().double().sum() for p in ag.parameters()] + [p.detach().double().abs().sum() for p in ag.parameters()])
allc = [torch.zeros_like(chk) for _ in range(world)]
dist.all_gather(allc, chk)
same = all(torch.equal(allc[0], c) for c in allc)
torch.cuda.synchronize(); t0 = time.perf_counter()
b = batch()
for i in range(50): ag.learn(b)
torch.cuda.synchronize(); us = (time.perf_counter() - t0) / 50 * 1e6
if rank == 0:
    print(f"ranks in sync: {same}; critic loss on a fixed batch {first:.3f} -> {last:.3f}; {us:.0f} us per update on {world} ranks")
assert same and last < first
dist.destroy_process_group()
