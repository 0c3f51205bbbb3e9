"""Data-parallel training through the update kernel: every rank owns an env shard and a replay ring, each update is cut
at the two gradient exchanges (three launches, NCCL all-reduce of the critics' / actors' flat gradients in between) and the
ranks' networks stay bit-identical.  The two-rank test needs two GPUs (`gpurun --gpus 2`); the one-rank test drives the
same host loop on one GPU."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


@pytest.mark.gpu
def test_segmented_host_loop_equals_one_launch():
    """FusedLearner.learn_from_ring's several-rank path (three launches per update) forced on one rank: the same
    parameters, bit for bit, as the one-launch path -- same draws, same noise keys, same arithmetic."""
    from marl_responsible_nav_b200 import maddpg
    from marl_responsible_nav_b200.learner import FusedLearner
    dev = torch.device("cuda", 0)
    hp = maddpg.preset("custom_fear_5")
    hp["MEMORY_SIZE"] = 640
    env = maddpg.make_env(hp, 64)
    tr = maddpg.BatchedTrainer(env, hp=hp, seed=0, fused_learner=False)
    tr.train(12, learn=False)
    la, lb = (FusedLearner(env, maddpg.BatchedMADDPG(2, 160, 9, hp=hp, device=dev, seed=8), seed=4) for _ in range(2))
    lb.force_segmented = True
    d0 = tr.ring._draws
    l1 = la.learn_from_ring(tr.ring, 6, 77).clone()
    tr.ring._draws = d0
    l2 = lb.learn_from_ring(tr.ring, 6, 77).clone()
    env.sync()
    assert torch.equal(l1, l2) and la.updates_done == lb.updates_done == 6
    for x, y in ((la.params, lb.params), (la.targets, lb.targets), (la.adam_m, lb.adam_m), (la.adam_v, lb.adam_v)):
        assert torch.equal(x, y)
    assert la.adam_steps.tolist() == lb.adam_steps.tolist() == [6.0] * 4


def _worker(rank, world, port, out_dir, exchange="peer"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    if exchange.startswith("peer-"):                   # the two forms of the in-kernel exchange: all-gather / owner sums and returns
        os.environ["GW_PEER_PROTOCOL"] = exchange[5:]
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from marl_responsible_nav_b200 import maddpg, sharding
    hp = maddpg.preset("custom_fear_5")
    hp["MEMORY_SIZE"] = 8192
    base, n = sharding.shard_range(512, rank, world)
    env = maddpg.make_env(hp, n, device=dev, env_id_base=base)
    tr = maddpg.BatchedTrainer(env, hp=hp, seed=3, global_envs=512, gradient_exchange=exchange.split("-")[0])
    how = tr.connect(0)
    tr.train(5)
    torch.cuda.synchronize()
    L = tr.learner
    chk = torch.stack([L.params.double().sum(), L.params.double().abs().sum(), L.targets.double().sum(), L.adam_m.double().abs().sum(),
                       L.adam_v.double().sum(), torch.tensor(float(tr.updates_done), device=dev, dtype=torch.float64)])
    allc = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(allc, chk)
    torch.save({"chk": [c.cpu() for c in allc], "finite": bool(torch.isfinite(L.params).all()), "updates": tr.updates_done,
                "kernel": L.kernel, "exchange": how, "timed_out": L.peer_timed_out()}, os.path.join(out_dir, f"{exchange}_rank{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.gpu
def test_two_rank_training_keeps_the_networks_identical(tmp_path):
    """Two ranks, the gradient exchanges: inside the update kernel over NVLink peer memory ("peer": one launch per block of
    updates, no NCCL call; as all-gather and as owner-sums-and-returns) and with two NCCL all-reduces per update ("nccl").  Ranks stay bit-identical either way, and with
    two ranks the two exchanges give the SAME parameters bit for bit (a sum of two terms has one order)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    res = {}
    for k, exchange in enumerate(("peer-ag", "peer-rs", "nccl")):
        port = 29800 + (os.getpid() % 1000) + k
        mp.spawn(_worker, args=(2, port, str(tmp_path), exchange), nprocs=2, join=True)
        r0 = torch.load(os.path.join(str(tmp_path), f"{exchange}_rank0.pt"))
        r1 = torch.load(os.path.join(str(tmp_path), f"{exchange}_rank1.pt"))
        assert r0["exchange"] == r1["exchange"] == exchange.split("-")[0] and not r0["timed_out"] and not r1["timed_out"]
        assert r0["finite"] and r1["finite"]
        assert r0["updates"] == r1["updates"] == 5 * (512 // 10)          # maddpg/agent.py:214-224 on the global env count
        assert all(torch.equal(a, b) for a, b in zip(r0["chk"], r0["chk"][1:]))   # both ranks' checksums, as gathered on rank 0
        assert all(torch.equal(a, b) for a, b in zip(r0["chk"], r1["chk"]))
        res[exchange] = r0["chk"][0]
    assert torch.equal(res["peer-ag"], res["nccl"]) and torch.equal(res["peer-rs"], res["nccl"])
