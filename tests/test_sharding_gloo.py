"""CPU, world_size 2 over gloo: the host-side sharding logic (contiguous global-id ranges, statistics all-reduce,
max-over-ranks timing).  Each rank advances its shard with the C oracle (same RNG spec as the kernels): the shards
must reproduce the single-process run env for env, and the all-reduced statistics must equal its totals."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
E_GLOBAL, STEPS = 1001, 12          # odd on purpose: uneven shards


def _worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import c_oracle
    from marl_responsible_nav_b200 import sharding
    base, n = sharding.shard_range(E_GLOBAL, rank, world)
    o = c_oracle.COracle("Level 3", num_envs=n, fear=True, auto_reset=True, max_steps=150, seed=5, env_id_base=base)
    o.reset()
    rng = np.random.default_rng(99)
    acts = rng.integers(0, 9, size=(STEPS, E_GLOBAL, 2)).astype(np.int8)
    pos = []
    for t in range(STEPS):
        o.step(acts[t, base:base + n])
        pos.append(o.positions.copy())
    total = sharding.allreduce_stats(o.stats())
    tmax = sharding.max_over_ranks(1.0 + rank)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), pos=np.stack(pos), base=base, n=n, fear=o.fear,
             total=np.array([total[k] for k in sharding.STAT_KEYS], dtype=np.float64), tmax=tmax)
    dist.destroy_process_group()


def test_shard_ranges_tile_exactly():
    from marl_responsible_nav_b200.sharding import shard_range
    for E in (1, 7, 4096, 1_000_000):
        for world in (1, 2, 3, 8):
            spans = [shard_range(E, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(n for _, n in spans) == E
            for (b0, n0), (b1, _) in zip(spans, spans[1:]):
                assert b0 + n0 == b1
            assert max(n for _, n in spans) - min(n for _, n in spans) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_two_rank_gloo_matches_single_process(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import c_oracle
    from marl_responsible_nav_b200 import sharding
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    ref = c_oracle.COracle("Level 3", num_envs=E_GLOBAL, fear=True, auto_reset=True, max_steps=150, seed=5)
    ref.reset()
    rng = np.random.default_rng(99)
    acts = rng.integers(0, 9, size=(STEPS, E_GLOBAL, 2)).astype(np.int8)
    pos = []
    for t in range(STEPS):
        ref.step(acts[t])
        pos.append(ref.positions.copy())
    pos = np.stack(pos)
    want = ref.stats()
    for r in range(2):
        d = np.load(os.path.join(str(tmp_path), f"rank{r}.npz"))
        b, n = int(d["base"]), int(d["n"])
        assert np.array_equal(d["pos"], pos[:, b:b + n])                  # shard == slice of the single-process run
        assert np.array_equal(d["fear"], ref.fear[b:b + n])
        got = dict(zip(sharding.STAT_KEYS, d["total"]))
        for k in sharding.STAT_KEYS[:8]:
            assert int(got[k]) == int(want[k]), k
        assert abs(got["return_sum"] - want["return_sum"]) < 1e-9
        assert abs(got["fear_sum"] - want["fear_sum"]) < 1e-9
        assert float(d["tmax"]) == 2.0                                    # max over ranks
