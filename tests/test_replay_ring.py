"""ReplayRing: index logic on CPU tensors (no kernels involved), and on the GPU fed by gw_step."""
import numpy as np
import pytest
import torch

from marl_responsible_nav_b200.replay import ReplayRing


def test_ring_indexing_cpu():
    E, L, OL = 6, 2, 8
    ring = ReplayRing(E, L, OL, capacity=5 * E, device="cpu")
    assert ring.T == 6
    log = {}
    ring.obs_slot(0).copy_(torch.full((E, L, OL), 0.0))
    for t in range(17):
        ended = torch.tensor([(t + e) % 5 == 0 for e in range(E)])
        b = ring.buffers_slot(t)
        b.shaped_reward.copy_(torch.full((E, L), float(t)))
        b.terminated.copy_((ended[:, None] & torch.tensor([True, False])[None, :]).to(torch.uint8))
        b.ended.copy_(ended.to(torch.uint8))
        ring.final_slot(t)[ended] = 1000.0 + t                        # written only where the episode ended
        ring.obs_slot(t + 1).copy_(torch.full((E, L, OL), float(t + 1)))
        ring.store_action(t, torch.full((E, L, 9), 0.5 * t))
        ring.advance()
        log[t] = ended.clone()
    assert len(ring) == (ring.T - 1) * E
    g = torch.Generator().manual_seed(0)
    batch = ring.sample(512, g)
    for i in range(512):
        t, e = int(batch["t"][i]), int(batch["env"][i])
        assert 17 - (ring.T - 1) <= t <= 16
        assert float(batch["state"][i, 0, 0]) == float(t)
        assert float(batch["reward"][i, 0]) == float(t) and float(batch["action"][i, 1, 3]) == 0.5 * t
        if log[t][e]:
            assert float(batch["next_state"][i, 0, 0]) == 1000.0 + t and int(batch["done"][i, 0]) == 1
        else:
            assert float(batch["next_state"][i, 0, 0]) == float(t + 1) and int(batch["done"][i, 0]) == 0
    empty = ReplayRing(E, L, OL, capacity=10, device="cpu")
    with pytest.raises(RuntimeError):
        empty.sample(4)


@pytest.mark.gpu
def test_ring_filled_by_step_kernel():
    from marl_responsible_nav_b200 import BatchedGridWorld
    E = 512
    env = BatchedGridWorld("Level 3", num_envs=E, fear=True, fear_weight=-5.0, auto_reset=True, seed=21)
    ring = ReplayRing(E, env.n_learners, env.obs_len, capacity=6 * E)
    env.reset(obs_out=ring.obs_slot(0))
    hist = []
    gen = torch.Generator(device="cuda").manual_seed(3)
    for t in range(20):
        before = ring.obs_slot(t).clone()
        acts = torch.randint(0, 9, (E, 2), generator=gen, device="cuda", dtype=torch.int8)
        out = env.step(acts, obs_out=ring.obs_slot(t + 1), final_obs_out=ring.final_slot(t), buffers=ring.buffers_slot(t))
        ring.store_action(t, torch.nn.functional.one_hot(acts.long(), 9).float())
        ring.advance()
        nxt = torch.where(out.ended.bool()[:, None, None], ring.final_slot(t), ring.obs_slot(t + 1))
        hist.append((before, acts.clone(), out.shaped_reward.clone(), nxt.clone(), out.terminated.clone()))
        assert torch.equal(out.shaped_reward, (-5.0 * out.fear + out.reward.double()).float())     # maddpg/agent.py:130
    batch = ring.sample(2048, gen)
    assert int(batch["t"].min()) >= 20 - (ring.T - 1)
    for i in range(0, 2048, 7):
        t, e = int(batch["t"][i]), int(batch["env"][i])
        before, acts, rew, nxt, term = hist[t]
        assert torch.equal(batch["state"][i], before[e]) and torch.equal(batch["next_state"][i], nxt[e])
        assert torch.equal(batch["reward"][i], rew[e]) and torch.equal(batch["done"][i], term[e])
        assert int(batch["action"][i, 0].argmax()) == int(acts[e, 0])
    # terminal observations never show the 0.5 spawn marker, fresh ones always do
    ended_rows = batch["ended"]
    if ended_rows.any():
        assert not (batch["next_state"][ended_rows] == 0.5).any()


def _fill_ring(env, ring, steps, seed=3):
    gen = torch.Generator(device="cuda").manual_seed(seed + ring.t)
    if ring.t == 0:
        env.reset(obs_out=ring.obs_slot(0))
    for t in range(ring.t, ring.t + steps):
        acts = torch.randint(0, 9, (env.num_envs, env.n_learners), generator=gen, device="cuda", dtype=torch.int8)
        env.step(acts, obs_out=ring.obs_slot(t + 1), final_obs_out=ring.final_slot(t), buffers=ring.buffers_slot(t))
        ring.store_action(t, torch.rand((env.num_envs, env.n_learners, 9), generator=gen, device="cuda"))
        ring.advance()
    return gen


@pytest.mark.gpu
@pytest.mark.parametrize("obs_dtype", [torch.float32, torch.bfloat16])
def test_fused_sampler_matches_torch_indexing(obs_dtype):
    """gw_replay_sample (one kernel) against ReplayRing.sample's PyTorch indexing on the same (time, env) pairs:
    every field bit for bit, for fp32 and bf16 rings, with a ring that has wrapped around."""
    from marl_responsible_nav_b200 import BatchedGridWorld
    E = 257                                                   # odd: rows of every alignment class
    env = BatchedGridWorld("Level 3", num_envs=E, fear=True, fear_weight=-5.0, auto_reset=True, seed=5, obs_dtype=obs_dtype)
    ring = ReplayRing(E, env.n_learners, env.obs_len, capacity=7 * E, obs_dtype=obs_dtype)
    gen = _fill_ring(env, ring, 23)
    assert ring.t > ring.T                                    # wrapped
    B = 1000
    ref = ring.sample(B, gen)
    assert ref["ended"].any() and not ref["ended"].all()      # both next_state sources are exercised
    got = ring.sample_fused(env, B, indices=(ref["t"], ref["env"]))
    assert torch.equal(got["t"], ref["t"]) and torch.equal(got["env"], ref["env"])
    for k in ("state", "action", "reward", "next_state"):
        assert got[k].dtype == torch.float32
        assert torch.equal(got[k], ref[k].float()), k
    assert torch.equal(got["done"], ref["done"].float())
    env.sync()


@pytest.mark.gpu
def test_fused_sampler_draws_uniform_valid_indices():
    from marl_responsible_nav_b200 import BatchedGridWorld
    E = 64
    env = BatchedGridWorld("Level 3", num_envs=E, fear=False, auto_reset=True, seed=9)
    ring = ReplayRing(E, env.n_learners, env.obs_len, capacity=8 * E)
    empty_err = pytest.raises(RuntimeError, match="empty")
    with empty_err:
        ring.sample_fused(env, 8)
    _fill_ring(env, ring, 5)                                  # 5 of 8 storable time steps so far
    B = 1 << 16
    a = ring.sample_fused(env, B, seed=1)
    t, e = a["t"].clone(), a["env"].clone()
    assert int(t.min()) == 0 and int(t.max()) == 4 and int(e.min()) == 0 and int(e.max()) == E - 1
    ct = torch.bincount(t, minlength=5).double() / B
    ce = torch.bincount(e, minlength=E).double() / B
    assert (ct - 1 / 5).abs().max() < 0.01 and (ce - 1 / E).abs().max() < 0.003      # ~6 sigma
    # gathered rows belong to the drawn indices
    chk = ring.sample_fused(env, B, indices=(t, e), out=ring.new_batch(B))
    assert torch.equal(chk["state"], ring.obs[t % ring.T, e]) and torch.equal(chk["reward"], ring.shaped_reward[t % ring.T, e])
    # the next draw differs, the same (seed, draw number) repeats
    b = ring.sample_fused(env, B, seed=1)
    assert not torch.equal(b["t"], t)
    ring._draws -= 1
    c = ring.sample_fused(env, B, seed=1)
    assert torch.equal(c["t"], b["t"]) and torch.equal(c["env"], b["env"])
    d = ring.sample_fused(env, B, seed=2)
    assert not torch.equal(d["env"], c["env"])
    _fill_ring(env, ring, 30)                                 # wrapped: only the newest T-1 steps may be drawn
    w = ring.sample_fused(env, B, seed=1)
    assert int(w["t"].min()) == ring.t - (ring.T - 1) and int(w["t"].max()) == ring.t - 1
    with pytest.raises(ValueError):
        ring.sample_fused(env, 16, out=ring.new_batch(8))
    env.sync()
