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
