"""GPU (-m gpu): gw_rollout -- T env steps per launch -- against the C oracle stepped T times on the same action stream,
bit for bit, for every output ring, and against the same steps taken one gw_step at a time."""
import numpy as np
import pytest

import c_oracle

pytestmark = pytest.mark.gpu

FIELDS = ("positions", "reward", "shaped_reward", "fear", "terminated", "truncated", "ended", "info")


def _rollout_vs_oracle(E, T, launches, seed, scenario="Level 3", slots=None, n_act=None, obs_bf16=False, **kw):
    import torch
    from marl_responsible_nav_b200 import BatchedGridWorld
    slots = slots or (T + 1)
    n_act = n_act or T
    env = BatchedGridWorld(scenario, num_envs=E, seed=seed, obs_dtype=torch.bfloat16 if obs_bf16 else torch.float32, **kw)
    ora = c_oracle.COracle(scenario, num_envs=E, seed=seed, threads=8, obs_bf16=obs_bf16, **kw)
    rings = env.new_rings(slots)
    env.reset(obs_out=rings.obs[0])
    ora.reset()
    L = ora.L
    rng = np.random.default_rng(seed + 7)
    t_abs = 0
    n_end = 0
    for _ in range(launches):
        acts = rng.integers(0, 9, size=(n_act, E, L)).astype(np.int8)
        first_action = int(rng.integers(0, n_act))
        env.rollout(torch.from_numpy(acts).cuda(), T, rings, first_slot=t_abs % slots, first_action=first_action)
        env.sync()
        got = {k: getattr(rings, k).cpu().numpy() for k in FIELDS + ("obs_code", "action_mask")}
        raw = (lambda t: t.view(torch.int16).cpu().numpy().view(np.uint16)) if obs_bf16 else (lambda t: t.cpu().numpy())
        obs, fin = raw(rings.obs), raw(rings.final_obs)
        # a ring shorter than the launch keeps the newest `slots` transitions only
        for k in range(T):
            ora.step(acts[(first_action + k) % n_act])
            st, so = (t_abs + k) % slots, (t_abs + k + 1) % slots
            ended = ora.ended.astype(bool)
            n_end += int(ended.sum())
            if k < T - slots + 1 and T >= slots:      # overwritten later in this launch
                continue
            for name in FIELDS:
                a, b = got[name][st], getattr(ora, name)
                assert np.array_equal(a, b), (name, k, np.flatnonzero((a != b).reshape(E, -1).any(1))[:5])
            assert np.array_equal(got["action_mask"][so], ora.action_mask), k
            assert np.array_equal(obs[so].reshape(ora.obs.shape), ora.obs), k
            if ended.any() and kw.get("auto_reset", True):
                assert np.array_equal(fin[st][ended].reshape(-1), ora.final_obs[ended].reshape(-1)), k
        t_abs += T
    assert np.array_equal(env.state_dict().numpy().view(np.uint32).reshape(E, 4), ora.state())
    sg, so_ = env.stats(), ora.stats()
    for k in ("env_steps", "agent_steps", "episodes", "episode_len_sum", "crashes", "apples", "fear_nonzero"):
        assert sg[k] == so_[k], (k, sg[k], so_[k])
    assert abs(sg["fear_sum"] - so_["fear_sum"]) <= 1e-9 * max(1.0, abs(so_["fear_sum"]))
    return n_end


def test_rollout_kernel_config2_4096():
    """BASELINE config[1] through gw_rollout: 4096 envs, FeAR on, 3 launches of 16 steps."""
    n_end = _rollout_vs_oracle(4096, 16, 3, seed=42, fear=True, fear_weight=-5.0, auto_reset=True, max_steps=150)
    assert n_end > 4096


def test_rollout_kernel_ring_wrap_and_odd_sizes():
    """Ring shorter than the launch count, action ring shorter than the launch, tile tail (E % 32 != 0)."""
    _rollout_vs_oracle(1002, 7, 4, seed=5, slots=5, n_act=3, fear=True, fear_weight=-10.0, auto_reset=True, max_steps=20)


def test_rollout_kernel_bf16_no_fear_large():
    _rollout_vs_oracle(16384, 6, 2, seed=66, obs_bf16=True, fear=False, auto_reset=True, max_steps=150)


def test_rollout_kernel_single_kind_and_level5():
    _rollout_vs_oracle(2048, 9, 2, seed=3, scenario="Level 5", fear=True, auto_reset=True, max_steps=150)
    _rollout_vs_oracle(2048, 9, 2, seed=4, env_kind="single", fear=True, auto_reset=True, max_steps=150)


def test_rollout_equals_step_by_step_and_interleaves():
    """gw_rollout and gw_step interleave freely on one handle: the random-word cache and the packed state are handed
    over both ways."""
    import torch
    from marl_responsible_nav_b200 import BatchedGridWorld
    E, T = 2048, 5
    kw = dict(num_envs=E, fear=True, fear_weight=-5.0, auto_reset=True, max_steps=30, seed=11)
    a, b = BatchedGridWorld("Level 3", **kw), BatchedGridWorld("Level 3", **kw)
    rings = a.new_rings(T + 1)
    a.reset(obs_out=rings.obs[0]); b.reset()
    gen = torch.Generator(device="cuda").manual_seed(3)
    for rnd in range(4):
        acts = torch.randint(0, 9, (T, E, 2), generator=gen, device="cuda", dtype=torch.int8)
        a.rollout(acts, T, rings, first_slot=0)
        for k in range(T):
            out = b.step(acts[k])
            assert torch.equal(rings.reward[k], out.reward) and torch.equal(rings.fear[k], out.fear), (rnd, k)
            assert torch.equal(rings.obs[k + 1].view_as(out.obs), out.obs), (rnd, k)
            assert torch.equal(rings.obs_code[k + 1], out.obs_code), (rnd, k)
        one = torch.randint(0, 9, (E, 2), generator=gen, device="cuda", dtype=torch.int8)
        oa, ob = a.step(one), b.step(one)                      # a plain step in between
        assert torch.equal(oa.obs, ob.obs) and torch.equal(oa.reward, ob.reward)
        a.reset(obs_out=rings.obs[0], mask=(oa.ended * 0 + (rnd % 2)).to(torch.uint8)) if rnd % 2 else None
        b.reset(mask=(ob.ended * 0 + 1).to(torch.uint8)) if rnd % 2 else None
    assert torch.equal(a.state_dict(), b.state_dict())
