"""GPU (-m gpu): the CUDA library, called through its C-ABI, vs the golden vectors recorded from the reference
and vs the C oracle on seeded device-RNG rollouts at the BASELINE.json batch sizes.

Bar: bit-exact for positions, crash/restricted flags, done flags, integer rewards, observations, masks and the
packed env state; FeAR and the single-env float reward are compared with == as well (north_star allows 1e-6 rel)."""
import numpy as np
import pytest

import c_oracle
import replay_checks as RC

pytestmark = pytest.mark.gpu


def make(**kw):
    return RC.GpuBackend("Level 3", **kw)


def test_update_cases_golden():
    assert RC.check_update_cases(make) > 6000


def test_fear_cases_golden():
    assert RC.check_fear_cases(make) == 500


def test_fear_matrix_and_feal_golden():
    assert RC.check_matrix_cases(make) == 160


def test_walls_and_oneways_golden():
    """f4: restricted paths (walls both ways, one-ways against their direction) enforced by the kernels' next-cell table,
    against what the reference computes when its paths are tuple-typed (wall_cases.npz, wall_ma_episodes.npz) and against
    the C oracle on a device-RNG rollout."""
    mk = lambda **kw: RC.GpuBackend(RC.wall_scenario(), **kw)
    assert RC.check_update_cases(mk, "wall_cases.npz", "u_") == 3000
    assert RC.check_fear_cases(mk, "wall_cases.npz", "f_") == 300
    assert RC.check_ma_episodes(mk, fixture="wall_ma_episodes.npz") > 4000
    _rollout_vs_oracle(2048, 25, seed=12, scenario=RC.wall_scenario(), fear=True, fear_weight=-5.0, auto_reset=True, max_steps=150)


def test_ma_episodes_golden():
    assert RC.check_ma_episodes(make) > 1500


def test_ma_episodes_golden_bf16():
    assert RC.check_ma_episodes(make, obs_bf16=True) > 1500


def test_ma_sessions_autoreset_golden():
    assert RC.check_ma_sessions_autoreset(make) > 1500


def test_single_episodes_golden():
    assert RC.check_single_episodes(make) > 1000


def _rollout_vs_oracle(E, steps, seed, scenario="Level 3", threads=8, masked=False, **kw):
    g = RC.GpuBackend(scenario, num_envs=E, seed=seed, **kw)
    o = c_oracle.COracle(scenario, num_envs=E, seed=seed, threads=threads, **kw)
    g.reset(); o.reset()
    assert np.array_equal(g.positions, o.positions)
    assert np.array_equal(g.obs, o.obs) and np.array_equal(g.action_mask, o.action_mask)
    rng = np.random.default_rng(seed + 1)
    L = o.L
    n_end = 0
    for t in range(steps):
        if masked:      # masked-uniform learner actions: fewer restricted moves, different crash rate
            u = rng.random((E, L, 9)) * o.action_mask
            la = u.argmax(-1).astype(np.int8)
        else:
            la = rng.integers(0, 9, size=(E, L)).astype(np.int8)
        g.step(la); o.step(la)
        for name in ("positions", "reward", "shaped_reward", "fear", "terminated", "truncated", "ended", "info",
                     "action_mask", "obs"):
            a, b = getattr(g, name), getattr(o, name)
            assert np.array_equal(a, b), (name, t, np.flatnonzero((a != b).reshape(E, -1).any(1))[:5])
        ended = o.ended.astype(bool)
        if ended.any() and kw.get("auto_reset", True):
            assert np.array_equal(g.final_obs[ended], o.final_obs[ended]), t
        n_end += int(ended.sum())
    assert np.array_equal(g.state(), o.state())
    sg, so = g.stats(), o.stats()
    for k in ("env_steps", "agent_steps", "episodes", "episode_len_sum", "crashes", "apples", "unresolved", "fear_nonzero"):
        assert sg[k] == so[k], (k, sg[k], so[k])
    assert abs(sg["return_sum"] - so["return_sum"]) < 1e-6
    assert abs(sg["fear_sum"] - so["fear_sum"]) <= 1e-9 * max(1.0, abs(so["fear_sum"]))
    return n_end, so


def test_rollout_config2_fear_4096():
    """BASELINE config[1]: custom_fear_5.yaml, 4096 envs, FeAR on, fp32 MLP obs, device RNG + auto-reset."""
    n_end, st = _rollout_vs_oracle(4096, 40, seed=42, fear=True, fear_weight=-5.0, auto_reset=True, max_steps=150)
    assert n_end > 4096 and st["fear_nonzero"] > 1000


def test_rollout_config3_cnn_16384_bf16():
    """BASELINE config[2]: 16384 envs, obs rendered in the CNN actor's input layout (same bytes), bf16."""
    n_end, _ = _rollout_vs_oracle(16384, 12, seed=66, fear=False, obs_bf16=True, auto_reset=True, max_steps=150)
    assert n_end > 4096


def test_rollout_masked_actions_and_time_limit():
    n_end, st = _rollout_vs_oracle(2048, 60, seed=0, fear=True, masked=True, auto_reset=True, max_steps=20)
    assert st["episodes"] > 2048


def test_rollout_level5_three_agents():
    _rollout_vs_oracle(2048, 25, seed=3, scenario="Level 5", fear=True, auto_reset=True, max_steps=150)


def test_rollout_gamemap_four_agents():
    _rollout_vs_oracle(2048, 25, seed=4, scenario="GameMap", n_agents=4, fear=True, auto_reset=True, max_steps=150)


@pytest.mark.parametrize("scenario", ["Level 3", "Level 5", "GameMap"])
@pytest.mark.parametrize("n_agents", [2, 3, 4])
def test_rollout_every_layout_and_agent_count(scenario, n_agents):
    """BASELINE config[3]: "all Scenarios.json layouts, 4 agents" -- every shipped layout with 2, 3 and 4 world agents
    (2 learners + NPCs), FeAR on, against the C oracle."""
    n_end, st = _rollout_vs_oracle(1024, 20, seed=30 + n_agents, scenario=scenario, n_agents=n_agents, fear=True, fear_weight=-5.0,
                                   auto_reset=True, max_steps=150)
    assert n_end > 0


def test_rollout_oracle_diff_at_1m_envs():
    """Config[3] size: the full 1 048 576-env batch against the C oracle (all host threads), two steps after the reset,
    every output compared with == (1.3 GB of observations per step)."""
    n_end, st = _rollout_vs_oracle(1 << 20, 2, seed=77, fear=True, fear_weight=-5.0, auto_reset=True, max_steps=150, threads=32)
    assert st["env_steps"] == 2 << 20


def test_rollout_single_env_kind():
    _rollout_vs_oracle(4096, 40, seed=5, env_kind="single", fear=True, auto_reset=True, max_steps=150)


def test_rollout_no_autoreset_sticky_flags():
    _rollout_vs_oracle(1024, 30, seed=6, fear=False, auto_reset=False, max_steps=0)


def test_rollout_every_kernel_variant():
    """gw_step picks its kernel by batch size (8-lanes-per-env kernel, thread-per-env with 32 / 128 / 256-env tiles,
    fixed-stride or dynamic tiles): every variant against the oracle, FeAR on, odd sizes for the tile tails."""
    _rollout_vs_oracle(6144, 10, seed=21, fear=True, fear_weight=-5.0)          # small-batch kernel, largest size
    _rollout_vs_oracle(6145 + 2000, 10, seed=22, fear=True, fear_weight=-5.0)   # thread per env, 32-env tiles
    _rollout_vs_oracle(40001, 6, seed=23, fear=True, fear_weight=-5.0, threads=16)      # 128-env tiles
    _rollout_vs_oracle(200003, 3, seed=24, fear=True, obs_bf16=True, threads=16)        # 256-env tiles, fixed stride
    _rollout_vs_oracle(4 * 592 * 256 + 77, 2, seed=25, fear=True, obs_bf16=True, threads=16)   # dynamic tiles


def test_large_batch_properties_1m():
    """Config[3] size (1M envs): size-independent properties instead of an oracle diff --
    agents stay on active cells and distinct, crashed agents do not move, masks match positions,
    sharding invariance: the same global env ids on two handles give the same trajectories."""
    import torch
    from marl_responsible_nav_b200 import BatchedGridWorld, builtin_scenario
    E = 1 << 20
    sc = builtin_scenario("Level 3")
    region = torch.as_tensor(sc.region, device="cuda").bool()
    env = BatchedGridWorld("Level 3", num_envs=E, fear=True, auto_reset=True, seed=9)
    half = BatchedGridWorld("Level 3", num_envs=E // 2, fear=True, auto_reset=True, seed=9, env_id_base=E // 2)
    out = env.reset(); outh = half.reset()
    assert torch.equal(out.positions[E // 2:], outh.positions)
    gen = torch.Generator(device="cuda").manual_seed(1)
    fresh = torch.zeros(E, dtype=torch.bool, device="cuda")    # env respawned after `positions` was written
    for t in range(5):
        prev = out.positions.clone().long()
        la = torch.randint(0, 9, (E, 2), generator=gen, device="cuda", dtype=torch.int8)
        out = env.step(la); outh = half.step(la[E // 2:].contiguous())
        pos = out.positions.long()
        assert region[pos[..., 0], pos[..., 1]].all()
        flat = pos[..., 0] * 16 + pos[..., 1]
        srt = flat.sort(dim=1).values
        assert (srt[:, 1:] != srt[:, :-1]).all()
        crash = torch.stack([(out.info >> i) & 1 for i in range(4)], 1).bool() & ~fresh[:, None]
        assert torch.equal(pos[crash], prev[crash])
        fresh = out.ended.bool()
        assert torch.equal(out.positions[E // 2:], outh.positions)
        assert torch.equal(out.fear[E // 2:], outh.fear) and torch.equal(out.obs[E // 2:], outh.obs)
        assert ((out.info >> 13) & 1).sum().item() == 0          # collision fix-point always resolves


def test_state_checkpoint_roundtrip():
    import torch
    from marl_responsible_nav_b200 import BatchedGridWorld
    a = BatchedGridWorld("Level 3", num_envs=512, fear=True, seed=11)
    a.reset()
    la = torch.randint(0, 9, (512, 2), device="cuda", dtype=torch.int8)
    for _ in range(5):
        a.step(la)
    snap = a.state_dict()
    b = BatchedGridWorld("Level 3", num_envs=512, fear=True, seed=11)
    b.load_state_dict(snap)
    for _ in range(5):
        oa, ob = a.step(la), b.step(la)
        assert torch.equal(oa.obs, ob.obs) and torch.equal(oa.positions, ob.positions) and torch.equal(oa.fear, ob.fear)


def test_compat_envs_match_reference_trace_under_seeds():
    """CustomMAEnv / CustomEnv with the reference's own seeding (ctor seed + random.seed + np.random.seed):
    identical trajectories to the traces recorded from the reference."""
    import random
    from marl_responsible_nav_b200 import CustomMAEnv, CustomEnv
    g = RC.npz("ma_episodes.npz")
    for sess in np.unique(g["ep_session"]):
        eps = np.flatnonzero(g["ep_session"] == sess)
        seed, fear = int(g["ep_seed"][eps[0]]), bool(g["ep_fear"][eps[0]])
        env = CustomMAEnv(render=False, fear=fear, seed=seed)
        random.seed(seed); np.random.seed(seed)
        for ep in eps[:6]:
            obs, info = env.reset()
            assert np.array_equal(np.stack([obs[a] for a in env.possible_agents]).astype(np.float32), g["ep_reset_obs"][ep])
            assert obs["agent_0"].dtype == np.float64 and obs["agent_0"].shape == (10, 16)
            assert info["fear"] == 0.0
            for t in range(int(g["ep_n_steps"][ep])):
                s = int(g["ep_first_step"][ep]) + t
                obs, rew, term, trunc, info = env.step(tuple(int(a) for a in g["learner_act"][s]))
                assert [rew[a] for a in env.possible_agents] == list(g["reward"][s])
                assert [term[a] for a in env.possible_agents] == list(g["term"][s])
                assert [trunc[a] for a in env.possible_agents] == list(g["trunc"][s])
                assert [float(info["fear"][a]) for a in env.possible_agents] == list(g["fear"][s])
                assert info["agent_crashes"] == g["crash_count"][s] and info["apples_caught"] == g["apples_caught"][s]
                assert np.array_equal(np.stack([obs[a] for a in env.possible_agents]).astype(np.float32), g["obs"][s])
                assert np.array_equal(np.stack([info[a]["action_mask"] for a in env.possible_agents]), g["mask"][s])
        assert env.step(()) == ({}, {}, {}, {}, {})
        assert env.num_agents == 2 and env.action_space.n == 9
    gs = RC.npz("single_episodes.npz")
    for sess in np.unique(gs["ep_session"]):
        eps = np.flatnonzero(gs["ep_session"] == sess)
        seed, fear = int(gs["ep_seed"][eps[0]]), bool(gs["ep_fear"][eps[0]])
        CustomEnv.rng = np.random.default_rng(seed)
        env = CustomEnv(render=False, fear=fear)
        random.seed(seed); np.random.seed(seed)
        for ep in eps[:6]:
            obs, _ = env.reset()
            assert np.array_equal(obs.astype(np.float32), gs["ep_reset_obs"][ep])
            for t in range(int(gs["ep_n_steps"][ep])):
                s = int(gs["ep_first_step"][ep]) + t
                obs, rew, term, trunc, info = env.step([int(gs["action"][s])])
                assert rew[0] == gs["reward"][s] and term[0] == gs["term"][s] and trunc == gs["trunc"][s]
                assert float(info["fear"]) == gs["fear"][s] and info["restricted"] == gs["restricted"][s]
                assert info["episode"]["r"] == gs["ep_r"][s] and info["episode"]["l"] == gs["ep_l"][s]
                assert np.array_equal(obs.astype(np.float32), gs["obs"][s])


def test_errors_are_loud():
    import torch
    from marl_responsible_nav_b200 import BatchedGridWorld
    env = BatchedGridWorld("Level 3", num_envs=8)
    with pytest.raises(RuntimeError, match="gw_reset first"):
        env.step(torch.zeros((8, 2), dtype=torch.int8, device="cuda"))
    env.reset()
    with pytest.raises(ValueError):
        env.step(torch.zeros((7, 2), dtype=torch.int8, device="cuda"))
    with pytest.raises(RuntimeError, match="n_agents"):                  # the packed layout holds up to 4 agents ...
        BatchedGridWorld("Level 3", num_envs=8, n_agents=7, layout="packed")
    assert type(BatchedGridWorld("Level 3", num_envs=8, n_agents=7)).__name__ == "GeneralGridWorld"   # ... seven take the general one ...
    with pytest.raises(RuntimeError, match="n_agents"):                  # ... which holds up to 16
        BatchedGridWorld("Level 3", num_envs=8, n_agents=17)


def test_step_host_matches_device_step():
    """gw_step_host (pinned host actions in, rewards / ended out, synchronised) == gw_step on the same inputs."""
    import torch
    from marl_responsible_nav_b200 import BatchedGridWorld
    E = 2048
    a = BatchedGridWorld("Level 3", num_envs=E, fear=True, fear_weight=-5.0, seed=13)
    b = BatchedGridWorld("Level 3", num_envs=E, fear=True, fear_weight=-5.0, seed=13)
    a.reset(); b.reset()
    acts = torch.randint(0, 9, (6, E, 2), dtype=torch.int8).pin_memory()
    rew = torch.empty((E, 2), dtype=torch.float32).pin_memory()
    shp = torch.empty((E, 2), dtype=torch.float32).pin_memory()
    end = torch.empty((E,), dtype=torch.uint8).pin_memory()
    for t in range(6):                                   # default: the kernel reads / writes the pinned buffers itself
        oa = a.step_host(acts[t], rew, end, host_shaped=shp)
        ob = b.step(acts[t].cuda())
        assert torch.equal(rew, ob.reward.cpu()) and torch.equal(end, ob.ended.cpu()) and torch.equal(shp, ob.shaped_reward.cpu())
        assert torch.equal(oa.obs, ob.obs) and torch.equal(oa.obs_code, ob.obs_code)
        assert oa.reward is rew and oa.ended is end and oa.shaped_reward is shp
    for t in range(3):                                   # memcpy variant: same results, device-side copies as well
        oa = a.step_host(acts[t], rew, end, host_shaped=shp, zero_copy=False)
        ob = b.step(acts[t].cuda())
        assert torch.equal(rew, ob.reward.cpu()) and torch.equal(end, ob.ended.cpu()) and torch.equal(shp, ob.shaped_reward.cpu())
        assert torch.equal(oa.obs, ob.obs) and torch.equal(oa.reward, ob.reward)
    ring = torch.empty((2,) + tuple(ob.obs.shape), dtype=ob.obs.dtype, device="cuda")
    for t in range(4):                                   # cached marshalling: alternating buffers still hit the right ones
        oa = a.step_host(acts[t], rew, end, obs_out=ring[t % 2])
        ob = b.step(acts[t].cuda())
        assert torch.equal(rew, ob.reward.cpu()) and torch.equal(ring[t % 2], ob.obs) and oa.obs.data_ptr() == ring[t % 2].data_ptr()
    st = a.stats()
    assert st["fear_tasks"] > 0 and st["fear_tasks"] == b.stats()["fear_tasks"]
    with pytest.raises(ValueError):
        a.step_host(acts[0].clone(), rew, end)          # not pinned


@pytest.mark.parametrize("E,fear,bf16", [(4096, True, False), (1001, True, True), (6144, False, False), (77, True, False),
                                         (12000, True, False)])      # 12000 envs: 375 tiles on 296 CTAs, some walk two tiles
def test_step_host_resident_kernel(E, fear, bf16):
    """gw_step_host mode 2 (resident kernel: doorbell / completion word in pinned host memory) == gw_step, step after
    step and in the final state; survives idle exits, relaunches, and other calls in between."""
    import time
    import torch
    from marl_responsible_nav_b200 import BatchedGridWorld
    kw = dict(num_envs=E, fear=fear, fear_weight=-5.0, seed=21, obs_dtype=torch.bfloat16 if bf16 else torch.float32)
    a = BatchedGridWorld("Level 3", **kw)
    b = BatchedGridWorld("Level 3", **kw)
    a.reset(); b.reset()
    T = 96
    acts = torch.randint(0, 9, (T, E, 2), dtype=torch.int8).pin_memory()
    dacts = acts.cuda()
    rew = [torch.empty((E, 2), dtype=torch.float32).pin_memory() for _ in range(2)]
    shp = torch.empty((E, 2), dtype=torch.float32).pin_memory()
    end = torch.empty((E,), dtype=torch.uint8).pin_memory()
    ring = torch.empty((4,) + tuple(a.buf.obs.shape), dtype=a.buf.obs.dtype, device="cuda")
    ref_obs = torch.empty_like(ring)
    exp = []
    for t in range(T):                                   # the expected results first: nothing else touches the GPU below
        ob = b.step(dacts[t])
        exp.append((ob.reward.cpu(), ob.shaped_reward.cpu(), ob.ended.cpu()))
        if t >= T - 4:
            ref_obs[t % 4].copy_(ob.obs)
    torch.cuda.synchronize()
    for t in range(T):
        oa = a.step_host(acts[t], rew[t % 2], end, host_shaped=shp, obs_out=ring[t % 4], resident=True)
        assert oa.reward is rew[t % 2]
        assert torch.equal(rew[t % 2], exp[t][0]) and torch.equal(shp, exp[t][1]) and torch.equal(end, exp[t][2]), t
        if t == 40:
            assert a.server_info()["running"] == 1
            time.sleep(0.02)                             # > idle time: the kernel leaves by itself, the next call relaunches it
            assert a.server_info()["running"] == 0
        if t == 60:
            a.sync()                                     # any other call stops it
            assert a.server_info()["running"] == 0
    info = a.server_info()
    assert 3 <= info["launches"] <= 6 and info["buffer_sets"] == T
    a.sync()
    assert torch.equal(ring, ref_obs)
    assert torch.equal(a.state_dict(), b.state_dict())
    sa, sb = a.stats(), b.stats()
    for k in sa:                                         # the two float sums are atomics: same terms, order may differ
        assert sa[k] == sb[k] if k not in ("fear_sum", "return_sum") else abs(sa[k] - sb[k]) <= 1e-9 * max(1.0, abs(sb[k])), k
    # the same pinned buffers rewritten by the host before every step: the kernel must see the new bytes, not a cached copy
    one = torch.empty((E, 2), dtype=torch.int8).pin_memory()
    for t in range(24):
        one.copy_(acts[T - 1 - t])
        a.step_host(one, rew[0], end, resident=True)
        ob = b.step(dacts[T - 1 - t])
        assert torch.equal(rew[0], ob.reward.cpu()) and torch.equal(end, ob.ended.cpu()), t
    # device steps and resident steps interleave on the same handle
    for t in range(8):
        if t % 2:
            oa = a.step_host(acts[t], rew[0], end, resident=True)
            r_a = rew[0].clone()
        else:
            r_a = a.step(dacts[t]).reward.cpu()
        assert torch.equal(r_a, b.step(dacts[t]).reward.cpu())
    assert torch.equal(a.state_dict(), b.state_dict())
    a.close(); b.close()


def test_step_host_resident_falls_back_when_launches_block():
    """Under CUDA_LAUNCH_BLOCKING=1 (as under a profiler) a launch only returns when the kernel has ended, so a resident
    kernel can never be rung: mode 2 must notice, step through the ordinary launch, and give the same results."""
    import subprocess, sys, os
    code = r"""
import torch
from marl_responsible_nav_b200 import BatchedGridWorld
E = 512
a = BatchedGridWorld("Level 3", num_envs=E, fear=True, fear_weight=-5.0, seed=5); b = BatchedGridWorld("Level 3", num_envs=E, fear=True, fear_weight=-5.0, seed=5)
a.reset(); b.reset()
acts = torch.randint(0, 9, (6, E, 2), dtype=torch.int8).pin_memory()
rew = torch.empty((E, 2), dtype=torch.float32).pin_memory(); end = torch.empty((E,), dtype=torch.uint8).pin_memory()
for t in range(6):
    a.step_host(acts[t], rew, end, resident=True)
    ob = b.step(acts[t].cuda())
    assert torch.equal(rew, ob.reward.cpu()) and torch.equal(end, ob.ended.cpu()), t
assert a.server_info()["running"] == -1, a.server_info()
assert torch.equal(a.state_dict(), b.state_dict())
print("OK")
"""
    env = dict(os.environ, CUDA_LAUNCH_BLOCKING="1")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "OK" in res.stdout, res.stdout + res.stderr


def test_pinned_io_block_matches_tensor_api():
    """PinnedIO (every input and output in one pinned host block, stepped by the resident kernel; what the E = 1 drop-in
    classes use) against the tensor API on the same seed, device-RNG mode, 96 envs."""
    import numpy as np
    import torch
    from marl_responsible_nav_b200 import BatchedGridWorld
    E = 96
    a = BatchedGridWorld("Level 3", num_envs=E, fear=True, fear_weight=-5.0, seed=31)
    b = BatchedGridWorld("Level 3", num_envs=E, fear=True, fear_weight=-5.0, seed=31)
    io = a.pinned_io()
    io.reset()
    ob = b.reset()
    assert np.array_equal(io.obs, ob.obs.cpu().numpy()) and np.array_equal(io.positions, ob.positions.cpu().numpy())
    assert np.array_equal(io.action_mask, ob.action_mask.cpu().numpy())
    rng = np.random.default_rng(3)
    for t in range(40):
        la = rng.integers(0, 9, size=(E, 2)).astype(np.int8)
        io.actions[:] = la
        io.step()
        ob = b.step(la)
        for name in ("obs", "reward", "shaped_reward", "fear", "terminated", "truncated", "ended", "action_mask", "positions"):
            assert np.array_equal(getattr(io, name), getattr(ob, name).cpu().numpy()), (t, name)
        assert np.array_equal(io.info, ob.info.cpu().numpy().view(np.uint32)), t
        assert np.array_equal(io.obs_code, ob.obs_code.cpu().numpy().view(np.uint64)), t
    assert torch.equal(a.state_dict(), b.state_dict())
