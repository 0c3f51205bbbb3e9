"""The GENERAL state layout (gww_*: grids other than 10 x 16, more than 4 agents).

Parity is pinned on fixtures recorded by RUNNING THE REFERENCE on a 20 x 28 map with 7 agents, walls and one-ways
(tests/golden/make_wide_golden.py; the reference's GWorld / Responsibility / env classes are data-driven).  CPU tests: the
Python restatement and the C oracle built on gww_config reproduce them.  GPU tests (-m gpu): the CUDA kernels of
csrc/gw_wide.cu, through the C-ABI, reproduce them too, equal the C oracle on seeded device-RNG rollouts (20 x 28 x 7 and
the layout's maximum, 64 x 64 x 16), and equal the PACKED layout's kernels bit for bit on Level 3."""
import numpy as np
import pytest

import c_oracle
import gridworld_oracle as PO
import replay_checks as RC

WIDE = RC.wide_scenario()


def make_c(**kw):
    return c_oracle.COracle(WIDE, **kw)


# ----------------------------------------------------------------------------- host side
def test_wide_scenario_is_loaded_from_the_reference_json_format():
    g = RC.npz("wide_cases.npz")
    assert WIDE.shape == (20, 28) and WIDE.n_agents == 7
    want = [((int(p[0][0]), int(p[0][1])), (int(p[1][0]), int(p[1][1]))) for p in g["restricted_paths"]]
    assert WIDE.blocked == want and len(want) == 110         # = GWorld.RestrictedPaths built by the reference from the same lists
    from marl_responsible_nav_b200 import _native as N
    assert not N.fits_packed_layout(WIDE)
    cfg = N.build_wide_config(WIDE, num_envs=3)
    assert (cfg.height, cfg.width, cfg.n_agents, cfg.n_blocked) == (20, 28, 7, 110)
    assert cfg.policy_map[5 * 28 + 2] == WIDE.policy_index[5, 2] and cfg.mdr_map[9 * 28 + 3] == 4
    with pytest.raises(ValueError):
        from marl_responsible_nav_b200.scenarios import Scenario
        N.build_wide_config(Scenario("big", np.ones((65, 8), np.int8), 2, np.zeros((65, 8), np.uint8), [([1, 1, 1], [1, 1, 1, 1])],
                                     np.zeros((65, 8), np.uint8)))


def _po_scenario():
    return PO.Scenario(region=WIDE.region, n_agents=7, policy_map=WIDE.policy_index.astype(np.int64),
                       mdr_map=WIDE.mdr_action.astype(np.int64), policies={}, name="wide", blocked=list(WIDE.blocked))


def test_python_oracle_wide_cases():
    g = RC.npz("wide_cases.npz")
    blocked = set(WIDE.blocked)
    for c in range(len(g["u_n"])):
        n = int(g["u_n"][c])
        locs = [tuple(int(v) for v in g["u_locs"][c, i]) for i in range(n)]
        acts = [int(a) for a in g["u_acts"][c, :n]]
        apples = {k: tuple(int(v) for v in g["u_apples"][c, k]) for k in range(2) if g["u_apples"][c, k, 0] >= 0}
        kw = dict(apples=apples, eaters=list(range(min(2, n)))) if apples else {}
        new_locs, crash, restr, caught, _ = PO.update_world(WIDE.region, locs, acts, blocked=blocked, **kw)
        assert new_locs == [tuple(int(v) for v in g["u_out_locs"][c, i]) for i in range(n)], c
        assert crash == list(g["u_crash"][c, :n]) and restr == list(g["u_restr"][c, :n]), c
    for c in range(0, len(g["f_n"]), 8):
        n = int(g["f_n"][c])
        locs = [tuple(int(v) for v in g["f_locs"][c, i]) for i in range(n)]
        lst = [(i, int(g["f_acts"][c, i])) for i in range(n) if g["f_in_list"][c, i]]
        actor = int(g["f_actor"][c])
        resp, n_mdr, n_act = PO.fear_one_actor(WIDE.region, locs, lst, [int(m) for m in g["f_mdr"][c, :n]], actor, blocked)
        assert np.array_equal(resp[actor], g["f_resp"][c, :n]) and float(np.sum(resp)) == float(g["f_fear_sum"][c]), c
        assert np.array_equal(n_mdr[actor], g["f_n_mdr"][c, :n]) and np.array_equal(n_act[actor], g["f_n_act"][c, :n]), c


def test_python_oracle_wide_episodes():
    """CustomMAEnv / CustomEnv of the reference on the 20 x 28 x 7 scenario (FeAR-off sessions in full, a FeAR session each)."""
    sc = _po_scenario()
    g = RC.npz("wide_ma_episodes.npz")
    fear_eps = np.flatnonzero(g["ep_fear"])[:3]
    for e in list(np.flatnonzero(~g["ep_fear"])) + list(fear_eps):
        env = PO.MAEnvOracle(sc, fear=bool(g["ep_fear"][e]))
        obs, masks = env.reset([tuple(int(v) for v in c) for c in g["ep_spawn"][e]])
        assert np.array_equal(np.array(obs, np.float32), g["ep_reset_obs"][e])
        s0 = int(g["ep_first_step"][e])
        for t in range(int(g["ep_n_steps"][e])):
            s = s0 + t
            assert env.mdr_of_agents() == [int(m) for m in g["mdr"][s]], (e, t)
            r = env.step(g["learner_act"][s], g["all_act"][s])
            assert r.locs == [tuple(int(v) for v in c) for c in g["locs"][s]], (e, t)
            assert r.rewards == list(g["reward"][s]) and r.fear == list(g["fear"][s]), (e, t)
            assert r.terminations == list(g["term"][s]) and r.truncations == list(g["trunc"][s]), (e, t)
            assert np.array_equal(np.array(r.obs, np.float32), g["obs"][s]), (e, t)
            assert np.array_equal(np.array(r.masks), g["mask"][s]), (e, t)
    g = RC.npz("wide_single_episodes.npz")
    for e in list(np.flatnonzero(~g["ep_fear"])) + list(np.flatnonzero(g["ep_fear"])[:3]):
        env = PO.SingleEnvOracle(sc, fear=bool(g["ep_fear"][e]))
        obs = env.reset([tuple(int(v) for v in c) for c in g["ep_spawn"][e]])
        assert np.array_equal(obs.astype(np.float32), g["ep_reset_obs"][e])
        s0 = int(g["ep_first_step"][e])
        for t in range(int(g["ep_n_steps"][e])):
            s = s0 + t
            r = env.step(int(g["action"][s]), g["all_act"][s])
            assert r.locs == [tuple(int(v) for v in c) for c in g["locs"][s]], (e, t)
            assert r.reward == g["reward"][s] and r.fear == g["fear"][s], (e, t)
            assert r.terminated == g["term"][s] and r.truncated == g["trunc"][s], (e, t)
            assert np.array_equal(r.obs.astype(np.float32), g["obs"][s]), (e, t)


def _check_all(make):
    assert RC.check_update_cases(make, "wide_cases.npz", "u_") == 2500          # 1..16 agents
    assert RC.check_fear_cases(make, "wide_cases.npz", "f_") == 160             # 5..12 agents, np.sum in numpy's pairwise order
    assert RC.check_matrix_cases(make, "wide_cases.npz", "m_") == 50
    assert RC.check_ma_episodes(make, fixture="wide_ma_episodes.npz") > 1300
    assert RC.check_ma_episodes(make, fixture="wide_ma_episodes.npz", obs_bf16=True) > 1300
    assert RC.check_single_episodes(make, fixture="wide_single_episodes.npz") > 900
    assert RC.check_ma_sessions_autoreset(make, fixture="wide_ma_episodes.npz", max_steps=60) > 1300


def test_c_oracle_wide_golden():
    _check_all(make_c)


def test_c_oracle_wide_equals_packed_on_level3():
    """The two builds of oracle/gw_oracle.c agree where both apply: Level 3 through gww_config = Level 3 through gw_config
    (device-RNG mode: same spawns, same NPC draws)."""
    a = c_oracle.COracle("Level 3", num_envs=257, seed=5, fear_weight=-5.0, layout="packed")
    b = c_oracle.COracle("Level 3", num_envs=257, seed=5, fear_weight=-5.0, layout="wide")
    a.reset(); b.reset()
    rng = np.random.default_rng(0)
    for t in range(40):
        act = rng.integers(0, 9, size=(257, 2)).astype(np.int8)
        a.step(act); b.step(act)
        for name in ("obs", "reward", "shaped_reward", "fear", "terminated", "truncated", "ended", "action_mask", "positions"):
            assert np.array_equal(getattr(a, name), getattr(b, name)), (t, name)
        assert np.array_equal(a.info & 0xFFFF, b.info & 0xFFFF), t
    assert a.stats() == b.stats()


# ----------------------------------------------------------------------------- GPU
def make_gpu(**kw):
    return RC.GpuBackend(WIDE, **kw)


def _compare_rollout(gpu, ora, steps, seed, names=("obs", "final_obs", "reward", "shaped_reward", "fear", "terminated", "truncated",
                                                   "ended", "action_mask", "positions", "info")):
    rng = np.random.default_rng(seed)
    E, L = ora.E, ora.L
    gpu.reset(); ora.reset()
    assert np.array_equal(gpu.obs, ora.obs) and np.array_equal(gpu.positions, ora.positions) and np.array_equal(gpu.action_mask, ora.action_mask)
    ended_total = 0
    for t in range(steps):
        act = rng.integers(0, 9, size=(E, L)).astype(np.int8)
        gpu.step(act); ora.step(act)
        for name in names:
            a, b = getattr(gpu, name), getattr(ora, name)
            if name == "final_obs":
                m = ora.ended.astype(bool)
                a, b = a[m], b[m]
            assert np.array_equal(a, b), (t, name)
        ended_total += int(ora.ended.sum())
    assert np.array_equal(gpu.state(), ora.state())
    gs, os_ = gpu.stats(), ora.stats()
    for k in ("env_steps", "agent_steps", "episodes", "episode_len_sum", "crashes", "apples", "unresolved", "fear_nonzero"):
        assert gs[k] == os_[k], k
    for k in ("return_sum", "fear_sum"):         # sums of doubles in different orders (the single env's returns are tenths)
        assert abs(gs[k] - os_[k]) <= 1e-9 * max(1.0, abs(os_[k])), k
    return ended_total


@pytest.mark.gpu
def test_general_layout_is_selected_when_the_scenario_does_not_fit():
    from marl_responsible_nav_b200 import BatchedGridWorld, GeneralGridWorld
    assert type(BatchedGridWorld("Level 3", num_envs=4)) is BatchedGridWorld
    w = BatchedGridWorld(WIDE, num_envs=4)
    assert type(w) is GeneralGridWorld and w.obs_len == 560 and w.n_agents == 7
    assert type(BatchedGridWorld("Level 3", num_envs=4, layout="general")) is GeneralGridWorld
    with pytest.raises(NotImplementedError):
        w.step_host(None, None)
    with pytest.raises(RuntimeError):                         # 5 agents do not fit the packed layout
        BatchedGridWorld("Level 3", num_envs=4, n_agents=5, layout="packed")


@pytest.mark.gpu
def test_general_gpu_reference_golden():
    """Everything the reference recorded on the 20 x 28 x 7 scenario, through the C-ABI on the GPU."""
    _check_all(make_gpu)


@pytest.mark.gpu
@pytest.mark.parametrize("env_kind,fear,bf16", [("multi", True, False), ("multi", False, True), ("single", True, False)])
def test_general_gpu_rollout_equals_c_oracle(env_kind, fear, bf16):
    kw = dict(num_envs=1001, env_kind=env_kind, fear=fear, fear_weight=-5.0, seed=11, max_steps=40, obs_bf16=bf16)
    ended = _compare_rollout(make_gpu(**kw), make_c(threads=8, **kw), steps=60, seed=3)
    assert ended > 500


@pytest.mark.gpu
def test_general_gpu_maximum_size_64x64_16_agents():
    """The layout's limits: a 64 x 64 map (random obstacles, walls), 16 agents, apples moved into the far corner region."""
    from marl_responsible_nav_b200.scenarios import Scenario, restricted_paths
    rng = np.random.default_rng(9)
    region = (rng.random((64, 64)) < 0.8).astype(np.int8)
    region[60:64, 58:64] = 1
    region[0, :] = 1
    pol = (rng.integers(0, 3, size=(64, 64))).astype(np.uint8)
    mdr = rng.integers(0, 5, size=(64, 64)).astype(np.uint8)
    cells = [(int(r), int(c)) for r, c in zip(*np.where(region == 1))]
    act = set(cells)
    pairs = [(a, (a[0], a[1] + 1)) for a in cells if (a[0], a[1] + 1) in act]
    idx = rng.permutation(len(pairs))[:600]
    sc = Scenario("max", region, 16, pol, [([1, 1, 1], [1, 1, 1, 1]), ([0, 1, 2], [1, 0, 2, 1]), ([1, 0, 3], [0, 1, 1, 0])], mdr,
                  blocked=restricted_paths((64, 64), walls=[list(pairs[int(i)]) for i in idx[:300]], oneways=[list(pairs[int(i)]) for i in idx[300:]]))
    kw = dict(num_envs=300, fear=True, fear_weight=-10.0, seed=2, max_steps=25, apples=((63, 63), (61, 59)), fear_radius=9)
    _compare_rollout(RC.GpuBackend(sc, **kw), c_oracle.COracle(sc, threads=8, **kw), steps=30, seed=4)


@pytest.mark.gpu
@pytest.mark.parametrize("bf16", [False, True])
def test_general_gpu_odd_sized_map_takes_the_cell_by_cell_row_writer(bf16):
    """7 x 9 cells: an env's observation block (2 x 63 values) is not a whole number of 16-byte granules, so it cannot leave as
    one bulk copy; the kernel writes such rows cell by cell."""
    from marl_responsible_nav_b200.scenarios import Scenario
    region = np.ones((7, 9), np.int8)
    region[3, 2:7] = 0
    sc = Scenario("odd", region, 5, np.zeros((7, 9), np.uint8), [([1, 2, 1], [1, 1, 1, 1])], np.full((7, 9), 4, np.uint8))
    kw = dict(num_envs=333, fear=True, fear_weight=-1.0, seed=6, max_steps=30, apples=((6, 8), (0, 0)), obs_bf16=bf16)
    _compare_rollout(RC.GpuBackend(sc, **kw), c_oracle.COracle(sc, threads=4, **kw), steps=40, seed=8)


@pytest.mark.gpu
def test_general_layout_equals_packed_layout_on_level3():
    """Two independent CUDA formulations of the same step -- the packed layout's pair-mask table / bit-parallel FeAR and the
    general layout's literal paths -- produce identical rollouts on Level 3 (device RNG; 4 096 envs, FeAR on)."""
    kw = dict(num_envs=4096, fear=True, fear_weight=-5.0, seed=21)
    a, b = RC.GpuBackend("Level 3", layout="packed", **kw), RC.GpuBackend("Level 3", layout="general", **kw)
    rng = np.random.default_rng(1)
    a.reset(); b.reset()
    assert np.array_equal(a.obs, b.obs)
    for t in range(50):
        act = rng.integers(0, 9, size=(4096, 2)).astype(np.int8)
        a.step(act); b.step(act)
        for name in ("obs", "reward", "shaped_reward", "fear", "terminated", "truncated", "ended", "action_mask", "positions"):
            assert np.array_equal(getattr(a, name), getattr(b, name)), (t, name)
        m = a.ended.astype(bool)
        assert np.array_equal(a.final_obs[m], b.final_obs[m]) and np.array_equal(a.info & 0xFFFF, b.info & 0xFFFF), t
    sa, sb = a.stats(), b.stats()
    for k in ("episodes", "episode_len_sum", "crashes", "apples", "fear_nonzero", "return_sum"):
        assert sa[k] == sb[k], k


@pytest.mark.gpu
def test_general_one_update_fear_count_equals_the_literal_nine(monkeypatch):
    """count_valid_fast (one update of the world without the affected agent) against the literal nine re-simulations on the
    GPU itself: a crowded 24 x 24 plaza with 12 agents and walls, where chains of collisions are the rule."""
    from marl_responsible_nav_b200.scenarios import Scenario, restricted_paths
    rng = np.random.default_rng(3)
    region = np.zeros((24, 24), np.int8)
    region[8:16, 8:16] = 1                                     # 64 active cells for 12 agents
    region[12, :] = 1
    cells = [(int(r), int(c)) for r, c in zip(*np.where(region == 1))]
    pairs = [(a, (a[0], a[1] + 1)) for a in cells if region[a[0], min(a[1] + 1, 23)] == 1 and a[1] < 23]
    idx = rng.permutation(len(pairs))[:24]
    sc = Scenario("plaza", region, 12, np.zeros((24, 24), np.uint8), [([1, 2, 2], [1, 1, 1, 1])], rng.integers(0, 9, size=(24, 24)).astype(np.uint8) % 5,
                  blocked=restricted_paths((24, 24), walls=[list(pairs[int(i)]) for i in idx[:12]], oneways=[list(pairs[int(i)]) for i in idx[12:]]))
    kw = dict(num_envs=512, fear=True, seed=8, max_steps=20, apples=((12, 0), (12, 23)), fear_radius=6)
    fast = RC.GpuBackend(sc, **kw)
    monkeypatch.setenv("GWW_FEAR_LITERAL", "1")
    lit = RC.GpuBackend(sc, **kw)
    acts = np.random.default_rng(0).integers(0, 9, size=(25, 512, 2)).astype(np.int8)
    monkeypatch.delenv("GWW_FEAR_LITERAL")
    fast.reset()
    monkeypatch.setenv("GWW_FEAR_LITERAL", "1")
    lit.reset()
    nz = 0
    for t in range(25):
        monkeypatch.delenv("GWW_FEAR_LITERAL")
        fast.step(acts[t])
        monkeypatch.setenv("GWW_FEAR_LITERAL", "1")
        lit.step(acts[t])
        assert np.array_equal(fast.fear, lit.fear) and np.array_equal(fast.positions, lit.positions), t
        nz += int((lit.fear != 0).sum())
    assert nz > 2000, nz
    ora = c_oracle.COracle(sc, threads=8, **kw)
    ora.reset()
    for t in range(3):
        ora.step(acts[t])
    # (the oracle walks the same trajectory: same seed, same actions)
    monkeypatch.delenv("GWW_FEAR_LITERAL")
    chk = RC.GpuBackend(sc, **kw)
    chk.reset()
    for t in range(3):
        chk.step(acts[t])
    assert np.array_equal(chk.fear, ora.fear) and np.array_equal(chk.positions, ora.positions)


@pytest.mark.gpu
def test_general_rollout_into_rings_equals_single_steps():
    """BatchedGridWorld.rollout's contract on the general layout (one launch per step, outputs straight into the ring slots)."""
    import torch
    from marl_responsible_nav_b200 import BatchedGridWorld
    kw = dict(num_envs=300, fear=True, fear_weight=-5.0, seed=13, max_steps=15)
    a, b = BatchedGridWorld(WIDE, **kw), BatchedGridWorld(WIDE, **kw)
    T, K = 7, 12
    acts = torch.randint(0, 9, (5, 300, 2), dtype=torch.int8, device="cuda")
    rings = a.new_rings(T)
    rings.obs[2] = a.reset().obs
    b.reset()
    a.rollout(acts, K, rings, first_slot=2, first_action=3)
    for k in range(K):
        o = b.step(acts[(3 + k) % 5])
        if k >= K - (T - 1):                                  # the newest T - 1 transitions are still in the ring
            slot, nxt = (2 + k) % T, (3 + k) % T
            for name in ("reward", "shaped_reward", "fear", "terminated", "truncated", "ended", "info", "positions"):
                assert torch.equal(getattr(rings, name)[slot], getattr(o, name)), (k, name)
            assert torch.equal(rings.obs[nxt], o.obs) and torch.equal(rings.action_mask[nxt], o.action_mask), k
    assert a.stats()["episodes"] == b.stats()["episodes"] > 0
    a.reset_stats()
    assert a.stats()["episodes"] == 0 and a.stats()["env_steps"] == 0


@pytest.mark.gpu
def test_training_loop_on_the_general_layout():
    """BatchedTrainer on the 20 x 28 x 7 world: PyTorch actors on the 560-cell observations, device replay ring written by
    gww_step and read by gw_replay_sample, the update as the CUDA graph of PyTorch / library kernels (the one-kernel update
    holds critic input rows up to about 800 floats; 2 x 569 here), at the batched cadence."""
    import torch
    from marl_responsible_nav_b200 import maddpg
    hp = maddpg.preset("custom_fear_5")
    hp["MEMORY_SIZE"] = 4096
    env = maddpg.make_env(hp, 256, scenario=WIDE)
    with pytest.warns(UserWarning, match="unsupported shape"):
        tr = maddpg.BatchedTrainer(env, hp=hp, seed=0, learn_cadence="batched")
    assert tr.fused is None and tr.learner is None and tr.agent.obs_dim == 560
    p0 = [p.detach().clone() for p in tr.agent.actors[0].parameters()]
    st = tr.train(40)
    assert st["env_steps"] == 40 * 256 and tr.updates_done >= 3
    assert any(not torch.equal(a, b) for a, b in zip(p0, tr.agent.actors[0].parameters()))
    assert all(bool(torch.isfinite(l.critic_loss).all()) for l in tr.losses)
    b = tr.ring.sample(64, tr.gen)
    assert b["state"].shape == (64, 2, 560)
    s, e = int(b["t"][0]) % tr.ring.T, int(b["env"][0])
    assert torch.equal(b["reward"][0], (-5.0 * tr.ring.fear[s, e] + tr.ring.reward[s, e].double()).float())


@pytest.mark.gpu
def test_general_state_roundtrip_and_errors():
    from marl_responsible_nav_b200 import BatchedGridWorld
    import torch
    env = BatchedGridWorld(WIDE, num_envs=64, seed=3)
    with pytest.raises(RuntimeError, match="gww_reset"):
        env.step(np.zeros((64, 2), np.int8))
    env.reset()
    acts = torch.randint(0, 9, (64, 2), dtype=torch.int8, device="cuda")
    for _ in range(5):
        env.step(acts)
    snap = env.state_dict()
    assert snap.numel() == 64 * 64
    o1 = env.step(acts).obs.clone()
    env.load_state_dict(snap)
    assert torch.equal(env.step(acts).obs, o1)
