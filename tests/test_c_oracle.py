"""CPU: the C restatement (oracle/gw_oracle.c) vs the golden vectors recorded from the reference."""
import numpy as np
import pytest

import c_oracle
import gridworld_oracle as PO
import replay_checks as RC


def make(**kw):
    return c_oracle.COracle("Level 3", **kw)


def test_update_cases():
    assert RC.check_update_cases(make) > 6000


def test_fear_cases():
    assert RC.check_fear_cases(make) == 500


def make_walls(**kw):
    return c_oracle.COracle(RC.wall_scenario(), **kw)


def test_walls_and_oneways_golden():
    """Restricted paths with the tuple semantics the reference intends: UpdateGWorld, FeAR and whole episodes recorded by
    running the reference with tuple-typed Walls / OneWays (tests/golden/make_wall_golden.py)."""
    assert RC.check_update_cases(make_walls, "wall_cases.npz", "u_") == 3000
    assert RC.check_fear_cases(make_walls, "wall_cases.npz", "f_") == 300
    assert RC.check_ma_episodes(make_walls, fixture="wall_ma_episodes.npz") > 4000


def test_fear_matrix_and_feal_cases():
    assert RC.check_matrix_cases(make) == 160


def test_python_oracle_fear_matrix_and_feal():
    g = RC.npz("matrix_cases.npz")
    region = c_oracle.builtin_scenario("Level 3").region
    for c in range(0, 160, 4):
        n = int(g["n"][c])
        locs = [tuple(int(v) for v in g["locs"][c, i]) for i in range(n)]
        lst = [(i, int(g["acts"][c, i])) for i in range(n) if g["in_list"][c, i]]
        mdr = [int(m) for m in g["mdr"][c, :n]]
        resp, nm, na = PO.fear_all_actors(region, locs, lst, mdr)
        assert np.array_equal(resp, g["fear"][c, :n, :n]) and np.array_equal(nm, g["fear_n_mdr"][c, :n, :n])
        fl, fm, fa = PO.feal(region, locs, lst, mdr)
        assert np.array_equal(fl, g["feal"][c, :n]) and np.array_equal(fa, g["feal_n_act"][c, :n])


def test_ma_episodes():
    assert RC.check_ma_episodes(make) > 1500


def test_ma_episodes_bf16_obs():
    assert RC.check_ma_episodes(make, obs_bf16=True) > 1500


def test_ma_sessions_autoreset():
    assert RC.check_ma_sessions_autoreset(make) > 1500


def test_single_episodes():
    assert RC.check_single_episodes(make) > 1000


def test_np_sum_order_matches_numpy():
    """fear = np.sum(Resp) (ma_customenv.py:252): the C oracle's pairwise order vs numpy itself on every triple of
    attainable Resp values placed in either learner's row (n = 4) and every pair (n = 3)."""
    vals = sorted({float(np.clip((m - a) / (m + 0.000001), -1, 1)) for m in range(10) for a in range(10)})
    rng = np.random.default_rng(0)
    o = make(num_envs=1)
    import ctypes as C
    lib = c_oracle.load()
    # exercise through gwo_fear_one_actor's fear_sum on synthetic rows is not possible (it recomputes Resp), so check the
    # documented association directly against numpy: first + (second + third)
    for _ in range(20000):
        a, b, c = rng.choice(vals, 3)
        for actor, cols in ((0, (1, 2, 3)), (1, (0, 2, 3))):
            m = np.zeros((4, 4)); m[actor, cols[0]], m[actor, cols[1]], m[actor, cols[2]] = a, b, c
            assert np.sum(m) == a + (b + c)
        for actor, cols in ((0, (1, 2)), (1, (0, 2))):
            m = np.zeros((3, 3)); m[actor, cols[0]], m[actor, cols[1]] = a, b
            assert np.sum(m) == a + b


def test_c_oracle_matches_python_oracle_random_rollout():
    """Device-RNG mode of the C oracle (spawn + NPC draws from Philox) cross-checked against the Python restatement
    fed with the C oracle's own draws (positions and info words reveal them)."""
    sc = c_oracle.builtin_scenario("Level 3")
    E = 64
    o = make(num_envs=E, fear=True, auto_reset=False, max_steps=0, seed=123)
    o.reset()
    psc = PO.Scenario(region=sc.region, n_agents=4, policy_map=sc.policy_index.astype(np.int64),
                      mdr_map=sc.mdr_action.astype(np.int64), policies={})
    envs = []
    for e in range(E):
        env = PO.MAEnvOracle(psc, fear=True)
        env.reset([tuple(int(v) for v in c) for c in o.positions[e]])
        envs.append(env)
    rng = np.random.default_rng(5)
    for t in range(6):
        la = rng.integers(0, 9, size=(E, 2)).astype(np.int8)
        npc = rng.integers(0, 9, size=(E, 4)).astype(np.int8)
        o.step(la, npc_actions=npc)
        for e in range(E):
            r = envs[e].step(la[e], npc[e])
            assert r.locs == [tuple(int(v) for v in c) for c in o.positions[e]]
            assert r.rewards == list(o.reward[e].astype(int))
            assert r.fear == list(o.fear[e])
            assert np.array_equal(np.array(r.obs, np.float32).reshape(2, 160), o.obs[e])


def test_device_rng_statistics():
    """Native mode: NPC action frequencies per policy region vs GeneratePolicy, spawn uniformity (chi-square)."""
    sc = c_oracle.builtin_scenario("Level 3")
    E = 20000
    o = make(num_envs=E, fear=False, auto_reset=True, max_steps=150, seed=7)
    o.reset()
    cells = sc.active_cells()
    idx_of = {c: i for i, c in enumerate(cells)}
    counts = np.zeros(len(cells))
    for e in range(E):
        for a in range(4):
            counts[idx_of[tuple(int(v) for v in o.positions[e, a])]] += 1
    exp = E * 4 / len(cells)
    chi2 = ((counts - exp) ** 2 / exp).sum()
    assert chi2 < 130, chi2            # 71 dof: mean 71, p(>130) ~ 1e-5
    # spawn cells are sorted row-major and distinct
    flat = o.positions[:, :, 0].astype(int) * 16 + o.positions[:, :, 1]
    assert (np.diff(flat, axis=1) > 0).all()
    # NPC actions: infer from displacement of NPC agent 2 when nothing blocks it -- instead compare against the expected
    # mixture 0.75*base + 0.25*perturbed through the move it made (blocked moves stay put and are counted as such)
    before = o.positions.copy()
    la = np.zeros((E, 2), np.int8)
    o.step(la)
    hist = {}
    for e in range(E):
        for a in (2, 3):
            c0 = tuple(int(v) for v in before[e, a])
            key = int(sc.policy_index[c0])
            if o.ended[e] or (o.info[e] & 0xF):          # skip crashed / respawned envs
                continue
            d = (int(o.positions[e, a, 0]) - c0[0], int(o.positions[e, a, 1]) - c0[1])
            hist.setdefault(key, {}).setdefault(d, 0)
            hist[key][d] += 1
    # region 0 (stay or one step, uniform): every unblocked single step must appear, never a double step
    moves0 = hist[0]
    assert all(abs(d[0]) + abs(d[1]) <= 1 for d in moves0)
    # outer ring top row (policy 1): moves right by two (75 % + 25 %/4) dominate
    m1 = hist[1]
    tot = sum(m1.values())
    assert m1.get((0, 2), 0) / tot > 0.6
