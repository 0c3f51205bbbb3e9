"""CPU: the Python oracle restatement vs the golden vectors recorded from the reference
(tests/golden/make_golden.py).  Bit-exact for cells / flags / integer rewards /
observations / masks; FeAR and the single-env float reward compared with == (they
are bit-equal) and, per north_star, would be allowed 1e-6 relative."""
import os

import numpy as np
import pytest

import gridworld_oracle as O


def _npz(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


@pytest.fixture(scope="module")
def level3(golden_dir):
    t = _npz(golden_dir, "scenario_tables.npz")
    keys = [int(k) for k in t["Level3_policy_keys"]]
    # weights are not stored in the fixture; rebuild equivalent integer weights from p (only ratios matter)
    pols = {}
    for i, k in enumerate(keys):
        pols[k] = (t["Level3_policy_base"][i], t["Level3_policy_perturbed"][i])
    sc = O.Scenario(region=t["Level3_region"].astype(np.int8), n_agents=int(t["Level3_n_agents"]),
                    policy_map=t["Level3_policy_map"].astype(np.int64),
                    mdr_map=t["Level3_mdr_action"].astype(np.int64), policies={}, name="Level 3")
    sc._golden_p = pols
    return sc


def test_update_world_golden(golden_dir, level3):
    g = _npz(golden_dir, "update_cases.npz")
    region = level3.region
    for c in range(len(g["n"])):
        n = int(g["n"][c])
        locs = [tuple(int(v) for v in g["locs"][c, i]) for i in range(n)]
        acts = [int(a) for a in g["acts"][c, :n]]
        apples = {k: tuple(int(v) for v in g["apples"][c, k]) for k in range(2) if g["apples"][c, k, 0] >= 0}
        if apples:
            out = O.update_world(region, locs, acts, apples=apples, eaters=list(range(min(2, n))))
        else:
            out = O.update_world(region, locs, acts)
        new_locs, crash, restr, caught, unresolved = out
        assert new_locs == [tuple(int(v) for v in g["out_locs"][c, i]) for i in range(n)], c
        assert crash == list(g["crash"][c, :n]), c
        assert restr == list(g["restr"][c, :n]), c
        cm = np.zeros((2, 2), np.int8)
        for idx, k in caught:
            cm[idx, k] += 1
        assert np.array_equal(cm, g["caught"][c]), c
        assert unresolved == 0


def test_fear_golden(golden_dir, level3):
    g = _npz(golden_dir, "fear_cases.npz")
    for c in range(len(g["n"])):
        n = int(g["n"][c])
        locs = [tuple(int(v) for v in g["locs"][c, i]) for i in range(n)]
        lst = [(i, int(g["acts"][c, i])) for i in range(n) if g["in_list"][c, i]]
        actor = int(g["actor"][c])
        resp, n_mdr, n_act = O.fear_one_actor(level3.region, locs, lst, [int(m) for m in g["mdr"][c, :n]], actor)
        assert np.array_equal(n_mdr[actor], g["n_mdr"][c, :n]), c
        assert np.array_equal(n_act[actor], g["n_act"][c, :n]), c
        assert np.array_equal(resp[actor], g["resp"][c, :n]), c          # bit-equal fp64
        assert float(np.sum(resp)) == float(g["fear_sum"][c]), c


def test_fear_kat_value(golden_dir, level3):
    # SURVEY Appendix B.2 inputs; the reference's own output (agent 3 in the corner has 5 valid moves)
    g = _npz(golden_dir, "fear_cases.npz")
    assert list(g["n_mdr"][0, 1:4]) == [4, 7, 5] and list(g["n_act"][0, 1:4]) == [3, 7, 5]
    assert abs(g["resp"][0, 1] - 0.24999994) < 1e-8


def test_ma_episodes_golden(golden_dir, level3):
    g = _npz(golden_dir, "ma_episodes.npz")
    for e in range(len(g["ep_seed"])):
        env = O.MAEnvOracle(level3, fear=bool(g["ep_fear"][e]))
        obs, masks = env.reset([tuple(int(v) for v in c) for c in g["ep_spawn"][e]])
        assert np.array_equal(np.array(obs, np.float32), g["ep_reset_obs"][e])
        assert np.array_equal(np.array(masks), g["ep_reset_mask"][e])
        s0 = int(g["ep_first_step"][e])
        for t in range(int(g["ep_n_steps"][e])):
            s = s0 + t
            assert env.mdr_of_agents() == [int(m) for m in g["mdr"][s]], (e, t)
            r = env.step(g["learner_act"][s], g["all_act"][s])
            assert r.locs == [tuple(int(v) for v in c) for c in g["locs"][s]], (e, t)
            assert r.rewards == list(g["reward"][s]), (e, t)
            assert r.terminations == list(g["term"][s]) and r.truncations == list(g["trunc"][s]), (e, t)
            assert r.fear == list(g["fear"][s]), (e, t)                 # bit-equal fp64
            assert r.agent_crashes == g["crash_count"][s] and r.apples_caught == g["apples_caught"][s]
            assert np.array_equal(np.array(r.obs, np.float32), g["obs"][s]), (e, t)
            assert np.array_equal(np.array(r.masks), g["mask"][s]), (e, t)


def test_single_episodes_golden(golden_dir, level3):
    g = _npz(golden_dir, "single_episodes.npz")
    for e in range(len(g["ep_seed"])):
        env = O.SingleEnvOracle(level3, fear=bool(g["ep_fear"][e]))
        obs = env.reset([tuple(int(v) for v in c) for c in g["ep_spawn"][e]])
        assert np.array_equal(obs.astype(np.float32), g["ep_reset_obs"][e])
        s0 = int(g["ep_first_step"][e])
        for t in range(int(g["ep_n_steps"][e])):
            s = s0 + t
            r = env.step(int(g["action"][s]), g["all_act"][s])
            assert r.locs == [tuple(int(v) for v in c) for c in g["locs"][s]], (e, t)
            assert r.reward == g["reward"][s], (e, t)                   # bit-equal fp64 (0.1 shaping)
            assert r.terminated == g["term"][s] and r.truncated == g["trunc"][s], (e, t)
            assert r.fear == g["fear"][s], (e, t)
            assert r.restricted == g["restricted"][s]
            assert r.episode_r == g["ep_r"][s] and r.episode_l == g["ep_l"][s]
            assert np.array_equal(r.obs.astype(np.float32), g["obs"][s]), (e, t)


def test_policy_tables_golden(golden_dir):
    t = _npz(golden_dir, "scenario_tables.npz")
    # GeneratePolicy restatement on the Level-3 weights implied by the fixture
    assert np.allclose(O.generate_policy([1, 1, 0], [1, 1, 1, 1]), t["Level3_policy_base"][0], rtol=0, atol=0)
    assert np.array_equal(O.generate_policy([0, 0, 1], [0, 0, 0, 1]), t["Level3_policy_base"][1])
    assert np.array_equal(O.generate_policy([0, 0, 1], None), t["Level3_policy_perturbed"][1])
    assert np.array_equal(O.generate_policy([0, 1, 0], [0, 0, 0, 1]), t["Level3_policy_base"][5])


def test_walls_and_oneways_golden(golden_dir, level3):
    """Restricted paths (walls / one-ways) with the tuple semantics the reference intends: the host-side path list equals
    the RestrictedPaths GWorld built, and the Python restatement reproduces what the reference's UpdateGWorld /
    FeAR_4_one_actor return with them (wall_cases.npz, recorded by tests/golden/make_wall_golden.py).  Without the paths
    a good share of the cases comes out differently: the fixture does exercise them."""
    from marl_responsible_nav_b200.scenarios import restricted_paths
    g = _npz(golden_dir, "wall_cases.npz")
    region = level3.region
    blocked_list = restricted_paths(region.shape, g["walls"].tolist(), g["oneways"].tolist())
    want = [((int(p[0][0]), int(p[0][1])), (int(p[1][0]), int(p[1][1]))) for p in g["restricted_paths"]]
    assert blocked_list == want and len(want) == 44
    blocked = set(blocked_list)
    differs = 0
    for c in range(len(g["u_n"])):
        n = int(g["u_n"][c])
        locs = [tuple(int(v) for v in g["u_locs"][c, i]) for i in range(n)]
        acts = [int(a) for a in g["u_acts"][c, :n]]
        apples = {k: tuple(int(v) for v in g["u_apples"][c, k]) for k in range(2) if g["u_apples"][c, k, 0] >= 0}
        kw = dict(apples=apples, eaters=list(range(min(2, n)))) if apples else {}
        new_locs, crash, restr, caught, _ = O.update_world(region, locs, acts, blocked=blocked, **kw)
        assert new_locs == [tuple(int(v) for v in g["u_out_locs"][c, i]) for i in range(n)], c
        assert crash == list(g["u_crash"][c, :n]) and restr == list(g["u_restr"][c, :n]), c
        cm = np.zeros((2, 2), np.int8)
        for idx, k in caught:
            cm[idx, k] += 1
        assert np.array_equal(cm, g["u_caught"][c]), c
        differs += O.update_world(region, locs, acts, **kw)[:3] != (new_locs, crash, restr)
    assert differs > 200, differs
    for c in range(0, len(g["f_n"]), 3):
        n = int(g["f_n"][c])
        locs = [tuple(int(v) for v in g["f_locs"][c, i]) for i in range(n)]
        lst = [(i, int(g["f_acts"][c, i])) for i in range(n) if g["f_in_list"][c, i]]
        actor = int(g["f_actor"][c])
        resp, n_mdr, n_act = O.fear_one_actor(region, locs, lst, [int(m) for m in g["f_mdr"][c, :n]], actor, blocked)
        assert np.array_equal(resp[actor], g["f_resp"][c, :n]), c
        assert np.array_equal(n_mdr[actor], g["f_n_mdr"][c, :n]) and np.array_equal(n_act[actor], g["f_n_act"][c, :n]), c
