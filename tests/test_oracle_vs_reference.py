"""CPU, build container only: live fuzz of both oracles against the reference's own code (skipped when
/root/reference is not mounted, e.g. on the GPU box; the committed golden vectors cover that case)."""
import numpy as np
import pytest

import _ref_loader
import gridworld_oracle as PO

pytestmark = pytest.mark.skipif(not _ref_loader.reference_available(), reason="reference not mounted")


@pytest.fixture(scope="module")
def ref():
    return _ref_loader.load_reference()


def _world(ref, region, locs):
    w = ref.grid_world.GWorld(region, Walls=[], OneWays=[])
    for loc in locs:
        assert w.AddAgent(ref.custom_agent.CustomAgent(), tuple(loc), printStatus=False)
    return w


def test_update_world_live_fuzz(ref):
    import c_oracle
    sc = c_oracle.builtin_scenario("Level 3")
    region = sc.region.astype(float)
    cells = sc.active_cells()
    rng = np.random.default_rng(2024)
    co = c_oracle.COracle("Level 3", num_envs=1, fear=False)
    C = 1500
    pos = np.zeros((C, 4, 2), np.int8); act = np.zeros((C, 4), np.int8); nn = np.zeros(C, np.int8)
    want = []
    for c in range(C):
        n = int(rng.integers(2, 5))
        anchor = cells[rng.integers(len(cells))]
        near = [x for x in cells if abs(x[0] - anchor[0]) + abs(x[1] - anchor[1]) <= 4]
        pool = near if len(near) >= n else cells
        locs = [pool[int(i)] for i in rng.choice(len(pool), size=n, replace=False)]
        acts = [int(a) for a in rng.integers(0, 9, size=n)]
        w = _world(ref, region, locs)
        crash, restr = w.UpdateGWorld(ActionID4Agents=list(enumerate(acts)))
        want.append(([tuple(int(v) for v in l) for l in w.AgentLocations], list(crash), list(restr)))
        got = PO.update_world(sc.region, locs, acts)
        assert (got[0], got[1], got[2]) == want[-1], c
        pos[c, :n] = locs; act[c, :n] = acts; nn[c] = n
    new_pos, crash, restr, _ = co.update_world(pos, act, n_agents=nn)
    for c, (locs, cr, rs) in enumerate(want):
        n = int(nn[c])
        assert [tuple(int(v) for v in p) for p in new_pos[c, :n]] == locs
        assert list(crash[c, :n].astype(bool)) == cr and list(restr[c, :n].astype(bool)) == rs


def test_fear_live_fuzz(ref):
    import c_oracle
    sc = c_oracle.builtin_scenario("Level 3")
    region = sc.region.astype(float)
    cells = sc.active_cells()
    rng = np.random.default_rng(7)
    co = c_oracle.COracle("Level 3", num_envs=1, fear=True)
    for c in range(40):
        anchor = cells[rng.integers(len(cells))]
        near = [x for x in cells if abs(x[0] - anchor[0]) + abs(x[1] - anchor[1]) <= 4]
        pool = near if len(near) >= 4 else cells
        locs = [pool[int(i)] for i in rng.choice(len(pool), size=4, replace=False)]
        acts = [int(a) for a in rng.integers(0, 9, size=4)]
        mdr = [int(a) for a in rng.integers(0, 5, size=4)]
        actor = int(rng.integers(0, 2))
        lst = PO.close_agents(locs, list(enumerate(acts)), actor, 5)
        w = _world(ref, region, locs)
        resp, n_mdr, n_act, _, _ = ref.Responsibility.FeAR_4_one_actor(w, lst, [[i, m] for i, m in enumerate(mdr)], actor)
        r2, m2, a2 = PO.fear_one_actor(sc.region, locs, lst, mdr, actor)
        assert np.array_equal(resp, r2) and np.array_equal(n_mdr, m2) and np.array_equal(n_act, a2)
        in_list = np.array([[any(i == a for a, _ in lst) for i in range(4)]])
        r3, m3, a3, fs = co.fear_one_actor(np.array([locs], np.int8), np.array([acts], np.int8), np.array([mdr], np.int8),
                                           np.array([actor], np.int8), in_list=in_list)
        assert np.array_equal(r3[0], resp[actor]) and fs[0] == np.sum(resp)
    ref.Responsibility.CountValidMovesOfAffected_tuple.cache_clear()


def test_general_layout_live_fuzz_random_maps(ref):
    """The general layout's oracles on RANDOM worlds: maps of random size (up to 40 x 40), random obstacles, random walls /
    one-ways (tuple-typed, the enforced semantics), 2..14 agents -- UpdateGWorld and FeAR_4_one_actor of the reference, live,
    against the Python restatement and the C oracle built on gww_config."""
    import c_oracle
    from marl_responsible_nav_b200.scenarios import Scenario, restricted_paths
    rng = np.random.default_rng(99)
    for world in range(12):
        H, W = int(rng.integers(5, 41)), int(rng.integers(5, 41))
        region = (rng.random((H, W)) < 0.85).astype(np.int8)
        region[0, :] = 1
        cells = [(int(r), int(c)) for r, c in zip(*np.where(region == 1))]
        act_set = set(cells)
        pairs = [(a, (a[0] + dr, a[1] + dc)) for a in cells for dr, dc in ((0, 1), (1, 0)) if (a[0] + dr, a[1] + dc) in act_set]
        idx = rng.permutation(len(pairs))
        nw = min(len(pairs) // 6, 60)
        walls = [[pairs[int(i)][0], pairs[int(i)][1]] for i in idx[:nw]]
        oneways = [[pairs[int(i)][1], pairs[int(i)][0]] if rng.random() < 0.5 else [pairs[int(i)][0], pairs[int(i)][1]] for i in idx[nw:2 * nw]]
        sc = Scenario(f"fuzz{world}", region, 4, np.zeros((H, W), np.uint8), [([1, 1, 1], [1, 1, 1, 1])], np.zeros((H, W), np.uint8),
                      blocked=restricted_paths((H, W), walls, oneways))
        co = c_oracle.COracle(sc, num_envs=1, fear=True, layout="wide", apples=((0, 0), (0, 1)))
        blocked = set(sc.blocked)
        C = 60
        pos = np.zeros((C, 16, 2), np.int8); act = np.zeros((C, 16), np.int8); nn = np.zeros(C, np.int8)
        want = []
        for c in range(C):
            n = int(rng.integers(2, min(15, len(cells))))
            anchor = cells[int(rng.integers(len(cells)))]
            near = [x for x in cells if abs(x[0] - anchor[0]) + abs(x[1] - anchor[1]) <= 3 + n // 3]
            pool = near if len(near) >= n else cells
            locs = [pool[int(i)] for i in rng.choice(len(pool), size=n, replace=False)]
            acts = [int(a) for a in rng.integers(0, 9, size=n)]
            w = ref.grid_world.GWorld(region.astype(float), Walls=[list(p) for p in walls], OneWays=[list(p) for p in oneways])
            for loc in locs:
                assert w.AddAgent(ref.custom_agent.CustomAgent(), tuple(loc), printStatus=False)
            if c == 0:
                assert [((int(p[0][0]), int(p[0][1])), (int(p[1][0]), int(p[1][1]))) for p in w.RestrictedPaths] == sc.blocked
            if c % 10 == 5 and n <= 8:                           # FeAR on the same world before it moves
                mdr = [int(a) for a in rng.integers(0, 5, size=n)]
                actor = int(rng.integers(0, n))
                lst = PO.close_agents(locs, list(enumerate(acts)), actor, 5)
                resp, n_mdr, n_act, _, _ = ref.Responsibility.FeAR_4_one_actor(w, lst, [[i, m] for i, m in enumerate(mdr)], actor)
                r2, m2, a2 = PO.fear_one_actor(region, locs, lst, mdr, actor, blocked)
                assert np.array_equal(resp, r2) and np.array_equal(n_mdr, m2) and np.array_equal(n_act, a2), (world, c)
                il = np.zeros((1, 16), np.uint8)
                for a, _ in lst:
                    il[0, a] = 1
                p1 = np.zeros((1, 16, 2), np.int8); a1 = np.zeros((1, 16), np.int8); m1 = np.zeros((1, 16), np.int8)
                p1[0, :n] = locs; a1[0, :n] = acts; m1[0, :n] = mdr
                r3, m3, a3, fs = co.fear_one_actor(p1, a1, m1, np.array([actor], np.int8), in_list=il, n_agents=np.array([n], np.int8))
                assert np.array_equal(r3[0, :n], resp[actor]) and np.array_equal(m3[0, :n], n_mdr[actor]) and fs[0] == np.sum(resp), (world, c)
            crash, restr = w.UpdateGWorld(ActionID4Agents=list(enumerate(acts)))
            want.append(([tuple(int(v) for v in l) for l in w.AgentLocations], list(crash), list(restr)))
            got = PO.update_world(region, locs, acts, blocked=blocked)
            assert (got[0], got[1], got[2]) == want[-1], (world, c)
            pos[c, :n] = locs; act[c, :n] = acts; nn[c] = n
        new_pos, crash, restr, _ = co.update_world(pos, act, n_agents=nn)
        for c, (locs, cr, rs) in enumerate(want):
            n = int(nn[c])
            assert [tuple(int(v) for v in p) for p in new_pos[c, :n]] == locs, (world, c)
            assert list(crash[c, :n].astype(bool)) == cr and list(restr[c, :n].astype(bool)) == rs, (world, c)
        ref.Responsibility.CountValidMovesOfAffected_tuple.cache_clear()
