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
