"""Golden fixtures for restricted paths (walls / one-ways), recorded by RUNNING THE REFERENCE with tuple-typed paths.

    python tests/golden/make_wall_golden.py        # build container (needs /root/reference), ~1 min

The reference's JSON loader hands `GWorld` lists of lists, which never equal the tuple paths `UpdateGWorld` tests
(`[old, new] in self.RestrictedPaths`, custom/grid_world.py:498), so walls are inert there; the conversion to tuples is
commented out in `LoadJsonScenario` (:654-665).  Here the same `GWorld` / `CustomMAEnv` code runs with the walls given as
tuples -- the semantics the product's `restricted_paths` / `walls="enforce"` implements -- and its outputs are recorded:
wall_cases.npz (UpdateGWorld and FeAR_4_one_actor cases) and wall_ma_episodes.npz (multi-agent episodes, FeAR on and off).
Some wall entries are deliberately invalid (not adjacent, outside the grid): GWorld.__init__ drops them (:43-60).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402  (loads the reference)

GW, CA, RESP, MA = MG.GW, MG.CA, MG.RESP, MG.MA
REGION, ACTIVE = MG.REGION, MG.ACTIVE


def pick_paths(seed=7, n_walls=16, n_oneways=12):
    rng = np.random.default_rng(seed)
    act = set(ACTIVE)
    pairs = [(a, (a[0] + dr, a[1] + dc)) for a in ACTIVE for dr, dc in ((0, 1), (1, 0)) if (a[0] + dr, a[1] + dc) in act]
    idx = rng.permutation(len(pairs))
    walls = [list(pairs[int(i)]) for i in idx[:n_walls]]
    oneways = []
    for i in idx[n_walls:n_walls + n_oneways]:
        a, b = pairs[int(i)]
        oneways.append([a, b] if rng.random() < 0.5 else [b, a])
    walls += [[(0, 0), (0, 2)], [(1, 1), (2, 2)], [(0, 15), (0, 16)]]          # dropped by GWorld.__init__: not neighbours / off grid
    oneways += [[(3, 3), (5, 3)]]
    return walls, oneways


WALLS, ONEWAYS = pick_paths()


def make_world(locs):
    w = GW.GWorld(REGION, Walls=[list(p) for p in WALLS], OneWays=[list(p) for p in ONEWAYS])
    for loc in locs:
        assert w.AddAgent(CA.CustomAgent(), tuple(loc), printStatus=False)
    return w


def cells_near_paths(rng, n):
    """n distinct active cells, most of them next to a wall / one-way so that the restriction is exercised."""
    ends = [c for p in WALLS[:16] + ONEWAYS[:12] for c in p]
    anchor = ends[int(rng.integers(len(ends)))]
    near = [c for c in ACTIVE if abs(c[0] - anchor[0]) + abs(c[1] - anchor[1]) <= 3]
    pool = near if (len(near) >= n and rng.random() < 0.85) else ACTIVE
    idx = rng.choice(len(pool), size=n, replace=False)
    return [pool[int(i)] for i in idx]


def gen(n_update=3000, n_fear=300, seed=11):
    rng = np.random.default_rng(seed)
    out = dict(walls=np.array(WALLS, np.int16), oneways=np.array(ONEWAYS, np.int16))
    # the RestrictedPaths GWorld built from them: (from, to) pairs, for the host-side test of restricted_paths
    w0 = make_world([ACTIVE[0], ACTIVE[1]])
    out["restricted_paths"] = np.array([[p[0], p[1]] for p in w0.RestrictedPaths], np.int16)
    u = dict(n=np.zeros(n_update, np.int8), locs=-np.ones((n_update, 4, 2), np.int8), acts=np.zeros((n_update, 4), np.int8),
             apples=-np.ones((n_update, 2, 2), np.int8), out_locs=-np.ones((n_update, 4, 2), np.int8),
             crash=np.zeros((n_update, 4), bool), restr=np.zeros((n_update, 4), bool), caught=np.zeros((n_update, 2, 2), np.int8))
    for c in range(n_update):
        n = int(rng.choice([2, 3, 4, 4]))
        locs = cells_near_paths(rng, n)
        acts = [int(a) for a in rng.integers(0, 9, size=n)]
        w = make_world(locs)
        al = [(i, a) for i, a in enumerate(acts)]
        u["n"][c] = n; u["locs"][c, :n] = locs; u["acts"][c, :n] = acts
        if rng.random() < 0.4:
            near = [x for x in ACTIVE if min(abs(x[0] - l[0]) + abs(x[1] - l[1]) for l in locs) <= 2]
            pick = rng.choice(len(near), size=2, replace=len(near) < 2)
            apples = {"apple_0": near[int(pick[0])], "apple_1": near[int(pick[1])]}
            for k, v in apples.items():
                u["apples"][c, int(k[-1])] = v
            crash, restr, _, caught = w.UpdateGWorld(ActionID4Agents=al, apples=dict(apples), apple_eaters=list(range(min(2, n))))
            for idx, key in caught:
                u["caught"][c, idx, int(key[-1])] += 1
        else:
            crash, restr = w.UpdateGWorld(ActionID4Agents=al)
        u["out_locs"][c, :n] = [tuple(int(v) for v in l) for l in w.AgentLocations]
        u["crash"][c, :n] = crash; u["restr"][c, :n] = restr
    out.update({"u_" + k: v for k, v in u.items()})
    f = dict(n=np.zeros(n_fear, np.int8), locs=-np.ones((n_fear, 4, 2), np.int8), acts=np.zeros((n_fear, 4), np.int8),
             mdr=np.zeros((n_fear, 4), np.int8), actor=np.zeros(n_fear, np.int8), in_list=np.zeros((n_fear, 4), bool),
             resp=np.zeros((n_fear, 4)), n_mdr=np.zeros((n_fear, 4), np.int8), n_act=np.zeros((n_fear, 4), np.int8),
             fear_sum=np.zeros(n_fear))
    for c in range(n_fear):
        n = int(rng.choice([3, 4, 4]))
        locs = cells_near_paths(rng, n)
        acts = [int(a) for a in rng.integers(0, 9, size=n)]
        mdr = [int(a) for a in rng.integers(0, 5, size=n)]
        actor = int(rng.integers(0, 2))
        in_list = [i == actor or abs(locs[i][0] - locs[actor][0]) + abs(locs[i][1] - locs[actor][1]) <= 5 for i in range(n)]
        w = make_world(locs)
        lst = [(i, acts[i]) for i in range(n) if in_list[i]]
        resp, n_mdr, n_act, _, _ = RESP.FeAR_4_one_actor(w, lst, [[i, mdr[i]] for i in range(n)], actor)
        f["n"][c] = n; f["locs"][c, :n] = locs; f["acts"][c, :n] = acts; f["mdr"][c, :n] = mdr; f["actor"][c] = actor
        f["in_list"][c, :n] = in_list; f["resp"][c, :n] = resp[actor]; f["n_mdr"][c, :n] = n_mdr[actor]; f["n_act"][c, :n] = n_act[actor]
        f["fear_sum"][c] = np.sum(resp)
        if c % 50 == 0:
            RESP.CountValidMovesOfAffected_tuple.cache_clear()
    RESP.CountValidMovesOfAffected_tuple.cache_clear()
    out.update({"f_" + k: v for k, v in f.items()})
    np.savez_compressed(os.path.join(HERE, "wall_cases.npz"), **out)
    print("wall_cases: restricted paths", len(out["restricted_paths"]), "| update cases", n_update, "restricted moves",
          int(u["restr"].sum()), "crashes", int(u["crash"].sum()), "| fear cases", n_fear, "nonzero", int((f["fear_sum"] != 0).sum()))


def gen_episodes():
    MA.Scenario["Map"]["Walls"] = [list(p) for p in WALLS]          # tuple-typed cells: what LoadJsonScenario's commented-out fix does
    MA.Scenario["Map"]["OneWays"] = [list(p) for p in ONEWAYS]
    try:
        MG.gen_ma_episodes("wall_ma_episodes.npz", [(3, False, 24), (9, True, 8), (21, True, 8)])
    finally:
        MA.Scenario["Map"]["Walls"], MA.Scenario["Map"]["OneWays"] = [], []


if __name__ == "__main__":
    gen()
    gen_episodes()
