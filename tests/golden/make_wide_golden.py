"""Golden fixtures for the GENERAL layout (grids other than 10x16, more than 4 agents), recorded by RUNNING THE REFERENCE.

    python tests/golden/make_wide_golden.py        # build container (needs /root/reference), a few minutes

`GWorld`, `Responsibility` and both env classes are data-driven: the map, the agent count, the policy / MdR regions and the
restricted paths all come from the scenario dict (custom/ma_customenv.py:27-28, :338-365; custom/customenv.py:25-27).  This
script writes a scenario in the reference's own JSON format (wide_scenarios.json: a 20 x 28 road grid with a plaza, 7 agents,
walls and one-ways), loads it with the reference's `LoadJsonScenario`, points the env modules at it (their module-level
`Scenario` / agent-count globals) and records what the reference computes:

  wide_cases.npz          UpdateGWorld (1..16 agents), FeAR_4_one_actor (5..12 agents), FeAR (all actors) and FeAL cases
  wide_ma_episodes.npz    CustomMAEnv episodes, FeAR off and on (apples stay at the reference's hard-coded cells)
  wide_single_episodes.npz CustomEnv episodes

Walls / one-ways are handed over with tuple cells (the semantics `walls="enforce"` implements, see make_wall_golden.py).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402  (loads the reference)

GW, CA, RESP, MA, SE = MG.GW, MG.CA, MG.RESP, MG.MA, MG.SE
NAME = "Wide 20x28"
JSON_PATH = os.path.join(HERE, "wide_scenarios.json")
H, W, N_AGENTS, PAD = 20, 28, 7, 16


def build_scenario_json():
    region = np.zeros((H, W), dtype=int)
    region[0, :] = region[H - 1, :] = 1
    region[:, 0] = region[:, W - 1] = 1
    for c in (5, 10, 15, 22):
        region[:, c] = 1
    for r in (5, 9, 14):
        region[r, :] = 1
    region[10:14, 11:15] = 1                       # a plaza: agents meet from every side
    assert region[9, 0] and region[5, 10] and region[9, 15]     # the envs' hard-coded apple cells
    pol = {
        "00": dict(slicex=[0, H, 0], slicey=[0, W, 0], stepWeights=[1, 1, 0], directionWeights=[1, 1, 1, 1]),
        "01": dict(slicex=[0, 1, 0], slicey=[0, W - 1, 0], stepWeights=[0, 0, 1], directionWeights=[0, 0, 0, 1]),
        "02": dict(slicex=[H - 1, H, 0], slicey=[1, W, 0], stepWeights=[0, 1, 1], directionWeights=[0, 0, 1, 0]),
        "03": dict(slicex=[1, H, 0], slicey=[0, 1, 0], stepWeights=[0, 0, 1], directionWeights=[1, 0, 0, 0]),
        "04": dict(slicex=[0, H - 1, 0], slicey=[W - 1, W, 0], stepWeights=[1, 1, 1], directionWeights=[0, 1, 0, 0]),
        "05": dict(slicex=[9, 10, 0], slicey=[1, W - 1, 0], stepWeights=[0, 1, 0], directionWeights=[0, 0, 1, 3]),
        "06": dict(slicex=[1, H - 1, 0], slicey=[15, 16, 0], stepWeights=[1, 2, 1], directionWeights=[2, 1, 0, 0]),
        "07": dict(slicex=[10, 14, 0], slicey=[11, 15, 0], stepWeights=[1, 3, 0], directionWeights=[1, 1, 1, 1]),
        "08": dict(slicex=[5, 6, 0], slicey=[2, 26, 2], stepWeights=[0, 1, 1], directionWeights=[0, 0, 1, 1]),   # stepped slice
    }
    mdr = {
        "00": dict(slicex=[0, H, 0], slicey=[0, W, 0], mdr=0),
        "01": dict(slicex=[0, 1, 0], slicey=[0, W - 1, 0], mdr=4),
        "02": dict(slicex=[H - 1, H, 0], slicey=[1, W, 0], mdr=3),
        "03": dict(slicex=[1, H, 0], slicey=[0, 1, 0], mdr=1),
        "04": dict(slicex=[0, H - 1, 0], slicey=[W - 1, W, 0], mdr=2),
        "05": dict(slicex=[9, 10, 0], slicey=[1, W - 1, 0], mdr=4),
        "06": dict(slicex=[1, H - 1, 0], slicey=[15, 16, 0], mdr=1),
    }
    rng = np.random.default_rng(5)
    act = {(int(x), int(y)) for x, y in zip(*np.where(region == 1))}
    pairs = [(a, (a[0] + dr, a[1] + dc)) for a in sorted(act) for dr, dc in ((0, 1), (1, 0)) if (a[0] + dr, a[1] + dc) in act]
    idx = rng.permutation(len(pairs))
    walls = [[list(pairs[int(i)][0]), list(pairs[int(i)][1])] for i in idx[:40]]
    oneways = []
    for i in idx[40:70]:
        a, b = pairs[int(i)]
        oneways.append([list(a), list(b)] if rng.random() < 0.5 else [list(b), list(a)])
    walls += [[[0, 0], [0, 2]], [[1, 1], [2, 2]], [[0, W - 1], [0, W]]]       # dropped by GWorld.__init__ (:43-60)
    sc = dict(AgentLocations=[], DirectionWeights=[1, 1, 1, 1], StepWeights=[1, 1, 1], N_Agents=N_AGENTS, N_Cases=1,
              N_iterations=10, SpecificAction4Agents=[], SpecificDirectionWeights4Agents=[], SpecificStepWeights4Agents=[],
              defaultAction="random", Map=dict(Region=region.astype(float).tolist(), Walls=walls, OneWays=oneways), Policies=pol, MdRs=mdr)
    with open(JSON_PATH, "w") as f:
        json.dump({NAME: sc}, f, separators=(",", ":"))
    return sc


def tupled(paths):
    return [[tuple(p[0]), tuple(p[1])] for p in paths]


build_scenario_json()      # Region as floats, like every shipped scenario (an int Region would make WorldState an int array: the 0.5 spawn marker truncates to 0)
SC = GW.LoadJsonScenario(json_filename=JSON_PATH, scenario_name=NAME)      # the reference's loader (slices become slice objects)
REGION = np.array(SC["Map"]["Region"])
ACTIVE = [(int(x), int(y)) for x, y in zip(*np.where(REGION == 1))]
WALLS, ONEWAYS = tupled(SC["Map"]["Walls"]), tupled(SC["Map"]["OneWays"])


def make_world(locs):
    w = GW.GWorld(REGION, Walls=[list(p) for p in WALLS], OneWays=[list(p) for p in ONEWAYS])
    for loc in locs:
        assert w.AddAgent(CA.CustomAgent(), tuple(loc), printStatus=False)
    return w


def clustered(rng, n, radius):
    anchor = ACTIVE[int(rng.integers(len(ACTIVE)))]
    near = [c for c in ACTIVE if abs(c[0] - anchor[0]) + abs(c[1] - anchor[1]) <= radius]
    pool = near if (len(near) >= n and rng.random() < 0.9) else ACTIVE
    idx = rng.choice(len(pool), size=n, replace=False)
    return [pool[int(i)] for i in idx]


def gen_cases(n_update=2500, n_fear=160, n_matrix=50, seed=17):
    rng = np.random.default_rng(seed)
    out = {}
    w0 = make_world([ACTIVE[0], ACTIVE[1]])
    out["restricted_paths"] = np.array([[p[0], p[1]] for p in w0.RestrictedPaths], np.int16)
    u = dict(n=np.zeros(n_update, np.int8), locs=-np.ones((n_update, PAD, 2), np.int8), acts=np.zeros((n_update, PAD), np.int8),
             apples=-np.ones((n_update, 2, 2), np.int8), out_locs=-np.ones((n_update, PAD, 2), np.int8),
             crash=np.zeros((n_update, PAD), bool), restr=np.zeros((n_update, PAD), bool), caught=np.zeros((n_update, 2, 2), np.int8))
    for c in range(n_update):
        n = int(rng.choice([2, 3, 5, 6, 7, 8, 10, 16])) if c % 25 else 1
        locs = clustered(rng, n, radius=int(rng.choice([3, 4, 5])) + n // 4)
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
    print("update cases", n_update, "crashes", int(u["crash"].sum()), "restricted", int(u["restr"].sum()), flush=True)

    f = dict(n=np.zeros(n_fear, np.int8), locs=-np.ones((n_fear, PAD, 2), np.int8), acts=np.zeros((n_fear, PAD), np.int8),
             mdr=np.zeros((n_fear, PAD), np.int8), actor=np.zeros(n_fear, np.int8), in_list=np.zeros((n_fear, PAD), bool),
             resp=np.zeros((n_fear, PAD)), n_mdr=np.zeros((n_fear, PAD), np.int8), n_act=np.zeros((n_fear, PAD), np.int8),
             fear_sum=np.zeros(n_fear))
    for c in range(n_fear):
        n = int(rng.choice([5, 6, 7, 9, 12]))
        locs = clustered(rng, n, radius=int(rng.choice([3, 4, 6])) + n // 4)
        acts = [int(a) for a in rng.integers(0, 9, size=n)]
        mdr = [int(a) for a in rng.integers(0, 5, size=n)]
        actor = int(rng.integers(0, n))
        if rng.random() < 0.7:
            in_list = [i == actor or abs(locs[i][0] - locs[actor][0]) + abs(locs[i][1] - locs[actor][1]) <= 5 for i in range(n)]
        else:
            in_list = [i == actor or bool(rng.random() < 0.6) for i in range(n)]
        w = make_world(locs)
        lst = [(i, acts[i]) for i in range(n) if in_list[i]]
        resp, n_mdr, n_act, _, _ = RESP.FeAR_4_one_actor(w, lst, [[i, mdr[i]] for i in range(n)], actor)
        f["n"][c] = n; f["locs"][c, :n] = locs; f["acts"][c, :n] = acts; f["mdr"][c, :n] = mdr; f["actor"][c] = actor
        f["in_list"][c, :n] = in_list; f["resp"][c, :n] = resp[actor]; f["n_mdr"][c, :n] = n_mdr[actor]; f["n_act"][c, :n] = n_act[actor]
        f["fear_sum"][c] = np.sum(resp)
        if c % 20 == 0:
            RESP.CountValidMovesOfAffected_tuple.cache_clear()
    RESP.CountValidMovesOfAffected_tuple.cache_clear()
    out.update({"f_" + k: v for k, v in f.items()})
    print("fear cases", n_fear, "nonzero", int((f["fear_sum"] != 0).sum()), flush=True)

    m = dict(n=np.zeros(n_matrix, np.int8), locs=-np.ones((n_matrix, PAD, 2), np.int8), acts=np.zeros((n_matrix, PAD), np.int8),
             mdr=np.zeros((n_matrix, PAD), np.int8), in_list=np.zeros((n_matrix, PAD), bool),
             fear=np.zeros((n_matrix, PAD, PAD)), fear_n_mdr=np.zeros((n_matrix, PAD, PAD), np.int8),
             fear_n_act=np.zeros((n_matrix, PAD, PAD), np.int8), feal=np.zeros((n_matrix, PAD)),
             feal_n_mdr=np.zeros((n_matrix, PAD), np.int8), feal_n_act=np.zeros((n_matrix, PAD), np.int8))
    for c in range(n_matrix):
        n = int(rng.choice([5, 6, 7]))
        locs = clustered(rng, n, radius=int(rng.choice([3, 4, 5])))
        acts = [int(a) for a in rng.integers(0, 9, size=n)]
        mdr = [int(a) for a in rng.integers(0, 5, size=n)]
        in_list = [True] * n if rng.random() < 0.7 else [bool(rng.random() < 0.7) for _ in range(n)]
        w = make_world(locs)
        lst = [(i, acts[i]) for i in range(n) if in_list[i]]
        mdrs = [[i, mdr[i]] for i in range(n)]
        resp, n_mdr, n_act, _, _ = RESP.FeAR(w, lst, mdrs)
        feal, fm, fa, _, _ = RESP.FeAL(w, lst, mdrs)
        m["n"][c] = n; m["locs"][c, :n] = locs; m["acts"][c, :n] = acts; m["mdr"][c, :n] = mdr; m["in_list"][c, :n] = in_list
        m["fear"][c, :n, :n] = resp; m["fear_n_mdr"][c, :n, :n] = n_mdr; m["fear_n_act"][c, :n, :n] = n_act
        m["feal"][c, :n] = feal; m["feal_n_mdr"][c, :n] = fm; m["feal_n_act"][c, :n] = fa
        if c % 10 == 0:
            RESP.CountValidMovesOfAffected_tuple.cache_clear()
    RESP.CountValidMovesOfAffected_tuple.cache_clear()
    out.update({"m_" + k: v for k, v in m.items()})
    print("matrix cases", n_matrix, "fear nonzero", int((m["fear"] != 0).any((1, 2)).sum()), flush=True)
    np.savez_compressed(os.path.join(HERE, "wide_cases.npz"), **out)


def gen_episodes():
    live = dict(SC)
    live["Map"] = dict(Region=SC["Map"]["Region"], Walls=[list(p) for p in WALLS], OneWays=[list(p) for p in ONEWAYS])
    saved = (MA.Scenario, MA.total_num_agents, SE.Scenario, SE.num_agents, MG.REGION)
    MA.Scenario, MA.total_num_agents, SE.Scenario, SE.num_agents = live, N_AGENTS, live, N_AGENTS
    MG.REGION = REGION                               # greedy_action looks at the map
    try:
        MG.gen_ma_episodes("wide_ma_episodes.npz", [(4, False, 24), (12, True, 6), (31, True, 6)], max_steps=60)
        MG.gen_single_episodes("wide_single_episodes.npz", [(6, False, 16), (15, True, 10)], max_steps=60)
    finally:
        MA.Scenario, MA.total_num_agents, SE.Scenario, SE.num_agents, MG.REGION = saved


if __name__ == "__main__":
    gen_cases()
    gen_episodes()
