"""Generate the golden fixtures under tests/golden/ by RUNNING THE REFERENCE.

Run in the build container (where /root/reference is mounted):

    python tests/golden/make_golden.py            # ~3-4 min, writes *.npz next to this file

The reference ships no tests or golden vectors, so parity is
pinned on outputs of its own code, recorded here with every input needed to
replay them (spawn cells, every agent's action, MdRs).  The fixtures are small
compressed .npz files and are committed; the GPU box never sees /root/reference.
"""
import contextlib
import io
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from _ref_loader import load_reference  # noqa: E402

REF = load_reference()
GW, CA, RESP = REF.grid_world, REF.custom_agent, REF.Responsibility
MA, SE = REF.ma_customenv, REF.customenv


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def load_scenario(name):
    cwd = os.getcwd()
    os.chdir(REF.root)
    try:
        return GW.LoadJsonScenario(scenario_name=name)
    finally:
        os.chdir(cwd)


REGION = np.array(load_scenario("Level 3")["Map"]["Region"])
ACTIVE = [(int(x), int(y)) for x, y in zip(*np.where(REGION == 1))]


def make_world(locs):
    w = GW.GWorld(REGION, Walls=[], OneWays=[])
    for loc in locs:
        ok = w.AddAgent(CA.CustomAgent(), tuple(loc), printStatus=False)
        assert ok
    return w


def clustered_cells(rng, n, p_cluster=0.75, radius=4):
    """n distinct active cells; with prob p_cluster they are drawn near a common anchor."""
    anchor = ACTIVE[rng.integers(len(ACTIVE))]
    if rng.random() < p_cluster:
        near = [c for c in ACTIVE if abs(c[0] - anchor[0]) + abs(c[1] - anchor[1]) <= radius]
        pool = near if len(near) >= n else ACTIVE
    else:
        pool = ACTIVE
    idx = rng.choice(len(pool), size=n, replace=False)
    return [pool[int(i)] for i in idx]


# --------------------------------------------------------------------------- #
KAT = [  # SURVEY.md Appendix B.1 (locs, actions, apples or None)
    ([(0, 0), (0, 1)], [4, 3], None), ([(0, 0), (0, 1)], [4, 4], None), ([(0, 0), (0, 1)], [8, 4], None),
    ([(0, 0), (0, 1)], [4, 8], None), ([(0, 0), (0, 2)], [4, 3], None), ([(0, 0), (0, 2)], [8, 0], None),
    ([(0, 0), (0, 1)], [4, 0], None), ([(0, 0), (0, 3)], [8, 7], None), ([(0, 0), (0, 4)], [8, 7], None),
    ([(0, 0), (0, 3)], [8, 3], None), ([(0, 4), (0, 5)], [4, 2], None), ([(0, 3), (0, 5)], [8, 2], None),
    ([(0, 4), (1, 5)], [4, 1], None), ([(0, 0), (0, 1)], [8, 8], None), ([(0, 0), (0, 2)], [8, 8], None),
    ([(2, 6), (9, 8)], [1, 0], None), ([(3, 5), (9, 8)], [7, 0], None), ([(2, 6), (9, 8)], [5, 0], None),
    ([(0, 0), (9, 8)], [1, 0], None), ([(0, 1), (9, 8)], [7, 0], None), ([(0, 0), (0, 1), (0, 2)], [4, 4, 0], None),
    ([(0, 2), (0, 1), (0, 0)], [0, 4, 4], None), ([(0, 0), (0, 1), (0, 3)], [4, 8, 7], None), ([(0, 0)], [4], None),
    ([(0, 0), (9, 8)], [8, 0], {"apple_0": (0, 1)}), ([(0, 0), (9, 8)], [0, 0], {"apple_0": (0, 0)}),
    ([(0, 0), (0, 1)], [4, 0], {"apple_0": (0, 1)}),
]


def gen_update_cases(n_fuzz=6000, seed=1):
    rng = np.random.default_rng(seed)
    cases = list(KAT)
    for _ in range(n_fuzz):
        n = int(rng.choice([2, 3, 4, 4, 4]))
        locs = clustered_cells(rng, n)
        acts = [int(a) for a in rng.integers(0, 9, size=n)]
        apples = None
        if rng.random() < 0.5:
            near = [c for c in ACTIVE if min(abs(c[0] - l[0]) + abs(c[1] - l[1]) for l in locs) <= 2]
            pick = rng.choice(len(near), size=2, replace=len(near) < 2)
            apples = {"apple_0": near[int(pick[0])], "apple_1": near[int(pick[1])]}
        cases.append((locs, acts, apples))
    C = len(cases)
    out = dict(n=np.zeros(C, np.int8), locs=-np.ones((C, 4, 2), np.int8), acts=np.zeros((C, 4), np.int8),
               apples=-np.ones((C, 2, 2), np.int8), out_locs=-np.ones((C, 4, 2), np.int8),
               crash=np.zeros((C, 4), bool), restr=np.zeros((C, 4), bool), caught=np.zeros((C, 2, 2), np.int8))
    for c, (locs, acts, apples) in enumerate(cases):
        n = len(locs)
        w = make_world(locs)
        al = [(i, a) for i, a in enumerate(acts)]
        out["n"][c] = n
        out["locs"][c, :n] = locs
        out["acts"][c, :n] = acts
        if apples is None:
            crash, restr = w.UpdateGWorld(ActionID4Agents=al)
        else:
            eaters = [i for i in range(min(2, n))]
            for k, v in apples.items():
                out["apples"][c, int(k[-1])] = v
            crash, restr, _, caught = w.UpdateGWorld(ActionID4Agents=al, apples=dict(apples), apple_eaters=eaters)
            for idx, key in caught:
                out["caught"][c, idx, int(key[-1])] += 1
        out["out_locs"][c, :n] = [tuple(int(v) for v in l) for l in w.AgentLocations]
        out["crash"][c, :n] = crash
        out["restr"][c, :n] = restr
    np.savez_compressed(os.path.join(HERE, "update_cases.npz"), **out)
    print("update_cases", C, "crash cases", int(out["crash"].any(1).sum()))


def gen_fear_cases(n_cases=500, seed=2):
    rng = np.random.default_rng(seed)
    out = dict(n=np.zeros(n_cases, np.int8), locs=-np.ones((n_cases, 4, 2), np.int8),
               acts=np.zeros((n_cases, 4), np.int8), mdr=np.zeros((n_cases, 4), np.int8),
               actor=np.zeros(n_cases, np.int8), in_list=np.zeros((n_cases, 4), bool),
               resp=np.zeros((n_cases, 4)), n_mdr=np.zeros((n_cases, 4), np.int8),
               n_act=np.zeros((n_cases, 4), np.int8), fear_sum=np.zeros(n_cases))
    # SURVEY Appendix B.2 as case 0
    fixed = ([(0, 0), (0, 3), (2, 5), (9, 15)], [8, 7, 0, 3], [4, 4, 0, 3], 0, [True] * 4)
    for c in range(n_cases):
        if c == 0:
            locs, acts, mdr, actor, in_list = fixed
            n = 4
        else:
            n = int(rng.choice([3, 4, 4]))
            locs = clustered_cells(rng, n, p_cluster=0.85, radius=int(rng.choice([3, 4, 6])))
            acts = [int(a) for a in rng.integers(0, 9, size=n)]
            mdr = [int(a) for a in rng.integers(0, 5, size=n)]
            actor = int(rng.integers(0, min(2, n)))
            if rng.random() < 0.7:      # the env's close list (radius 5 around the actor)
                in_list = [i == actor or abs(locs[i][0] - locs[actor][0]) + abs(locs[i][1] - locs[actor][1]) <= 5
                           for i in range(n)]
            else:                       # arbitrary subset, actor always present
                in_list = [i == actor or bool(rng.random() < 0.6) for i in range(n)]
        w = make_world(locs)
        lst = [(i, acts[i]) for i in range(n) if in_list[i]]
        resp, n_mdr, n_act, _, _ = RESP.FeAR_4_one_actor(w, lst, [[i, mdr[i]] for i in range(n)], actor)
        out["n"][c] = n
        out["locs"][c, :n] = locs
        out["acts"][c, :n] = acts
        out["mdr"][c, :n] = mdr
        out["actor"][c] = actor
        out["in_list"][c, :n] = in_list
        out["resp"][c, :n] = resp[actor]
        out["n_mdr"][c, :n] = n_mdr[actor]
        out["n_act"][c, :n] = n_act[actor]
        out["fear_sum"][c] = np.sum(resp)
    RESP.CountValidMovesOfAffected_tuple.cache_clear()
    np.savez_compressed(os.path.join(HERE, "fear_cases.npz"), **out)
    print("fear_cases", n_cases, "nonzero", int((out["fear_sum"] != 0).sum()))


def gen_matrix_cases(n_cases=160, seed=3):
    """Responsibility.FeAR (all actors) and Responsibility.FeAL on full and partial action lists."""
    rng = np.random.default_rng(seed)
    out = dict(n=np.zeros(n_cases, np.int8), locs=-np.ones((n_cases, 4, 2), np.int8), acts=np.zeros((n_cases, 4), np.int8),
               mdr=np.zeros((n_cases, 4), np.int8), in_list=np.zeros((n_cases, 4), bool),
               fear=np.zeros((n_cases, 4, 4)), fear_n_mdr=np.zeros((n_cases, 4, 4), np.int8),
               fear_n_act=np.zeros((n_cases, 4, 4), np.int8), feal=np.zeros((n_cases, 4)),
               feal_n_mdr=np.zeros((n_cases, 4), np.int8), feal_n_act=np.zeros((n_cases, 4), np.int8))
    for c in range(n_cases):
        n = int(rng.choice([3, 4, 4]))
        locs = clustered_cells(rng, n, p_cluster=0.9, radius=int(rng.choice([3, 4, 5])))
        acts = [int(a) for a in rng.integers(0, 9, size=n)]
        mdr = [int(a) for a in rng.integers(0, 5, size=n)]
        in_list = [True] * n if rng.random() < 0.7 else [bool(rng.random() < 0.7) for _ in range(n)]
        w = make_world(locs)
        lst = [(i, acts[i]) for i in range(n) if in_list[i]]
        mdrs = [[i, mdr[i]] for i in range(n)]
        resp, n_mdr, n_act, _, _ = RESP.FeAR(w, lst, mdrs)
        feal, fm, fa, _, _ = RESP.FeAL(w, lst, mdrs)
        out["n"][c] = n; out["locs"][c, :n] = locs; out["acts"][c, :n] = acts; out["mdr"][c, :n] = mdr
        out["in_list"][c, :n] = in_list
        out["fear"][c, :n, :n] = resp; out["fear_n_mdr"][c, :n, :n] = n_mdr; out["fear_n_act"][c, :n, :n] = n_act
        out["feal"][c, :n] = feal; out["feal_n_mdr"][c, :n] = fm; out["feal_n_act"][c, :n] = fa
        if c % 40 == 0:
            RESP.CountValidMovesOfAffected_tuple.cache_clear()
    RESP.CountValidMovesOfAffected_tuple.cache_clear()
    np.savez_compressed(os.path.join(HERE, "matrix_cases.npz"), **out)
    print("matrix_cases", n_cases, "fear nonzero", int((out["fear"] != 0).any((1, 2)).sum()), "feal<1", int((out["feal"] < 0.999).any(1).sum()))


_TARGET = [None, (-1, 0), (1, 0), (0, -1), (0, 1), (-2, 0), (2, 0), (0, -2), (0, 2)]


def greedy_action(rng, loc, apple, mask=None):
    """Head for the apple (80 %), else uniform: gives the traces apple catches and shaping rewards."""
    if apple is None or rng.random() < 0.2:
        return int(rng.integers(0, 9))
    best, best_d = [0], abs(loc[0] - apple[0]) + abs(loc[1] - apple[1])
    for a in range(1, 9):
        t = (loc[0] + _TARGET[a][0], loc[1] + _TARGET[a][1])
        if not (0 <= t[0] < REGION.shape[0] and 0 <= t[1] < REGION.shape[1]) or REGION[t] == 0:
            continue
        mid = (loc[0] + _TARGET[a][0] // 2, loc[1] + _TARGET[a][1] // 2)
        if a >= 5 and REGION[mid] == 0:
            continue
        d = abs(t[0] - apple[0]) + abs(t[1] - apple[1])
        if d < best_d:
            best, best_d = [a], d
        elif d == best_d:
            best.append(a)
    return int(best[rng.integers(len(best))])


def pick_learner_actions(rng, masks, masked, locs=None, apples=None):
    acts = []
    if masked == "greedy":
        return [greedy_action(rng, locs[k], apples.get(f"apple_{k}")) for k in range(len(masks))]
    for m in masks:
        if masked:
            valid = np.flatnonzero(m)
            acts.append(int(valid[rng.integers(len(valid))]))
        else:
            acts.append(int(rng.integers(0, 9)))
    return acts


def gen_ma_episodes(fname, sessions, max_steps=150):
    """sessions: list of (seed, fear, n_episodes).  For every session the env's spawn
    stream is default_rng(seed) (ctor) and the two global streams are seeded with
    `seed` right after construction, so the whole session is reproducible by a
    host-side mirror that makes the same RNG calls."""
    ep = dict(seed=[], fear=[], spawn=[], reset_obs=[], reset_mask=[], first_step=[], n_steps=[], session=[])
    st = dict(learner_act=[], all_act=[], mdr=[], locs=[], reward=[], term=[], trunc=[], fear=[],
              crash_count=[], apples_caught=[], obs=[], mask=[])
    for s_idx, (seed, fear, n_eps) in enumerate(sessions):
        with quiet():
            env = MA.CustomMAEnv(render=False, fear=fear, seed=seed)
        random.seed(seed)
        np.random.seed(seed)
        arng = np.random.default_rng(10_000 + seed)     # learner actions: private stream
        for e in range(n_eps):
            obs, info = env.reset()
            ep["seed"].append(seed); ep["fear"].append(fear); ep["session"].append(s_idx)
            ep["spawn"].append([tuple(int(v) for v in l) for l in env.World.AgentLocations])
            ep["reset_obs"].append([obs[a] for a in env.possible_agents])
            ep["reset_mask"].append([info[a]["action_mask"] for a in env.possible_agents])
            ep["first_step"].append(len(st["reward"]))
            masked = ["greedy", True, False][int(arng.integers(0, 3))]
            masks = [info[a]["action_mask"] for a in env.possible_agents]
            n_steps = 0
            for t in range(max_steps):
                la = pick_learner_actions(arng, masks, masked, env.World.AgentLocations, env.apples)
                obs, rew, term, trunc, info = env.step(tuple(la))
                n_steps += 1
                st["learner_act"].append(la)
                st["all_act"].append([int(a) for _, a in env.Action4Agents])
                st["mdr"].append([int(m) for _, m in env.MdR4Agents])
                st["locs"].append([tuple(int(v) for v in l) for l in env.World.AgentLocations])
                st["reward"].append([rew[a] for a in env.possible_agents])
                st["term"].append([term[a] for a in env.possible_agents])
                st["trunc"].append([trunc[a] for a in env.possible_agents])
                st["fear"].append([float(info["fear"][a]) for a in env.possible_agents])
                st["crash_count"].append(info["agent_crashes"])
                st["apples_caught"].append(info["apples_caught"])
                st["obs"].append([obs[a] for a in env.possible_agents])
                masks = [info[a]["action_mask"] for a in env.possible_agents]
                st["mask"].append(masks)
                if all(term.values()) or all(trunc.values()):
                    break
            ep["n_steps"].append(n_steps)
        RESP.CountValidMovesOfAffected_tuple.cache_clear()
    out = dict(
        ep_seed=np.array(ep["seed"], np.int64), ep_fear=np.array(ep["fear"], bool),
        ep_session=np.array(ep["session"], np.int32),
        ep_spawn=np.array(ep["spawn"], np.int8), ep_reset_obs=np.array(ep["reset_obs"], np.float32),
        ep_reset_mask=np.array(ep["reset_mask"], np.int8), ep_first_step=np.array(ep["first_step"], np.int64),
        ep_n_steps=np.array(ep["n_steps"], np.int32),
        learner_act=np.array(st["learner_act"], np.int8), all_act=np.array(st["all_act"], np.int8),
        mdr=np.array(st["mdr"], np.int8), locs=np.array(st["locs"], np.int8),
        reward=np.array(st["reward"], np.int32), term=np.array(st["term"], bool),
        trunc=np.array(st["trunc"], bool), fear=np.array(st["fear"], np.float64),
        crash_count=np.array(st["crash_count"], np.int8), apples_caught=np.array(st["apples_caught"], np.int8),
        obs=np.array(st["obs"], np.float32), mask=np.array(st["mask"], np.int8))
    assert np.array_equal(out["obs"].astype(np.float64), np.array(st["obs"], np.float64))   # small values: exact in fp32
    np.savez_compressed(os.path.join(HERE, fname), **out)
    print(fname, "episodes", len(ep["seed"]), "steps", len(st["reward"]),
          "fear!=0 steps", int((out["fear"] != 0).any(1).sum()), "crash steps", int((out["crash_count"] > 0).sum()),
          "apples", int(out["apples_caught"].sum()))


def gen_single_episodes(fname, sessions, max_steps=150):
    ep = dict(seed=[], fear=[], spawn=[], reset_obs=[], first_step=[], n_steps=[], session=[])
    st = dict(action=[], all_act=[], locs=[], reward=[], term=[], trunc=[], fear=[], restricted=[], obs=[],
              ep_r=[], ep_l=[])
    for s_idx, (seed, fear, n_eps) in enumerate(sessions):
        SE.rng = np.random.default_rng(seed)            # module-level, unseeded in the reference (customenv.py:18)
        with quiet():
            env = SE.CustomEnv(render=False, fear=fear)
        random.seed(seed)
        np.random.seed(seed)
        arng = np.random.default_rng(20_000 + seed)
        for e in range(n_eps):
            obs, _ = env.reset()
            captured = []
            orig = env.World.SelectActionsForAll

            def wrapped(*a, _orig=orig, **k):
                r = _orig(*a, **k)
                captured.append([int(x) for _, x in r])
                return r

            env.World.SelectActionsForAll = wrapped
            ep["seed"].append(seed); ep["fear"].append(fear); ep["session"].append(s_idx)
            ep["spawn"].append([tuple(int(v) for v in l) for l in env.World.AgentLocations])
            ep["reset_obs"].append(np.array(obs))
            ep["first_step"].append(len(st["reward"]))
            n_steps = 0
            greedy = bool(arng.random() < 0.6)
            for t in range(max_steps):
                a = (greedy_action(arng, env.World.AgentLocations[0], env.apples.get("apple_0")) if greedy
                     else int(arng.integers(0, 9)))
                obs, rew, term, trunc, info = env.step([a])
                n_steps += 1
                st["action"].append(a)
                st["all_act"].append(captured[-1])
                st["locs"].append([tuple(int(v) for v in l) for l in env.World.AgentLocations])
                st["reward"].append(float(rew[0])); st["term"].append(bool(term[0])); st["trunc"].append(bool(trunc))
                st["fear"].append(float(info["fear"])); st["restricted"].append(bool(info["restricted"]))
                st["obs"].append(np.array(obs))
                st["ep_r"].append(float(info["episode"]["r"])); st["ep_l"].append(int(info["episode"]["l"]))
                if term[0] or trunc:
                    break
            ep["n_steps"].append(n_steps)
        RESP.CountValidMovesOfAffected_tuple.cache_clear()
    out = dict(
        ep_seed=np.array(ep["seed"], np.int64), ep_fear=np.array(ep["fear"], bool),
        ep_session=np.array(ep["session"], np.int32), ep_spawn=np.array(ep["spawn"], np.int8),
        ep_reset_obs=np.array(ep["reset_obs"], np.float32), ep_first_step=np.array(ep["first_step"], np.int64),
        ep_n_steps=np.array(ep["n_steps"], np.int32),
        action=np.array(st["action"], np.int8), all_act=np.array(st["all_act"], np.int8),
        locs=np.array(st["locs"], np.int8), reward=np.array(st["reward"], np.float64),
        term=np.array(st["term"], bool), trunc=np.array(st["trunc"], bool), fear=np.array(st["fear"], np.float64),
        restricted=np.array(st["restricted"], bool), obs=np.array(st["obs"], np.float32),
        ep_r=np.array(st["ep_r"], np.float64), ep_l=np.array(st["ep_l"], np.int32))
    np.savez_compressed(os.path.join(HERE, fname), **out)
    print(fname, "episodes", len(ep["seed"]), "steps", len(st["reward"]),
          "fear!=0", int((out["fear"] != 0).sum()), "crashes", int(out["term"].sum()), "apples", int(out["trunc"].sum()))


def gen_scenario_tables():
    out = {}
    for name in ("GameMap", "Level 5", "Level 3"):
        sc = load_scenario(name)
        tag = name.replace(" ", "")
        region = np.array(sc["Map"]["Region"])
        out[f"{tag}_region"] = region.astype(np.int8)
        out[f"{tag}_n_agents"] = np.int64(sc["N_Agents"])
        if "Policies" not in sc:
            continue
        pmap = np.zeros(region.shape, dtype=int)          # ma_customenv.py:346-354
        for key in sc["Policies"].keys():
            pmap[sc["Policies"][key]["slicex"], sc["Policies"][key]["slicey"]] = key
        mmap = np.zeros(region.shape, dtype=int)          # :357-365
        for key in sc["MdRs"].keys():
            mmap[sc["MdRs"][key]["slicex"], sc["MdRs"][key]["slicey"]] = key
        mdr_action = np.zeros(region.shape, dtype=int)
        for x in range(region.shape[0]):
            for y in range(region.shape[1]):
                mdr_action[x, y] = sc["MdRs"][str(mmap[x, y]).zfill(2)]["mdr"]
        keys = sorted(int(k) for k in sc["Policies"].keys())
        base = np.zeros((len(keys), 9)); pert = np.zeros((len(keys), 9))
        for i, k in enumerate(keys):
            pol = sc["Policies"][str(k).zfill(2)]
            base[i] = CA.GeneratePolicy(StepWeights=pol["stepWeights"], DirectionWeights=pol["directionWeights"])
            pert[i] = CA.GeneratePolicy(StepWeights=pol["stepWeights"], DirectionWeights=None)
        out[f"{tag}_policy_map"] = pmap.astype(np.int8)
        out[f"{tag}_mdr_action"] = mdr_action.astype(np.int8)
        out[f"{tag}_policy_keys"] = np.array(keys, np.int8)
        out[f"{tag}_policy_base"] = base
        out[f"{tag}_policy_perturbed"] = pert
    np.savez_compressed(os.path.join(HERE, "scenario_tables.npz"), **out)
    print("scenario_tables", sorted(out.keys()))


if __name__ == "__main__":
    which = sys.argv[1:] or ["scenario", "update", "fear", "matrix", "ma", "single"]
    if "scenario" in which:
        gen_scenario_tables()
    if "update" in which:
        gen_update_cases()
    if "fear" in which:
        gen_fear_cases()
    if "matrix" in which:
        gen_matrix_cases()
    if "ma" in which:
        gen_ma_episodes("ma_episodes.npz", [(0, False, 40), (42, False, 40), (66, False, 40),
                                            (0, True, 12), (42, True, 12), (66, True, 12), (7, True, 12)])
    if "single" in which:
        gen_single_episodes("single_episodes.npz", [(0, False, 30), (42, False, 30), (66, True, 10), (5, True, 10)])
