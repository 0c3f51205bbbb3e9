"""Shared replay checks: run a backend over the golden traces recorded from the reference and compare.

A "backend" exposes the numpy attributes of oracle/c_oracle.py::COracle (obs, reward, fear, ...).
`tests/test_c_oracle.py` runs these checks on the C oracle (CPU); `tests/test_gpu_parity.py` runs the very
same checks on the CUDA library through the C-ABI (via GpuBackend below).
"""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def npz(name):
    return np.load(os.path.join(GOLDEN, name))


class GpuBackend:
    """BatchedGridWorld -> numpy attribute view (same surface as COracle)."""

    def __init__(self, scenario="Level 3", num_envs=1, **kw):
        import torch
        from marl_responsible_nav_b200 import BatchedGridWorld
        self.torch = torch
        obs_bf16 = kw.pop("obs_bf16", False)
        kw.pop("threads", None)
        self.obs_bf16 = obs_bf16
        self.env = BatchedGridWorld(scenario, num_envs=num_envs, obs_dtype=torch.bfloat16 if obs_bf16 else torch.float32, **kw)
        self.E = num_envs
        shape = (num_envs, self.env.n_learners, self.env.obs_len)
        self._final = torch.zeros(shape, dtype=self.env.obs_dtype, device=self.env.device)

    def _np_obs(self, t):
        if self.obs_bf16:
            return t.view(self.torch.int16).cpu().numpy().view(np.uint16)
        return t.cpu().numpy()

    def _pull(self, out, step):
        self.obs = self._np_obs(out.obs)
        self.action_mask = out.action_mask.cpu().numpy()
        self.positions = out.positions.cpu().numpy()
        if step:
            self.final_obs = self._np_obs(self._final)
            self.reward = out.reward.cpu().numpy()
            self.shaped_reward = out.shaped_reward.cpu().numpy()
            self.fear = out.fear.cpu().numpy()
            self.terminated = out.terminated.cpu().numpy()
            self.truncated = out.truncated.cpu().numpy()
            self.ended = out.ended.cpu().numpy()
            self.info = out.info.cpu().numpy().view(np.uint32)

    def reset(self, mask=None, spawn=None):
        m = None if mask is None else self.torch.as_tensor(np.asarray(mask, dtype=np.uint8), device=self.env.device)
        self._pull(self.env.reset(mask=m, spawn=spawn), False)
        return self

    def step(self, actions, npc_actions=None, spawn=None):
        self._pull(self.env.step(np.asarray(actions, np.int8), npc_actions=npc_actions, spawn=spawn,
                                 final_obs_out=self._final), True)
        return self

    def stats(self):
        return self.env.stats()

    def state(self):
        return self.env.state_dict().numpy().view(np.uint32).reshape(self.E, -1)      # 4 words (packed layout) or 16 (general)

    def update_world(self, positions, actions, n_agents=None, apples=None):
        return tuple(t.cpu().numpy() for t in self.env.update_world(positions, actions, n_agents, apples))

    def fear_one_actor(self, positions, actions, mdr, actor, in_list=None, n_agents=None):
        r = self.env.fear_one_actor(positions, actions, mdr, actor, in_list, n_agents)
        fsum = getattr(self.env, "last_fear_sum", None)                          # general layout: np.sum of the matrix as well
        return tuple(t.cpu().numpy() for t in r) + (None if fsum is None else fsum.cpu().numpy(),)

    def fear_matrix(self, positions, actions, mdr, in_list=None, n_agents=None):
        return tuple(t.cpu().numpy() for t in self.env.fear_matrix(positions, actions, mdr, in_list, n_agents))

    def feal(self, positions, actions, mdr, in_list=None, n_agents=None):
        return tuple(t.cpu().numpy() for t in self.env.feal(positions, actions, mdr, in_list, n_agents))


def bf16_to_f32(u16):
    return (u16.astype(np.uint32) << 16).view(np.float32)


def obs_f32(backend, arr):
    return bf16_to_f32(arr) if getattr(backend, "obs_bf16", False) else arr


# ----------------------------------------------------------------------------- operator level
def wall_scenario():
    """Level 3 with the walls / one-ways of tests/golden/wall_cases.npz (recorded from the reference run with tuple-typed
    paths, make_wall_golden.py), enforced."""
    from marl_responsible_nav_b200 import builtin_scenario
    g = npz("wall_cases.npz")
    return builtin_scenario("Level 3", walls=g["walls"].tolist(), oneways=g["oneways"].tolist())


def wide_scenario():
    """The 20 x 28, 7-agent scenario of tests/golden/wide_scenarios.json (the reference's JSON format) with its walls /
    one-ways enforced -- what make_wide_golden.py ran the reference on.  Does not fit the packed layout."""
    from marl_responsible_nav_b200.scenarios import load_scenario_json
    return load_scenario_json(os.path.join(GOLDEN, "wide_scenarios.json"), "Wide 20x28", walls="enforce")


def _prefixed(name, prefix):
    g = npz(name)
    return {k[len(prefix):]: g[k] for k in g.files if k.startswith(prefix)} if prefix else g


def check_update_cases(make_backend, fixture="update_cases.npz", prefix=""):
    g = _prefixed(fixture, prefix)
    b = make_backend(num_envs=1, fear=False)
    pos = np.where(g["locs"] < 0, 0, g["locs"]).astype(np.int8)
    new_pos, crash, restr, caught = b.update_world(pos, g["acts"], n_agents=g["n"], apples=g["apples"])
    assert np.array_equal(new_pos, g["out_locs"])
    assert np.array_equal(crash.astype(bool), g["crash"])
    assert np.array_equal(restr.astype(bool), g["restr"])
    assert np.array_equal(caught, g["caught"])
    return len(g["n"])


def check_fear_cases(make_backend, fixture="fear_cases.npz", prefix=""):
    g = _prefixed(fixture, prefix)
    b = make_backend(num_envs=1, fear=True)
    pos = np.where(g["locs"] < 0, 0, g["locs"]).astype(np.int8)
    resp, n_mdr, n_act, fsum = b.fear_one_actor(pos, g["acts"], g["mdr"], g["actor"], in_list=g["in_list"], n_agents=g["n"])
    assert np.array_equal(n_mdr, g["n_mdr"])
    assert np.array_equal(n_act, g["n_act"])
    assert np.array_equal(resp, g["resp"])            # fp64, bit-equal (north_star allows 1e-6 relative)
    if fsum is not None:
        assert np.array_equal(fsum, g["fear_sum"])
    return len(g["n"])


def check_matrix_cases(make_backend, fixture="matrix_cases.npz", prefix=""):
    """Responsibility.FeAR (all actors) and Responsibility.FeAL, full and partial action lists."""
    g = _prefixed(fixture, prefix)
    b = make_backend(num_envs=1, fear=True)
    pos = np.where(g["locs"] < 0, 0, g["locs"]).astype(np.int8)
    resp, n_mdr, n_act = b.fear_matrix(pos, g["acts"], g["mdr"], in_list=g["in_list"], n_agents=g["n"])
    assert np.array_equal(n_mdr, g["fear_n_mdr"]) and np.array_equal(n_act, g["fear_n_act"])
    assert np.array_equal(resp, g["fear"])                    # fp64 bit-equal
    feal, fm, fa = b.feal(pos, g["acts"], g["mdr"], in_list=g["in_list"], n_agents=g["n"])
    assert np.array_equal(fm, g["feal_n_mdr"]) and np.array_equal(fa, g["feal_n_act"])
    assert np.array_equal(feal, g["feal"])
    return len(g["n"])


# ----------------------------------------------------------------------------- multi-agent episodes
def check_ma_episodes(make_backend, obs_bf16=False, fixture="ma_episodes.npz"):
    """All golden episodes of one FeAR setting run side by side, one env per episode."""
    g = npz(fixture)
    total = 0
    H, W = g["obs"].shape[-2:]                      # 10 x 16 for the shipped scenarios; the general layout's fixtures differ
    NA = g["all_act"].shape[1]
    for fear in (False, True):
        eps = np.flatnonzero(g["ep_fear"] == fear)
        E, T = len(eps), int(g["ep_n_steps"][eps].max())
        b = make_backend(num_envs=E, fear=fear, auto_reset=False, max_steps=0, obs_bf16=obs_bf16)
        b.reset(spawn=g["ep_spawn"][eps])
        assert np.array_equal(obs_f32(b, b.obs).reshape(E, 2, H, W), g["ep_reset_obs"][eps])
        assert np.array_equal(b.action_mask, g["ep_reset_mask"][eps])
        first, nst = g["ep_first_step"][eps], g["ep_n_steps"][eps]
        for t in range(T):
            live = np.flatnonzero(nst > t)
            idx = first[live] + t
            la = np.zeros((E, 2), np.int8)
            npc = np.zeros((E, NA), np.int8)
            la[live], npc[live] = g["learner_act"][idx], g["all_act"][idx]
            b.step(la, npc_actions=npc)
            assert np.array_equal(b.positions[live], g["locs"][idx]), (fear, t)
            assert np.array_equal(b.reward[live], g["reward"][idx].astype(np.float32)), (fear, t)
            assert np.array_equal(b.terminated[live].astype(bool), g["term"][idx]), (fear, t)
            assert np.array_equal(b.truncated[live].astype(bool), g["trunc"][idx]), (fear, t)
            assert np.array_equal(b.fear[live], g["fear"][idx]), (fear, t)                   # fp64 bit-equal
            assert np.array_equal((b.info[live] >> 8) & 3, g["crash_count"][idx]), (fear, t)
            assert np.array_equal((b.info[live] >> 10) & 3, g["apples_caught"][idx]), (fear, t)
            assert np.array_equal(obs_f32(b, b.obs)[live].reshape(-1, 2, H, W), g["obs"][idx]), (fear, t)
            assert np.array_equal(b.action_mask[live], g["mask"][idx]), (fear, t)
            total += len(live)
    return total


def check_ma_sessions_autoreset(make_backend, fixture="ma_episodes.npz", max_steps=150):
    """One env per recorded session with auto-reset: episodes follow each other inside gw_step, the next
    episode's recorded spawn is offered at every step and consumed when the episode ends."""
    g = npz(fixture)
    H, W = g["obs"].shape[-2:]
    NA = g["all_act"].shape[1]
    sessions = np.unique(g["ep_session"])
    n_steps_total = 0
    for fear in (False, True):
        sess = [s for s in sessions if bool(g["ep_fear"][g["ep_session"] == s][0]) == fear]
        E = len(sess)
        eps_of = [np.flatnonzero(g["ep_session"] == s) for s in sess]
        b = make_backend(num_envs=E, fear=fear, auto_reset=True, max_steps=max_steps)
        b.reset(spawn=np.stack([g["ep_spawn"][e[0]] for e in eps_of]))
        cur_ep = [0] * E            # index into eps_of[i]
        cur_t = [0] * E
        done = [False] * E
        while not all(done):
            la = np.zeros((E, 2), np.int8); npc = np.zeros((E, NA), np.int8); spawn = np.zeros((E, NA, 2), np.int8)
            for i in range(E):
                if done[i]:
                    spawn[i] = g["ep_spawn"][eps_of[i][0]]
                    continue
                ep = eps_of[i][cur_ep[i]]
                s = g["ep_first_step"][ep] + cur_t[i]
                la[i], npc[i] = g["learner_act"][s], g["all_act"][s]
                nxt = eps_of[i][cur_ep[i] + 1] if cur_ep[i] + 1 < len(eps_of[i]) else eps_of[i][0]
                spawn[i] = g["ep_spawn"][nxt]
            b.step(la, npc_actions=npc, spawn=spawn)
            for i in range(E):
                if done[i]:
                    continue
                ep = eps_of[i][cur_ep[i]]
                s = g["ep_first_step"][ep] + cur_t[i]
                last = cur_t[i] + 1 == g["ep_n_steps"][ep]
                assert np.array_equal(b.positions[i], g["locs"][s])
                assert np.array_equal(b.reward[i], g["reward"][s].astype(np.float32))
                assert np.array_equal(b.fear[i], g["fear"][s])
                assert bool(b.ended[i]) == bool(last), (i, cur_ep[i], cur_t[i])
                if last:
                    assert np.array_equal(b.final_obs[i].reshape(2, H, W), g["obs"][s])
                    nxt = eps_of[i][cur_ep[i] + 1] if cur_ep[i] + 1 < len(eps_of[i]) else eps_of[i][0]
                    assert np.array_equal(b.obs[i].reshape(2, H, W), g["ep_reset_obs"][nxt])
                    assert np.array_equal(b.action_mask[i], g["ep_reset_mask"][nxt])
                    cur_ep[i] += 1
                    cur_t[i] = 0
                    if cur_ep[i] >= len(eps_of[i]):
                        done[i] = True
                else:
                    assert np.array_equal(b.obs[i].reshape(2, H, W), g["obs"][s])
                    assert np.array_equal(b.action_mask[i], g["mask"][s])
                    cur_t[i] += 1
                n_steps_total += 1
        st = b.stats()
        want_eps = sum(len(e) for e in eps_of)
        assert st["episodes"] >= want_eps            # envs that finished early keep cycling their first episode
    return n_steps_total


# ----------------------------------------------------------------------------- single-agent episodes
def check_single_episodes(make_backend, fixture="single_episodes.npz"):
    g = npz(fixture)
    H, W = g["obs"].shape[-2:]
    NA = g["all_act"].shape[1]
    total = 0
    for fear in (False, True):
        eps = np.flatnonzero(g["ep_fear"] == fear)
        E, T = len(eps), int(g["ep_n_steps"][eps].max())
        b = make_backend(num_envs=E, fear=fear, env_kind="single", auto_reset=False, max_steps=0)
        b.reset(spawn=g["ep_spawn"][eps])
        assert np.array_equal(b.obs.reshape(E, H, W), g["ep_reset_obs"][eps])
        first, nst = g["ep_first_step"][eps], g["ep_n_steps"][eps]
        for t in range(T):
            live = np.flatnonzero(nst > t)
            idx = first[live] + t
            la = np.zeros((E, 1), np.int8); npc = np.zeros((E, NA), np.int8)
            la[live, 0], npc[live] = g["action"][idx], g["all_act"][idx]
            b.step(la, npc_actions=npc)
            assert np.array_equal(b.positions[live], g["locs"][idx]), (fear, t)
            assert np.array_equal(b.reward[live, 0], g["reward"][idx].astype(np.float32)), (fear, t)
            assert np.array_equal(b.terminated[live, 0].astype(bool), g["term"][idx]), (fear, t)
            assert np.array_equal(b.truncated[live, 0].astype(bool), g["trunc"][idx]), (fear, t)
            assert np.array_equal(b.fear[live, 0], g["fear"][idx]), (fear, t)
            assert np.array_equal(((b.info[live] >> 4) & 1).astype(bool), g["restricted"][idx]), (fear, t)
            assert np.array_equal(b.obs[live].reshape(-1, H, W), g["obs"][idx]), (fear, t)
            total += len(live)
    return total
