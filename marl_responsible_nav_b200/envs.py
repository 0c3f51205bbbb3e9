"""Single-environment classes with the reference's env API, executed on the GPU.

`CustomMAEnv` mirrors custom/ma_customenv.py:71-506 (PettingZoo-parallel style: dicts keyed by
"agent_k"), `CustomEnv` mirrors custom/customenv.py:45-401 (gymnasium style, one learner).  They are
E = 1 front-ends of BatchedGridWorld so that main_custom.py / maddpg/agent.py / customeval.py can
use them unchanged.  The host part mirrors the reference's RNG consumption exactly -- the spawn draw
on `default_rng(seed)` (ma_customenv.py:97,376-377), one `random.random()` (+ `random.shuffle`) per
agent and one legacy `np.random.choice` per agent per step (:441-443, custom_agent.py:31) -- and hands
the drawn spawn cells / NPC actions to the kernels, so under the same seeds the trajectories are
identical to the reference's.  World dynamics, FeAR, rewards, observations and masks all come from
the CUDA library; nothing is computed on the CPU, and construction fails without a GPU.
"""
from __future__ import annotations

import random
from typing import Dict, Optional, Union

import numpy as np
import torch

from .batched import BatchedGridWorld
from .scenarios import Scenario, builtin_scenario

N_DISCRETE_ACTIONS = 9          # custom/ma_customenv.py:18
N_INTELLIGENT_AGENTS = 2        # custom/ma_customenv.py:19

try:                                                     # optional: real space classes when installed
    from gymnasium.spaces import Box, Discrete           # type: ignore
except Exception:                                        # pragma: no cover - not installed in the build image
    class Discrete:                                      # the callers only read `.n` (util.py:28)
        def __init__(self, n):
            self.n = int(n)

        def __repr__(self):
            return f"Discrete({self.n})"

    class Box:
        def __init__(self, low, high, shape, dtype):
            self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype

        def __repr__(self):
            return f"Box({self.low}, {self.high}, {self.shape}, {self.dtype})"


class _HostMirror:
    """RNG-order mirror of setup_env / setup_step (host side, no dynamics)."""

    def __init__(self, scenario: Scenario):
        self.sc = scenario
        self.cells = scenario.active_cells()

    def spawn(self, rng: np.random.Generator, n_agents: int):
        idx = rng.choice(len(self.cells), size=n_agents, replace=False, shuffle=False)   # ma_customenv.py:376
        idx.sort()                                                                        # :377
        return [self.cells[int(i)] for i in idx]

    def npc_actions(self, locs):
        pols = []
        for loc in locs:                                       # ma_customenv.py:435-450
            perturbed = random.random() < 0.25
            if perturbed:
                random.shuffle([0, 0, 0, 1])                   # result unused (None) but the stream advances
            pols.append(self.sc.npc_policy(loc, perturbed))
        return [int(np.random.choice(np.arange(N_DISCRETE_ACTIONS), p=p)) for p in pols]   # custom_agent.py:31


class CustomMAEnv:
    """custom/ma_customenv.py::CustomMAEnv on the GPU (2 learners + NPCs, 'Level 3' by default)."""

    metadata = {"name": "gridworld_b200_ma"}

    def __init__(self, render: bool = False, fear: bool = True, seed: Optional[int] = None,
                 scenario: Union[str, Scenario] = "Level 3", device="cuda"):
        if render:
            raise NotImplementedError("pygame rendering is not part of the accelerated path")
        self.possible_agents = ["agent_" + str(r) for r in range(N_INTELLIGENT_AGENTS)]
        self.agents = self.possible_agents[:]
        self.action_space = Discrete(N_DISCRETE_ACTIONS)       # instance attribute, as in the reference (:90)
        self.rendering = render
        self.fear = fear
        self.rng = np.random.default_rng(seed=seed)            # :97
        self.scenario = builtin_scenario(scenario) if isinstance(scenario, str) else scenario
        self._mirror = _HostMirror(self.scenario)
        self._gw = BatchedGridWorld(self.scenario, num_envs=1, device=device, env_kind="multi", fear=fear,
                                    n_learners=N_INTELLIGENT_AGENTS, max_steps=0, auto_reset=False)
        self._io = self._gw.pinned_io()                       # inputs / outputs in pinned host memory, written by the kernels
        self._io.replay = True                                 # spawn cells and NPC draws come from the host RNG mirror
        self._locs = None
        self.num_moves = 0

    @property
    def num_agents(self) -> int:                               # pettingzoo.ParallelEnv.num_agents (util.py:25)
        return len(self.agents)

    @property
    def max_num_agents(self) -> int:
        return len(self.possible_agents)

    def observation_space(self, agent):
        return Box(low=-1.0, high=16.0, shape=self.scenario.shape, dtype=np.float64)   # :115-116

    def render(self):
        raise NotImplementedError("pygame rendering is not part of the accelerated path")

    def close(self):
        pass

    def _obs_dict(self) -> Dict[str, np.ndarray]:
        o = self._io.obs[0].astype(np.float64).reshape(len(self.possible_agents), *self.scenario.shape)
        return {a: o[k] for k, a in enumerate(self.possible_agents)}

    def _mask_dict(self):
        m = self._io.action_mask[0].copy()
        return {a: {"action_mask": m[k]} for k, a in enumerate(self.possible_agents)}

    def reset(self, seed=None, options=None):
        # `seed` / `options` are ignored, exactly like the reference (:169,:183)
        spawn = self._mirror.spawn(self.rng, self._gw.n_agents)
        self._io.spawn[0] = spawn
        self._io.reset()
        self._locs = spawn
        self.agents = self.possible_agents[:]
        self.num_moves = 0
        self.terminations = {a: False for a in self.agents}
        self.truncation = {a: False for a in self.agents}
        self.observations = self._obs_dict()
        info = {"fear": 0.0}
        info.update(self._mask_dict())
        return self.observations, info

    def step(self, actions):
        if actions is None or len(actions) == 0:               # :229-231
            return {}, {}, {}, {}, {}
        if self._locs is None:
            raise RuntimeError("step() called before reset()")
        npc = self._mirror.npc_actions(self._locs)             # setup_step (:233), all agents draw
        self.num_moves += 1
        io = self._io
        for k in range(len(self.possible_agents)):
            io.actions[0, k] = int(np.asarray(actions[k]).item())
        io.npc_actions[0] = npc
        io.step()                                              # one library call; the arrays below are host memory
        self._locs = [(int(r), int(c)) for r, c in io.positions[0]]
        rew, term, trunc, fear = io.reward[0], io.terminated[0], io.truncated[0], io.fear[0]
        info_bits = int(io.info[0])
        self.observations = self._obs_dict()
        self.rewards = {a: int(rew[k]) for k, a in enumerate(self.agents)}
        self.terminations = {a: bool(term[k]) for k, a in enumerate(self.agents)}
        self.truncation = {a: bool(trunc[k]) for k, a in enumerate(self.agents)}
        info = {"fear": {a: np.float64(fear[k]) if self.fear else 0.0 for k, a in enumerate(self.agents)},
                "agent_crashes": (info_bits >> 8) & 3, "apples_caught": (info_bits >> 10) & 3}
        info.update(self._mask_dict())
        return self.observations, self.rewards, self.terminations, self.truncation, info

    def get_action_mask(self, agent):                          # :467-506 (mask of the current position)
        k = int(agent[-1])
        return {"action_mask": self._io.action_mask[0, k].copy()}


class CustomEnv:
    """custom/customenv.py::CustomEnv on the GPU (learner = agent 0, three NPCs, apple at (9,15))."""

    # module-level generator in the reference (customenv.py:18); exposed so callers/tests can seed it
    rng = np.random.default_rng()

    def __init__(self, render: bool = False, fear: bool = True, scenario: Union[str, Scenario] = "Level 3",
                 device="cuda"):
        if render:
            raise NotImplementedError("pygame rendering is not part of the accelerated path")
        self.scenario = builtin_scenario(scenario) if isinstance(scenario, str) else scenario
        self.action_space = Discrete(N_DISCRETE_ACTIONS)
        self.observation_space = Box(low=-1.0, high=16.0, shape=self.scenario.shape, dtype=np.float64)
        self.num_agents = 1                                    # :57
        self.fear = fear
        self.rendering = render
        self._mirror = _HostMirror(self.scenario)
        self._gw = BatchedGridWorld(self.scenario, num_envs=1, device=device, env_kind="single", fear=fear,
                                    n_learners=1, max_steps=0, auto_reset=False)
        self._io = self._gw.pinned_io()
        self._io.replay = True
        self._locs = None
        self._apple_left = False

    def reset(self, seed=None, options=None):
        spawn = self._mirror.spawn(type(self).rng, self._gw.n_agents)      # :235-236
        self._io.spawn[0] = spawn
        self._io.reset()
        self._locs = spawn
        self._apple_left = True
        self.episode_reward = 0
        self.episode_length = 0
        self.observation = self._io.obs[0, 0].astype(np.float64).reshape(self.scenario.shape)
        return self.observation, {}

    def step(self, action):
        if self._locs is None:
            raise RuntimeError("step() called before reset()")
        npc = self._mirror.npc_actions(self._locs)                         # :83-103
        if not self._apple_left:
            raise StopIteration                                            # next(iter({})) in the reference (:132)
        io = self._io
        io.actions[0, 0] = int(np.asarray(action[0]).item())
        io.npc_actions[0] = npc
        io.step()
        self._locs = [(int(r), int(c)) for r, c in io.positions[0]]
        bits = int(io.info[0])
        terminated = bool(io.terminated[0, 0])
        truncated = bool(io.truncated[0, 0])
        # the reference's reward is a Python int unless the +0.1 shaping fired (:126-158); rebuild it from the
        # kernel's flags so that the value is bit-identical in fp64 (the fp32 tensor is the batched output)
        reward = 0
        if terminated:
            reward -= 10
        if (bits >> 10) & 3:
            reward += 20
            self._apple_left = False
        if (bits >> 14) & 1:
            reward += 0.1
        self.episode_reward += reward
        self.episode_length += 1
        self.observation = io.obs[0, 0].astype(np.float64).reshape(self.scenario.shape)
        info = {"episode": {"r": self.episode_reward, "l": self.episode_length},
                "restricted": bool((bits >> 4) & 1),
                "fear": np.float64(io.fear[0, 0])}
        return self.observation, [reward], [terminated], truncated, info

    def render(self, mode="human"):
        raise NotImplementedError("pygame rendering is not part of the accelerated path")

    def close(self):
        pass
