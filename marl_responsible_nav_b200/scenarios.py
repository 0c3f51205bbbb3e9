"""Scenario tables (the reference keeps them in custom/Scenarios.json).

`builtin_scenario(name)` rebuilds the three shipped layouts ("Level 3", "Level 5", "GameMap")
from their structure -- one 10x16 map made of an outer ring road and an inner ring road
joined by four spokes, with clockwise traffic -- and `load_scenario_json(path, name)` reads
any file in the reference's own JSON format (custom/grid_world.py:621-674), so a user can
point the environment at the reference's Scenarios.json unchanged.  Both produce the same
`Scenario`; tests/ pin them against tables recorded from the reference.
"""
from __future__ import annotations

import json
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

# action ids, custom/custom_agent.py:140-150
STAY, UP1, DOWN1, LEFT1, RIGHT1, UP2, DOWN2, LEFT2, RIGHT2 = range(9)
# directionWeights order is (Up, Down, Left, Right), custom/custom_agent.py:181-192
_UP, _DOWN, _LEFT, _RIGHT = [1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]


def policy_probs(step_weights: Sequence[float], direction_weights: Optional[Sequence[float]]) -> np.ndarray:
    """GeneratePolicy (custom/custom_agent.py:181-197): p[9] over (Stay, 4 x one step, 4 x two steps).
    `direction_weights=None` is what the 25 % perturbation produces (random.shuffle returns None,
    custom/ma_customenv.py:441-443) and means uniform directions."""
    dw = [1, 1, 1, 1] if direction_weights is None else list(direction_weights)
    p = [step_weights[0]]
    for sw in step_weights[1:]:
        p = p + [sw * x for x in dw]
    p = np.array(p)
    return p / p.sum()


@dataclass
class Scenario:
    name: str
    region: np.ndarray                       # int8 [H, W], 1 = active
    n_agents: int
    policy_index: np.ndarray                 # uint8 [H, W] -> row of `policies`
    policies: List[Tuple[List[float], List[float]]]   # (stepWeights[3], directionWeights[4])
    mdr_action: np.ndarray                   # uint8 [H, W] Move-de-Rigueur action id per cell
    policy_keys: List[int] = field(default_factory=list)   # original JSON keys, for reporting
    blocked: List[Tuple[Tuple[int, int], Tuple[int, int]]] = field(default_factory=list)   # restricted paths (from, to), see restricted_paths

    def blocked_set(self):
        return set(self.blocked) if self.blocked else None

    @property
    def shape(self) -> Tuple[int, int]:
        return tuple(self.region.shape)

    def active_cells(self) -> List[Tuple[int, int]]:
        xs, ys = np.where(self.region == 1)              # row-major, custom/ma_customenv.py:373
        return [(int(x), int(y)) for x, y in zip(xs, ys)]

    def map_rows(self) -> List[int]:
        return [int(sum(1 << c for c in range(self.region.shape[1]) if self.region[r, c] == 1))
                for r in range(self.region.shape[0])]

    def npc_policy(self, cell, perturbed: bool) -> np.ndarray:
        sw, dw = self.policies[int(self.policy_index[cell[0], cell[1]])]
        return policy_probs(sw, None if perturbed else dw)


def _paint(shape, regions):
    """Later regions overwrite earlier ones, as the reference's slice assignments do."""
    m = np.zeros(shape, dtype=np.int64)
    for value, (x0, x1), (y0, y1) in regions:
        m[x0:x1, y0:y1] = value
    return m


def _ring_region(h=10, w=16, inner_rows=(2, 7), inner_cols=(5, 10)) -> np.ndarray:
    region = np.zeros((h, w), dtype=np.int8)
    region[0, :] = region[h - 1, :] = 1                  # outer ring, horizontal roads
    region[:, 0] = region[:, w - 1] = 1                  # outer ring, vertical roads
    region[:, inner_cols[0]] = region[:, inner_cols[1]] = 1      # spokes / inner ring verticals
    region[inner_rows[0], inner_cols[0]:inner_cols[1] + 1] = 1   # inner ring horizontals
    region[inner_rows[1], inner_cols[0]:inner_cols[1] + 1] = 1
    return region


def _ring_scenario(name: str, n_agents: int, outer_sw, inner_sw, corner_sw, rest_sw) -> Scenario:
    h, w = 10, 16
    region = _ring_region(h, w)
    # (key, rows, cols, stepWeights, directionWeights): outer ring clockwise, inner ring clockwise, corners turn
    zones = [
        (0, (0, h), (0, w), rest_sw, [1, 1, 1, 1]),
        (1, (0, 1), (0, w - 1), outer_sw, _RIGHT),
        (2, (h - 1, h), (1, w), outer_sw, _LEFT),
        (3, (1, h), (0, 1), outer_sw, _UP),
        (4, (0, h - 1), (w - 1, w), outer_sw, _DOWN),
        (5, (7, 8), (5, 11), inner_sw, _RIGHT),
        (6, (2, 3), (6, 11), inner_sw, _LEFT),
        (7, (3, 8), (10, 11), inner_sw, _UP),
        (8, (2, 7), (5, 6), inner_sw, _DOWN),
        (9, (0, 1), (0, 1), corner_sw, _RIGHT),
        (10, (h - 1, h), (w - 1, w), corner_sw, _LEFT),
        (11, (h - 1, h), (0, 1), corner_sw, _UP),
        (12, (0, 1), (w - 1, w), corner_sw, _DOWN),
    ]
    policy_index = _paint((h, w), [(k, xs, ys) for k, xs, ys, _, _ in zones]).astype(np.uint8)
    policies = [(list(sw), list(dw)) for _, _, _, sw, dw in zones]
    mdr = _paint((h, w), [(STAY, (0, h), (0, w)), (RIGHT1, (0, 1), (0, w - 1)), (LEFT1, (h - 1, h), (1, w)),
                          (UP1, (1, h), (0, 1)), (DOWN1, (0, h - 1), (w - 1, w))]).astype(np.uint8)
    return Scenario(name=name, region=region, n_agents=n_agents, policy_index=policy_index, policies=policies,
                    mdr_action=mdr, policy_keys=[z[0] for z in zones])


def restricted_paths(shape, walls=(), oneways=()) -> List[Tuple[Tuple[int, int], Tuple[int, int]]]:
    """GWorld.RestrictedPaths (custom/grid_world.py:32-86) as (from, to) cell pairs: a wall [a, b] blocks a -> b and b -> a,
    a one-way [a, b] blocks b -> a; entries whose cells are not inside the grid or not 4-neighbours are dropped, as there.
    In the reference the paths only take effect when their cells are TUPLES (`[old, new] in RestrictedPaths`, :498, with
    tuple locations); the JSON loader hands over lists, which never match (the tuple conversion is commented out, :654-665).
    This function gives the intended, tuple-typed semantics."""
    h, w = shape
    out = []

    def ok(a, b):
        (x1, y1), (x2, y2) = a, b
        if not (x1 < h and x2 < h and y1 < w and y2 < w) or min(x1, y1, x2, y2) < 0:
            return False
        return (x1 == x2 and abs(y1 - y2) == 1) or (y1 == y2 and abs(x1 - x2) == 1)

    for wall in walls:
        a, b = (int(wall[0][0]), int(wall[0][1])), (int(wall[1][0]), int(wall[1][1]))
        if ok(a, b):
            out += [(a, b), (b, a)]
    for way in oneways:
        a, b = (int(way[0][0]), int(way[0][1])), (int(way[1][0]), int(way[1][1]))
        if ok(a, b):
            out.append((b, a))
    return out


def builtin_scenario(name: str = "Level 3", n_agents: Optional[int] = None, walls=(), oneways=()) -> Scenario:
    """The three layouts of custom/Scenarios.json.  `n_agents` overrides the scenario's N_Agents; `walls` / `oneways`
    (lists of [cell, cell]) add restricted paths with the semantics of `restricted_paths`."""
    if name == "Level 3":        # 4 agents; outer ring at speed 2, inner ring at speed 1, elsewhere stay-or-one-step
        sc = _ring_scenario(name, 4, outer_sw=[0, 0, 1], inner_sw=[0, 1, 0], corner_sw=[0, 0, 1], rest_sw=[1, 1, 0])
    elif name == "Level 5":      # 3 agents; ring traffic may also stay or move one cell
        sc = _ring_scenario(name, 3, outer_sw=[1, 1, 1], inner_sw=[1, 1, 1], corner_sw=[0, 0, 1], rest_sw=[1, 1, 0])
    elif name == "GameMap":
        # The reference's GameMap has no 'MdRs' / 'Policies' (its envs raise KeyError on it).
        # Extension: MdR = Stay everywhere, one uniform policy from the scenario-level weights.
        region = _ring_region()
        sc = Scenario(name=name, region=region, n_agents=3, policy_index=np.zeros(region.shape, np.uint8),
                      policies=[([1, 1, 1], [1, 1, 1, 1])], mdr_action=np.zeros(region.shape, np.uint8),
                      policy_keys=[0])
    else:
        raise KeyError(f"unknown built-in scenario {name!r} (have 'Level 3', 'Level 5', 'GameMap')")
    if n_agents is not None:
        sc.n_agents = int(n_agents)
    sc.blocked = restricted_paths(sc.region.shape, walls, oneways)
    return sc


def _as_slice(arg) -> slice:
    return slice(arg[0], arg[1], None if arg[2] == 0 else arg[2])       # custom/grid_world.py:633-639


def load_scenario_json(path: str, name: str = "Level 3", n_agents: Optional[int] = None, walls: str = "refuse") -> Scenario:
    """Read one scenario from a file in the reference's Scenarios.json format.
    walls: what to do with non-empty Map.Walls / Map.OneWays -- "refuse" (default: raise, the caller must choose),
    "inert" (the reference's actual behaviour: JSON lists never equal its tuple paths, custom/grid_world.py:498, so they
    restrict nothing) or "enforce" (the intended semantics, see `restricted_paths`)."""
    with open(path) as f:
        sc = json.load(f)[name]
    region = (np.array(sc["Map"]["Region"]) == 1).astype(np.int8)
    blocked = []
    if sc["Map"].get("Walls") or sc["Map"].get("OneWays"):
        if walls == "enforce":
            blocked = restricted_paths(region.shape, sc["Map"].get("Walls") or (), sc["Map"].get("OneWays") or ())
        elif walls != "inert":
            raise ValueError("the scenario has Walls / OneWays: pass walls='inert' (what the reference does with a JSON file: "
                             "nothing, its list paths never match) or walls='enforce' (the intended restricted paths)")
    if sc.get("AgentLocations"):
        raise ValueError("fixed AgentLocations are not supported; all shipped scenarios spawn randomly")
    policy_index = np.zeros(region.shape, dtype=np.uint8)
    policies: List[Tuple[List[float], List[float]]] = []
    keys: List[int] = []
    if "Policies" in sc:
        key_map = np.zeros(region.shape, dtype=np.int64)
        for key, pol in sc["Policies"].items():                         # JSON order, later overwrite earlier
            key_map[_as_slice(pol["slicex"]), _as_slice(pol["slicey"])] = int(key)
            keys.append(int(key))
            policies.append((list(pol["stepWeights"]), list(pol["directionWeights"])))
        if 0 not in keys:                                               # cells not covered keep key 0
            raise ValueError("scenario has no policy '00' covering the default region")
        lut = {k: i for i, k in enumerate(keys)}
        policy_index = np.vectorize(lambda k: lut[int(k)])(key_map).astype(np.uint8)
    else:
        keys = [0]
        policies = [(list(sc.get("StepWeights", [1, 1, 1])), list(sc.get("DirectionWeights", [1, 1, 1, 1])))]
    mdr_action = np.zeros(region.shape, dtype=np.uint8)
    if "MdRs" in sc:
        key_map = np.zeros(region.shape, dtype=np.int64)
        mdr_of = {0: 0}
        for key, m in sc["MdRs"].items():
            key_map[_as_slice(m["slicex"]), _as_slice(m["slicey"])] = int(key)
            mdr_of[int(key)] = int(m["mdr"])
        mdr_action = np.vectorize(lambda k: mdr_of[int(k)])(key_map).astype(np.uint8)
    return Scenario(name=name, region=region, n_agents=int(n_agents if n_agents is not None else sc["N_Agents"]),
                    policy_index=policy_index, policies=policies, mdr_action=mdr_action, policy_keys=keys, blocked=blocked)
