"""B200-native batched grid-world step / FeAR / observation path.

Drop-in for the hot path of Henweiz/MARL-Responsible-Nav (`custom/ma_customenv.py`,
`custom/customenv.py`, `custom/grid_world.py`, `custom/Responsibility.py`): the same
`reset` / `step` env API, executed by hand-written sm_100a CUDA kernels behind a thin
C-ABI (`include/gridworld_b200.h`, `csrc/libgridworld_b200.so`).

There is no CPU implementation in this package: constructing an environment without
the compiled library or without a CUDA device raises.
"""
from .scenarios import Scenario, builtin_scenario, load_scenario_json, policy_probs  # noqa: F401
from .batched import BatchedGridWorld, GeneralGridWorld, StepOutput  # noqa: F401
from .envs import CustomMAEnv, CustomEnv  # noqa: F401
from .replay import ReplayRing  # noqa: F401
from .actor import FusedActor  # noqa: F401
from . import sharding  # noqa: F401
from . import _native  # noqa: F401

__all__ = ["BatchedGridWorld", "GeneralGridWorld", "StepOutput", "CustomMAEnv", "CustomEnv", "Scenario", "builtin_scenario",
           "load_scenario_json", "policy_probs", "ReplayRing", "FusedActor", "sharding"]
__version__ = "0.1.0"
