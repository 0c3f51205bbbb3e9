"""Import the reference's hot-path modules from /root/reference (build container only).

The reference needs three packages that are not installed here (pettingzoo,
gymnasium, pygame); they only provide base classes / space descriptors, so tiny
in-memory stand-ins are enough.  Nothing is copied: the
modules are imported from where they lie, with the working directory switched to
the reference root because `LoadJsonScenario` opens 'custom/Scenarios.json'
relative to cwd (custom/grid_world.py:621).

Used by `make_golden.py` (fixture generator) and by the optional live-fuzz test;
never by the product package and never on the GPU box (no /root/reference there).
"""
import contextlib
import io
import os
import sys
import types

REF_ROOT = os.environ.get("GW_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "custom", "grid_world.py"))


def _install_stubs():
    if "pettingzoo" not in sys.modules:
        pz = types.ModuleType("pettingzoo")

        class ParallelEnv:                       # only `num_agents` is used (util.py:25)
            @property
            def num_agents(self):
                return len(self.agents)

        class AECEnv:
            pass

        pz.ParallelEnv, pz.AECEnv = ParallelEnv, AECEnv
        pzu = types.ModuleType("pettingzoo.utils")
        pzu.agent_selector = object
        pz.utils = pzu
        sys.modules["pettingzoo"], sys.modules["pettingzoo.utils"] = pz, pzu
    if "gymnasium" not in sys.modules:
        gym = types.ModuleType("gymnasium")

        class Env:
            pass

        gym.Env = Env
        sp = types.ModuleType("gymnasium.spaces")

        class Discrete:
            def __init__(self, n):
                self.n = n

        class Box:
            def __init__(self, low, high, shape, dtype):
                self.low, self.high, self.shape, self.dtype = low, high, shape, dtype

        sp.Discrete, sp.Box = Discrete, Box
        gym.spaces = sp
        sys.modules["gymnasium"], sys.modules["gymnasium.spaces"] = gym, sp
    if "pygame" not in sys.modules:
        sys.modules["pygame"] = types.ModuleType("pygame")


def load_reference():
    """Returns a namespace with the reference modules (grid_world, custom_agent,
    Responsibility, ma_customenv, customenv)."""
    if not reference_available():
        raise RuntimeError(f"reference not found at {REF_ROOT}")
    sys.dont_write_bytecode = True          # the reference tree is read-only
    _install_stubs()
    cwd = os.getcwd()
    os.chdir(REF_ROOT)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    try:
        with contextlib.redirect_stdout(io.StringIO()):      # modules print 'N_Agents : 4' on import
            from custom import grid_world, custom_agent, Responsibility, ma_customenv, customenv
    finally:
        os.chdir(cwd)
    return types.SimpleNamespace(grid_world=grid_world, custom_agent=custom_agent,
                                 Responsibility=Responsibility, ma_customenv=ma_customenv,
                                 customenv=customenv, root=REF_ROOT)
