"""numpy front-end of oracle/_build/libgw_oracle.so (TEST INFRASTRUCTURE ONLY, see gw_oracle.c).

Mirrors the tensor API of marl_responsible_nav_b200.BatchedGridWorld with host numpy arrays so the
parity tests can feed both sides the same inputs.  Imports only the POD struct definitions
(`GwConfig`, `GwIO`, `GwStats`, `build_config`) from the product package; the product never imports
this file.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from marl_responsible_nav_b200 import _native as N
from marl_responsible_nav_b200.scenarios import builtin_scenario

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "libgw_oracle.so")
LIB_PATH_WIDE = os.path.join(HERE, "_build", "libgw_oracle_wide.so")     # the same source built on gww_config (general layout)
_libs = {}


def build(force=False):
    src, hdr = os.path.join(HERE, "gw_oracle.c"), os.path.join(HERE, "..", "include", "gridworld_b200.h")
    stale = any(not os.path.exists(lp) or os.path.getmtime(lp) < max(os.path.getmtime(src), os.path.getmtime(hdr))
                for lp in (LIB_PATH, LIB_PATH_WIDE))
    if force or stale:
        res = subprocess.run(["make", "-C", HERE] + (["-B"] if force else []), capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("building the C oracle failed:\n" + res.stdout + res.stderr)
    return LIB_PATH


def load(wide=False):
    if wide not in _libs:
        build()
        lib = C.CDLL(LIB_PATH_WIDE if wide else LIB_PATH)
        vp, i64 = C.c_void_p, C.c_int64
        lib.gwo_create.argtypes = [C.POINTER(N.GwwConfig if wide else N.GwConfig), C.POINTER(vp)]
        lib.gwo_destroy.argtypes = [vp]
        lib.gwo_reset.argtypes = [vp, vp, C.POINTER(N.GwIO)]
        lib.gwo_step.argtypes = [vp, C.POINTER(N.GwIO), C.c_int]
        lib.gwo_get_stats.argtypes = [vp, C.POINTER(N.GwStats)]
        lib.gwo_get_state.argtypes = [vp, vp]
        lib.gwo_update_world.argtypes = [vp, i64] + [vp] * 8
        lib.gwo_fear_one_actor.argtypes = [vp, i64] + [vp] * 10
        lib.gwo_fear_matrix.argtypes = [vp, i64] + [vp] * 8
        lib.gwo_feal.argtypes = [vp, i64] + [vp] * 8
        _libs[wide] = lib
    return _libs[wide]


def _p(a):
    return C.c_void_p(a.ctypes.data) if a is not None else None


def _i8(x):
    return None if x is None else np.ascontiguousarray(np.asarray(x), dtype=np.int8)


class COracle:
    def __init__(self, scenario="Level 3", num_envs=1, threads=1, layout=None, **kw):
        """layout: "packed" (gw_config), "wide" (gww_config: the general layout) or None = packed when the scenario fits."""
        self.scenario = builtin_scenario(scenario) if isinstance(scenario, str) else scenario
        self.obs_bf16 = bool(kw.pop("obs_bf16", False))
        if layout is None:
            layout = "packed" if N.fits_packed_layout(self.scenario, kw.get("n_agents")) else "wide"
        self.wide = layout == "wide"
        self.PAD = N.GWW_MAX_AGENTS if self.wide else N.GW_MAX_AGENTS      # per-agent dimension of the operator-level arrays
        self.lib = load(self.wide)
        self.cfg = (N.build_wide_config if self.wide else N.build_config)(self.scenario, num_envs=num_envs, obs_bf16=self.obs_bf16, **kw)
        self.threads = threads
        h = C.c_void_p()
        rc = self.lib.gwo_create(C.byref(self.cfg), C.byref(h))
        if rc != 0:
            raise RuntimeError(f"gwo_create failed ({rc})")
        self._h = h
        E, L, A = self.cfg.num_envs, self.cfg.n_learners, self.cfg.n_agents
        self.E, self.L, self.A = E, L, A
        self.obs_len = self.cfg.height * self.cfg.width
        odt = np.uint16 if self.obs_bf16 else np.float32
        self.obs = np.zeros((E, L, self.obs_len), odt)
        self.final_obs = np.zeros((E, L, self.obs_len), odt)
        self.action_mask = np.zeros((E, L, 9), np.int8)
        self.positions = np.zeros((E, A, 2), np.int8)
        self.reward = np.zeros((E, L), np.float32)
        self.shaped_reward = np.zeros((E, L), np.float32)
        self.fear = np.zeros((E, L), np.float64)
        self.terminated = np.zeros((E, L), np.uint8)
        self.truncated = np.zeros((E, L), np.uint8)
        self.ended = np.zeros((E,), np.uint8)
        self.info = np.zeros((E,), np.uint32)

    def __del__(self):
        if getattr(self, "_h", None):
            self.lib.gwo_destroy(self._h)
            self._h = None

    def _io(self, actions=None, npc=None, spawn=None):
        io = N.GwIO()
        self._keep = (_i8(actions), _i8(npc), _i8(spawn))
        io.learner_actions, io.npc_actions, io.spawn = (a.ctypes.data if a is not None else None for a in self._keep)
        for name in ("obs", "final_obs", "reward", "shaped_reward", "fear", "terminated", "truncated", "ended",
                     "action_mask", "positions", "info"):
            setattr(io, name, getattr(self, name).ctypes.data)
        return io

    def reset(self, mask=None, spawn=None):
        io = self._io(spawn=spawn)
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        rc = self.lib.gwo_reset(self._h, _p(m), C.byref(io))
        assert rc == 0, rc
        return self

    def step(self, actions, npc_actions=None, spawn=None):
        io = self._io(actions, npc_actions, spawn)
        rc = self.lib.gwo_step(self._h, C.byref(io), self.threads)
        assert rc == 0, rc
        return self

    def stats(self):
        s = N.GwStats()
        self.lib.gwo_get_stats(self._h, C.byref(s))
        return {name: getattr(s, name) for name, _ in N.GwStats._fields_}

    def state(self):
        out = np.zeros((self.E, 16 if self.wide else 4), np.uint32)
        self.lib.gwo_get_state(self._h, _p(out))
        return out

    def update_world(self, positions, actions, n_agents=None, apples=None):
        pos, act, nper, app = _i8(positions), _i8(actions), _i8(n_agents), _i8(apples)
        Cn = pos.shape[0]
        new_pos = np.zeros((Cn, self.PAD, 2), np.int8)
        crash = np.zeros((Cn, self.PAD), np.uint8)
        restr = np.zeros((Cn, self.PAD), np.uint8)
        caught = np.zeros((Cn, 2, 2), np.int8)
        rc = self.lib.gwo_update_world(self._h, Cn, _p(nper), _p(pos), _p(act), _p(app), _p(new_pos), _p(crash),
                                       _p(restr), _p(caught))
        assert rc == 0
        return new_pos, crash, restr, caught

    def fear_one_actor(self, positions, actions, mdr, actor, in_list=None, n_agents=None):
        pos, act, md, ac, nper = _i8(positions), _i8(actions), _i8(mdr), _i8(actor), _i8(n_agents)
        il = None if in_list is None else np.ascontiguousarray(in_list, dtype=np.uint8)
        Cn = pos.shape[0]
        resp = np.zeros((Cn, self.PAD))
        n_mdr = np.zeros((Cn, self.PAD), np.int8)
        n_act = np.zeros((Cn, self.PAD), np.int8)
        fsum = np.zeros(Cn)
        rc = self.lib.gwo_fear_one_actor(self._h, Cn, _p(nper), _p(pos), _p(act), _p(md), _p(ac), _p(il), _p(resp),
                                         _p(n_mdr), _p(n_act), _p(fsum))
        assert rc == 0
        return resp, n_mdr, n_act, fsum

    def _matrix_call(self, fn, shape_tail, positions, actions, mdr, in_list, n_agents):
        pos, act, md, nper = _i8(positions), _i8(actions), _i8(mdr), _i8(n_agents)
        il = None if in_list is None else np.ascontiguousarray(in_list, dtype=np.uint8)
        Cn = pos.shape[0]
        val = np.zeros((Cn,) + shape_tail)
        n_mdr = np.zeros((Cn,) + shape_tail, np.int8)
        n_act = np.zeros((Cn,) + shape_tail, np.int8)
        rc = fn(self._h, Cn, _p(nper), _p(pos), _p(act), _p(md), _p(il), _p(val), _p(n_mdr), _p(n_act))
        assert rc == 0
        return val, n_mdr, n_act

    def fear_matrix(self, positions, actions, mdr, in_list=None, n_agents=None):
        return self._matrix_call(self.lib.gwo_fear_matrix, (self.PAD, self.PAD), positions, actions, mdr, in_list, n_agents)

    def feal(self, positions, actions, mdr, in_list=None, n_agents=None):
        return self._matrix_call(self.lib.gwo_feal, (self.PAD,), positions, actions, mdr, in_list, n_agents)
