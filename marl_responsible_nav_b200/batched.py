"""BatchedGridWorld: E independent grid worlds advanced per kernel launch (tensor API).

Host side is PyTorch only for device memory and streams; every computation happens in
csrc/libgridworld_b200.so through the C-ABI of include/gridworld_b200.h.  No CPU path.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _native as N
from .scenarios import Scenario, builtin_scenario

MA_APPLES = ((9, 0), (5, 10))        # custom/ma_customenv.py:422
SINGLE_APPLE = ((9, 15),)            # custom/customenv.py:334


@dataclass
class StepOutput:
    """Device tensors written by one gw_reset / gw_step call (views of the env's own buffers unless
    the caller passed its own)."""
    obs: torch.Tensor                       # [E, n_learners, H*W] (mlp) or [E, n_learners, 1, H, W] (cnn)
    action_mask: torch.Tensor               # int8 [E, n_learners, 9]
    positions: torch.Tensor                 # int8 [E, n_agents, 2]
    reward: Optional[torch.Tensor] = None        # f32 [E, n_learners]
    shaped_reward: Optional[torch.Tensor] = None  # f32 [E, n_learners]  FeAR_weight*fear + reward
    fear: Optional[torch.Tensor] = None          # f64 [E, n_learners]
    terminated: Optional[torch.Tensor] = None    # u8 [E, n_learners]
    truncated: Optional[torch.Tensor] = None     # u8 [E, n_learners]
    ended: Optional[torch.Tensor] = None         # u8 [E]
    info: Optional[torch.Tensor] = None          # int32 [E] packed (see gridworld_b200.h)
    final_obs: Optional[torch.Tensor] = None     # like obs; rows valid where ended != 0
    obs_code: Optional[torch.Tensor] = None      # int64 [E] compact form of obs (input of FusedActor)


def _check(t: Optional[torch.Tensor], name, dtype, shape, device):
    if t is None:
        return None
    if t.device != device or t.dtype != dtype or tuple(t.shape) != tuple(shape) or not t.is_contiguous():
        raise ValueError(f"{name}: expected contiguous {dtype} tensor of shape {tuple(shape)} on {device}, "
                         f"got {t.dtype} {tuple(t.shape)} on {t.device}")
    return t


class BatchedGridWorld:
    """E grid worlds on one GPU.  Two state layouts sit behind the same tensor API and are chosen here, when the world
    is created: the PACKED layout (16 bytes per env; the shipped scenarios: W = 16, H <= 16, <= 4 agents) with every
    kernel of the library, and the GENERAL layout (`GeneralGridWorld`, gww_* in gridworld_b200.h: grids up to 64 x 64, up
    to 16 agents) for scenario files that do not fit.  `layout=None` picks packed whenever the scenario fits."""
    _P, _wide, PAD = "gw_", False, N.GW_MAX_AGENTS       # C-ABI prefix, error channel, per-agent length of the operator-level arrays

    def __new__(cls, scenario: Union[str, Scenario] = "Level 3", *args, layout: Optional[str] = None, **kw):
        if cls is BatchedGridWorld:
            if layout not in (None, "packed", "general"):
                raise ValueError("layout must be None, 'packed' or 'general'")
            sc = builtin_scenario(scenario) if isinstance(scenario, str) else scenario
            if layout == "general" or (layout is None and not N.fits_packed_layout(sc, kw.get("n_agents"))):
                return object.__new__(GeneralGridWorld)
        return object.__new__(cls)

    def _f(self, name):
        return getattr(self.lib, self._P + name)

    def _check_rc(self, rc, what):
        if rc:
            N.check(rc, self._h, self._P + what, wide=self._wide)

    def _create_handle(self, **kw):
        cfg = N.build_config(self.scenario, **kw)
        probe = N.GwConfig()
        N.check(self.lib.gw_default_config(C.byref(probe)), None, "gw_default_config")
        if probe.struct_size != C.sizeof(N.GwConfig):
            raise RuntimeError("gw_config layout mismatch between ctypes and the library")
        h = C.c_void_p()
        N.check(self.lib.gw_create(C.byref(cfg), C.byref(h)), None, "gw_create")
        return cfg, h

    def __init__(self, scenario: Union[str, Scenario] = "Level 3", num_envs: int = 1, device="cuda",
                 env_kind: str = "multi", fear: bool = True, fear_weight: float = 0.0, fear_radius: int = 5,
                 n_agents: Optional[int] = None, n_learners: Optional[int] = None,
                 apples: Optional[Sequence[Tuple[int, int]]] = None, max_steps: int = 150,
                 auto_reset: bool = True, obs_dtype: torch.dtype = torch.float32, obs_layout: str = "mlp",
                 seed: int = 0, env_id_base: int = 0, perturb_prob: float = 0.25, layout: Optional[str] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("BatchedGridWorld needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.lib = N.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("BatchedGridWorld runs on CUDA devices only")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.scenario = builtin_scenario(scenario) if isinstance(scenario, str) else scenario
        sc = self.scenario
        self.kind = {"multi": N.GW_ENV_MULTI, "single": N.GW_ENV_SINGLE}[env_kind]
        self.n_agents = int(n_agents if n_agents is not None else sc.n_agents)
        self.n_learners = int(n_learners if n_learners is not None else (2 if env_kind == "multi" else 1))
        self.num_envs = int(num_envs)
        self.env_id_base = int(env_id_base)
        self.H, self.W = sc.shape
        self.obs_len = self.H * self.W
        if obs_dtype not in (torch.float32, torch.bfloat16):
            raise ValueError("obs_dtype must be torch.float32 or torch.bfloat16")
        if obs_layout not in ("mlp", "cnn"):
            raise ValueError("obs_layout must be 'mlp' ([E,L,H*W]) or 'cnn' ([E,L,1,H,W])")
        self.obs_dtype, self.obs_layout = obs_dtype, obs_layout
        if apples is None:
            apples = MA_APPLES[:self.n_learners] if env_kind == "multi" else SINGLE_APPLE
        self.apples = tuple(tuple(a) if a is not None else None for a in apples)

        self.cfg, self._h = self._create_handle(
            num_envs=self.num_envs, env_kind=env_kind, fear=fear, fear_weight=fear_weight, fear_radius=fear_radius,
            n_agents=self.n_agents, n_learners=self.n_learners, apples=self.apples, max_steps=max_steps, auto_reset=auto_reset,
            obs_bf16=(obs_dtype == torch.bfloat16), seed=seed, env_id_base=env_id_base, perturb_prob=perturb_prob,
            device=self.device.index)
        self.fear, self.fear_weight, self.auto_reset, self.max_steps = bool(fear), float(fear_weight), bool(auto_reset), int(max_steps)

        E, L, A, dev = self.num_envs, self.n_learners, self.n_agents, self.device
        self._obs_shape = (E, L, self.obs_len)
        self.buf = StepOutput(
            obs=torch.empty(self._obs_shape, dtype=obs_dtype, device=dev),
            action_mask=torch.empty((E, L, N.GW_N_ACTIONS), dtype=torch.int8, device=dev),
            positions=torch.empty((E, A, 2), dtype=torch.int8, device=dev),
            reward=torch.zeros((E, L), dtype=torch.float32, device=dev),
            shaped_reward=torch.zeros((E, L), dtype=torch.float32, device=dev),
            fear=torch.zeros((E, L), dtype=torch.float64, device=dev),
            terminated=torch.zeros((E, L), dtype=torch.uint8, device=dev),
            truncated=torch.zeros((E, L), dtype=torch.uint8, device=dev),
            ended=torch.zeros((E,), dtype=torch.uint8, device=dev),
            info=torch.zeros((E,), dtype=torch.int32, device=dev),
            final_obs=None,
            obs_code=torch.zeros((E,), dtype=torch.int64, device=dev))
        self._io_step, self._io_reset = self._make_io(True), self._make_io(False)
        self._host_calls = {}                                      # step_host: (buffer pointers, mode) -> (library token, StepOutput)

    # ------------------------------------------------------------------ helpers
    def close(self):
        h = getattr(self, "_h", None)
        if h:
            self._h = None
            self._f("destroy")(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        try:                                                       # raw handle of the current stream, without building a Stream object
            return C.c_void_p(torch._C._cuda_getCurrentRawStream(self.device.index))
        except AttributeError:
            return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _view(self, obs: torch.Tensor) -> torch.Tensor:
        if self.obs_layout == "cnn":
            return obs.view(self.num_envs, self.n_learners, 1, self.H, self.W)
        return obs

    def _io(self, obs, final_obs, actions=None, npc_actions=None, spawn=None, outputs=True, buffers=None) -> N.GwIO:
        E, L, A, dev = self.num_envs, self.n_learners, self.n_agents, self.device
        io = self._io_step if outputs else self._io_reset          # pointers to the env-owned buffers are filled once
        if outputs:
            self._point_outputs(io, buffers if buffers is not None else self.buf)
        io.learner_actions = _check(actions, "actions", torch.int8, (E, L), dev).data_ptr() if actions is not None else None
        io.npc_actions = _check(npc_actions, "npc_actions", torch.int8, (E, A), dev).data_ptr() if npc_actions is not None else None
        io.spawn = _check(spawn, "spawn", torch.int8, (E, A, 2), dev).data_ptr() if spawn is not None else None
        io.obs = obs.data_ptr()
        io.final_obs = final_obs.data_ptr() if final_obs is not None else None
        return io

    def _make_io(self, outputs: bool) -> N.GwIO:
        io, b = N.GwIO(), self.buf
        io.action_mask, io.positions = b.action_mask.data_ptr(), b.positions.data_ptr()
        io.obs_code = b.obs_code.data_ptr()
        if outputs:
            self._point_outputs(io, b)
        return io

    _OUT_SPECS = (("reward", torch.float32, True), ("shaped_reward", torch.float32, True), ("fear", torch.float64, True),
                  ("terminated", torch.uint8, True), ("truncated", torch.uint8, True), ("ended", torch.uint8, False),
                  ("info", torch.int32, False))

    def _point_outputs(self, io: N.GwIO, b):
        """Scalar outputs go to the env's own buffers or to caller tensors of the same shapes (e.g. a replay-ring slot)."""
        if b is getattr(self, "_last_out_owner", None):
            return
        ptrs = getattr(b, "_gw_ptrs", None)                        # checked once per buffer object (a replay-ring slot comes back)
        if ptrs is None or ptrs[0] is not self:
            E, L, dev = self.num_envs, self.n_learners, self.device
            vals = []
            for name, dtype, per_learner in self._OUT_SPECS:
                t = getattr(b, name)
                if b is not self.buf:
                    _check(t, name, dtype, (E, L) if per_learner else (E,), dev)
                vals.append(t.data_ptr())
            ptrs = (self, tuple(vals))
            try:
                b._gw_ptrs = ptrs
            except AttributeError:
                pass
        (io.reward, io.shaped_reward, io.fear, io.terminated, io.truncated, io.ended, io.info) = ptrs[1]
        self._last_out_owner = b if b is self.buf else None

    def _obs_arg(self, t: Optional[torch.Tensor], name: str) -> Optional[torch.Tensor]:
        if t is None:
            return None
        if t.dtype != self.obs_dtype or t.device != self.device or not t.is_contiguous() or t.numel() != self.num_envs * self.n_learners * self.obs_len:
            raise ValueError(f"{name}: need a contiguous {self.obs_dtype} tensor with {self._obs_shape} elements on {self.device}")
        if t.data_ptr() % 16:
            raise ValueError(f"{name}: must be 16-byte aligned")
        return t

    @staticmethod
    def _as_i8(x, device) -> Optional[torch.Tensor]:
        if x is None:
            return None
        if isinstance(x, torch.Tensor):
            if x.dtype == torch.int8 and x.device == device and x.is_contiguous():
                return x                                           # hot path: no copy, no launch
            return x.to(device=device, dtype=torch.int8).contiguous()
        return torch.as_tensor(np.asarray(x)).to(device=device, dtype=torch.int8).contiguous()

    # ------------------------------------------------------------------ API
    def reset(self, mask: Optional[torch.Tensor] = None, spawn=None, obs_out: Optional[torch.Tensor] = None) -> StepOutput:
        """CustomMAEnv.reset for every env (or those with mask != 0).  `spawn` [E, n_agents, 2] replays recorded
        spawn cells (sorted row-major like the reference); None draws them on the device."""
        spawn = self._as_i8(spawn, self.device)
        obs = self._obs_arg(obs_out, "obs_out") if obs_out is not None else self.buf.obs
        io = self._io(obs, None, None, None, spawn, outputs=False)
        mptr = None
        if mask is not None:
            mask = _check(mask.to(torch.uint8) if mask.dtype == torch.bool else mask, "mask", torch.uint8, (self.num_envs,), self.device)
            mptr = mask.data_ptr()
        self._check_rc(self._f("reset")(self._h, mptr, C.byref(io), self._stream()), "reset")
        return StepOutput(obs=self._view(obs), action_mask=self.buf.action_mask, positions=self.buf.positions,
                          obs_code=self.buf.obs_code)

    def step(self, actions, npc_actions=None, spawn=None, obs_out: Optional[torch.Tensor] = None,
             final_obs_out: Optional[torch.Tensor] = None, buffers=None) -> StepOutput:
        """One env.step for all envs.  actions: int8 [E, n_learners].  `npc_actions` [E, n_agents] / `spawn`
        replay recorded NPC draws / respawn cells; None = device RNG.  `buffers`: object with reward / shaped_reward /
        fear / terminated / truncated / ended / info tensors to write instead of the env's own (ReplayRing slot)."""
        actions = self._as_i8(actions, self.device)
        npc_actions = self._as_i8(npc_actions, self.device)
        spawn = self._as_i8(spawn, self.device)
        obs = self._obs_arg(obs_out, "obs_out") if obs_out is not None else self.buf.obs
        fin = self._obs_arg(final_obs_out, "final_obs_out")
        io = self._io(obs, fin, actions, npc_actions, spawn, buffers=buffers)
        self._check_rc(self._f("step")(self._h, C.byref(io), self._stream()), "step")
        b = buffers if buffers is not None else self.buf
        m = self.buf
        return StepOutput(obs=self._view(obs), action_mask=m.action_mask, positions=m.positions, reward=b.reward,
                          shaped_reward=b.shaped_reward, fear=b.fear, terminated=b.terminated, truncated=b.truncated,
                          ended=b.ended, info=b.info, final_obs=self._view(fin) if fin is not None else None,
                          obs_code=m.obs_code)

    # ---- device-side rollout: T steps per launch (gw_rollout)
    _RING_SPECS = (("obs", None, ("L", "O")), ("final_obs", None, ("L", "O")), ("reward", torch.float32, ("L",)),
                   ("shaped_reward", torch.float32, ("L",)), ("fear", torch.float64, ("L",)), ("terminated", torch.uint8, ("L",)),
                   ("truncated", torch.uint8, ("L",)), ("ended", torch.uint8, ()), ("info", torch.int32, ()),
                   ("positions", torch.int8, ("A", 2)), ("obs_code", torch.int64, ()), ("action_mask", torch.int8, ("L", 9)))

    def new_rings(self, slots: int, fields: Optional[Sequence[str]] = None) -> "SimpleNamespace":
        """Time-major output rings for `rollout`: one tensor [slots, E, ...] per output of gw_step (`fields` = subset;
        "obs" is always there).  A ReplayRing has the same attribute names and works as well."""
        from types import SimpleNamespace
        dims = {"L": self.n_learners, "O": self.obs_len, "A": self.n_agents}
        ns = SimpleNamespace()
        for name, dtype, tail in self._RING_SPECS:
            if fields is not None and name not in fields and name != "obs":
                setattr(ns, name, None)
                continue
            shape = (int(slots), self.num_envs) + tuple(dims.get(d, d) for d in tail)
            setattr(ns, name, torch.zeros(shape, dtype=dtype or self.obs_dtype, device=self.device))
        return ns

    def rollout(self, actions: torch.Tensor, steps: int, rings, first_slot: int = 0, first_action: int = 0,
                npc_actions: Optional[torch.Tensor] = None, spawn: Optional[torch.Tensor] = None):
        """`steps` consecutive env steps in ONE kernel launch (gw_rollout): the inner loop of MADDPGAgent.train
        (maddpg/agent.py:85-197) for actions that are already on the device.  actions: int8 [A, E, n_learners]; step k
        plays actions[(first_action + k) % A] and writes its transition into slot (first_slot + k) % T of `rings`
        (time-major tensors [T, E, ...] named like StepOutput's fields: `new_rings`, or a ReplayRing) and the observation /
        obs_code / action mask that follow it into the slot after.  Replay mode: npc_actions [A, E, n_agents], spawn
        [A, E, n_agents, 2] with the same indexing.  Bit-identical to `steps` calls of step()."""
        E, L, A, dev = self.num_envs, self.n_learners, self.n_agents, self.device
        actions = self._as_i8(actions, dev)
        if actions.dim() != 3 or tuple(actions.shape[1:]) != (E, L):
            raise ValueError(f"actions: expected int8 [A, {E}, {L}]")
        n_act = actions.shape[0]
        T = int(rings.obs.shape[0])
        dims = {"L": L, "O": self.obs_len, "A": A}
        io = N.GwIO()
        io.learner_actions = actions.data_ptr()
        for name, t, tail in (("npc_actions", npc_actions, (A,)), ("spawn", spawn, (A, 2))):
            if t is not None:
                t = self._as_i8(t, dev)
                if tuple(t.shape) != (n_act, E) + tail:
                    raise ValueError(f"{name}: expected int8 {(n_act, E) + tail}")
                setattr(io, name, t.data_ptr())
        keep = [actions, npc_actions, spawn]
        for name, dtype, tail in self._RING_SPECS:
            t = getattr(rings, name, None)
            if t is None:
                continue
            if name in ("obs", "final_obs"):
                if t.dtype != self.obs_dtype or t.device != dev or not t.is_contiguous() or t.shape[0] != T or t[0].numel() != E * L * self.obs_len:
                    raise ValueError(f"rings.{name}: need a contiguous {self.obs_dtype} tensor [T, E, L, {self.obs_len}] on {dev}")
            else:
                _check(t, "rings." + name, dtype, (T, E) + tuple(dims.get(d, d) for d in tail), dev)
            setattr(io, name, t.data_ptr())
        plan = N.GwRolloutPlan(C.sizeof(N.GwRolloutPlan), int(steps), T, int(first_slot) % T, n_act, int(first_action) % n_act)
        N.check(self.lib.gw_rollout(self._h, C.byref(io), C.byref(plan), self._stream()), self._h, "gw_rollout")
        return rings

    def step_host(self, host_actions: torch.Tensor, host_reward: torch.Tensor, host_ended: Optional[torch.Tensor] = None,
                  host_shaped: Optional[torch.Tensor] = None, obs_out: Optional[torch.Tensor] = None,
                  zero_copy: bool = True, resident: Optional[bool] = None) -> StepOutput:
        """Host-driven step in one library call (the reference's calling pattern, maddpg/agent.py:121-131): pinned int8
        actions [E, L] in, pinned f32 rewards [E, L] (and u8 ended [E], f32 shaped rewards) out, stream synchronised on
        return.  Observations and masks stay on the device.
        zero_copy (default): the kernel itself loads the actions from the pinned buffer and stores rewards / flags into
        the pinned buffers over PCIe; the returned StepOutput then carries the host tensors for those fields.
        zero_copy=False: cudaMemcpyAsync H2D / D2H around the kernel and device-side copies of the outputs as well.
        resident (default: on for zero_copy with up to 8192 envs, where it is measured to win: 20 us against 28 us per
        step at 8192 envs, level at 16384; the library accepts it up to 24576 envs and ignores it beyond): the step is
        served by a kernel that stays on the GPU between calls (doorbell and completion word in pinned host memory, no
        launch and no stream synchronisation per step; gridworld_b200.h, GW_HOST_RESIDENT).  The host buffers are valid
        on return as before; GPU work queued behind it on the stream (reading obs, say) starts once the kernel has left:
        at the next call of any other method of this object (`sync()` included), or by itself after 1 ms without a step.
        Argument checking and marshalling are done once per distinct set of buffers."""
        obs = obs_out if obs_out is not None else self.buf.obs
        if zero_copy and (resident or (resident is None and self.num_envs <= 8192)):
            zero_copy = 2
        key = (host_actions.data_ptr(), host_reward.data_ptr(), 0 if host_ended is None else host_ended.data_ptr(),
               0 if host_shaped is None else host_shaped.data_ptr(), obs.data_ptr(), zero_copy)
        cache = self._host_calls
        ent = cache.get(key)
        if ent is None:
            if len(cache) >= 4096:                                 # tokens live in the library: forget both sides together
                cache.clear()
                N.check(self.lib.gw_host_call_reset(self._h), self._h, "gw_host_call_reset")
            ent = cache[key] = self._prepare_host_call(host_actions, host_reward, host_ended, host_shaped, obs_out, zero_copy)
        token, out = ent
        rc = self.lib.gw_host_call_run(self._h, token, self._stream())
        if rc:
            N.check(rc, self._h, "gw_step_host")
        return out

    def _prepare_host_call(self, host_actions, host_reward, host_ended, host_shaped, obs_out, zero_copy):
        E, L = self.num_envs, self.n_learners
        for t, name, dtype, shape in ((host_actions, "host_actions", torch.int8, (E, L)), (host_reward, "host_reward", torch.float32, (E, L)),
                                      (host_ended, "host_ended", torch.uint8, (E,)), (host_shaped, "host_shaped", torch.float32, (E, L))):
            if t is None:
                continue
            if t.is_cuda or not t.is_pinned() or not t.is_contiguous() or t.dtype != dtype or tuple(t.shape) != shape:
                raise ValueError(f"{name} must be a contiguous pinned host tensor of dtype {dtype} and shape {shape}")
        if not hasattr(self, "_host_act_dev"):
            self._host_act_dev = torch.empty((E, L), dtype=torch.int8, device=self.device)
        obs = self._obs_arg(obs_out, "obs_out") if obs_out is not None else self.buf.obs
        io = N.GwIO()
        C.memmove(C.byref(io), C.byref(self._io(obs, None, self._host_act_dev, None, None)), C.sizeof(N.GwIO))   # own copy
        p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        b = self.buf
        pick = (lambda host, dev: host if (zero_copy and host is not None) else dev)
        out = StepOutput(obs=self._view(obs), action_mask=b.action_mask, positions=b.positions, reward=pick(host_reward, b.reward),
                         shaped_reward=pick(host_shaped, b.shaped_reward), fear=b.fear, terminated=b.terminated,
                         truncated=b.truncated, ended=pick(host_ended, b.ended), info=b.info, obs_code=b.obs_code)
        token = C.c_int()
        N.check(self.lib.gw_host_call_prepare(self._h, C.byref(io), p(host_actions), p(host_reward), p(host_shaped), p(host_ended),
                                              int(zero_copy), C.byref(token)), self._h, "gw_host_call_prepare")
        return (int(token.value), out)

    def pinned_io(self) -> "PinnedIO":
        """Host-visible input / output block for host-driven use of a small batch (the E = 1 drop-in classes)."""
        if getattr(self, "_pinned_io", None) is None:
            self._pinned_io = PinnedIO(self)
        return self._pinned_io

    def sync(self):
        """Stream synchronisation (a resident step kernel is told to leave first)."""
        self._check_rc(self._f("sync")(self._h, self._stream()), "sync")

    def server_info(self) -> Dict[str, int]:
        """Resident step kernel: running now, launches so far, relaunches after an idle exit, registered buffer sets."""
        r, n = C.c_int(), C.c_int()
        a, b = C.c_uint64(), C.c_uint64()
        N.check(self.lib.gw_server_info(self._h, C.byref(r), C.byref(a), C.byref(b), C.byref(n)), self._h, "gw_server_info")
        return {"running": int(r.value), "launches": int(a.value), "relaunches": int(b.value), "buffer_sets": int(n.value)}

    def stats(self) -> Dict[str, float]:
        s = N.GwStats()
        self._check_rc(self._f("get_stats")(self._h, C.byref(s), self._stream()), "get_stats")
        return {name: getattr(s, name) for name, _ in N.GwStats._fields_}

    def reset_stats(self):
        N.check(self.lib.gw_reset_stats(self._h, self._stream()), self._h, "gw_reset_stats")

    def launch_count(self) -> int:
        n = C.c_uint64()
        self._check_rc(self._f("launch_count")(self._h, C.byref(n)), "launch_count")
        return int(n.value)

    def state_dict(self) -> torch.Tensor:
        """Packed per-env state (16 B/env) as a uint8 CPU tensor (checkpoint / resume)."""
        nbytes = int(self._f("state_bytes")(self._h))
        out = torch.empty(nbytes, dtype=torch.uint8)
        self._check_rc(self._f("get_state")(self._h, C.c_void_p(out.data_ptr()), 0, self._stream()), "get_state")
        return out

    def load_state_dict(self, state: torch.Tensor):
        nbytes = int(self._f("state_bytes")(self._h))
        state = state.contiguous()
        if state.dtype != torch.uint8 or state.numel() != nbytes:
            raise ValueError(f"state must be a uint8 tensor of {nbytes} bytes")
        self._check_rc(self._f("set_state")(self._h, C.c_void_p(state.data_ptr()), int(state.is_cuda), self._stream()), "set_state")

    # ------------------------------------------------------------------ operator-level entry points
    def update_world(self, positions, actions, n_agents=None, apples=None):
        """GWorld.UpdateGWorld for C independent cases.  positions [C,4,2], actions [C,4] (int8, padded)."""
        dev = self.device
        pos, act = self._as_i8(positions, dev), self._as_i8(actions, dev)
        Cn = pos.shape[0]
        nper, app = self._as_i8(n_agents, dev), self._as_i8(apples, dev)
        new_pos = torch.empty((Cn, self.PAD, 2), dtype=torch.int8, device=dev)
        crash = torch.empty((Cn, self.PAD), dtype=torch.uint8, device=dev)
        restr = torch.empty((Cn, self.PAD), dtype=torch.uint8, device=dev)
        caught = torch.zeros((Cn, 2, 2), dtype=torch.int8, device=dev)
        p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        self._check_rc(self._f("update_world")(self._h, Cn, p(nper), p(pos), p(act), p(app), p(new_pos), p(crash), p(restr),
                                               p(caught), self._stream()), "update_world")
        return new_pos, crash, restr, caught

    def fear_one_actor(self, positions, actions, mdr, actor, in_list=None, n_agents=None):
        """Responsibility.FeAR_4_one_actor for C independent cases -> (resp f64 [C,4], n_mdr, n_act int8 [C,4])."""
        dev = self.device
        pos, act, md, ac = (self._as_i8(x, dev) for x in (positions, actions, mdr, actor))
        Cn = pos.shape[0]
        nper = self._as_i8(n_agents, dev)
        il = None if in_list is None else torch.as_tensor(np.asarray(in_list)).to(device=dev, dtype=torch.uint8).contiguous()
        resp = torch.empty((Cn, self.PAD), dtype=torch.float64, device=dev)
        n_mdr = torch.empty((Cn, self.PAD), dtype=torch.int8, device=dev)
        n_act = torch.empty((Cn, self.PAD), dtype=torch.int8, device=dev)
        p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        if self._wide:                                             # gww_fear_one_actor also returns np.sum of the matrix (info["fear"])
            self.last_fear_sum = torch.empty((Cn,), dtype=torch.float64, device=dev)
            self._check_rc(self.lib.gww_fear_one_actor(self._h, Cn, p(nper), p(pos), p(act), p(md), p(ac), p(il), p(resp), p(n_mdr),
                                                       p(n_act), p(self.last_fear_sum), self._stream()), "fear_one_actor")
            return resp, n_mdr, n_act
        N.check(self.lib.gw_fear_one_actor(self._h, Cn, p(nper), p(pos), p(act), p(md), p(ac), p(il), p(resp), p(n_mdr),
                                           p(n_act), self._stream()), self._h, "gw_fear_one_actor")
        return resp, n_mdr, n_act

    def _matrix_op(self, fn, name, tail, positions, actions, mdr, in_list, n_agents):
        dev = self.device
        pos, act, md = (self._as_i8(x, dev) for x in (positions, actions, mdr))
        Cn = pos.shape[0]
        nper = self._as_i8(n_agents, dev)
        il = None if in_list is None else torch.as_tensor(np.asarray(in_list)).to(device=dev, dtype=torch.uint8).contiguous()
        val = torch.empty((Cn,) + tail, dtype=torch.float64, device=dev)
        n_mdr = torch.empty((Cn,) + tail, dtype=torch.int8, device=dev)
        n_act = torch.empty((Cn,) + tail, dtype=torch.int8, device=dev)
        p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        self._check_rc(fn(self._h, Cn, p(nper), p(pos), p(act), p(md), p(il), p(val), p(n_mdr), p(n_act), self._stream()), name)
        return val, n_mdr, n_act

    def fear_matrix(self, positions, actions, mdr, in_list=None, n_agents=None):
        """Responsibility.FeAR (all actors) for C cases -> (resp f64 [C,4,4], n_mdr, n_act int8 [C,4,4])."""
        return self._matrix_op(self._f("fear_matrix"), "fear_matrix", (self.PAD, self.PAD), positions, actions, mdr, in_list, n_agents)

    def feal(self, positions, actions, mdr, in_list=None, n_agents=None):
        """Responsibility.FeAL for C cases -> (feal f64 [C,4], n_mdr, n_act int8 [C,4])."""
        return self._matrix_op(self._f("feal"), "feal", (self.PAD,), positions, actions, mdr, in_list, n_agents)


class GeneralGridWorld(BatchedGridWorld):
    """The GENERAL state layout (gww_* in include/gridworld_b200.h, csrc/gw_wide.cu): the same reset / step / rollout /
    operator-level API for grids up to 64 x 64 with up to 16 agents.  What is tied to the packed layout is not offered here: the
    multi-step rollout KERNEL (`rollout` issues one launch per step), the resident host-driven step, `obs_code` and the fused
    actor (networks sized for 160 cells)."""
    _P, _wide, PAD = "gww_", True, N.GWW_MAX_AGENTS

    def _create_handle(self, **kw):
        cfg = N.build_wide_config(self.scenario, **kw)
        probe = N.GwwConfig()
        N.check(self.lib.gww_default_config(C.byref(probe)), None, "gww_default_config", wide=True)
        if probe.struct_size != C.sizeof(N.GwwConfig):
            raise RuntimeError("gww_config layout mismatch between ctypes and the library")
        h = C.c_void_p()
        N.check(self.lib.gww_create(C.byref(cfg), C.byref(h)), None, "gww_create", wide=True)
        return cfg, h

    def _obs_arg(self, t, name):
        if t is None:
            return None
        if t.dtype != self.obs_dtype or t.device != self.device or not t.is_contiguous() or t.numel() != self.num_envs * self.n_learners * self.obs_len:
            raise ValueError(f"{name}: need a contiguous {self.obs_dtype} tensor with {self._obs_shape} elements on {self.device}")
        return t

    def _packed_only(self, *a, **k):
        raise NotImplementedError("this entry point exists for the packed layout only (W = 16, H <= 16, <= 4 agents); "
                                  "the general layout offers reset / step / state / stats and the operator-level calls")

    step_host = pinned_io = server_info = _packed_only

    @property
    def service_env(self) -> BatchedGridWorld:
        """A one-env world on the packed layout and the same device: the handle that the library's replay-sampling and
        update kernels (gw_replay_sample, gw_learner_*: they take a gw_handle for their device, stream and error state) are
        created on when the training loop runs on a general-layout world."""
        if getattr(self, "_service_env", None) is None:
            self._service_env = BatchedGridWorld("Level 3", num_envs=1, device=self.device, fear=False, layout="packed")
        return self._service_env

    def reset_stats(self):
        self._check_rc(self.lib.gww_reset_stats(self._h, self._stream()), "reset_stats")

    def rollout(self, actions: torch.Tensor, steps: int, rings, first_slot: int = 0, first_action: int = 0,
                npc_actions: Optional[torch.Tensor] = None, spawn: Optional[torch.Tensor] = None):
        """The same contract as BatchedGridWorld.rollout (time-major rings, step k's transition in slot (first_slot + k) % T,
        the observation / action mask that follow it in the slot after), served by `steps` gww_step launches whose outputs
        point straight into the ring slots -- the general layout has no multi-step kernel.  `rings.obs_code` is not written."""
        E, L, A, dev = self.num_envs, self.n_learners, self.n_agents, self.device
        actions = self._as_i8(actions, dev)
        if actions.dim() != 3 or tuple(actions.shape[1:]) != (E, L):
            raise ValueError(f"actions: expected int8 [A, {E}, {L}]")
        n_act, T = actions.shape[0], int(rings.obs.shape[0])
        dims = {"L": L, "O": self.obs_len, "A": A}
        extra = {}
        for name, t, tail in (("npc_actions", npc_actions, (A,)), ("spawn", spawn, (A, 2))):
            if t is not None:
                t = self._as_i8(t, dev)
                if tuple(t.shape) != (n_act, E) + tail:
                    raise ValueError(f"{name}: expected int8 {(n_act, E) + tail}")
                extra[name] = t
        fields = {}
        for name, dtype, tail in self._RING_SPECS:
            t = getattr(rings, name, None)
            if t is None or name == "obs_code":
                continue
            if name in ("obs", "final_obs"):
                if t.dtype != self.obs_dtype or t.device != dev or not t.is_contiguous() or t.shape[0] != T or t[0].numel() != E * L * self.obs_len:
                    raise ValueError(f"rings.{name}: need a contiguous {self.obs_dtype} tensor [T, E, L, {self.obs_len}] on {dev}")
            else:
                _check(t, "rings." + name, dtype, (T, E) + tuple(dims.get(d, d) for d in tail), dev)
            fields[name] = t
        stream = self._stream()
        for k in range(int(steps)):
            slot, nxt, a = (first_slot + k) % T, (first_slot + k + 1) % T, (first_action + k) % n_act
            io = N.GwIO()
            io.learner_actions = actions[a].data_ptr()
            for name, t in extra.items():
                setattr(io, name, t[a].data_ptr())
            for name, t in fields.items():
                setattr(io, name, t[nxt if name in ("obs", "action_mask") else slot].data_ptr())
            self._check_rc(self.lib.gww_step(self._h, C.byref(io), stream), "step")
        return rings


class PinnedIO:
    """Every input and output of reset / step in pinned HOST memory, read and written by the kernels themselves over PCIe
    (unified addressing), exposed as numpy arrays: one library call per step, no device-to-host copy, no tensor op.  The
    step goes through `gw_step_host` with the resident kernel (include/gridworld_b200.h, GW_HOST_RESIDENT), so after the
    first step there is no kernel launch and no stream synchronisation either.  Meant for the single-env front-ends
    (`CustomMAEnv`, `CustomEnv`: ~45 us per step instead of ~300 us with tensor copies); at a few KB per env it is the
    wrong tool for large batches, whose observations belong in HBM.

    Fill `actions` (and, in replay mode, `npc_actions` / `spawn`), call `step()` / `reset()`, read the other arrays."""

    def __init__(self, env: BatchedGridWorld):
        if env.obs_dtype != torch.float32:
            raise ValueError("PinnedIO exposes float32 observations")
        if env.num_envs > 8192:
            raise ValueError("PinnedIO is meant for small batches (<= 8192 envs)")
        self.env = env
        E, L, A = env.num_envs, env.n_learners, env.n_agents
        spec = (("actions", np.int8, (E, L)), ("npc_actions", np.int8, (E, A)), ("spawn", np.int8, (E, A, 2)),
                ("obs", np.float32, (E, L, env.obs_len)), ("final_obs", np.float32, (E, L, env.obs_len)),
                ("reward", np.float32, (E, L)), ("shaped_reward", np.float32, (E, L)), ("fear", np.float64, (E, L)),
                ("terminated", np.uint8, (E, L)), ("truncated", np.uint8, (E, L)), ("ended", np.uint8, (E,)),
                ("action_mask", np.int8, (E, L, N.GW_N_ACTIONS)), ("positions", np.int8, (E, A, 2)), ("info", np.uint32, (E,)),
                ("obs_code", np.uint64, (E,)))
        off, offs = 0, {}
        for name, dt, shape in spec:
            offs[name] = off
            off += -(-int(np.prod(shape)) * np.dtype(dt).itemsize // 64) * 64          # 64-byte slots: every alignment rule holds
        self._block = torch.zeros(off, dtype=torch.uint8).pin_memory()                   # one pinned allocation
        raw = self._block.numpy()
        base = self._block.data_ptr()
        ptr = {}
        for name, dt, shape in spec:
            n = int(np.prod(shape)) * np.dtype(dt).itemsize
            setattr(self, name, raw[offs[name]:offs[name] + n].view(dt).reshape(shape))
            ptr[name] = base + offs[name]
        self._ptr = ptr
        self.replay = False                     # True: `npc_actions` / `spawn` are inputs (recorded draws); False: device RNG
        self._io = {}

    def _make_io(self, step: bool, replay: bool) -> N.GwIO:
        io, p = N.GwIO(), self._ptr
        for name in ("obs", "action_mask", "positions", "obs_code"):
            setattr(io, name, p[name])
        if step:
            for name in ("final_obs", "reward", "shaped_reward", "fear", "terminated", "truncated", "ended", "info"):
                setattr(io, name, p[name])
            io.learner_actions = p["actions"]
            if replay:
                io.npc_actions = p["npc_actions"]
        if replay:
            io.spawn = p["spawn"]
        return io

    def _get_io(self, step: bool):
        key = (step, bool(self.replay))
        io = self._io.get(key)
        if io is None:
            io = self._io[key] = self._make_io(*key)
        return io

    def reset(self):
        """gw_reset for all envs (spawn cells from `spawn` in replay mode); the arrays are valid on return."""
        env = self.env
        N.check(env.lib.gw_reset(env._h, None, C.byref(self._get_io(False)), env._stream()), env._h, "gw_reset")
        env.sync()

    def step(self):
        """One step with `actions` (and `npc_actions` / `spawn` in replay mode); the arrays are valid on return."""
        env = self.env
        rc = env.lib.gw_step_host(env._h, C.byref(self._get_io(True)), self._ptr["actions"], None, None, None, 2, env._stream())
        if rc:
            N.check(rc, env._h, "gw_step_host")
