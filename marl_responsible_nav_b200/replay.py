"""Device-resident replay ring (the reference uses AgileRL's MultiAgentReplayBuffer, maddpg/agent.py:68-73,190-197).

Layout: time-major ring of T slots x E envs.  `gw_step` writes the observation of time t+1 straight into slot
(t+1) % T (`obs_out=ring.obs_slot(t+1)`), so a transition's `next_state` is the neighbouring slot and no observation
is stored twice; where an episode ended and the env was re-spawned inside the step, `next_state` is the terminal
observation the kernel wrote to `final_slot(t)` (only those rows are ever written).  Rewards, done flags and the
`ended` marker are written by the same kernel launch into per-slot buffers (`buffers_slot(t)`); the actor's continuous
action vector is the only thing the caller copies in.

Fields follow the reference: state, action (the continuous 9-vector, maddpg/agent.py:187-192), reward
(FeAR_weight*fear + reward, :128-131), next_state, done (= terminations, :186,195).
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict, Optional

import torch


class ReplayRing:
    def __init__(self, num_envs: int, n_learners: int, obs_len: int, capacity: int, action_dim: int = 9,
                 device="cuda", obs_dtype: torch.dtype = torch.float32):
        """capacity: number of (env, time) transitions to keep, as MEMORY_SIZE in configs/custom*.yaml."""
        E, L = int(num_envs), int(n_learners)
        self.E, self.L, self.obs_len, self.action_dim = E, L, int(obs_len), int(action_dim)
        self.T = max(3, -(-int(capacity) // E) + 1)            # +1: slot t+1 holds the next observation
        self.device = torch.device(device)
        dev, T = self.device, self.T
        self.obs = torch.zeros((T, E, L, self.obs_len), dtype=obs_dtype, device=dev)
        self.final_obs = torch.zeros((T, E, L, self.obs_len), dtype=obs_dtype, device=dev)
        self.action = torch.zeros((T, E, L, self.action_dim), dtype=torch.float32, device=dev)
        self.reward = torch.zeros((T, E, L), dtype=torch.float32, device=dev)          # env reward
        self.shaped_reward = torch.zeros((T, E, L), dtype=torch.float32, device=dev)   # what the trainer stores
        self.fear = torch.zeros((T, E, L), dtype=torch.float64, device=dev)
        self.terminated = torch.zeros((T, E, L), dtype=torch.uint8, device=dev)
        self.truncated = torch.zeros((T, E, L), dtype=torch.uint8, device=dev)
        self.ended = torch.zeros((T, E), dtype=torch.uint8, device=dev)
        self.info = torch.zeros((T, E), dtype=torch.int32, device=dev)
        self.t = 0                                              # number of transitions recorded per env so far
        self._draws, self._prepared, self._c_view = 0, {}, None  # sample_fused: draw counter, marshalled argument sets, ring view
        # per-slot views, built once: slicing nine tensors per env step costs more host time than the step kernel takes
        self._obs_v, self._final_v, self._action_v = list(self.obs.unbind(0)), list(self.final_obs.unbind(0)), list(self.action.unbind(0))
        self._slot_v = [SimpleNamespace(reward=self.reward[s], shaped_reward=self.shaped_reward[s], fear=self.fear[s],
                                        terminated=self.terminated[s], truncated=self.truncated[s], ended=self.ended[s],
                                        info=self.info[s]) for s in range(T)]

    # ---- views handed to BatchedGridWorld.reset / step
    def obs_slot(self, t: int) -> torch.Tensor:
        return self._obs_v[t % self.T]

    def final_slot(self, t: int) -> torch.Tensor:
        return self._final_v[t % self.T]

    def buffers_slot(self, t: int) -> SimpleNamespace:
        return self._slot_v[t % self.T]

    def action_slot(self, t: int) -> torch.Tensor:
        """[E, L, action_dim] f32: where the actor writes its continuous actions of time t (FusedActor.forward(cont_out=...))."""
        return self._action_v[t % self.T]

    def store_action(self, t: int, cont_actions: torch.Tensor):
        self._action_v[t % self.T].copy_(cont_actions.reshape(self.E, self.L, self.action_dim))

    def advance(self):
        """Call once per env.step after the transition of time self.t has been written."""
        self.t += 1

    def __len__(self) -> int:
        return min(self.t, self.T - 1) * self.E

    # ---- sampling in one kernel (csrc/gw_replay.cu): draws the indices and gathers the batch on the device
    def _view(self):
        import ctypes as C
        from . import _native as N
        v = self._c_view
        if v is None:
            v = N.GwReplayView()
            v.struct_size = C.sizeof(N.GwReplayView)
            v.obs_dtype = {torch.float32: N.GW_OBS_F32, torch.bfloat16: N.GW_OBS_BF16}[self.obs.dtype]
            v.slots, v.num_envs, v.n_learners, v.obs_len, v.action_dim = self.T, self.E, self.L, self.obs_len, self.action_dim
            v.obs, v.final_obs, v.action = self.obs.data_ptr(), self.final_obs.data_ptr(), self.action.data_ptr()
            v.reward, v.terminated, v.ended = self.shaped_reward.data_ptr(), self.terminated.data_ptr(), self.ended.data_ptr()
            self._c_view = v
        return v

    def _same_device(self, t: torch.Tensor) -> bool:
        return t.device.type == self.device.type and (self.device.index is None or t.device.index is None
                                                       or t.device.index == self.device.index)

    _FIELDS = ("state", "action", "reward", "next_state", "done")

    def new_batch(self, batch_size: int) -> Dict[str, torch.Tensor]:
        """f32 batch tensors of the shapes `sample_fused` fills (what BatchedMADDPG.learn consumes)."""
        B, L, dev = int(batch_size), self.L, self.device
        f = dict(dtype=torch.float32, device=dev)
        return {"state": torch.empty((B, L, self.obs_len), **f), "action": torch.empty((B, L, self.action_dim), **f),
                "reward": torch.empty((B, L), **f), "next_state": torch.empty((B, L, self.obs_len), **f),
                "done": torch.empty((B, L), **f), "t": torch.empty((B,), dtype=torch.int64, device=dev),
                "env": torch.empty((B,), dtype=torch.int64, device=dev)}

    def sample_fused(self, env, batch_size: int, seed: int = 0, out: Optional[Dict[str, torch.Tensor]] = None,
                     indices=None) -> Dict[str, torch.Tensor]:
        """`memory.sample(BATCH_SIZE)` (maddpg/agent.py:209-211) as ONE kernel launch: gw_replay_sample draws
        `batch_size` (time, env) pairs with Philox(seed, sample, draw number) -- uniform over the stored transitions,
        with replacement -- and gathers the five fields into `out` (f32; `new_batch` makes one).  `indices` =
        (t_abs, env) int64 device tensors to gather instead of drawing.  `env` is the BatchedGridWorld that owns the
        library handle; there is no PyTorch fallback."""
        import ctypes as C
        from . import _native as N
        if self.device.type != "cuda":
            raise RuntimeError("sample_fused needs the CUDA library (ring on a CUDA device); there is no CPU path")
        B = int(batch_size)
        out = self.new_batch(B) if out is None else out
        p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        prep = self._prepared.get(id(out))
        if (prep is None or prep[0] is not out or prep[1] != B
                or any(out[k].data_ptr() != q for k, q in zip(self._FIELDS, prep[4]))):   # checked once per set of batch tensors
            for k, shape in (("state", (B, self.L, self.obs_len)), ("action", (B, self.L, self.action_dim)), ("reward", (B, self.L)),
                             ("next_state", (B, self.L, self.obs_len)), ("done", (B, self.L))):
                t = out[k]
                if t.dtype != torch.float32 or tuple(t.shape) != shape or not t.is_contiguous() or not self._same_device(t):
                    raise ValueError(f"sample_fused: out[{k!r}] must be a contiguous f32 tensor of shape {shape} on {self.device}")
            for k in ("t", "env"):
                t = out.get(k)
                if t is not None and (t.dtype != torch.int64 or tuple(t.shape) != (B,) or not t.is_contiguous() or not self._same_device(t)):
                    raise ValueError(f"sample_fused: out[{k!r}] must be a contiguous int64 vector of length {B} on {self.device}")
            prep = (out, B, tuple(p(out.get(k)) for k in ("state", "action", "reward", "next_state", "done", "t", "env")),
                    C.byref(self._view()), tuple(out[k].data_ptr() for k in self._FIELDS))
            if len(self._prepared) > 64:
                self._prepared.clear()
            self._prepared[id(out)] = prep
        t_in = env_in = None
        if indices is not None:
            t_in, env_in = (x.to(device=self.device, dtype=torch.int64).contiguous() for x in indices)
            if t_in.shape != (B,) or env_in.shape != (B,):
                raise ValueError("sample_fused: indices must be two int64 vectors of length batch_size")
        self._draws += indices is None
        N.check(env.lib.gw_replay_sample(env._h, prep[3], self.t, B, int(seed) & (2 ** 64 - 1), self._draws,
                                         p(t_in), p(env_in), *prep[2], env._stream()), env._h, "gw_replay_sample")
        return out

    # ---- sampling (uniform over the stored transitions, like random.sample over the deque)
    def sample(self, batch_size: int, generator: Optional[torch.Generator] = None) -> Dict[str, torch.Tensor]:
        n_t = min(self.t, self.T - 1)
        if n_t == 0:
            raise RuntimeError("replay ring is empty")
        dev = self.device
        k = torch.randint(0, n_t, (batch_size,), device=dev, generator=generator)
        e = torch.randint(0, self.E, (batch_size,), device=dev, generator=generator)
        t_abs = (self.t - 1) - k                                # the newest n_t time steps are valid
        s = t_abs % self.T
        s1 = (t_abs + 1) % self.T
        ended = self.ended[s, e].bool()
        nxt = torch.where(ended[:, None, None], self.final_obs[s, e], self.obs[s1, e])
        return {"state": self.obs[s, e], "action": self.action[s, e], "reward": self.shaped_reward[s, e],
                "next_state": nxt, "done": self.terminated[s, e], "ended": ended, "t": t_abs, "env": e}
