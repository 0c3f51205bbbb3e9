"""FusedActor: the actors' forward pass of the rollout loop in one CUDA kernel (K5; csrc/gw_actor.cu).

Replaces the `agent.get_action(states, training, infos)` call of maddpg/agent.py:109-113 for the batched env.  The
kernel never reads the rendered observation: it rebuilds the first layer from the 8-byte `obs_code` that gw_step
writes next to it (template + <= 5 special cells), runs the 128x128 layer on the tensor cores (tcgen05.mma, bf16
inputs, fp32 accumulation in TMEM) and finishes LayerNorm / ReLU / 128->9 / Gumbel-softmax / exploration noise /
action mask / arg-max in registers.  AgileRL is not available here, so there is no parity claim against it; the test
compares against the same network evaluated by PyTorch in fp32 (tolerance of one bf16 GEMM).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence

import numpy as np
import torch
import torch.nn as nn

from . import _native as N


def _linear_norm_layers(actor: nn.Sequential):
    lin = [m for m in actor if isinstance(m, nn.Linear)]
    ln = [m for m in actor if isinstance(m, nn.LayerNorm)]
    if len(lin) != 3 or len(ln) != 2 or lin[0].out_features != 128 or lin[1].out_features != 128 or lin[2].out_features != 9:
        raise ValueError("FusedActor supports the reference's actor shape only: Linear-LN-ReLU-Linear-LN-ReLU-Linear(9), hidden 128")
    return lin, ln


class FusedActor:
    def __init__(self, env, actors: Sequence[nn.Sequential], seed: int = 0):
        self.env, self.lib, self.seed = env, env.lib, int(seed)
        self.E, self.L = env.num_envs, env.n_learners
        if len(actors) != self.L:
            raise ValueError("one actor per learner")
        self._h = C.c_void_p()
        w, keep = self._pack(actors)
        N.check(self.lib.gw_actor_create(env._h, w, self.L, C.byref(self._h)), env._h, "gw_actor_create")
        del keep
        dev = env.device
        self.cont = torch.empty((self.E, self.L, 9), dtype=torch.float32, device=dev)
        self.ids = torch.empty((self.E, self.L), dtype=torch.int8, device=dev)
        self.step = 0

    def _pack(self, actors):
        arr = (N.GwActorWeights * self.L)()
        keep: List[np.ndarray] = []
        for k, actor in enumerate(actors):
            lin, ln = _linear_norm_layers(actor)
            if lin[0].in_features != self.env.obs_len:
                raise ValueError("actor input size does not match the observation length")
            vals = [lin[0].weight, lin[0].bias, ln[0].weight, ln[0].bias, lin[1].weight, lin[1].bias, ln[1].weight,
                    ln[1].bias, lin[2].weight, lin[2].bias]
            for name, t in zip(("w1", "b1", "ln1_g", "ln1_b", "w2", "b2", "ln2_g", "ln2_b", "w3", "b3"), vals):
                a = np.ascontiguousarray(t.detach().float().cpu().numpy())
                keep.append(a)
                setattr(arr[k], name, a.ctypes.data)
        return arr, keep

    def update(self, actors: Sequence[nn.Sequential]):
        """Refresh the kernel's packed copy of the weights (after a learn step).  Parameters that live on the env's
        device are packed there by a kernel, in stream order; anything else goes through the host."""
        tensors = []
        for actor in actors:
            lin, ln = _linear_norm_layers(actor)
            tensors.append([lin[0].weight, lin[0].bias, ln[0].weight, ln[0].bias, lin[1].weight, lin[1].bias, ln[1].weight,
                            ln[1].bias, lin[2].weight, lin[2].bias])
        if len(actors) == self.L and all(t.device == self.env.device and t.dtype == torch.float32 and t.is_contiguous()
                                         for ts in tensors for t in ts):
            arr = (N.GwActorWeights * self.L)()
            for k, ts in enumerate(tensors):
                for name, t in zip(("w1", "b1", "ln1_g", "ln1_b", "w2", "b2", "ln2_g", "ln2_b", "w3", "b3"), ts):
                    setattr(arr[k], name, t.data_ptr())
            N.check(self.lib.gw_actor_update_device(self._h, arr, self.L, self.env._stream()), self.env._h, "gw_actor_update_device")
            return
        w, keep = self._pack(actors)
        N.check(self.lib.gw_actor_update(self._h, w, self.L, self.env._stream()), self.env._h, "gw_actor_update")
        del keep

    def forward(self, obs_code: torch.Tensor, action_mask: torch.Tensor = None, training: bool = True,
                expl_noise: float = 0.1, mean_noise: float = 0.0, gumbel: bool = None, cont_out: torch.Tensor = None):
        """-> (cont_actions f32 [E, L, 9], action_ids int8 [E, L]); both are views of buffers reused by the next call.
        training=True: Gumbel noise of the output activation + Gaussian exploration noise.  training=False: no noise at
        all, unless gumbel=True, which is the reference's evaluation mode (its GumbelSoftmax activation samples on every
        forward, `training=False` only drops the exploration noise)."""
        mode = 1 if training else (2 if gumbel else 0)
        cont = self.cont
        if cont_out is not None:                           # e.g. the replay ring's action slot: no copy afterwards
            if cont_out.dtype != torch.float32 or tuple(cont_out.shape) != (self.E, self.L, 9) or not cont_out.is_contiguous() \
                    or cont_out.device != self.cont.device:
                raise ValueError("cont_out must be a contiguous float32 [E, L, 9] tensor on the env's device")
            cont = cont_out
        if obs_code.dtype != torch.int64 or obs_code.numel() != self.E or not obs_code.is_cuda:
            raise ValueError("obs_code must be the int64 [E] tensor written by BatchedGridWorld.step / reset")
        mptr = None
        if action_mask is not None:
            if action_mask.dtype != torch.int8 or tuple(action_mask.shape) != (self.E, self.L, 9) or not action_mask.is_contiguous():
                raise ValueError("action_mask must be a contiguous int8 [E, L, 9] tensor")
            mptr = action_mask.data_ptr()
        N.check(self.lib.gw_actor_forward(self._h, self.E, obs_code.data_ptr(), mptr, cont.data_ptr(),
                                          self.ids.data_ptr(), mode, float(expl_noise), float(mean_noise),
                                          self.seed, self.step, self.env._stream()), self.env._h, "gw_actor_forward")
        self.step += 1
        return cont, self.ids

    def close(self):
        if getattr(self, "_h", None):
            h, self._h = self._h, None
            self.lib.gw_actor_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
