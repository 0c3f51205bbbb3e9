"""FusedLearner: `agent.learn(experiences)` (maddpg/agent.py:209-213, :218-224) as ONE kernel (csrc/gw_maddpg.cu).

The reference learns `num_envs // LEARN_STEP` times after every vector step (maddpg/agent.py:214-224); with 4 096
environments that is 409 updates per env step, so the update -- not the env -- is the cost of training.  Round 1 ran
it as a CUDA graph of ~85 dependent PyTorch / cuBLAS kernels (321 us).  `gw_learner_update` runs whole updates inside one
persistent cooperative kernel: batch draw + gather from the device replay ring, target actors, TD target, critic
forward / backward, Adam, actor loss through the updated critic, actor backward, Adam, soft target update, for all agents,
`updates` times per launch.

This class owns the flat fp32 vectors the kernel works on (parameters, target parameters, Adam moments, gradients) and
re-seats the `nn.Parameter`s of a `BatchedMADDPG` -- and the moments of its torch optimisers -- as VIEWS into them, so
everything that reads the modules (checkpoints, FusedActor.update, get_action, evaluate) sees the kernel's weights
without a copy.  There is no PyTorch fallback: without the CUDA library the constructor raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import torch
import torch.distributed as dist

from . import _native as N


class FusedLearner:
    phase_names = ("gather", "l1", "l2", "heads", "ct_l2", "td", "c_bwd2", "c_ln1", "c_dw1", "adam_c", "c2_l1", "c2_l2", "aloss",
                   "c2_dh1", "act_bwd", "a_bwd2", "a_ln1", "a_dw1", "adam_a")   # csrc/gw_maddpg.cu enum Phase
    n_phases = len(phase_names)

    def __init__(self, env, agent, batch_size: Optional[int] = None, seed: int = 0):
        """env: the BatchedGridWorld that owns the library handle and the stream; agent: a BatchedMADDPG on env.device
        with the reference's network shapes (hidden [128, 128], 9 actions)."""
        if env.device.type != "cuda":
            raise RuntimeError("FusedLearner needs the CUDA library (env on a CUDA device); there is no CPU path")
        self.env, self.lib, self.agent = env, env.lib, agent
        hp = agent.hp
        self.n, self.obs_len, self.act_dim = agent.n, agent.obs_dim, agent.act_dim
        self.B = int(hp["BATCH_SIZE"] if batch_size is None else batch_size)
        cfg = N.GwLearnerConfig()
        cfg.struct_size = C.sizeof(N.GwLearnerConfig)
        cfg.n_agents, cfg.obs_len, cfg.action_dim, cfg.batch = self.n, self.obs_len, self.act_dim, self.B
        cfg.lr_actor, cfg.lr_critic = float(hp["LR_ACTOR"]), float(hp["LR_CRITIC"])
        cfg.gamma, cfg.tau = float(hp["GAMMA"]), float(hp["TAU"])
        cfg.beta1, cfg.beta2, cfg.adam_eps, cfg.ln_eps = 0.9, 0.999, 1e-8, 1e-5
        cfg.seed = int(seed) & (2 ** 64 - 1)
        lay = N.GwLearnerLayout()
        lay.struct_size = C.sizeof(N.GwLearnerLayout)
        rc = self.lib.gw_learner_layout_of(C.byref(cfg), C.byref(lay))
        if rc != 0:
            raise ValueError("FusedLearner: unsupported shape (1..2 agents, obs_len a multiple of 16 whose critic input row "
                             "fits the kernel's shared-memory staging (about 400 cells for two learners), 9 actions, "
                             "BATCH_SIZE a multiple of 32 and <= 512)")
        self.cfg, self.layout = cfg, lay
        dev, P = env.device, int(lay.param_floats)
        f = dict(dtype=torch.float32, device=dev)
        self.params, self.targets = torch.zeros(P, **f), torch.zeros(P, **f)
        self.adam_m, self.adam_v, self.grads = torch.zeros(P, **f), torch.zeros(P, **f), torch.zeros(P, **f)
        self.adam_steps = torch.zeros(2 * self.n, **f)
        self.scratch = torch.zeros(int(lay.scratch_bytes), dtype=torch.uint8, device=dev)
        self.offsets = [int(lay.net_offset[k]) for k in range(2 * self.n)]
        self.sizes = [int(lay.net_params[k]) for k in range(2 * self.n)]
        self.critic_lo = self.offsets[self.n]                    # grads[:critic_lo] = all actors, grads[critic_lo:] = all critics
        self._adopt(agent)
        buf = N.GwLearnerBuffers()
        for name in ("params", "targets", "adam_m", "adam_v", "grads", "adam_steps", "scratch"):
            setattr(buf, name, getattr(self, name).data_ptr())
        self._h = C.c_void_p()
        N.check(self.lib.gw_learner_create(env._h, C.byref(cfg), C.byref(buf), C.byref(self._h)), env._h, "gw_learner_create")
        self.updates_done = 0
        self._loss_buf: Optional[torch.Tensor] = None
        agent.learner = self                                     # the agent's parameters live in this learner's vectors from now on:
                                                                 # `agent.learn` routes here, checkpoints export its step counters
        self.force_segmented = False                             # tests: take the several-rank path (three launches per update) on one rank
        self.peer_world = 1                                      # > 1 once connect_peers() succeeded: gradients exchanged inside the kernel

    # ---- the modules' parameters and the optimisers' moments become views of the flat vectors
    def _adopt(self, agent):
        nets = list(agent.actors) + list(agent.critics)
        tnets = list(agent.actor_targets) + list(agent.critic_targets)
        opts = list(agent.actor_opt) + list(agent.critic_opt)
        want = ("weight", "bias")
        self._step_tensors: List[List[torch.Tensor]] = []
        with torch.no_grad():
            for k, (net, tnet, opt) in enumerate(zip(nets, tnets, opts)):
                ps, tps = list(net.parameters()), list(tnet.parameters())
                if sum(p.numel() for p in ps) != self.sizes[k] or len(ps) != 10 or not all(n.split(".")[-1] in want for n, _ in net.named_parameters()):
                    raise ValueError("FusedLearner supports the reference's networks only: Linear-LayerNorm-ReLU x2 - Linear, hidden 128")
                off, steps = self.offsets[k], []
                for p, tp in zip(ps, tps):
                    m = p.numel()
                    for flat, t in ((self.params, p), (self.targets, tp)):
                        view = flat[off:off + m].view_as(t)
                        view.copy_(t.data)
                        t.data = view
                    st = opt.state.get(p)
                    mv, vv = self.adam_m[off:off + m].view_as(p), self.adam_v[off:off + m].view_as(p)
                    if st is not None and "exp_avg" in st:         # resumed from a checkpoint: keep the moments and the step count
                        mv.copy_(st["exp_avg"])
                        vv.copy_(st["exp_avg_sq"])
                        self.adam_steps[k] = float(st["step"])
                        step = st["step"] if torch.is_tensor(st["step"]) else torch.tensor(float(st["step"]))
                    else:
                        step = torch.zeros((), dtype=torch.float32, device=p.device if opt.param_groups[0].get("capturable") else "cpu")
                    opt.state[p] = {"step": step, "exp_avg": mv, "exp_avg_sq": vv}
                    steps.append(step)
                    off += m
                self._step_tensors.append(steps)

    KERNELS = {"auto": N.GW_LEARN_KERNEL_AUTO, "phase": N.GW_LEARN_KERNEL_PHASE, "cluster": N.GW_LEARN_KERNEL_CLUSTER}

    def set_kernel(self, kind: str):
        """"cluster" (csrc/gw_maddpg_cluster.cu: row-block clusters, 3xTF32 tensor-core tiles, 4 grid barriers per update),
        "phase" (csrc/gw_maddpg.cu: fp32 FMA, 19 grid-wide steps, every supported shape) or "auto" (cluster where the
        shape and the device allow it).  Raises if "cluster" is not possible for this learner."""
        N.check(self.lib.gw_learner_set_kernel(self._h, self.KERNELS[kind]), self.env._h, "gw_learner_set_kernel")

    @property
    def kernel(self) -> str:
        return {v: k for k, v in self.KERNELS.items()}[self.lib.gw_learner_kernel(self._h)]

    def connect_peers(self) -> bool:
        """Data-parallel run, one process per GPU of one node: exchange the gradients INSIDE the update kernel over NVLink
        peer memory instead of two NCCL all-reduces per update (gw_learner_peer_export / _connect: every rank's exchange
        block is shared as a CUDA IPC handle, gathered here with torch.distributed).  Collective: every rank calls it.
        Returns False (nothing changed, the NCCL path stays) when there is one rank, more than 8, or the cluster kernel is
        not the active one."""
        if not (dist.is_available() and dist.is_initialized()) or self.peer_world > 1:
            return self.peer_world > 1
        world, rank = dist.get_world_size(), dist.get_rank()
        dev = self.params.device

        def all_ok(flag: bool) -> bool:                            # every rank or none (a rank that failed must not leave the others waiting)
            t = torch.tensor([1 if flag else 0], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            return bool(int(t.item()))

        mine = N.GwPeerHandle()
        ok = 2 <= world <= N.GW_MAX_PEERS and self.kernel == "cluster"
        if ok:
            ok = self.lib.gw_learner_peer_export(self._h, C.byref(mine)) == 0
        if not all_ok(ok):
            return False
        t = torch.tensor(list(bytes(mine.bytes)), dtype=torch.uint8, device=dev)
        gathered = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(gathered, t)
        arr = (N.GwPeerHandle * world)()
        for q, g in enumerate(gathered):
            C.memmove(arr[q].bytes, bytes(g.cpu().tolist()), 64)
        ok = self.lib.gw_learner_peer_connect(self._h, rank, world, arr) == 0      # fails without peer access between the GPUs
        if not all_ok(ok):                                         # (all_reduce: also the barrier after the mapping)
            self.lib.gw_learner_peer_disable(self._h)
            return False
        self.peer_world = world
        return True

    def peer_timed_out(self) -> bool:
        """True if a peer failed to arrive at an in-kernel exchange in time (the update then ran without it)."""
        w, e = C.c_int32(), C.c_int32()
        N.check(self.lib.gw_learner_peer_status(self._h, C.byref(w), C.byref(e)), self.env._h, "gw_learner_peer_status")
        return bool(e.value)

    def export_steps(self):
        """Write the kernel's per-network Adam step counters into the torch optimisers' per-parameter `step` entries
        (the moments are shared storage already): call before reading `optimizer.state_dict()` (checkpoints)."""
        vals = self.adam_steps.tolist()
        for k, steps in enumerate(self._step_tensors):
            for s in steps:
                s.fill_(vals[k])

    # ---- one update on an explicit batch (tests, the reference's `learn(experiences)` call shape)
    def learn(self, batch: Dict[str, torch.Tensor], gumbel_next: Optional[torch.Tensor] = None,
              gumbel_cur: Optional[torch.Tensor] = None, segment: int = N.GW_LEARN_ALL, grad_scale: float = 1.0,
              losses: Optional[torch.Tensor] = None) -> torch.Tensor:
        """batch: state / next_state [B, n, obs_len], action [B, n, 9], reward / done [B, n] (device; cast to f32 here).
        gumbel_*: [B, n, 9] noise to use instead of the kernel's own draws.  Returns losses f32 [1, 2, n]
        (actor loss, critic loss per agent), a device tensor."""
        B, n = self.B, self.n
        shapes = {"state": (B, n, self.obs_len), "action": (B, n, self.act_dim), "reward": (B, n),
                  "next_state": (B, n, self.obs_len), "done": (B, n)}
        keep, gb = [], N.GwLearnBatch()
        for k, shape in shapes.items():
            t = batch[k]
            if tuple(t.shape) != shape or t.device != self.params.device:
                raise ValueError(f"FusedLearner.learn: batch[{k!r}] must have shape {shape} on {self.params.device}")
            t = t.to(torch.float32).contiguous()
            keep.append(t)
            setattr(gb, k, t.data_ptr())
        for name, g in (("gumbel_next", gumbel_next), ("gumbel_cur", gumbel_cur)):
            if g is not None:
                if tuple(g.shape) != (B, n, self.act_dim):
                    raise ValueError(f"{name} must be [B, n, 9]")
                g = g.to(device=self.params.device, dtype=torch.float32).contiguous()
                keep.append(g)
                setattr(gb, name, g.data_ptr())
        losses = self._losses(1) if losses is None else losses
        N.check(self.lib.gw_learner_update(self._h, C.byref(gb), None, 0, 0, 0, 1, int(segment), float(grad_scale),
                                           losses.data_ptr(), self.env._stream()), self.env._h, "gw_learner_update")
        self._keep = keep                                        # inputs stay alive until the next call (stream order)
        if segment in (N.GW_LEARN_ALL, N.GW_LEARN_FINISH):
            self.updates_done += 1
        return losses

    def _losses(self, updates: int) -> torch.Tensor:
        if self._loss_buf is None or self._loss_buf.shape[0] < updates:
            self._loss_buf = torch.zeros((max(updates, 16), 2, self.n), dtype=torch.float32, device=self.params.device)
        return self._loss_buf[:updates]

    # ---- `updates` whole updates in one launch, batches drawn from the ring inside the kernel
    def learn_from_ring(self, ring, updates: int, sample_seed: int) -> torch.Tensor:
        """Update u draws its batch exactly like `ring.sample_fused(seed=sample_seed)` would on its u-th call from
        here (same Philox key and draw number) and gathers it inside the kernel.  Several ranks: after `connect_peers`
        still ONE launch (the gradients are exchanged inside the kernel over NVLink peer memory); otherwise every update is
        cut at the two gradient exchanges (three launches, two NCCL all-reduces on the flat gradient vector's critic /
        actor part).  Returns losses f32 [updates, 2, n] on the device."""
        updates = int(updates)
        losses = self._losses(updates)
        view = C.byref(ring._view())
        seed = int(sample_seed) & (2 ** 64 - 1)
        stream, h = self.env._stream(), self.env._h
        multi = dist.is_available() and dist.is_initialized()
        world = dist.get_world_size() if multi else 1
        if (world == 1 or self.peer_world == world) and not self.force_segmented:   # one launch; several ranks: in-kernel exchange
            N.check(self.lib.gw_learner_update(self._h, None, view, ring.t, seed, ring._draws + 1, updates, N.GW_LEARN_ALL, 1.0,
                                               losses.data_ptr(), stream), h, "gw_learner_update")
        else:
            g_act, g_cri, scale = self.grads[:self.critic_lo], self.grads[self.critic_lo:], 1.0 / world
            for u in range(updates):
                lp = losses[u].data_ptr()
                for seg, flat in ((N.GW_LEARN_CRITIC_GRADS, g_cri), (N.GW_LEARN_ACTOR_GRADS, g_act), (N.GW_LEARN_FINISH, None)):
                    N.check(self.lib.gw_learner_update(self._h, None, view, ring.t, seed, ring._draws + 1 + u, 1, seg, scale,
                                                       lp, stream), h, "gw_learner_update")
                    if flat is not None and multi:
                        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        ring._draws += updates
        self.updates_done += updates
        return losses

    def debug_tensor(self, name: str, index: int = 0) -> torch.Tensor:
        """A copy of one of the kernel's intermediate tensors of the last update (flat f32), see gw_learner_debug_ptr."""
        ptr, cnt = C.c_void_p(), C.c_int64()
        N.check(self.lib.gw_learner_debug_ptr(self._h, name.encode(), int(index), C.byref(ptr), C.byref(cnt)), self.env._h,
                "gw_learner_debug_ptr")
        base = self.scratch.data_ptr()
        off = ptr.value - base
        return self.scratch[off:off + 4 * cnt.value].view(torch.float32).clone()

    def close(self):
        if getattr(self, "_h", None):
            h, self._h = self._h, None
            self.lib.gw_learner_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
