"""Batched MADDPG on the device: the trainer-side twin of maddpg/agent.py for E parallel environments.

The reference delegates the algorithm to AgileRL 1.0.15 (`MADDPG`, `MultiAgentReplayBuffer`), a third-party package
that is neither vendored in the reference tree nor installed here, so **parity with it is unpinned**: this module
follows what the reference itself pins down --

  * network shapes, read from the shipped checkpoints: actor
    Linear(160,128)-LayerNorm-ReLU-Linear(128,128)-LayerNorm-ReLU-Linear(128,9)-GumbelSoftmax (38 793 parameters),
    critic Linear(2*160+2*9,128)-LayerNorm-ReLU-Linear(128,128)-LayerNorm-ReLU-Linear(128,1) (60 545 parameters);
  * hyper-parameters and their names (configs/custom*.yaml: EXPL_NOISE, GAMMA, TAU, LR_ACTOR, LR_CRITIC, BATCH_SIZE,
    LEARN_STEP, MEMORY_SIZE, FeAR_weight, WITH_FEAR, TRAIN_STEPS);
  * the loop of MADDPGAgent.train (maddpg/agent.py:77-252): act -> env.step -> reward = FeAR_weight*fear + reward
    (:128-131) -> store (state, continuous action, reward, next_state, termination) (:190-197) -> learn every
    LEARN_STEP environment steps once BATCH_SIZE transitions exist (:201-222);

and the textbook MADDPG update (centralised critics on all states and actions, actor loss -Q, soft target update).
Everything runs on the GPU: observations are written by gw_step straight into the replay ring, the actors read them
from there, nothing crosses to the host inside the loop.  Multi-GPU: env shards per rank, gradients all-reduced.
"""
from __future__ import annotations

import copy
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

DEFAULT_HP = {           # configs/custom.yaml:1-26 (custom_fear_*.yaml differ in WITH_FEAR / FeAR_weight / MAX_EPISODES)
    "CUSTOM_ENV": True, "SEED": 42, "MAX_EPISODES": 30000, "TRAIN_STEPS": 150, "CHANNELS_LAST": False,
    "DISCRETE_ACTIONS": True, "ARCH": "mlp", "O_U_NOISE": False, "EXPL_NOISE": 0.1, "MEAN_NOISE": 0.0, "THETA": 0.15,
    "DT": 0.01, "LR_ACTOR": 0.001, "LR_CRITIC": 0.001, "GAMMA": 0.98, "MEMORY_SIZE": 200000, "LEARN_STEP": 10,
    "TAU": 0.01, "BATCH_SIZE": 128, "WITH_FEAR": False, "FeAR_weight": -2.0,
}


def preset(name: str = "custom") -> Dict:
    """Hyper-parameter sets of configs/custom.yaml and configs/custom_fear_{1,3,5,10}.yaml."""
    hp = dict(DEFAULT_HP)
    if name == "custom":
        return hp
    if name.startswith("custom_fear_"):
        w = float(name.rsplit("_", 1)[1])
        hp.update(WITH_FEAR=True, FeAR_weight=-w, MAX_EPISODES=15000)
        return hp
    raise KeyError(name)


def load_yaml_config(path: str) -> Dict:
    """Read a reference-style YAML (same keys as configs/custom*.yaml); missing keys take the defaults."""
    import yaml
    with open(path) as f:
        hp = dict(DEFAULT_HP)
        hp.update(yaml.safe_load(f) or {})
    return hp


class GumbelSoftmax(nn.Module):
    """Output activation of the reference's actors (`mlp_output_activation='GumbelSoftmax'`): softmax of the logits
    perturbed with fresh Gumbel noise on every forward."""

    def forward(self, logits: torch.Tensor) -> torch.Tensor:
        u = torch.rand_like(logits).clamp_(1e-20, 1.0)
        g = -torch.log(-torch.log(u) + 1e-20)
        return F.softmax(logits + g, dim=-1)


class TrainOps:
    """The library's trainer-side kernels (csrc/gw_train_ops.cu) bound to one env handle (device, stream, error string)."""
    WIDTH = 128

    def __init__(self, env, linear_backward: bool = True):
        import ctypes as C
        from . import _native as N
        self.C, self.N, self.env, self.lib = C, N, env, env.lib
        self.use_linear_backward = bool(linear_backward)    # forward_mlp: Linear backward through gw_linear_backward

    def ln_relu_forward(self, x, gamma, beta, eps, y, mean, rstd):
        p, env = self.C.c_void_p, self.env
        self.N.check(self.lib.gw_ln_relu_forward(env._h, x.shape[0], x.shape[1], p(x.data_ptr()), p(gamma.data_ptr()),
                                                 p(beta.data_ptr()), float(eps), p(y.data_ptr()), p(mean.data_ptr()),
                                                 p(rstd.data_ptr()), env._stream()), env._h, "gw_ln_relu_forward")

    def linear_backward(self, dy, x, weight, dw, db, dx):
        p, env = self.C.c_void_p, self.env
        self.N.check(self.lib.gw_linear_backward(env._h, x.shape[0], x.shape[1], weight.shape[0], p(dy.data_ptr()),
                                                 p(x.data_ptr()), x.stride(0), p(weight.data_ptr()), p(dw.data_ptr()),
                                                 p(db.data_ptr()), p(dx.data_ptr()) if dx is not None else None,
                                                 env._stream()), env._h, "gw_linear_backward")

    def ln_relu_backward(self, dy, x, mean, rstd, gamma, beta, dx, dgamma, dbeta):
        p, env = self.C.c_void_p, self.env
        self.N.check(self.lib.gw_ln_relu_backward(env._h, x.shape[0], x.shape[1], p(dy.data_ptr()), p(x.data_ptr()),
                                                  p(mean.data_ptr()), p(rstd.data_ptr()), p(gamma.data_ptr()),
                                                  p(beta.data_ptr()), p(dx.data_ptr()), p(dgamma.data_ptr()),
                                                  p(dbeta.data_ptr()), env._stream()), env._h, "gw_ln_relu_backward")


class _Linear(torch.autograd.Function):
    """torch.nn.functional.linear whose backward is ONE library kernel (dW, db and dx; PyTorch: two GEMMs + a reduction)."""

    @staticmethod
    def forward(ctx, x, weight, bias, ops):
        ctx.save_for_backward(x, weight)
        ctx.ops = ops
        return F.linear(x, weight, bias)

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dy = dy.contiguous()
        dw = torch.empty_like(weight)
        db = torch.empty((weight.shape[0],), dtype=weight.dtype, device=weight.device)
        dx = torch.empty((x.shape[0], x.shape[1]), dtype=x.dtype, device=x.device) if ctx.needs_input_grad[0] else None
        ctx.ops.linear_backward(dy, x, weight, dw, db, dx)
        return dx, dw, db, None


class _LayerNormReLU(torch.autograd.Function):
    """relu(layer_norm(x)) over the last dimension (128): one library kernel forward, one backward."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps, ops):
        x = x.contiguous()
        y = torch.empty_like(x)
        mean = torch.empty((x.shape[0],), dtype=torch.float32, device=x.device)
        rstd = torch.empty_like(mean)
        ops.ln_relu_forward(x, gamma, beta, eps, y, mean, rstd)
        ctx.save_for_backward(x, gamma, beta, mean, rstd)
        ctx.ops = ops
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, beta, mean, rstd = ctx.saved_tensors
        dy = dy.contiguous()
        dx, dgamma, dbeta = torch.empty_like(x), torch.empty_like(gamma), torch.empty_like(beta)
        ctx.ops.ln_relu_backward(dy, x, mean, rstd, gamma, beta, dx, dgamma, dbeta)
        return dx, dgamma, dbeta, None, None


def forward_mlp(net: nn.Sequential, x: torch.Tensor, ops: Optional[TrainOps]) -> torch.Tensor:
    """net(x), with every LayerNorm(128) -> ReLU pair run by the fused kernels and every Linear's backward by the
    one-launch kernel when `ops` is given (CUDA, fp32).
    The modules and their parameter names stay what the checkpoints and the actor kernel expect."""
    if ops is None or not x.is_cuda:
        return net(x)
    mods, i = list(net), 0
    while i < len(mods):
        m = mods[i]
        if (isinstance(m, nn.LayerNorm) and i + 1 < len(mods) and isinstance(mods[i + 1], nn.ReLU) and m.elementwise_affine
                and tuple(m.normalized_shape) == (TrainOps.WIDTH,) and x.dtype == torch.float32 and x.dim() == 2):
            x = _LayerNormReLU.apply(x, m.weight, m.bias, m.eps, ops)
            i += 2
        elif (ops.use_linear_backward and isinstance(m, nn.Linear) and m.bias is not None and torch.is_grad_enabled() and x.dtype == torch.float32
              and x.dim() == 2 and x.stride(1) == 1 and x.stride(0) >= x.shape[1] and m.weight.is_contiguous()):
            x = _Linear.apply(x, m.weight, m.bias, ops)
            i += 1
        else:
            x = m(x)
            i += 1
    return x


def mlp(in_dim: int, hidden: Sequence[int], out_dim: int, out_act: Optional[nn.Module]) -> nn.Sequential:
    layers: List[nn.Module] = []
    d = in_dim
    for h in hidden:
        layers += [nn.Linear(d, h), nn.LayerNorm(h), nn.ReLU()]
        d = h
    layers.append(nn.Linear(d, out_dim))
    if out_act is not None:
        layers.append(out_act)
    return nn.Sequential(*layers)


@dataclass
class LearnStats:
    """Per-agent losses of one update, kept as device tensors (reading them would synchronise the stream)."""
    actor_loss: torch.Tensor
    critic_loss: torch.Tensor


class BatchedMADDPG:
    def __init__(self, n_agents: int = 2, obs_dim: int = 160, act_dim: int = 9, hidden: Sequence[int] = (128, 128),
                 hp: Optional[Dict] = None, device="cuda", seed: int = 0):
        self.hp = dict(DEFAULT_HP if hp is None else hp)
        self.n, self.obs_dim, self.act_dim = n_agents, obs_dim, act_dim
        self.device = torch.device(device)
        g = torch.Generator().manual_seed(seed)
        state = torch.random.get_rng_state()
        torch.manual_seed(int(torch.randint(0, 2**31 - 1, (1,), generator=g)))
        crit_in = n_agents * (obs_dim + act_dim)
        self.actors = [mlp(obs_dim, hidden, act_dim, GumbelSoftmax()).to(self.device) for _ in range(n_agents)]
        self.critics = [mlp(crit_in, hidden, 1, None).to(self.device) for _ in range(n_agents)]
        torch.random.set_rng_state(state)
        self.actor_targets = [copy.deepcopy(a) for a in self.actors]
        self.critic_targets = [copy.deepcopy(c) for c in self.critics]
        for net in self.actor_targets + self.critic_targets:
            for p in net.parameters():
                p.requires_grad_(False)
        cap = self.device.type == "cuda"                   # step counters on the device: the update can live in a CUDA graph
        kw = dict(capturable=True, fused=True) if cap else {}   # fused: one kernel per optimiser step instead of ~10
        self.actor_opt = [torch.optim.Adam(a.parameters(), lr=self.hp["LR_ACTOR"], **kw) for a in self.actors]
        self.critic_opt = [torch.optim.Adam(c.parameters(), lr=self.hp["LR_CRITIC"], **kw) for c in self.critics]
        self._graph = None                                 # (CUDAGraph, static batch, static LearnStats) once captured
        self._segments = None                              # several ranks: the update as a chain of graphs (see _learn_segmented)
        self._eager_learns = 0
        self.parallel_agents = True                        # one-GPU graph: the agents' updates are parallel branches (see _learn)
        self._side_streams: Dict[object, List[torch.cuda.Stream]] = {}
        self.ops: Optional[TrainOps] = None                # fused LayerNorm+ReLU kernels of the library (attach_ops)
        self.force_segmented = False                       # tests: take the multi-rank path in a one-rank process group
        self.learner = None                                # FusedLearner (learner.py): the update as one kernel, see attach_learner

    def attach_learner(self, env, seed: int = 0):
        """Run every update through gw_learner_update (csrc/gw_maddpg.cu).  The parameters of the modules and the moments
        of the optimisers become views into the learner's flat vectors; `learn` routes there from now on."""
        from .learner import FusedLearner
        self.learner = FusedLearner(env, self, seed=seed)
        self._graph, self._segments, self._eager_learns = None, None, 0
        return self.learner

    def attach_ops(self, env, linear_backward: bool = False):
        """Run the update's LayerNorm + ReLU pairs through the library's kernels (`env` provides handle and stream);
        `linear_backward`: also every Linear's backward as one launch (measured: no gain over cuBLAS's three kernels at
        batch 128 -- 145.0 vs 144.8 M agent-steps/s -- hence off by default).  Invalidates captured graphs."""
        self.ops = TrainOps(env, linear_backward=linear_backward) if env is not None else None
        self._graph, self._segments, self._eager_learns = None, None, 0

    def _fwd(self, net: nn.Sequential, x: torch.Tensor) -> torch.Tensor:
        return forward_mlp(net, x, self.ops)

    def parameters(self):
        for net in self.actors + self.critics:
            yield from net.parameters()

    def broadcast_parameters(self, src: int = 0):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            for net in self.actors + self.critics + self.actor_targets + self.critic_targets:
                for p in net.parameters():
                    dist.broadcast(p.data, src)

    # ---- acting (maddpg/agent.py:109-122): continuous 9-vector + masked arg-max action id
    @torch.no_grad()
    def get_action(self, obs: torch.Tensor, action_mask: Optional[torch.Tensor] = None, training: bool = True):
        """obs [E, n, obs_dim] (any float dtype), action_mask int8 [E, n, 9] -> (cont [E, n, 9] f32, ids int8 [E, n])."""
        cont = []
        for k, actor in enumerate(self.actors):
            a = actor(obs[:, k].float())
            if training:                                   # Gaussian exploration, EXPL_NOISE / MEAN_NOISE
                a = (a + self.hp["MEAN_NOISE"] + self.hp["EXPL_NOISE"] * torch.randn_like(a)).clamp_(0.0, 1.0)
            cont.append(a)
        cont = torch.stack(cont, dim=1)
        scores = cont if action_mask is None else cont.masked_fill(action_mask == 0, float("-inf"))
        return cont, scores.argmax(dim=-1).to(torch.int8)

    # ---- learning (textbook MADDPG; AgileRL's exact variant is not available here: parity unpinned)
    def learn(self, batch: Dict[str, torch.Tensor], graph: bool = True) -> LearnStats:
        """One MADDPG update.  On a single GPU the whole update (4 forward/backward passes, 4 Adam steps, soft update:
        ~250 small kernels, 6 ms of launches in eager mode) is captured into one CUDA graph after three eager calls
        and replayed from then on; with several ranks it is a chain of graphs cut at the gradient all-reduces."""
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        if self.learner is not None and not multi:         # one kernel launch (several ranks: FusedLearner.learn_from_ring)
            last = self.learner.learn(batch)[0].clone()
            return LearnStats(last[0], last[1])
        if not graph or self.device.type != "cuda":
            return self._learn(batch)
        if multi or self.force_segmented:
            return self._learn_segmented(batch)
        keys = ("state", "action", "reward", "next_state", "done")
        if self._graph is not None:
            g, static, out, shapes = self._graph
            if all(tuple(batch[k].shape) == shapes[k] for k in keys):
                for k in keys:
                    if batch[k] is not static[k]:          # the sampler may have written into the graph's inputs itself
                        static[k].copy_(batch[k])
                g.replay()
                return LearnStats(out.actor_loss.clone(), out.critic_loss.clone())
            self._graph, self._eager_learns = None, 0      # another batch shape: start over
        if self._eager_learns < 3:                         # warm-up (allocator, cuBLAS handles, Adam state) with real updates
            self._eager_learns += 1
            return self._learn(batch)
        static = {k: batch[k].float().clone() for k in keys}
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = self._learn(static, parallel=True)       # branches unless parallel_agents is off
        self._graph = (g, static, out, {k: tuple(batch[k].shape) for k in keys})
        g.replay()                                         # capture only records: this is the update for `batch`
        return LearnStats(out.actor_loss.clone(), out.critic_loss.clone())

    def static_inputs(self) -> Optional[Dict[str, torch.Tensor]]:
        """The captured update's input tensors (None before the capture): a sampler that fills them directly saves
        the five copies per update."""
        if self._graph is not None:
            return self._graph[1]
        if self._segments is not None:
            return self._segments["static"]
        return None

    # ---- several ranks: the same update cut into CUDA graphs at the gradient all-reduces
    def _learn_segmented(self, batch: Dict[str, torch.Tensor]) -> LearnStats:
        """Data-parallel update.  Capturing the NCCL all-reduces inside one graph hung (2 ranks, torch 2.11), so the update
        is cut where gradients are exchanged:  [critics forward/backward] AR [critic steps, actors forward/backward] AR
        [actor steps, soft update]  -- three graphs with the agents as parallel branches and two eager all-reduces, each
        on ONE flat gradient buffer (all critics / all actors; the parameters' .grad are views into it: no cat / split
        kernels).  (First version: per agent, five graphs and four all-reduces, 0.98 ms per update on two ranks.)"""
        keys = ("state", "action", "reward", "next_state", "done")
        seg = self._segments
        if seg is not None and not all(tuple(batch[k].shape) == seg["shapes"][k] for k in keys):
            seg = self._segments = None
            self._eager_learns = 0
        if seg is None:
            if self._eager_learns < 3:                     # warm-up with real (eager) updates
                self._eager_learns += 1
                return self._learn(batch)
            seg = self._segments = self._capture_segments({k: batch[k].float().clone() for k in keys},
                                                           {k: tuple(batch[k].shape) for k in keys})
        else:
            for k in keys:
                if batch[k] is not seg["static"][k]:
                    seg["static"][k].copy_(batch[k])
        world = dist.get_world_size()
        for g, flat in zip(seg["graphs"], seg["reduce_after"]):
            g.replay()
            if flat is not None:
                dist.all_reduce(flat, op=dist.ReduceOp.SUM)
                flat.div_(world)
        return LearnStats(seg["a_loss"].clone(), seg["c_loss"].clone())

    def _flat_grads(self, nets: Sequence[nn.Module]):
        """One flat gradient buffer for all of `nets`; every parameter's .grad becomes a view into it (kept from then
        on).  Returns (buffer, [the slice of each net])."""
        ps = [list(net.parameters()) for net in nets]
        flat = torch.zeros(sum(p.numel() for l in ps for p in l), dtype=ps[0][0].dtype, device=ps[0][0].device)
        off, parts = 0, []
        for l in ps:
            lo = off
            for p in l:
                p.grad = flat[off:off + p.numel()].view_as(p)
                off += p.numel()
            parts.append(flat[lo:off])
        return flat, parts

    def _fork_join(self, fns, key="agents", enabled: bool = True):
        """Record every fn of `fns`: the first on the current stream, the others on forked streams (parallel branches of
        the graph being captured), joined before returning; in turn when branching is off.  `key` names the set of side
        streams, so that an inner fork inside one agent's branch does not share streams with another agent's."""
        if not (enabled and self.parallel_agents and len(fns) > 1 and self.device.type == "cuda"):
            for fn in fns:
                fn()
            return
        cur = torch.cuda.current_stream(self.device)
        side = self._side_streams.setdefault(key, [])
        while len(side) < len(fns) - 1:
            side.append(torch.cuda.Stream(self.device))
        side = side[:len(fns) - 1]
        for st in side:                                     # fork here, before the first branch's work is recorded
            st.wait_stream(cur)
        for i, fn in enumerate(fns):
            with torch.cuda.stream(cur if i == 0 else side[i - 1]):
                fn()
        for st in side:                                     # join
            cur.wait_stream(st)

    def _branches(self, fn, enabled: bool = True):
        """fn(k) for every agent, as parallel branches."""
        self._fork_join([lambda k=k: fn(k) for k in range(self.n)], "agents", enabled)

    def _capture_segments(self, static: Dict[str, torch.Tensor], shapes) -> Dict:
        """Three graphs: [all critics forward / backward] AR [critic steps, all actors forward / backward] AR [actor
        steps, soft update]; inside each the agents are parallel branches (an agent's update reads nothing another
        agent's writes, see _learn), and each exchange is ONE all-reduce over the flat gradients of all critics /
        all actors."""
        tau, n = self.hp["TAU"], self.n
        s, a, r, s2, done = (static[k] for k in ("state", "action", "reward", "next_state", "done"))
        B = s.shape[0]
        fa_all, fa = self._flat_grads(self.actors)
        fc_all, fc = self._flat_grads(self.critics)
        a_loss = torch.zeros(n, device=self.device)
        c_loss = torch.zeros(n, device=self.device)
        ctx = {}

        def prologue():
            ctx["flat_s"], flat_s2 = s.reshape(B, -1), s2.reshape(B, -1)
            with torch.no_grad():
                a2 = [None] * n
                self._branches(lambda j: a2.__setitem__(j, self._fwd(self.actor_targets[j], s2[:, j])))
                ctx["crit_in2"] = torch.cat([flat_s2, torch.stack(a2, dim=1).reshape(B, -1)], dim=1)
                ctx["crit_in"] = torch.cat([ctx["flat_s"], a.reshape(B, -1)], dim=1)
            fc_all.zero_()

        def critic_pass(k):                                # [critic forward / backward]; the target is an inner branch
            box = {}
            self._fork_join([lambda: box.__setitem__("q", self._fwd(self.critics[k], ctx["crit_in"]).squeeze(-1)),
                             lambda: box.__setitem__("target", self._td_target(k, ctx["crit_in2"], r, done))], ("inner", k))
            loss = F.mse_loss(box["q"], box["target"])
            loss.backward()
            c_loss[k] = loss.detach()

        def actor_pass(k):                                 # [critic step, actor forward / backward]
            self.critic_opt[k].step()
            a_new = a.clone()
            a_new[:, k] = self._fwd(self.actors[k], s[:, k])
            loss = -self._fwd(self.critics[k], torch.cat([ctx["flat_s"], a_new.reshape(B, -1)], dim=1)).mean()
            loss.backward()                                # also reaches the critic's gradients: zeroed before their next use
            a_loss[k] = loss.detach()

        def finish():
            with torch.no_grad():
                src = [p for net in self.actors + self.critics for p in net.parameters()]
                dst = [p for net in self.actor_targets + self.critic_targets for p in net.parameters()]
                torch._foreach_lerp_(dst, src, tau)

        pieces = [                                         # (functions run in this graph, buffer to all-reduce after it)
            ([prologue, lambda: self._branches(critic_pass)], fc_all),
            ([fa_all.zero_, lambda: self._branches(actor_pass)], fa_all),
            ([lambda: self._branches(lambda k: self.actor_opt[k].step()), finish], None),
        ]
        torch.cuda.synchronize(self.device)
        graphs, pool = [], None
        for fns, _ in pieces:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=pool):
                for fn in fns:
                    fn()
            pool = g.pool() if pool is None else pool
            graphs.append(g)
        seg = {"graphs": graphs, "reduce_after": [f for _, f in pieces], "static": static, "shapes": shapes,
               "a_loss": a_loss, "c_loss": c_loss, "ctx": ctx, "flat": (fa, fc)}
        # capture only records: replay now, with the exchanges, so that this call is the update for `batch`
        return seg

    def _learn(self, batch: Dict[str, torch.Tensor], parallel: bool = False) -> LearnStats:
        """`parallel` (only while capturing the one-GPU graph): agent k's critic / actor update touches nothing agent j's
        reads -- own critic, own actor, own optimisers; the other agents' actions come from the batch -- so the agents'
        passes are recorded on forked streams and become parallel branches of the graph.  Same kernels, same operands,
        same results as the sequential order; the ~120 launch-latency-sized kernels per agent overlap."""
        tau = self.hp["TAU"]
        s, a, r = batch["state"].float(), batch["action"].float(), batch["reward"].float()
        s2, done = batch["next_state"].float(), batch["done"].float()
        B = s.shape[0]
        flat_s, flat_s2 = s.reshape(B, -1), s2.reshape(B, -1)
        with torch.no_grad():
            a2 = [None] * self.n
            self._branches(lambda k: a2.__setitem__(k, self._fwd(self.actor_targets[k], s2[:, k])), parallel)
            crit_in2 = torch.cat([flat_s2, torch.stack(a2, dim=1).reshape(B, -1)], dim=1)
            crit_in = torch.cat([flat_s, a.reshape(B, -1)], dim=1)
        a_losses: List[Optional[torch.Tensor]] = [None] * self.n
        c_losses: List[Optional[torch.Tensor]] = [None] * self.n

        def agent_pass(k: int):
            box = {}

            def critic_update():                           # the TD target is an inner branch next to the critic's forward
                self._fork_join([lambda: box.__setitem__("q", self._fwd(self.critics[k], crit_in).squeeze(-1)),
                                 lambda: box.__setitem__("target", self._td_target(k, crit_in2, r, done))], ("inner", k), parallel)
                c_loss = F.mse_loss(box["q"], box["target"])
                self.critic_opt[k].zero_grad(set_to_none=True)
                c_loss.backward()
                self._allreduce_grads(self.critics[k])
                self.critic_opt[k].step()
                box["c_loss"] = c_loss.detach()

            # the actor's forward needs nothing of the critic's update: a third branch (its backward follows it there)
            self._fork_join([critic_update, lambda: box.__setitem__("a_k", self._fwd(self.actors[k], s[:, k]))], ("actor", k), parallel)
            c_loss = box["c_loss"]
            a_new = a.clone()
            a_new[:, k] = box["a_k"]
            a_loss = -self._fwd(self.critics[k], torch.cat([flat_s, a_new.reshape(B, -1)], dim=1)).mean()
            self.actor_opt[k].zero_grad(set_to_none=True)
            a_loss.backward()
            self._allreduce_grads(self.actors[k])
            self.actor_opt[k].step()
            a_losses[k], c_losses[k] = a_loss.detach(), c_loss.detach()

        self._branches(agent_pass, parallel)
        with torch.no_grad():                              # soft update, TAU (one multi-tensor kernel, not 40 small ones)
            src = [p for net in self.actors + self.critics for p in net.parameters()]
            dst = [p for net in self.actor_targets + self.critic_targets for p in net.parameters()]
            torch._foreach_lerp_(dst, src, tau)
        return LearnStats(torch.stack(a_losses), torch.stack(c_losses))

    @torch.no_grad()
    def _td_target(self, k: int, crit_in2: torch.Tensor, r: torch.Tensor, done: torch.Tensor) -> torch.Tensor:
        """r + GAMMA (1 - done) Q_target(next state, target actors' actions) for agent k."""
        q2 = self._fwd(self.critic_targets[k], crit_in2).squeeze(-1)
        return r[:, k] + self.hp["GAMMA"] * (1.0 - done[:, k]) * q2

    @staticmethod
    def _allreduce_grads(net: nn.Module):
        """Data parallel over env shards: average the gradients of one network with a single flat all-reduce
        (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return
        grads = [p.grad for p in net.parameters() if p.grad is not None]
        flat = torch.cat([g.reshape(-1) for g in grads])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat /= dist.get_world_size()
        off = 0
        for g in grads:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()


class BatchedTrainer:
    """MADDPGAgent.train (maddpg/agent.py:77-252) for E environments at once, everything on the device."""

    def __init__(self, env, agent: Optional[BatchedMADDPG] = None, hp: Optional[Dict] = None,
                 updates_per_learn: int = 1, seed: int = 0, fused_actor: bool = True, fused_sampler: bool = True, fused_ops: bool = True,
                 fused_linear_bwd: bool = False, learn_cadence: str = "reference", global_envs: Optional[int] = None,
                 fused_learner: bool = True, gradient_exchange: str = "peer"):
        """learn_cadence: "reference" = the rule of maddpg/agent.py:199-224 applied to the GLOBAL env count (see
        `learn_schedule`); "batched" = one block of `updates_per_learn` updates every LEARN_STEP vector steps whatever
        E is (round 1's loop: 1 update per LEARN_STEP * E transitions, a much lower update-to-data ratio than the
        reference's).  global_envs: environments over all ranks (default: this shard x world size); the schedule and
        the BATCH_SIZE gate are computed from it so that every rank takes the same decisions.
        fused_learner: the update as ONE kernel that also draws its batches from the ring (learner.py, csrc/gw_maddpg.cu);
        off = round 1's CUDA graph of PyTorch / library kernels (fused_sampler / fused_ops choose its pieces).
        gradient_exchange (several ranks, fused learner): "peer" = inside the update kernel over NVLink peer memory
        (FusedLearner.connect_peers, called by `broadcast_parameters`-time setup in `connect`), "nccl" = two all-reduces per
        update between the kernel's three segments."""
        from .replay import ReplayRing
        self.env = env
        # the library's replay / update kernels hang their device, stream and error state on a packed-layout handle; a world on
        # the general layout (larger map, more agents) lends them a one-env companion handle on the same device
        self.svc = getattr(env, "service_env", env)
        self.hp = dict(DEFAULT_HP if hp is None else hp)
        if learn_cadence not in ("reference", "batched"):
            raise ValueError("learn_cadence must be 'reference' or 'batched'")
        self.learn_cadence = learn_cadence
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.global_envs = int(global_envs) if global_envs is not None else env.num_envs * world
        self.min_shard = self.global_envs // world              # the smallest shard (sharding.shard_range): rank-invariant gate
        self.updates_done = 0
        self.agent = agent or BatchedMADDPG(env.n_learners, env.obs_len, 9, hp=self.hp, device=env.device, seed=seed)
        self.ring = ReplayRing(env.num_envs, env.n_learners, env.obs_len, self.hp["MEMORY_SIZE"], 9,
                               device=env.device, obs_dtype=env.obs_dtype)
        if gradient_exchange not in ("peer", "nccl"):
            raise ValueError("gradient_exchange must be 'peer' or 'nccl'")
        self.gradient_exchange = gradient_exchange
        self.learner = None
        if fused_learner and env.device.type == "cuda":
            try:
                self.learner = self.agent.learner if self.agent.learner is not None else self.agent.attach_learner(self.svc, seed=seed + 3)
            except ValueError as exc:                          # a shape the update kernel does not hold (e.g. 560-cell observations)
                import warnings
                warnings.warn(f"{exc}; the update runs as the CUDA graph of PyTorch / library kernels instead")
        if self.learner is None and fused_ops and env.device.type == "cuda" and self.agent.ops is None:
            self.agent.attach_ops(self.svc, linear_backward=fused_linear_bwd)
        self.updates_per_learn = int(updates_per_learn)
        self.gen = torch.Generator(device=env.device).manual_seed(seed + 1 + int(getattr(env, "env_id_base", 0)))
        self.fused_sampler = bool(fused_sampler) and env.device.type == "cuda"      # csrc/gw_replay.cu
        # replay draws: every rank samples its own shard, so the draw streams must differ between ranks (the actor's and the
        # env's noise are keyed by the global env id instead and do not depend on the sharding)
        self.sample_seed, self._batch = (seed + 1) ^ (int(getattr(env, "env_id_base", 0)) * 0x9E3779B97F4A7C15 & (2 ** 63 - 1)), None
        self.t = 0
        self.out = env.reset(obs_out=self.ring.obs_slot(0))
        self.losses: List[LearnStats] = []
        # acting: the fused CUDA kernel (csrc/gw_actor.cu, tensor cores) or the PyTorch modules (same weights)
        self.fused = None
        if fused_actor and env.device.type == "cuda" and env.obs_len == self.agent.obs_dim and not getattr(env, "_wide", False):
            from .actor import FusedActor
            self.fused = FusedActor(env, self.agent.actors, seed=seed + 2)

    def connect(self, src: int = 0) -> str:
        """Several ranks (collective, call once after construction): rank `src`'s networks to everybody and, with
        gradient_exchange="peer", the in-kernel gradient exchange over peer memory.  Returns what exchanges the gradients:
        "peer", "nccl", or "none" (one rank)."""
        self.agent.broadcast_parameters(src)
        if self.fused is not None:
            self.fused.update(self.agent.actors)
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return "none"
        if self.learner is not None and self.gradient_exchange == "peer" and self.learner.connect_peers():
            return "peer"
        return "nccl"

    def _sample(self, batch_size: int) -> Dict[str, torch.Tensor]:
        """One batch for `learn`.  On the GPU: one gw_replay_sample launch (indices drawn and rows gathered by the
        kernel), written straight into the captured update's input tensors once they exist."""
        if not self.fused_sampler:
            return self.ring.sample(batch_size, self.gen)
        dst = self.agent.static_inputs()
        if dst is not None and tuple(dst["state"].shape) == (batch_size, self.ring.L, self.ring.obs_len):
            return self.ring.sample_fused(self.svc, batch_size, seed=self.sample_seed, out=dst)
        if self._batch is None or self._batch["state"].shape[0] != batch_size:
            self._batch = self.ring.new_batch(batch_size)
        return self.ring.sample_fused(self.svc, batch_size, seed=self.sample_seed, out=self._batch)

    def learn_schedule(self, t: int) -> int:
        """Number of updates after vector step `t` (0-based), before the BATCH_SIZE gate.
        reference (maddpg/agent.py:199-224, num_envs = all environments):
          LEARN_STEP > num_envs: one update every LEARN_STEP // num_envs steps (:201-213);
          otherwise num_envs // LEARN_STEP updates after EVERY step (:214-224).
        batched: `updates_per_learn` updates every LEARN_STEP steps."""
        ls, E = int(self.hp["LEARN_STEP"]), self.global_envs
        if self.learn_cadence == "batched":
            return self.updates_per_learn if t % ls == 0 else 0
        if ls > E:
            return 1 if t % (ls // E) == 0 else 0
        return E // ls

    def train(self, env_steps: int, learn: bool = True) -> Dict[str, float]:
        """`env_steps` synchronous steps of all E envs (auto-reset replaces the reference's per-episode outer loop;
        the TRAIN_STEPS cap is the env's max_steps).  Returns the episode statistics gathered meanwhile."""
        env, ring, agent, hp = self.env, self.ring, self.agent, self.hp
        env.reset_stats()
        for _ in range(env_steps):
            t = self.t
            obs = ring.obs_slot(t)
            if self.fused is not None:                     # the kernel writes the continuous actions into the ring itself
                cont, ids = self.fused.forward(self.out.obs_code, self.out.action_mask, training=True,
                                               expl_noise=hp["EXPL_NOISE"], mean_noise=hp["MEAN_NOISE"],
                                               cont_out=ring.action_slot(t))
            else:
                cont, ids = agent.get_action(obs, self.out.action_mask, training=True)
                ring.store_action(t, cont)
            self.out = env.step(ids, obs_out=ring.obs_slot(t + 1), final_obs_out=ring.final_slot(t),
                                buffers=ring.buffers_slot(t))
            ring.advance()
            self.t += 1
            # maddpg/agent.py:199-224: learn once `len(memory) >= BATCH_SIZE`; the count of stored transitions is taken
            # on the smallest shard so that all ranks start in the same step (their all-reduces pair up)
            n_upd = self.learn_schedule(t) if learn else 0
            if n_upd and min(ring.t, ring.T - 1) * self.min_shard >= hp["BATCH_SIZE"]:
                self._learn_block(n_upd)
        return env.stats()


    def _learn_block(self, n_upd: int):
        """`n_upd` consecutive updates (sample + learn each), then refresh the actor kernel's packed weights."""
        agent, hp = self.agent, self.hp
        if self.learner is not None:                         # one launch for the whole block (csrc/gw_maddpg.cu)
            last = self.learner.learn_from_ring(self.ring, n_upd, self.sample_seed)[n_upd - 1].clone()
            self.losses.append(LearnStats(last[0], last[1]))
        else:
            for _ in range(n_upd):
                self.losses.append(agent.learn(self._sample(hp["BATCH_SIZE"])))
        if len(self.losses) > 1024:                          # keep the tail only: the reference cadence makes hundreds per step
            del self.losses[:512]
        self.updates_done += n_upd
        if self.fused is not None:
            self.fused.update(agent.actors)                  # the kernel keeps its own packed copy of the weights


def make_env(hp: Dict, num_envs: int, device="cuda", scenario="Level 3", obs_dtype=torch.float32, seed: Optional[int] = None,
             env_id_base: int = 0, env_kind: str = "multi"):
    """util.create_custom_ma_env (util.py:21-30) for a batch: WITH_FEAR / FeAR_weight / TRAIN_STEPS / SEED from the config.
    env_kind "single": the one-learner env of custom/customenv.py (what the reference's shipped checkpoints were trained on)."""
    from .batched import BatchedGridWorld
    return BatchedGridWorld(scenario, num_envs=num_envs, device=device, env_kind=env_kind, fear=bool(hp["WITH_FEAR"]),
                            fear_weight=float(hp["FeAR_weight"]) if hp["WITH_FEAR"] else 0.0,
                            max_steps=int(hp["TRAIN_STEPS"]), auto_reset=True, obs_dtype=obs_dtype,
                            obs_layout="cnn" if hp.get("ARCH", "mlp") == "cnn" else "mlp",
                            seed=int(hp["SEED"] if seed is None else seed), env_id_base=env_id_base)
