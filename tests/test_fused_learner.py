"""gw_learner_update (csrc/gw_maddpg.cu): the MADDPG update of maddpg/agent.py:209-224 as one persistent kernel, against
the same update written with plain PyTorch fp32 modules and autograd (the fp32 reference the tier asks for: this is a
floating-point kernel).  Tolerances are stated where they are used."""
import copy
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from marl_responsible_nav_b200 import _native as N
from marl_responsible_nav_b200 import maddpg


def test_learner_structs_and_layout_on_cpu():
    """Host arithmetic only (no GPU): the layout of the flat parameter vector is torch's parameter order per network,
    actors first, and the sizes are the reference checkpoints' (SURVEY 2.2)."""
    lib = N.load()
    cfg = N.GwLearnerConfig()
    cfg.struct_size = C.sizeof(N.GwLearnerConfig)
    cfg.n_agents, cfg.obs_len, cfg.action_dim, cfg.batch = 2, 160, 9, 128
    lay = N.GwLearnerLayout()
    lay.struct_size = C.sizeof(N.GwLearnerLayout)
    assert lib.gw_learner_layout_of(C.byref(cfg), C.byref(lay)) == 0
    assert lay.n_nets == 4 and list(lay.net_params)[:4] == [38793, 38793, 60545, 60545]
    offs = list(lay.net_offset)[:4]
    assert offs[0] == 0 and all(o % 4 == 0 for o in offs) and offs == sorted(offs)
    assert lay.param_floats >= 2 * 38793 + 2 * 60545 and lay.scratch_bytes > 0 and lay.scratch_bytes % 256 == 0
    ag = maddpg.BatchedMADDPG(2, 160, 9, device="cpu")
    assert sum(p.numel() for p in ag.actors[0].parameters()) == lay.net_params[0]
    assert sum(p.numel() for p in ag.critics[0].parameters()) == lay.net_params[2]
    for bad in (dict(batch=100), dict(batch=1024), dict(n_agents=3), dict(action_dim=5), dict(obs_len=150)):
        c2 = N.GwLearnerConfig.from_buffer_copy(cfg)
        for k, v in bad.items():
            setattr(c2, k, v)
        assert lib.gw_learner_layout_of(C.byref(c2), C.byref(lay)) != 0
    cfg.n_agents = 1                                                    # the single-learner checkpoint: critic 128 x 169
    assert lib.gw_learner_layout_of(C.byref(cfg), C.byref(lay)) == 0 and list(lay.net_params)[:2] == [38793, 128 * 169 + 128 * 3 + 128 * 128 + 128 * 3 + 128 + 1]


def _batch(dev, B, n, seed, obs_len=160):
    g = torch.Generator(device=dev).manual_seed(seed)
    u = lambda *s: torch.rand(*s, device=dev, generator=g)
    gum = lambda: -torch.log(-torch.log(u(B, n, 9).clamp_(1e-20, 1.0)) + 1e-20)
    batch = {"state": (u(B, n, obs_len) < 0.1).float() * u(B, n, obs_len), "next_state": (u(B, n, obs_len) < 0.1).float() * u(B, n, obs_len),
             "action": u(B, n, 9), "reward": torch.randn(B, n, device=dev, generator=g), "done": (u(B, n) < 0.2).float()}
    return batch, gum(), gum()


def _torch_update(ag, batch, g_next, g_cur):
    """One update of `ag` with plain modules + autograd; returns (critic grads, actor grads, losses [2, n])."""
    hp, n = ag.hp, ag.n
    s, a, r, s2, d = (batch[k] for k in ("state", "action", "reward", "next_state", "done"))
    B = s.shape[0]
    act = lambda net, x, g: F.softmax(net[:-1](x) + g, dim=-1)
    with torch.no_grad():
        a2 = torch.stack([act(ag.actor_targets[k], s2[:, k], g_next[:, k]) for k in range(n)], dim=1)
        in2 = torch.cat([s2.reshape(B, -1), a2.reshape(B, -1)], dim=1)
    cin = torch.cat([s.reshape(B, -1), a.reshape(B, -1)], dim=1)
    cg, agr, losses = [], [], torch.zeros(2, n)
    for k in range(n):
        with torch.no_grad():
            y = r[:, k] + hp["GAMMA"] * (1 - d[:, k]) * ag.critic_targets[k](in2).squeeze(-1)
        c_loss = F.mse_loss(ag.critics[k](cin).squeeze(-1), y)
        ag.critic_opt[k].zero_grad(set_to_none=True)
        c_loss.backward()
        cg.append([p.grad.clone() for p in ag.critics[k].parameters()])
        ag.critic_opt[k].step()
        a_new = a.clone()
        a_new[:, k] = act(ag.actors[k], s[:, k], g_cur[:, k])
        a_loss = -ag.critics[k](torch.cat([s.reshape(B, -1), a_new.reshape(B, -1)], dim=1)).mean()
        ag.actor_opt[k].zero_grad(set_to_none=True)
        a_loss.backward()
        agr.append([p.grad.clone() for p in ag.actors[k].parameters()])
        ag.actor_opt[k].step()
        losses[0, k], losses[1, k] = float(a_loss.detach()), float(c_loss.detach())
    with torch.no_grad():
        for net, tnet in zip(ag.actors + ag.critics, ag.actor_targets + ag.critic_targets):
            for p, tp in zip(net.parameters(), tnet.parameters()):
                tp.lerp_(p, hp["TAU"])
    return cg, agr, losses


def _pair(dev, n, B, seed=5, obs_len=160, kernel="auto"):
    from marl_responsible_nav_b200 import BatchedGridWorld
    from marl_responsible_nav_b200.learner import FusedLearner
    env = BatchedGridWorld("Level 3", num_envs=32, fear=False, seed=1)
    hp = dict(maddpg.DEFAULT_HP, BATCH_SIZE=B)
    ref = maddpg.BatchedMADDPG(n, obs_len, 9, hp=hp, device=dev, seed=seed)
    ag = maddpg.BatchedMADDPG(n, obs_len, 9, hp=hp, device=dev, seed=seed)
    with torch.no_grad():                                              # LayerNorm gains / biases off their 1 / 0 initial values
        g = torch.Generator(device=dev).manual_seed(seed)
        for net in ref.actors + ref.critics:
            for name, p in net.named_parameters():
                if p.dim() == 1:
                    p.add_(0.1 * torch.randn(p.shape, device=dev, generator=g))
        for dst, src in zip(ag.actors + ag.critics + ag.actor_targets + ag.critic_targets,
                            ref.actors + ref.critics + ref.actors + ref.critics):
            dst.load_state_dict(src.state_dict())
        for dst, src in zip(ref.actor_targets + ref.critic_targets, ref.actors + ref.critics):
            dst.load_state_dict(src.state_dict())
    learner = FusedLearner(env, ag, batch_size=B, seed=3)
    learner.set_kernel(kernel)
    return env, ref, ag, learner


def _rel(a, b):
    return float((a - b).abs().max()) / max(float(b.abs().max()), 1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("n,B,kernel", [(2, 128, "cluster"), (2, 32, "cluster"), (2, 64, "cluster"), (2, 128, "phase"), (1, 128, "phase"),
                                        (2, 32, "phase"), (2, 256, "phase")])
def test_one_update_matches_autograd(n, B, kernel):
    """Gradients within 1e-4 (relative to the tensor's largest entry) of fp32 autograd, losses within 1e-5 relative, and
    after the Adam steps + soft update the parameters / targets / moments agree (Adam divides by sqrt(v): the first step
    moves every weight by ~lr whatever its gradient, so parameters are compared to 2e-5 absolute = 2 % of lr)."""
    dev = torch.device("cuda", 0)
    if kernel == "phase" and B % 32:
        pytest.skip("the phase kernel tiles the batch by 32")
    env, ref, ag, learner = _pair(dev, n, B, kernel=kernel)
    assert learner.kernel == kernel
    batch, g_next, g_cur = _batch(dev, B, n, seed=B + n)
    cg, agr, want_loss = _torch_update(ref, batch, g_next, g_cur)
    losses = learner.learn(batch, g_next, g_cur)
    env.sync()
    got = learner.grads
    for k in range(n):
        for nets, grads, base in ((ag.actors, agr, 0), (ag.critics, cg, n)):
            off = learner.offsets[base + k]
            for p, want in zip(nets[k].parameters(), grads[k]):
                mine = got[off:off + p.numel()].view_as(p)
                assert _rel(mine, want) < 1e-4, (k, base, tuple(p.shape), _rel(mine, want))
                off += p.numel()
    assert torch.allclose(losses[0].cpu(), want_loss, rtol=1e-5, atol=1e-6), (losses[0].cpu(), want_loss)
    for mine, theirs in zip(ag.actors + ag.critics + ag.actor_targets + ag.critic_targets,
                            ref.actors + ref.critics + ref.actor_targets + ref.critic_targets):
        for p, q in zip(mine.parameters(), theirs.parameters()):
            assert float((p - q).abs().max()) < 2e-5, float((p - q).abs().max())
    for k, (opt_m, opt_r) in enumerate(zip(ag.actor_opt + ag.critic_opt, ref.actor_opt + ref.critic_opt)):
        for p, q in zip(opt_m.param_groups[0]["params"], opt_r.param_groups[0]["params"]):
            assert _rel(opt_m.state[p]["exp_avg"], opt_r.state[q]["exp_avg"]) < 1e-4          # shared storage with the kernel's moments
            assert _rel(opt_m.state[p]["exp_avg_sq"], opt_r.state[q]["exp_avg_sq"]) < 2e-4
    assert learner.adam_steps.tolist() == [1.0] * (2 * n)
    learner.export_steps()
    assert all(float(st["step"]) == 1.0 for opt in ag.actor_opt + ag.critic_opt for st in opt.state.values())


@pytest.mark.gpu
@pytest.mark.parametrize("kernel", ["cluster", "phase"])
def test_intermediates_match_torch(kernel):
    """The kernel's activations against the modules: target actions, Q, TD target (debug tensors)."""
    dev = torch.device("cuda", 0)
    env, ref, ag, learner = _pair(dev, 2, 128, kernel=kernel)
    batch, g_next, g_cur = _batch(dev, 128, 2, seed=9)
    s, a, r, s2, d = (batch[k] for k in ("state", "action", "reward", "next_state", "done"))
    with torch.no_grad():
        a2 = torch.stack([F.softmax(ref.actor_targets[k][:-1](s2[:, k]) + g_next[:, k], dim=-1) for k in range(2)], dim=1)
        cin = torch.cat([s.reshape(128, -1), a.reshape(128, -1)], dim=1)
        in2 = torch.cat([s2.reshape(128, -1), a2.reshape(128, -1)], dim=1)
        q = [ref.critics[k](cin).squeeze(-1) for k in range(2)]
        y = [r[:, k] + 0.98 * (1 - d[:, k]) * ref.critic_targets[k](in2).squeeze(-1) for k in range(2)]
        anew = [F.softmax(ref.actors[k][:-1](s[:, k]) + g_cur[:, k], dim=-1) for k in range(2)]
    learner.learn(batch, g_next, g_cur)
    env.sync()
    assert torch.allclose(learner.debug_tensor("a2").view(128, 2, 9), a2, rtol=1e-4, atol=1e-6)
    for k in range(2):
        assert torch.allclose(learner.debug_tensor("q", k), q[k], rtol=1e-4, atol=1e-5)
        assert torch.allclose(learner.debug_tensor("y", k), y[k], rtol=1e-4, atol=1e-5)
        assert torch.allclose(learner.debug_tensor("anew", k).view(128, 9), anew[k], rtol=1e-4, atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("kernel", ["cluster", "phase"])
def test_several_updates_track_autograd(kernel):
    """Eight consecutive updates on changing batches: the critic losses follow the PyTorch run (1e-3 relative; Adam's
    normalised steps amplify rounding differences) and the parameters stay within 3e-4."""
    dev = torch.device("cuda", 0)
    env, ref, ag, learner = _pair(dev, 2, 128, kernel=kernel)
    for it in range(8):
        batch, g_next, g_cur = _batch(dev, 128, 2, seed=100 + it)
        _, _, want = _torch_update(ref, batch, g_next, g_cur)
        got = learner.learn(batch, g_next, g_cur)
        env.sync()
        assert torch.allclose(got[0, 1].cpu(), want[1], rtol=1e-3, atol=1e-5), (it, got[0, 1].cpu(), want[1])
    worst = max(float((p - q).abs().max()) for m, t in zip(ag.actors + ag.critics, ref.actors + ref.critics)
                for p, q in zip(m.parameters(), t.parameters()))
    assert worst < 3e-4, worst
    assert learner.adam_steps.tolist() == [8.0] * 4


def _filled_ring(env_n=64, steps=12, obs_dtype=torch.float32, seed=0):
    hp = maddpg.preset("custom_fear_5")
    hp["MEMORY_SIZE"] = env_n * 10                                      # the ring wraps during the fill
    env = maddpg.make_env(hp, env_n, obs_dtype=obs_dtype)
    tr = maddpg.BatchedTrainer(env, hp=hp, seed=seed, fused_learner=False)
    tr.train(steps, learn=False)
    return env, tr, hp


@pytest.mark.gpu
@pytest.mark.parametrize("kernel", ["cluster", "phase"])
@pytest.mark.parametrize("obs_dtype", [torch.float32, torch.bfloat16])
def test_ring_mode_equals_explicit_batches(obs_dtype, kernel):
    """U updates in ONE launch, each drawing + gathering its batch inside the kernel, against U one-update launches on
    the batches `ring.sample_fused` returns for the same draw numbers: bit-identical parameters, targets and losses
    (same kernel, same inputs, same noise keys)."""
    from marl_responsible_nav_b200.learner import FusedLearner
    dev = torch.device("cuda", 0)
    env, tr, hp = _filled_ring(obs_dtype=obs_dtype)
    ring, U = tr.ring, 5
    agents = [maddpg.BatchedMADDPG(2, 160, 9, hp=hp, device=dev, seed=8) for _ in range(2)]
    la, lb = (FusedLearner(env, ag, seed=4) for ag in agents)
    la.set_kernel(kernel); lb.set_kernel(kernel)
    draws0 = ring._draws
    got = la.learn_from_ring(ring, U, sample_seed=77).clone()
    assert ring._draws == draws0 + U and la.updates_done == U
    ring._draws = draws0
    want = []
    for u in range(U):
        b = ring.sample_fused(env, hp["BATCH_SIZE"], seed=77)
        want.append(lb.learn(b).clone())
    env.sync()
    assert torch.equal(got, torch.cat(want, dim=0))
    for x, y in ((la.params, lb.params), (la.targets, lb.targets), (la.adam_m, lb.adam_m), (la.adam_v, lb.adam_v), (la.grads, lb.grads)):
        assert torch.equal(x, y)
    assert torch.isfinite(la.params).all() and not torch.equal(la.params, la.targets)


@pytest.mark.gpu
@pytest.mark.parametrize("kernel", ["cluster", "phase"])
def test_segments_equal_the_whole_update(kernel):
    """The three pieces of a data-parallel update (critic gradients | critic steps + actor gradients | actor steps) run
    back to back with grad_scale 1 give exactly the one-launch update."""
    dev = torch.device("cuda", 0)
    env, ref, ag, la = _pair(dev, 2, 128, kernel=kernel)
    env2, _, ag2, lb = _pair(dev, 2, 128, kernel=kernel)
    for it in range(3):
        batch, g_next, g_cur = _batch(dev, 128, 2, seed=40 + it)
        l1 = la.learn(batch, g_next, g_cur).clone()
        l2 = torch.zeros_like(l1)
        for seg in (N.GW_LEARN_CRITIC_GRADS, N.GW_LEARN_ACTOR_GRADS, N.GW_LEARN_FINISH):
            lb.learn(batch, g_next, g_cur, segment=seg, losses=l2)
        env.sync(); env2.sync()
        assert torch.equal(l1, l2)
        for x, y in ((la.params, lb.params), (la.targets, lb.targets), (la.adam_m, lb.adam_m), (la.adam_v, lb.adam_v), (la.grads, lb.grads)):
            assert torch.equal(x, y), it
    assert la.adam_steps.tolist() == lb.adam_steps.tolist() == [3.0] * 4


@pytest.mark.gpu
def test_trainer_learns_through_the_fused_kernel():
    """BatchedTrainer at the reference cadence (E // LEARN_STEP updates after every vector step) through the kernel:
    update counts follow the schedule, the actor kernel's weights follow the learner's, losses are finite."""
    hp = maddpg.preset("custom_fear_5")
    hp["MEMORY_SIZE"] = 20000
    env = maddpg.make_env(hp, 256)
    tr = maddpg.BatchedTrainer(env, hp=hp, seed=0)
    assert tr.learner is not None and tr.learner.kernel == "cluster"
    p0 = tr.learner.params.clone()
    tr.train(6)
    env.sync()
    assert tr.updates_done == 6 * (256 // 10) and tr.learner.updates_done == tr.updates_done
    assert tr.learner.adam_steps.tolist() == [float(tr.updates_done)] * 4
    assert torch.isfinite(tr.learner.params).all() and not torch.equal(p0, tr.learner.params)
    last = tr.losses[-1]
    assert torch.isfinite(last.critic_loss).all() and torch.isfinite(last.actor_loss).all()


@pytest.mark.gpu
def test_checkpoint_round_trip_through_the_learner(tmp_path):
    """Train through the kernel, save in the reference's checkpoint format (maddpg/agent.py:255-266), load, adopt into a new
    learner: parameters, targets, Adam moments and step counters are the kernel's, bit for bit -- and the two learners stay
    identical when both continue on the same batches."""
    from marl_responsible_nav_b200 import checkpoint
    from marl_responsible_nav_b200.learner import FusedLearner
    dev = torch.device("cuda", 0)
    env, _, ag, la = _pair(dev, 2, 128)
    for it in range(3):
        batch, g_next, g_cur = _batch(dev, 128, 2, seed=70 + it)
        la.learn(batch, g_next, g_cur)
    env.sync()
    path = str(tmp_path / "MADDPG.pt")
    checkpoint.save_reference_checkpoint(ag, path, steps=[3])
    ag2 = checkpoint.load_reference_checkpoint(path, device=dev, hp=ag.hp)
    lb = FusedLearner(env, ag2, batch_size=128, seed=3)
    lb.updates_done = la.updates_done
    for x, y in ((la.params, lb.params), (la.targets, lb.targets), (la.adam_m, lb.adam_m), (la.adam_v, lb.adam_v)):
        assert torch.equal(x, y)
    assert lb.adam_steps.tolist() == la.adam_steps.tolist() == [3.0] * 4
    batch, g_next, g_cur = _batch(dev, 128, 2, seed=99)
    l1, l2 = la.learn(batch, g_next, g_cur).clone(), lb.learn(batch, g_next, g_cur).clone()
    env.sync()
    assert torch.equal(l1, l2) and torch.equal(la.params, lb.params) and torch.equal(la.adam_v, lb.adam_v)


@pytest.mark.gpu
def test_learner_refuses_what_it_cannot_run():
    """Error behaviour of the update entry points: integer status + message, nothing launched."""
    from marl_responsible_nav_b200.learner import FusedLearner
    from marl_responsible_nav_b200.replay import ReplayRing
    dev = torch.device("cuda", 0)
    env, _, ag, learner = _pair(dev, 2, 128)
    batch, _, _ = _batch(dev, 128, 2, seed=1)
    with pytest.raises(ValueError, match="shape"):
        learner.learn({k: v[:64] for k, v in batch.items()})
    empty = ReplayRing(32, 2, 160, capacity=320, device=dev)
    with pytest.raises(RuntimeError, match="empty"):
        learner.learn_from_ring(empty, 1, 5)
    other = ReplayRing(32, 2, 144, capacity=320, device=dev)
    other.advance()
    with pytest.raises(RuntimeError, match="shape"):
        learner.learn_from_ring(other, 1, 5)
    with pytest.raises(ValueError, match="unsupported shape"):
        FusedLearner(env, maddpg.BatchedMADDPG(2, 160, 9, hp=dict(maddpg.DEFAULT_HP, BATCH_SIZE=100), device=dev), batch_size=100)
    _, _, _, big = _pair(dev, 2, 512, kernel="phase")           # 64 row blocks: more clusters than fit at once
    with pytest.raises(RuntimeError, match="cluster kernel needs"):
        big.set_kernel("cluster")
    single = FusedLearner(env, maddpg.BatchedMADDPG(1, 160, 9, device=dev), seed=1)
    assert single.kernel == "phase"                            # one learner: the critic's 169-float rows are not 8-byte aligned
    assert learner.updates_done == 0 and torch.isfinite(learner.params).all()
