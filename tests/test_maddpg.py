"""MADDPG twin: shapes proven by the reference's checkpoints, update sanity, gradient all-reduce (gloo),
and the batched training loop on the GPU."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from marl_responsible_nav_b200 import maddpg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_network_shapes_match_reference_checkpoints():
    ag = maddpg.BatchedMADDPG(2, 160, 9, device="cpu")
    assert sum(p.numel() for p in ag.actors[0].parameters()) == 38793          # 20480+128+128+128+16384+128+128+128+1152+9
    assert sum(p.numel() for p in ag.critics[0].parameters()) == 60545         # critic input 2*160 + 2*9 = 338
    assert [tuple(p.shape) for p in ag.actors[0].parameters()][0] == (128, 160)
    single = maddpg.BatchedMADDPG(1, 160, 9, device="cpu")
    assert [tuple(p.shape) for p in single.critics[0].parameters()][0] == (128, 169)   # single-learner checkpoint: 128 x 169


def test_presets_match_reference_configs():
    hp = maddpg.preset("custom_fear_5")
    assert hp["WITH_FEAR"] is True and hp["FeAR_weight"] == -5.0 and hp["GAMMA"] == 0.98 and hp["TAU"] == 0.01
    assert hp["BATCH_SIZE"] == 128 and hp["LEARN_STEP"] == 10 and hp["MEMORY_SIZE"] == 200000 and hp["TRAIN_STEPS"] == 150
    assert maddpg.preset("custom")["WITH_FEAR"] is False and maddpg.preset("custom_fear_10")["FeAR_weight"] == -10.0
    ref = "/root/reference/configs/custom_fear_5.yaml"
    if os.path.exists(ref):
        assert maddpg.load_yaml_config(ref) == hp


def test_get_action_respects_mask_and_learn_reduces_critic_loss():
    torch.manual_seed(0)
    ag = maddpg.BatchedMADDPG(2, 160, 9, device="cpu", seed=1)
    obs = torch.randn(64, 2, 160)
    mask = torch.ones(64, 2, 9, dtype=torch.int8)
    mask[:, :, 5:] = 0
    cont, ids = ag.get_action(obs, mask, training=True)
    assert cont.shape == (64, 2, 9) and ids.dtype == torch.int8 and int(ids.max()) <= 4
    assert float(cont.min()) >= 0.0 and float(cont.max()) <= 1.0
    batch = {"state": obs, "action": cont, "reward": torch.randn(64, 2), "next_state": torch.randn(64, 2, 160),
             "done": torch.zeros(64, 2, dtype=torch.uint8)}
    first = ag.learn(batch).critic_loss
    for _ in range(60):
        last = ag.learn(batch).critic_loss
    assert float(last.sum()) < 0.5 * float(first.sum())


def _grad_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)
    ag = maddpg.BatchedMADDPG(2, 160, 9, device="cpu", seed=7 + rank)      # different initial weights per rank ...
    ag.broadcast_parameters(0)                                              # ... until rank 0's are broadcast
    g = torch.Generator().manual_seed(rank)                                 # different data shard per rank
    batch = {"state": torch.randn(32, 2, 160, generator=g), "action": torch.rand(32, 2, 9, generator=g),
             "reward": torch.randn(32, 2, generator=g), "next_state": torch.randn(32, 2, 160, generator=g),
             "done": torch.zeros(32, 2, dtype=torch.uint8)}
    torch.manual_seed(5)                                                    # same Gumbel noise on both ranks
    for _ in range(3):
        ag.learn(batch)
    torch.save([p.detach().clone() for p in ag.parameters()], os.path.join(out, f"params{rank}.pt"))
    dist.destroy_process_group()


def test_gradient_allreduce_keeps_ranks_in_sync(tmp_path):
    port = 29700 + (os.getpid() % 1000)
    mp.spawn(_grad_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    p0 = torch.load(os.path.join(str(tmp_path), "params0.pt"))
    p1 = torch.load(os.path.join(str(tmp_path), "params1.pt"))
    assert all(torch.allclose(a, b, atol=1e-6) for a, b in zip(p0, p1))


@pytest.mark.gpu
def test_batched_training_loop_runs_on_gpu():
    hp = maddpg.preset("custom_fear_5")
    hp["MEMORY_SIZE"] = 20000
    env = maddpg.make_env(hp, 512)
    tr = maddpg.BatchedTrainer(env, hp=hp, seed=0)
    st = tr.train(40)
    assert st["env_steps"] == 40 * 512 and st["episodes"] > 0
    assert len(tr.losses) >= 3 and all(bool(torch.isfinite(l.critic_loss).all()) for l in tr.losses)
    b = tr.ring.sample(256, tr.gen)
    assert b["state"].shape == (256, 2, 160) and b["action"].shape == (256, 2, 9)
    # stored reward is FeAR_weight * fear + env reward (maddpg/agent.py:130)
    s = int(b["t"][0]) % tr.ring.T
    e = int(b["env"][0])
    want = (-5.0 * tr.ring.fear[s, e] + tr.ring.reward[s, e].double()).float()
    assert torch.equal(b["reward"][0], want)


@pytest.mark.gpu
def test_segmented_graph_update_on_one_rank(tmp_path):
    """The multi-rank update (three CUDA graphs cut at the two gradient all-reduces, flat gradient buffers) run in a
    one-rank NCCL group: it learns (critic loss on a fixed batch falls), moves the target networks, and keeps every
    .grad a view of its network's flat buffer.  scripts/check_segmented_learn.py is the two-rank version (ranks stay
    bit-identical; run under torchrun)."""
    dev = torch.device("cuda", 0)
    dist.init_process_group("nccl", init_method=f"file://{tmp_path}/pg", rank=0, world_size=1, device_id=dev)
    try:
        ag = maddpg.BatchedMADDPG(2, 160, 9, device=dev, seed=5)
        ag.force_segmented = True
        g = torch.Generator(device=dev).manual_seed(1)
        batch = {"state": torch.randn(128, 2, 160, device=dev, generator=g), "next_state": torch.randn(128, 2, 160, device=dev, generator=g),
                 "action": torch.rand(128, 2, 9, device=dev, generator=g), "reward": torch.randn(128, 2, device=dev, generator=g),
                 "done": torch.zeros(128, 2, device=dev)}
        tgt0 = [p.clone() for p in ag.critic_targets[0].parameters()]
        losses = [float(ag.learn(batch).critic_loss.sum()) for _ in range(40)]
        assert ag._segments is not None and len(ag._segments["graphs"]) == 3
        assert losses[-1] < 0.2 * losses[0], (losses[0], losses[-1])
        assert any(not torch.equal(a, b) for a, b in zip(tgt0, ag.critic_targets[0].parameters()))
        fa, fc = ag._segments["flat"]
        for net, flat in zip(ag.actors + ag.critics, fa + fc):
            lo, hi = flat.data_ptr(), flat.data_ptr() + flat.numel() * 4
            assert all(lo <= p.grad.data_ptr() < hi for p in net.parameters())
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_parallel_agent_branches_equal_sequential_graph():
    """One-GPU update graph with the agents' passes as parallel branches (forked streams during capture) against the
    same graph recorded sequentially: same initial weights, same batches, same random stream -> the same weights."""
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(1)
    batches = [{"state": torch.randn(128, 2, 160, device=dev, generator=g), "next_state": torch.randn(128, 2, 160, device=dev, generator=g),
                "action": torch.rand(128, 2, 9, device=dev, generator=g), "reward": torch.randn(128, 2, device=dev, generator=g),
                "done": (torch.rand(128, 2, device=dev, generator=g) < 0.1).float()} for _ in range(4)]
    results = []
    for parallel in (False, True):
        ag = maddpg.BatchedMADDPG(2, 160, 9, device=dev, seed=5)
        ag.parallel_agents = parallel
        torch.manual_seed(11)
        torch.cuda.manual_seed(11)
        losses = [ag.learn(batches[i % 4]) for i in range(12)]
        assert ag._graph is not None                       # three eager updates, then the captured graph
        assert sorted(map(str, ag._side_streams)) == (sorted(map(str, ["agents", ("inner", 0), ("inner", 1), ("actor", 0), ("actor", 1)])) if parallel else [])
        torch.cuda.synchronize()
        results.append(([p.detach().clone() for p in ag.parameters()] +
                        [p.detach().clone() for net in ag.actor_targets + ag.critic_targets for p in net.parameters()],
                        torch.stack([l.critic_loss for l in losses])))
    (p_seq, l_seq), (p_par, l_par) = results
    assert torch.allclose(l_seq, l_par, rtol=1e-5, atol=1e-6)
    worst = max(float((a - b).abs().max()) for a, b in zip(p_seq, p_par))
    assert worst <= 1e-6, worst


@pytest.mark.gpu
@pytest.mark.parametrize("rows", [128, 1, 77, 300])
def test_fused_layernorm_relu_matches_torch(rows):
    """gw_ln_relu_forward / _backward (one kernel each) against torch.nn.LayerNorm + ReLU with autograd, fp32: rows = 128
    is the update's batch (one CTA writes dgamma / dbeta), 300 takes the several-CTA path (atomic column sums)."""
    from marl_responsible_nav_b200 import BatchedGridWorld
    dev = torch.device("cuda", 0)
    env = BatchedGridWorld("Level 3", num_envs=32, fear=False, seed=1)
    ops = maddpg.TrainOps(env)
    g = torch.Generator(device=dev).manual_seed(rows)
    ln = torch.nn.LayerNorm(128).to(dev)
    with torch.no_grad():
        ln.weight.copy_(torch.randn(128, device=dev, generator=g))
        ln.bias.copy_(0.3 * torch.randn(128, device=dev, generator=g))
    x = (3.0 * torch.randn(rows, 128, device=dev, generator=g) + 1.5).requires_grad_(True)
    up = torch.randn(rows, 128, device=dev, generator=g)
    ref = torch.relu(ln(x))
    ref.backward(up)
    want = (ref.detach().clone(), x.grad.clone(), ln.weight.grad.clone(), ln.bias.grad.clone())
    x.grad = None
    ln.zero_grad(set_to_none=True)
    got = maddpg.forward_mlp(torch.nn.Sequential(ln, torch.nn.ReLU()), x, ops)
    got.backward(up)
    env.sync()
    assert torch.allclose(got, want[0], rtol=1e-5, atol=1e-5)
    assert torch.equal(got == 0, want[0] == 0) or float(((got == 0) != (want[0] == 0)).float().mean()) < 1e-4   # same ReLU mask
    assert torch.allclose(x.grad, want[1], rtol=1e-4, atol=1e-5)
    assert torch.allclose(ln.weight.grad, want[2], rtol=1e-4, atol=1e-4)
    assert torch.allclose(ln.bias.grad, want[3], rtol=1e-4, atol=1e-4)


@pytest.mark.gpu
def test_update_with_fused_ops_tracks_the_torch_modules():
    """The same update (eager) with and without the fused LayerNorm+ReLU kernels, from the same weights and noise:
    parameters stay together to fp32 rounding over several updates."""
    from marl_responsible_nav_b200 import BatchedGridWorld
    dev = torch.device("cuda", 0)
    env = BatchedGridWorld("Level 3", num_envs=32, fear=False, seed=1)
    g = torch.Generator(device=dev).manual_seed(2)
    batch = {"state": torch.randn(128, 2, 160, device=dev, generator=g), "next_state": torch.randn(128, 2, 160, device=dev, generator=g),
             "action": torch.rand(128, 2, 9, device=dev, generator=g), "reward": torch.randn(128, 2, device=dev, generator=g),
             "done": torch.zeros(128, 2, device=dev)}
    res = []
    for fused in (False, True):
        ag = maddpg.BatchedMADDPG(2, 160, 9, device=dev, seed=5)
        if fused:
            ag.attach_ops(env)
        torch.manual_seed(3)
        torch.cuda.manual_seed(3)
        for _ in range(5):
            st = ag.learn(batch, graph=False)
        torch.cuda.synchronize()
        res.append(([p.detach().clone() for p in ag.parameters()], st.critic_loss.clone()))
    worst = max(float((a - b).abs().max()) for a, b in zip(res[0][0], res[1][0]))
    assert worst < 2e-4, worst                              # Adam's normalised steps amplify rounding differences of tiny gradients
    assert torch.allclose(res[0][1], res[1][1], rtol=1e-3)


@pytest.mark.gpu
@pytest.mark.parametrize("batch,n_in,n_out,strided,x_grad", [(128, 338, 128, False, False), (128, 338, 128, False, True),
                                                              (128, 128, 128, False, True), (128, 128, 1, False, True),
                                                              (128, 160, 128, True, False), (128, 128, 9, False, True),
                                                              (77, 45, 33, True, True), (1, 128, 128, False, True)])
def test_fused_linear_backward_matches_torch(batch, n_in, n_out, strided, x_grad):
    """gw_linear_backward (dW, db, dx in one launch) against autograd of torch.nn.functional.linear, fp32: the shapes of
    the update (critic 338-128-128-1, actor 160-128-128-9, the actor's input a strided slice of the batch) and ragged ones."""
    from marl_responsible_nav_b200 import BatchedGridWorld
    dev = torch.device("cuda", 0)
    env = BatchedGridWorld("Level 3", num_envs=32, fear=False, seed=1)
    ops = maddpg.TrainOps(env)
    g = torch.Generator(device=dev).manual_seed(batch + n_in)
    lin = torch.nn.Linear(n_in, n_out).to(dev)
    base = torch.randn(batch, 2, n_in, device=dev, generator=g)
    x = (base[:, 1] if strided else base[:, 1].contiguous()).detach().requires_grad_(x_grad)
    up = torch.randn(batch, n_out, device=dev, generator=g)
    lin(x).backward(up)
    want = (lin.weight.grad.clone(), lin.bias.grad.clone(), x.grad.clone() if x_grad else None)
    lin.zero_grad(set_to_none=True)
    x.grad = None
    y = maddpg.forward_mlp(torch.nn.Sequential(lin), x, ops)
    assert torch.equal(y, lin(x))
    y.backward(up)
    env.sync()
    tol = dict(rtol=1e-4, atol=1e-4)
    assert torch.allclose(lin.weight.grad, want[0], **tol) and torch.allclose(lin.bias.grad, want[1], **tol)
    if x_grad:
        assert torch.allclose(x.grad, want[2], **tol)
    else:
        assert x.grad is None


def test_host_side_of_the_fused_update_paths_on_cpu():
    """No GPU: forward_mlp without the library is the module itself, branching is off (functions run in turn), the
    update exposes no graph inputs, and the fused sampler refuses a CPU ring (there is no CPU path)."""
    from marl_responsible_nav_b200.replay import ReplayRing
    torch.manual_seed(0)
    net = maddpg.mlp(160, (128, 128), 9, None)
    x = torch.randn(5, 160)
    assert torch.equal(maddpg.forward_mlp(net, x, None), net(x))
    ag = maddpg.BatchedMADDPG(2, 160, 9, device="cpu", seed=1)
    assert ag.ops is None and ag.static_inputs() is None
    order = []
    ag._fork_join([lambda: order.append("a"), lambda: order.append("b"), lambda: order.append("c")], "k")
    ag._branches(lambda k: order.append(k))
    assert order == ["a", "b", "c", 0, 1] and ag._side_streams == {}
    ring = ReplayRing(4, 2, 160, capacity=40, device="cpu")
    b = ring.new_batch(8)
    assert b["state"].shape == (8, 2, 160) and b["done"].dtype == torch.float32 and b["t"].dtype == torch.int64
    ring.advance()
    with pytest.raises(RuntimeError, match="no CPU path"):
        ring.sample_fused(None, 8)


def test_learn_schedule_follows_the_reference_cadence():
    """maddpg/agent.py:199-224: LEARN_STEP > num_envs -> one update every LEARN_STEP // num_envs steps; otherwise
    num_envs // LEARN_STEP updates after every step.  `batched` is round 1's rule (a block every LEARN_STEP steps)."""
    from types import SimpleNamespace
    sched = maddpg.BatchedTrainer.learn_schedule

    def counts(E, ls, steps, cadence="reference", upl=1):
        tr = SimpleNamespace(hp={"LEARN_STEP": ls}, global_envs=E, learn_cadence=cadence, updates_per_learn=upl)
        return [sched(tr, t) for t in range(steps)]

    assert counts(1, 10, 21) == [1 if t % 10 == 0 else 0 for t in range(21)]          # the reference's own run: num_envs = 1
    assert counts(4, 10, 6) == [1 if t % 2 == 0 else 0 for t in range(6)]             # 10 // 4 = 2
    assert counts(10, 10, 3) == [1, 1, 1]                                               # not (LEARN_STEP > num_envs): 10 // 10
    assert counts(4096, 10, 3) == [409, 409, 409]
    assert counts(4096, 10, 21, "batched", 3) == [3 if t % 10 == 0 else 0 for t in range(21)]
    # update-to-data ratio of the reference rule: one update per ~LEARN_STEP transitions whatever E is
    for E in (1, 4, 64, 4096):
        n = sum(counts(E, 10, 200))
        assert abs(n / (200 * E) - 0.1) < 0.03
