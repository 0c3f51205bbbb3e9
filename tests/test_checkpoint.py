"""The reference's checkpoint format: layout pinned by the files the reference ships, round trip through
our writer, and the reference's trained Level-3 policy run through the fused actor kernel and the batched env."""
import glob
import os

import numpy as np
import pytest
import torch

from marl_responsible_nav_b200 import checkpoint, maddpg

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_single_actor_level3.npz")
REF = os.environ.get("GW_REFERENCE", "/root/reference")


def _rebuild_reference_file(tmp_path):
    """A checkpoint in the reference's format whose actor is the shipped Level-3 policy (from the fixture)."""
    g = np.load(GOLD)
    agent = maddpg.BatchedMADDPG(1, 160, 9, device="cpu", seed=3)
    path = str(tmp_path / "MADDPG.pt")
    checkpoint.save_reference_checkpoint(agent, path, steps=[int(s) for s in g["scalar/steps"]])
    ck = torch.load(path, map_location="cpu", weights_only=False)
    for k in ("actors_state_dict", "actor_targets_state_dict"):
        for name in list(ck[k][0].keys()):
            ck[k][0][name] = torch.from_numpy(g["param/" + name])
    torch.save(ck, path)
    return path, g


def test_writer_has_the_reference_layout(tmp_path):
    path, g = _rebuild_reference_file(tmp_path)
    ck = torch.load(path, map_location="cpu", weights_only=False)
    assert sorted(ck.keys()) == [str(k) for k in g["keys"]]                      # every top-level key of the shipped file, no other
    assert list(ck["actors_state_dict"][0].keys()) == [k[len("param/"):] for k in g.files if k.startswith("param/")]
    assert tuple(ck["critics_state_dict"][0]["feature_net.linear_layer_0.weight"].shape) == (128, 169)
    assert ck["actors_init_dict"][0]["mlp_output_activation"] == "GumbelSoftmax" and ck["critics_init_dict"][0]["mlp_output_activation"] is None


def test_round_trip_and_logits_of_the_shipped_policy(tmp_path):
    path, g = _rebuild_reference_file(tmp_path)
    agent = checkpoint.load_reference_checkpoint(path, device="cpu")
    assert agent.n == 1 and agent.obs_dim == 160 and agent.act_dim == 9 and agent.steps[-1] == int(g["scalar/steps"][-1])
    assert agent.hp["GAMMA"] == float(g["scalar/gamma"]) and agent.hp["BATCH_SIZE"] == int(g["scalar/batch_size"])
    with torch.no_grad():
        logits = agent.actors[0][:-1](torch.from_numpy(g["probe_obs"]))          # everything but the Gumbel-softmax
    np.testing.assert_allclose(logits.numpy(), g["probe_logits"], rtol=1e-4, atol=1e-4)
    # save -> load -> save is stable, optimiser moments included
    two = maddpg.BatchedMADDPG(2, 160, 9, device="cpu", seed=1)
    batch = {"state": torch.randn(32, 2, 160), "next_state": torch.randn(32, 2, 160), "action": torch.rand(32, 2, 9),
             "reward": torch.randn(32, 2), "done": torch.zeros(32, 2)}
    two.learn(batch)
    p2 = str(tmp_path / "two.pt")
    checkpoint.save_reference_checkpoint(two, p2)
    back = checkpoint.load_reference_checkpoint(p2, device="cpu")
    for a, b in zip(list(two.parameters()), list(back.parameters())):
        assert torch.equal(a, b)
    sa, sb = two.critic_opt[0].state_dict()["state"], back.critic_opt[0].state_dict()["state"]
    assert sa.keys() == sb.keys() and all(torch.equal(sa[k]["exp_avg"], sb[k]["exp_avg"]) for k in sa)


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models", "custom")), reason="reference not mounted")
def test_every_shipped_custom_checkpoint_loads():
    files = sorted(glob.glob(os.path.join(REF, "models", "custom", "**", "*.pt"), recursive=True))
    assert len(files) >= 10
    x = torch.from_numpy(np.load(GOLD)["probe_obs"])
    for f in files:
        ck = torch.load(f, map_location="cpu", weights_only=False)
        agent = checkpoint.load_reference_checkpoint(f, device="cpu")
        assert agent.n == ck["n_agents"] == 1 and agent.steps == ck["steps"]
        sd = ck["actors_state_dict"][0]
        with torch.no_grad():
            h = torch.nn.functional.linear(x, sd["feature_net.linear_layer_0.weight"], sd["feature_net.linear_layer_0.bias"])
            ours = agent.actors[0][0](x)
        assert torch.equal(h, ours)
        assert torch.equal(agent.critic_targets[0][6].weight, ck["critic_targets_state_dict"][0]["feature_net.linear_layer_output.weight"])


# scripts/eval_reference_policy_cpu.py (reference env + reference policy, CPU, build container; profiles/r1d_reference_policy_eval.txt)
REF_POLICY_EVAL = {"destinations_per_episode": 0.4978, "crashes_per_episode": 0.5018, "mean_episode_len": 14.66}   # 10 000 episodes


@pytest.mark.gpu
def test_shipped_policy_through_the_fused_actor_and_the_batched_env(tmp_path):
    """The reference's trained single-learner policy drives 2048 of our single-learner envs (device-side RNG for spawns
    and NPCs).  No random stream is shared with the reference, so the comparison is statistical: destinations, crashes
    and episode length per episode match what the reference's own env gives for the same policy, with the fused actor
    kernel as well as with the PyTorch modules.  (A uniformly random policy reaches the destination in ~10 % of the
    episodes; masking the policy's actions, which the single env does not do, shortens episodes to ~9.4 steps.)"""
    from marl_responsible_nav_b200 import evaluate
    path, _ = _rebuild_reference_file(tmp_path)
    agent = checkpoint.load_reference_checkpoint(path, device="cuda")
    res = {k: evaluate.evaluate(agent, num_envs=2048, episodes=16384, fused_actor=(k == "fused")) for k in ("fused", "torch")}
    for r in res.values():
        assert r["env_kind"] == "single" and r["episodes"] == 16384
        assert abs(r["destinations_per_episode"] - REF_POLICY_EVAL["destinations_per_episode"]) < 0.03, r
        assert abs(r["crashes_per_episode"] - REF_POLICY_EVAL["crashes_per_episode"]) < 0.03, r
        assert abs(r["mean_episode_len"] / REF_POLICY_EVAL["mean_episode_len"] - 1.0) < 0.06, r
