"""Batched counterpart of customeval.py: roll a trained MADDPG policy through E environments at once.

    python -m marl_responsible_nav_b200.evaluate --checkpoint models/custom/single/level3/fear/Single_MADDPG_4k.pt \
        --episodes 10000 --envs 4096 [--fear] [--torch-actor]

Reads the reference's own checkpoint files (checkpoint.py), acts greedily (`training=False`, customeval.py:90-94: no
exploration noise; the Gumbel-softmax output layer still draws its noise, as in the reference), steps the batched env
and prints the three totals customeval.py:134-136 prints -- destinations reached, crashes, steps -- over the requested
number of episodes (every env plays the same number of episodes to their end), plus mean return / FeAR.  A one-agent checkpoint runs on the single-learner env
(custom/customenv.py), a two-agent one on the multi-agent env (custom/ma_customenv.py).
"""
import argparse
import json
import time

import torch

from . import checkpoint, maddpg
from .batched import BatchedGridWorld


def evaluate(agent: maddpg.BatchedMADDPG, num_envs: int = 4096, episodes: int = 1000, fear: bool = False, max_steps: int = 150,
             scenario: str = "Level 3", seed: int = 42, device="cuda", fused_actor: bool = True):
    kind = "single" if agent.n == 1 else "multi"
    env = BatchedGridWorld(scenario, num_envs=num_envs, device=device, env_kind=kind, fear=fear, max_steps=max_steps,
                           auto_reset=True, seed=seed)
    if env.obs_len != agent.obs_dim or env.n_learners != agent.n:
        raise ValueError(f"checkpoint ({agent.n} agents, {agent.obs_dim} inputs) does not fit the {kind} env "
                         f"({env.n_learners} learners, {env.obs_len} cells)")
    fused = None
    if fused_actor:
        from .actor import FusedActor
        fused = FusedActor(env, agent.actors, seed=seed + 2)
    # Every env plays exactly `k` episodes to their end and only those count: stopping at a total instead would leave the
    # long episodes (a policy that stands still until the step cap) unfinished and under-counted.
    k = -(-int(episodes) // num_envs)
    dev = env.device
    played = torch.zeros(num_envs, dtype=torch.int32, device=dev)
    acc = torch.zeros(5, dtype=torch.float64, device=dev)            # destinations, crashes, steps, return, fear
    vals = torch.zeros((5, num_envs), dtype=torch.float64, device=dev)
    vals[2] = 1.0
    out = env.reset()
    torch.cuda.synchronize(dev)
    t0, steps = time.perf_counter(), 0
    check_every = max(1, max_steps // 10)
    # the multi-agent env hands the policy an action mask (info[agent]["action_mask"], ma_customenv.py:324-332); the
    # single-learner env has none (customenv.py:165-183), so its policy may well walk into a wall (restricted move)
    masked = kind == "multi"
    while True:
        for _ in range(check_every):
            mask = out.action_mask if masked else None
            if fused is not None:
                _, ids = fused.forward(out.obs_code, mask, training=False, gumbel=True)   # as the torch module below
            else:
                _, ids = agent.get_action(out.obs, mask, training=False)
            out = env.step(ids)
            live = played < k
            info = out.info
            vals[0] = (info >> 10) & 3                      # apples rewarded in this step
            vals[1] = (info >> 8) & 3                       # learner crashes
            vals[3] = out.reward.sum(dim=1)
            if fear:
                vals[4] = out.fear.sum(dim=1)
            acc += (vals * live).sum(dim=1)                 # vals[2] stays 1: steps
            played += (out.ended != 0) & live
        steps += check_every
        if bool((played >= k).all()):                       # synchronises; every `check_every` steps only
            break
    el = time.perf_counter() - t0
    dest, crashes, n_steps, ret, fear_sum = (float(v) for v in acc.cpu())
    eps = k * num_envs
    res = {"episodes": eps, "destinations_reached": int(dest), "crashes": int(crashes), "total_steps": int(n_steps),
           "mean_return": ret / eps, "mean_episode_len": n_steps / eps, "crashes_per_episode": crashes / eps,
           "destinations_per_episode": dest / eps, "fear_sum": fear_sum, "env_kind": kind, "num_envs": num_envs,
           "agent_steps_per_s": num_envs * env.n_learners * steps / el, "actor": "fused" if fused is not None else "torch"}
    if fused is not None:
        fused.close()
    env.close()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--checkpoint", required=True, help="a MADDPG.pt written by the reference (or by checkpoint.save_reference_checkpoint)")
    ap.add_argument("--episodes", type=int, default=100, help="finished episodes to count (customeval.py: eval_episodes = 100)")
    ap.add_argument("--envs", type=int, default=1024)
    ap.add_argument("--fear", action="store_true", help="also compute FeAR for every step (reported, never shapes the evaluation)")
    ap.add_argument("--max-steps", type=int, default=150, help="TRAIN_STEPS of the config")
    ap.add_argument("--scenario", default="Level 3")
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--torch-actor", action="store_true", help="act with the PyTorch modules instead of the fused kernel")
    a = ap.parse_args()
    agent = checkpoint.load_reference_checkpoint(a.checkpoint, device="cuda")
    r = evaluate(agent, a.envs, a.episodes, a.fear, a.max_steps, a.scenario, a.seed, fused_actor=not a.torch_actor)
    print(f"Total destination reached: {r['destinations_reached']} across {r['episodes']} episodes")
    print(f"Total crashes: {r['crashes']} across {r['episodes']} episodes")
    print(f"Total steps: {r['total_steps']} across {r['episodes']} episodes")
    print(json.dumps(r))


if __name__ == "__main__":
    main()
