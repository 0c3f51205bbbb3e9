"""Batched counterpart of main_custom.py: `python -m marl_responsible_nav_b200.train --config custom_fear_5 --envs 4096`.

Same configuration keys as configs/custom*.yaml (a YAML path or one of the preset names); POSIX paths (the reference
hard-codes Windows-style "configs\\custom.yaml", main_custom.py:39).  Launch with torchrun for several GPUs: every rank
owns a contiguous shard of the environments, gradients and episode statistics are all-reduced over NCCL.
"""
import argparse
import json
import os
import time

import torch
import torch.distributed as dist

from . import maddpg, sharding


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="custom", help="preset (custom, custom_fear_1/3/5/10) or path to a YAML file")
    ap.add_argument("--envs", type=int, default=4096, help="environments over all GPUs")
    ap.add_argument("--steps", type=int, default=1500, help="synchronous env steps (the reference: MAX_EPISODES x TRAIN_STEPS)")
    ap.add_argument("--report", type=int, default=150)
    ap.add_argument("--env-kind", choices=("multi", "single"), default="multi", help="multi: custom/ma_customenv.py (2 learners); single: custom/customenv.py")
    ap.add_argument("--learn-cadence", choices=("reference", "batched"), default="reference",
                    help="reference: maddpg/agent.py:199-224 on the global env count (envs // LEARN_STEP updates after every "
                         "vector step); batched: --updates-per-learn updates every LEARN_STEP vector steps")
    ap.add_argument("--updates-per-learn", type=int, default=1, help="batched cadence only")
    ap.add_argument("--torch-actor", action="store_true", help="act with the PyTorch modules instead of the fused kernel")
    ap.add_argument("--torch-sampler", action="store_true", help="sample the replay ring with PyTorch indexing instead of gw_replay_sample")
    ap.add_argument("--torch-ops", action="store_true", help="LayerNorm + ReLU of the update by PyTorch instead of gw_ln_relu_forward / _backward")
    ap.add_argument("--torch-learner", action="store_true", help="the update as round 1's CUDA graph of PyTorch / library kernels instead of the one-kernel gw_learner_update")
    ap.add_argument("--gradient-exchange", choices=("peer", "nccl"), default="peer",
                    help="several GPUs: peer = inside the update kernel over NVLink peer memory; nccl = two all-reduces per update")
    ap.add_argument("--scenario-json", default=None, help="a file in the reference's Scenarios.json format (any map up to 64 x 64, up to 16 agents: "
                                                          "what does not fit the packed 10 x 16 x 4 layout runs on the general one)")
    ap.add_argument("--scenario", default="Level 3", help="scenario name (built-in, or a key of --scenario-json)")
    ap.add_argument("--walls", choices=("refuse", "inert", "enforce"), default="refuse", help="what to do with Walls / OneWays of --scenario-json")
    ap.add_argument("--save", default=None, help="write the agents here in the reference's checkpoint format (maddpg/agent.py:255-266)")
    ap.add_argument("--load", default=None, help="resume from a checkpoint of the reference / of --save (maddpg/agent.py:268-283)")
    a = ap.parse_args()
    hp = maddpg.load_yaml_config(a.config) if os.path.exists(a.config) else maddpg.preset(a.config)
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    base, n = sharding.shard_range(a.envs, rank, world)
    scenario = a.scenario
    if a.scenario_json:
        from .scenarios import load_scenario_json
        scenario = load_scenario_json(a.scenario_json, a.scenario, walls=a.walls)
    env = maddpg.make_env(hp, n, device=dev, scenario=scenario, env_id_base=base, env_kind=a.env_kind)
    agent = None
    if a.load:
        from . import checkpoint
        agent = checkpoint.load_reference_checkpoint(a.load, device=dev, hp=hp)
    trainer = maddpg.BatchedTrainer(env, agent=agent, hp=hp, updates_per_learn=a.updates_per_learn, seed=hp["SEED"],
                                    fused_actor=not a.torch_actor, fused_sampler=not a.torch_sampler, fused_ops=not a.torch_ops,
                                    learn_cadence=a.learn_cadence, global_envs=a.envs, fused_learner=not a.torch_learner,
                                    gradient_exchange=a.gradient_exchange)
    exchange = trainer.connect(0)
    if rank == 0 and world > 1:
        print(json.dumps({"gradient_exchange": exchange}), flush=True)
    done = 0
    while done < a.steps:
        k = min(a.report, a.steps - done)
        torch.cuda.synchronize()
        upd0 = trainer.updates_done
        t0 = time.perf_counter()
        st = trainer.train(k)
        torch.cuda.synchronize()
        el = sharding.max_over_ranks(time.perf_counter() - t0)
        st = sharding.allreduce_stats(st)
        done += k
        if rank == 0:
            eps = max(1, st["episodes"])
            last = trainer.losses[-1] if trainer.losses else None
            print(json.dumps({"env_steps": done, "agent_steps_per_s": st["agent_steps"] / el, "episodes": st["episodes"],
                              "mean_return": st["return_sum"] / eps, "mean_len": st["episode_len_sum"] / eps,
                              "crashes_per_episode": st["crashes"] / eps, "apples_per_episode": st["apples"] / eps,
                              "fear_sum": st["fear_sum"], "updates": trainer.updates_done,
                              "updates_per_s": (trainer.updates_done - upd0) / el,
                              "critic_loss": None if last is None else float(last.critic_loss.sum()),
                              "actor_loss": None if last is None else float(last.actor_loss.sum())}), flush=True)
    if a.save and rank == 0:
        from . import checkpoint
        checkpoint.save_reference_checkpoint(trainer.agent, a.save, steps=[done * a.envs])
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
