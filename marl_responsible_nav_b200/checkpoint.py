"""The reference's checkpoint format, read and written.

`MADDPGAgent.save_checkpoint / load_checkpoint / load_wo_memory` (maddpg/agent.py:255-283) delegate to AgileRL's
`MADDPG.save_checkpoint`, which `torch.save`s ONE dict.  Its layout is pinned by the files the reference ships
(models/custom/**/*.pt, all read by tests/test_checkpoint.py when the reference is mounted):

    actors_state_dict / actor_targets_state_dict / critics_state_dict / critic_targets_state_dict
        one OrderedDict per agent, keys  feature_net.linear_layer_{0,1}.{weight,bias},
        feature_net.layer_norm_{0,1}.{weight,bias}, feature_net.linear_layer_output.{weight,bias}
    actors_init_dict / actor_targets_init_dict / critics_init_dict / critic_targets_init_dict
        constructor arguments (num_inputs, num_outputs, hidden_size [128, 128], mlp_activation 'ReLU',
        mlp_output_activation 'GumbelSoftmax' | None, layer_norm True, ...)
    actor_optimizers_state_dict / critic_optimizers_state_dict      torch Adam state dicts
    n_agents, agent_ids, state_dims [(160,)], action_dims [9], total_state_dims, total_actions, discrete_actions,
    net_config, batch_size, lr_actor, lr_critic, learn_step, gamma, tau, expl_noise, mean_noise, theta, dt, steps, scores ...

Only plain Python / numpy / torch objects are stored, so neither side needs the other's classes to read a file.
The layer order inside `feature_net` is Linear -> LayerNorm -> ReLU (x2) -> Linear -> output activation, which is the
order of `maddpg.mlp`; `_SEQ_TO_NAME` maps the positions.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn as nn

from . import maddpg

# position in maddpg.mlp's nn.Sequential -> AgileRL module name
_SEQ_TO_NAME = {0: "linear_layer_0", 1: "layer_norm_0", 3: "linear_layer_1", 4: "layer_norm_1", 6: "linear_layer_output"}
_NETS = (("actors", "actors"), ("actor_targets", "actor_targets"), ("critics", "critics"), ("critic_targets", "critic_targets"))


def _to_reference_names(net: nn.Sequential) -> "OrderedDict[str, torch.Tensor]":
    out = OrderedDict()
    for key, val in net.state_dict().items():
        idx, leaf = key.split(".", 1)
        out[f"feature_net.{_SEQ_TO_NAME[int(idx)]}.{leaf}"] = val.detach().cpu().clone()
    return out


def _from_reference_names(sd: Dict[str, torch.Tensor]) -> "OrderedDict[str, torch.Tensor]":
    back = {v: k for k, v in _SEQ_TO_NAME.items()}
    out = OrderedDict()
    for key, val in sd.items():
        parts = key.split(".")
        if len(parts) != 3 or parts[0] != "feature_net" or parts[1] not in back:
            raise ValueError(f"unexpected parameter '{key}' (only the reference's Linear-LayerNorm-ReLU MLPs are supported)")
        out[f"{back[parts[1]]}.{parts[2]}"] = val
    return out


def _init_dict(num_inputs: int, num_outputs: int, hidden: List[int], out_act: Optional[str]) -> Dict:
    return {"num_inputs": num_inputs, "num_outputs": num_outputs, "hidden_size": list(hidden), "num_atoms": 51,
            "mlp_activation": "ReLU", "mlp_output_activation": out_act, "min_hidden_layers": 1, "max_hidden_layers": 3,
            "min_mlp_nodes": 64, "max_mlp_nodes": 500, "layer_norm": True, "init_layers": True, "output_vanish": True,
            "support": None, "rainbow": False, "noise_std": 0.5, "device": torch.device("cpu"), "accelerator": None}


def _portable_optimizer_state(opt: torch.optim.Optimizer) -> Dict:
    """Adam state as the reference writes it: CPU tensors, no device-side step counters (`capturable` is how this
    package keeps the update inside a CUDA graph; a CPU run of the reference could not step such an optimiser)."""
    sd = opt.state_dict()
    state = {k: {n: (v.detach().cpu().clone() if torch.is_tensor(v) else v) for n, v in st.items()} for k, st in sd["state"].items()}
    groups = []
    for g in sd["param_groups"]:
        g = dict(g)
        for key, plain in (("capturable", False), ("fused", None), ("foreach", None)):
            if key in g:
                g[key] = plain
        groups.append(g)
    return {"state": state, "param_groups": groups}


_DILL_TYPES = {"dict": dict, "list": list, "tuple": tuple, "int": int, "float": float, "bool": bool, "str": str, "bytes": bytes,
               "NoneType": type(None), "type": type, "set": set, "frozenset": frozenset, "complex": complex, "slice": slice}


def _dill_load_type(name):
    """Stand-in for `dill._dill._load_type` (the reference saves with `pickle_module=dill`, which spells builtin types this
    way): builtin value types only."""
    if name not in _DILL_TYPES:
        raise ValueError(f"type '{name}' is not allowed in a checkpoint")
    return _DILL_TYPES[name]


def _dill_create_array(f, args, state, npdict=None):
    """Stand-in for `dill._dill._create_array`: numpy's own reconstructor, then the array state."""
    if getattr(f, "__name__", "") != "_reconstruct" or not getattr(f, "__module__", "").startswith("numpy"):
        raise ValueError("array constructor not allowed in a checkpoint")
    arr = f(*args)
    arr.__setstate__(state)
    return arr


def _safe_globals():
    """The only non-tensor types the reference's checkpoint dict holds (see the module docstring): allow-listed for
    `torch.load(weights_only=True)`, so that a `.pt` file given on the command line cannot run code on load."""
    import collections
    allow = [nn.MSELoss, torch.device, collections.OrderedDict, np.dtype, np.ndarray]
    for name in ("scalar", "_reconstruct"):                           # files written under numpy 1.x name numpy.core, 2.x numpy._core
        fn = None
        for mod in ("numpy._core.multiarray", "numpy.core.multiarray"):
            try:
                fn = getattr(__import__(mod, fromlist=[name]), name)
                break
            except (ImportError, AttributeError):
                continue
        if fn is not None:
            allow += [(fn, f"numpy.core.multiarray.{name}"), (fn, f"numpy._core.multiarray.{name}")]
    allow += [(_dill_load_type, "dill._dill._load_type"), (_dill_create_array, "dill._dill._create_array")]
    allow += [type(np.dtype(t)) for t in ("int64", "float64", "float32", "int32", "bool")]
    return allow


def _load_restricted(path: str):
    try:
        with torch.serialization.safe_globals(_safe_globals()):
            return torch.load(path, map_location="cpu", weights_only=True)
    except Exception as exc:                                          # name the type instead of silently unpickling it
        raise ValueError(f"{path}: cannot be read with the restricted unpickler ({type(exc).__name__}: {str(exc)[:300]}); only the "
                         "types of the reference's MADDPG checkpoints are allowed") from exc


def load_reference_checkpoint(path: str, device="cuda", hp: Optional[Dict] = None) -> maddpg.BatchedMADDPG:
    """A BatchedMADDPG with the networks, target networks, optimiser moments and hyper-parameters of a checkpoint written
    by the reference (or by `save_reference_checkpoint`)."""
    ck = _load_restricted(path)
    if not isinstance(ck, dict) or "actors_state_dict" not in ck:
        raise ValueError(f"{path}: not a MADDPG checkpoint of the reference")
    if ck.get("arch", ck.get("net_config", {}).get("arch", "mlp")) != "mlp":
        raise ValueError("only MLP checkpoints are supported (the reference's CNN path never ran on the custom env)")
    n = int(ck["n_agents"])
    obs_dim = int(np.prod(ck["state_dims"][0]))
    act_dim = int(ck["action_dims"][0])
    hidden = [int(h) for h in ck["actors_init_dict"][0]["hidden_size"]]
    h = dict(maddpg.DEFAULT_HP if hp is None else hp)
    for ours, theirs in (("GAMMA", "gamma"), ("TAU", "tau"), ("LR_ACTOR", "lr_actor"), ("LR_CRITIC", "lr_critic"),
                         ("BATCH_SIZE", "batch_size"), ("LEARN_STEP", "learn_step")):
        if theirs in ck:
            h[ours] = type(h[ours])(ck[theirs])
    for ours, theirs in (("EXPL_NOISE", "expl_noise"), ("MEAN_NOISE", "mean_noise")):
        if theirs in ck:                                             # per agent [1, 9] arrays with one value
            h[ours] = float(np.asarray(ck[theirs][0]).reshape(-1)[0])
    agent = maddpg.BatchedMADDPG(n, obs_dim, act_dim, hidden=hidden, hp=h, device=device)
    for ours, theirs in _NETS:
        for net, sd in zip(getattr(agent, ours), ck[f"{theirs}_state_dict"]):
            net.load_state_dict(_from_reference_names(sd))
    for opts, key in ((agent.actor_opt, "actor_optimizers_state_dict"), (agent.critic_opt, "critic_optimizers_state_dict")):
        for opt, sd in zip(opts, ck.get(key, [])):
            try:
                mine = [{k: g[k] for k in ("capturable", "fused", "foreach") if k in g} for g in opt.param_groups]
                cap = [m.get("capturable", False) for m in mine]
                opt.load_state_dict(sd)
                for g, m in zip(opt.param_groups, mine):             # keep this side's choices (CUDA graphs, fused step), not the file's
                    g.update(m)
                if any(cap):
                    for st in opt.state.values():
                        if torch.is_tensor(st.get("step")):
                            st["step"] = st["step"].to(device=agent.device, dtype=torch.float32)
            except (ValueError, KeyError) as exc:                    # a fresh optimiser is a valid way to resume -- but say so
                import warnings
                warnings.warn(f"{path}: optimiser state of '{key}' not restored ({exc}); Adam moments restart from zero")
    agent.steps = list(ck.get("steps", [0]))
    agent.scores = list(ck.get("scores", []))
    return agent


def save_reference_checkpoint(agent: maddpg.BatchedMADDPG, path: str, steps: Optional[List[int]] = None) -> None:
    """Write `agent` as the dict AgileRL's `MADDPG.load_checkpoint` expects: the reference reads it with
    `MADDPGAgent.load_wo_memory` (maddpg/agent.py:279-283).  Its `load_checkpoint` (:268-277) additionally wants `memory.pkl`
    and `steps.txt` next to the file -- the pickled AgileRL replay buffer and env, which this package does not produce."""
    n, hp = agent.n, agent.hp
    if getattr(agent, "learner", None) is not None:
        agent.learner.export_steps()                                 # the kernel counts Adam steps per network
    hidden = [m.out_features for m in agent.actors[0] if isinstance(m, nn.Linear)][:-1]
    crit_in = n * (agent.obs_dim + agent.act_dim)
    noise = lambda v: [np.full((1, agent.act_dim), float(v)) for _ in range(n)]
    ck = {
        "O_U_noise": False, "accelerator": None, "action_dims": [np.int64(agent.act_dim)] * n, "actor_networks": None,
        "agent_ids": [f"agent_{i}" for i in range(n)], "algo": "MADDPG", "arch": "mlp", "batch_size": int(hp["BATCH_SIZE"]),
        "criterion": nn.MSELoss(), "critic_networks": None, "current_noise": noise(0.0), "device": torch.device("cpu"),
        "discrete_actions": True, "dt": 0.01, "expl_noise": noise(hp["EXPL_NOISE"]), "fitness": [], "gamma": float(hp["GAMMA"]),
        "index": 0, "learn_step": int(hp["LEARN_STEP"]), "lr_actor": float(hp["LR_ACTOR"]), "lr_critic": float(hp["LR_CRITIC"]),
        "max_action": None, "mean_noise": noise(hp["MEAN_NOISE"]), "min_action": None, "multi": True, "mut": None, "n_agents": n,
        "net_config": {"arch": "mlp", "hidden_size": list(hidden), "mlp_output_activation": "GumbelSoftmax"}, "one_hot": False,
        "scores": list(getattr(agent, "scores", [])), "state_dims": [(agent.obs_dim,)] * n,
        "steps": list(steps if steps is not None else getattr(agent, "steps", [0])), "tau": float(hp["TAU"]), "theta": 0.15,
        "total_actions": np.int64(agent.act_dim * n), "total_state_dims": agent.obs_dim * n, "vect_noise_dim": 1,
    }
    for ours, theirs in _NETS:
        actor = ours.startswith("actor")
        ck[f"{theirs}_init_dict"] = [_init_dict(agent.obs_dim if actor else crit_in, agent.act_dim if actor else 1, hidden,
                                                "GumbelSoftmax" if actor else None) for _ in range(n)]
        ck[f"{theirs}_state_dict"] = [_to_reference_names(net) for net in getattr(agent, ours)]
    ck["actor_optimizers_state_dict"] = [_portable_optimizer_state(o) for o in agent.actor_opt]
    ck["critic_optimizers_state_dict"] = [_portable_optimizer_state(o) for o in agent.critic_opt]
    torch.save(ck, path)
