"""Fixture from the reference's shipped checkpoints (run in the build container, where /root/reference is mounted):

    python tests/golden/make_checkpoint_golden.py

Writes tests/golden/ref_single_actor_level3.npz: the ACTOR of models/custom/single/level3/wo_fear/Single_MADDPG.pt (the
reference's longest Level-3 training run, 607 466 steps; 38 793 parameters) under the file's own parameter names, the
logits that a plain numpy forward of those parameters gives on 32 seeded observations, and the scalar fields the
loader maps.  The GPU box has no /root/reference: tests/test_checkpoint.py uses this file there to rebuild a checkpoint
in the reference's format, load it, and run the trained policy through the fused actor kernel.
"""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("GW_REFERENCE", "/root/reference")
SRC = os.path.join(REF, "models/custom/single/level3/wo_fear/Single_MADDPG.pt")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_single_actor_level3.npz")


def numpy_logits(sd, x):
    def ln(h, g, b):
        m = h.mean(-1, keepdims=True)
        v = ((h - m) ** 2).mean(-1, keepdims=True)
        return (h - m) / np.sqrt(v + 1e-5) * g + b
    p = {k.split("feature_net.")[1]: v.double().numpy() for k, v in sd.items()}
    h = x @ p["linear_layer_0.weight"].T + p["linear_layer_0.bias"]
    h = np.maximum(ln(h, p["layer_norm_0.weight"], p["layer_norm_0.bias"]), 0)
    h = h @ p["linear_layer_1.weight"].T + p["linear_layer_1.bias"]
    h = np.maximum(ln(h, p["layer_norm_1.weight"], p["layer_norm_1.bias"]), 0)
    return h @ p["linear_layer_output.weight"].T + p["linear_layer_output.bias"]


def main():
    ck = torch.load(SRC, map_location="cpu", weights_only=False)
    sd = ck["actors_state_dict"][0]
    rng = np.random.default_rng(7)
    x = rng.choice(np.array([-1.0, 0.0, 1.0, 2.0, 3.0, 4.0, 9.0, 10.0]), size=(32, 160), p=[.5, .4, .02, .02, .02, .02, .01, .01])
    arrays = {"param/" + k: v.numpy() for k, v in sd.items()}
    arrays["probe_obs"] = x.astype(np.float32)
    arrays["probe_logits"] = numpy_logits(sd, x)
    for k in ("gamma", "tau", "lr_actor", "lr_critic", "batch_size", "learn_step", "n_agents"):
        arrays["scalar/" + k] = np.asarray(ck[k])
    arrays["scalar/expl_noise"] = np.asarray(ck["expl_noise"][0])
    arrays["scalar/steps"] = np.asarray(ck["steps"])
    arrays["keys"] = np.array(sorted(ck.keys()))
    np.savez_compressed(OUT, **arrays)
    print(OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    sys.exit(main())
