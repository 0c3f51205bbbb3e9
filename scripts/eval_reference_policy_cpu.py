"""Build-container only: the reference's OWN single-learner env (custom/customenv.py, imported from /root/reference) driven
by the reference's OWN trained policy (models/custom/single/level3/wo_fear/Single_MADDPG.pt), customeval.py style
(training=False, arg-max of the Gumbel-softmax output, episode ends on termination / truncation or after 150 steps).
Prints destinations / crashes / steps per episode -- the numbers `python -m marl_responsible_nav_b200.evaluate` (or
tests/test_checkpoint.py on the GPU) reports for the same policy in the batched CUDA env with device-side RNG.  The two
runs share no random stream, so they agree statistically, not step by step."""
import contextlib, io, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from _ref_loader import load_reference
from marl_responsible_nav_b200 import checkpoint

REF = load_reference()
SE = REF.customenv
episodes = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
ck = sys.argv[2] if len(sys.argv) > 2 else "/root/reference/models/custom/single/level3/wo_fear/Single_MADDPG.pt"
agent = checkpoint.load_reference_checkpoint(ck, device="cpu")
torch.manual_seed(0); np.random.seed(0)
import random; random.seed(0)
SE.rng = np.random.default_rng(0)
with contextlib.redirect_stdout(io.StringIO()):
    env = SE.CustomEnv(render=False, fear=False)
dest = crash = steps = 0
ret = 0.0
t0 = time.time()
for ep in range(episodes):
    with contextlib.redirect_stdout(io.StringIO()):
        obs, _ = env.reset()
    for t in range(150):
        with torch.no_grad():
            a = int(agent.actors[0](torch.from_numpy(np.asarray(obs, np.float32).reshape(1, -1))).argmax())
        with contextlib.redirect_stdout(io.StringIO()):
            obs, rew, term, trunc, info = env.step([a])
        steps += 1; ret += float(rew[0])
        crash += int(bool(term[0])); dest += int(bool(trunc) and not bool(term[0]))
        if term[0] or trunc:
            break
print(f"reference env + reference policy, {episodes} episodes ({time.time() - t0:.0f} s on one core): "
      f"destinations/episode {dest / episodes:.4f}  crashes/episode {crash / episodes:.4f}  mean length {steps / episodes:.3f}  "
      f"mean return {ret / episodes:.3f}")
