import torch, torch.nn.functional as F, sys
sys.path.insert(0, "/root/repo")
from marl_responsible_nav_b200 import BatchedGridWorld, FusedActor, maddpg
E = 20000
env = BatchedGridWorld("Level 3", num_envs=E, fear=False, auto_reset=True, seed=5)
agent = maddpg.BatchedMADDPG(2, 160, 9, device="cuda", seed=3)
for a in agent.actors:
    for m in a:
        if isinstance(m, torch.nn.LayerNorm):
            torch.nn.init.normal_(m.weight, 1.0, 0.2); torch.nn.init.normal_(m.bias, 0.0, 0.2)
fused = FusedActor(env, agent.actors, seed=1)
out = env.reset()
bf = lambda t: t.to(torch.bfloat16).float()
def emu(actor, obs, r1, rh1, r2, rh2, r3):
    l1, n1, _, l2, n2, _, l3 = list(actor)[:7]
    z1 = F.linear(obs.float(), bf(l1.weight) if r1 else l1.weight, l1.bias)
    h1 = torch.relu(n1(z1)); h1 = bf(h1) if rh1 else h1
    z2 = F.linear(h1, bf(l2.weight) if r2 else l2.weight, l2.bias)
    h2 = torch.relu(n2(z2)); h2 = bf(h2) if rh2 else h2
    return torch.softmax(F.linear(h2, bf(l3.weight) if r3 else l3.weight, l3.bias), -1)
gen = torch.Generator(device="cuda").manual_seed(0)
for t in range(3):
    cont, ids = fused.forward(out.obs_code, out.action_mask, training=False)
    for k in range(2):
        for name, cfgs in (("fp32", (0,0,0,0,0)), ("all bf16", (1,1,1,1,1)), ("w1 fp32", (0,1,1,1,1)), ("w1,w3 fp32", (0,1,1,1,0)), ("w2 only", (0,1,1,0,0))):
            ref = emu(agent.actors[k], out.obs[:, k], *cfgs)
            err = (cont[:, k] - ref).abs().max().item()
            masked = ref.masked_fill(out.action_mask[:, k] == 0, -1.0)
            dis = masked.argmax(-1) != ids[:, k].long()
            top2 = masked.topk(2, -1).values
            gap = (top2[:, 0] - top2[:, 1])[dis]
            print(t, k, name, f"max|dp| {err:.2e}  argmax disagree {dis.float().mean().item():.4f}  max gap where disagree {gap.max().item() if gap.numel() else 0:.2e}")
    out = env.step(torch.randint(0, 9, (E, 2), generator=gen, device="cuda", dtype=torch.int8))
