"""Phase times (SM clocks) of gram_tail, built with -DBRL_GRAM_TIMING (debug): stamps land behind spec in the scratch."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brl_b200 import ops
dev, B, total = "cuda:0", 1024, 4096
g = torch.Generator().manual_seed(0)
logits = torch.randn((B, 38), generator=g).to(dev)
value = torch.randn(B, generator=g).to(dev)
mask = (torch.rand((total, 38), generator=g) < 0.5); mask[:, 0] = True
mask = mask.to(torch.uint8).to(dev)
index = torch.randperm(total, generator=g)[:B].to(torch.int32).to(dev)
action = torch.zeros(total, dtype=torch.int32, device=dev)
z = torch.zeros(total, device=dev)
dl, dv = torch.empty((B, 38), device=dev), torch.empty(B, device=dev)
stats = torch.zeros(8, device=dev)
acc = ops.ppo_scratch(dev)
for _ in range(3):
    ops.ppo_loss(logits, value, index, mask, action, z, z, z, z, dl, dv, stats, acc, clip_eps=0.2, ent_coef=0.01, vf_coef=0.5, illegal_l2_coef=0.3)
torch.cuda.synchronize()
st = acc.view(torch.int64)[56:62].cpu().tolist()
print("stamps", st)
print("phase clocks: sum-partials %d, zero+trace+fill %d, squarings %d, start+power %d, final %d" % tuple(st[i + 1] - st[i] for i in range(5)))
print("stat", float(stats[6]))
