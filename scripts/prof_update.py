"""A few PPO minibatch gradients (brl_ppo_grad) at ppo.py's minibatch size, for ncu.  tune from argv."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brl_b200 import ops  # noqa: E402
from brl_b200.models import init_params  # noqa: E402
from brl_b200.optim import flatten_params  # noqa: E402

tune = int(sys.argv[1]) if len(sys.argv) > 1 else 0
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev, B, total = "cuda:0", 1024, 8192 * 4
g = torch.Generator().manual_seed(0)
obs = (torch.rand((total, 480), generator=g) < 0.05).to(torch.bfloat16).to(dev)
mask = (torch.rand((total, 38), generator=g) < 0.5)
mask[:, 0] = True
mask = mask.to(torch.uint8).to(dev)
action = torch.zeros(total, dtype=torch.int32, device=dev)
old_lp = (-torch.rand(total, generator=g) * 3).to(dev)
old_v = (torch.randn(total, generator=g) * 0.3).to(dev)
adv = torch.randn(total, generator=g).to(dev)
tgt = (torch.randn(total, generator=g) * 0.3).to(dev)
perm = torch.randperm(total, generator=g).to(torch.int32).to(dev)
flat_p, _ = flatten_params(init_params(1, dev))
blob = ops.mlp_pack_train(flat_p)
scratch = ops.mlp_train_scratch(B, dev)
flat_g = torch.empty_like(flat_p)
stats = torch.zeros(8, dtype=torch.float32, device=dev)
acc = ops.ppo_scratch(dev)
cfg = dict(clip_eps=0.2, ent_coef=0.01, vf_coef=0.5, illegal_l2_coef=0.0, value_clipping=True, reward_scaling=False, masked_policy=True)
for i in range(iters):
    ops.ppo_grad(obs, blob, scratch, perm[i * B:(i + 1) * B], mask, action, old_lp, old_v, adv, tgt, flat_g, stats, acc, tune=tune, **cfg)
torch.cuda.synchronize()
print("ok", float(stats[0]))
