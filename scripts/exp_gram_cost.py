"""What does the spectral-norm illegal_action_loss cost per optimizer step?  Interleaved A/B in one process:
illegal_stat=True (partials before the loss head, tail beside the bias sums) vs illegal_stat=False (no statistic) vs illegal_l2_coef = 0.3 (norm needed before
the loss gradient: stand-alone kernel on the critical path)."""
import json
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brl_b200 import ops  # noqa: E402
from brl_b200.models import init_params  # noqa: E402
from brl_b200.optim import AdamWithClip, flatten_params  # noqa: E402

dev, B, total = "cuda:0", 1024, 8192 * 32
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
params = init_params(1, dev)
opt = AdamWithClip(1e-4, eps=1e-5, max_grad_norm=0.5)
flat_p, _ = flatten_params(params)
state = [opt.init(params)]
blob = ops.mlp_pack_train(flat_p)
scratch = ops.mlp_train_scratch(B, dev)
flat_g = torch.empty_like(flat_p)
stats = torch.zeros(8, dtype=torch.float32, device=dev)
acc = ops.ppo_scratch(dev)
nmb = total // B


def run(stat, coef, iters=256):
    cfg = dict(clip_eps=0.2, ent_coef=0.01, vf_coef=0.5, illegal_l2_coef=coef, value_clipping=True, reward_scaling=False,
               masked_policy=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(iters):
        mb = i % nmb
        ops.ppo_grad(obs, blob, scratch, perm[mb * B:(mb + 1) * B], mask, action, old_lp, old_v, adv, tgt, flat_g, stats, acc,
                     illegal_stat=stat, **cfg)
        state[0] = opt.update_mlp_(flat_p, flat_g, state[0], acc[14:15], blob)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


cases = {"deferred_stat": (True, 0.0), "no_stat": (False, 0.0), "coef_0.3_on_critical_path": (True, 0.3)}
for c in cases.values():
    run(*c, iters=50)
res = {k: [] for k in cases}
for rep in range(5):
    for k, c in cases.items():
        res[k].append(run(*c))
print(json.dumps({k: {"ms_per_step_median": statistics.median(v), "all": [round(x, 5) for x in v]} for k, v in res.items()}))
