"""A few policy forwards (brl_policy_act: the fused persistent CTA-pair launch k_mlp_fused) at 8192 envs, for ncu."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brl_b200 import ops  # noqa: E402
from brl_b200.models import init_params, make_forward_pass  # noqa: E402

dev, n = "cuda:0", int(sys.argv[1]) if len(sys.argv) > 1 else 8192
prec = sys.argv[2] if len(sys.argv) > 2 else "tc"
g = torch.Generator().manual_seed(0)
obs = ops.obs_to_bf16((torch.rand((n, 480), generator=g) < 0.05).to(torch.float32).to(dev))
mask = torch.ones((n, 38), dtype=torch.uint8, device=dev)
action = torch.empty(n, dtype=torch.int32, device=dev)
lp, val = torch.empty(n, device=dev), torch.empty(n, device=dev)
fp = make_forward_pass(precision=prec)
params = init_params(1, dev)
for i in range(6):
    fp.act(params, obs, mask, action, lp, val, sample=True, seed=i)
torch.cuda.synchronize()
print("ok", int(action.sum()))
