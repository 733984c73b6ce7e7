"""Per-element error of the minibatch gradient: tcgen05 split path and library fp32 path vs float64 autograd."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brl_b200 import ops  # noqa: E402
from brl_b200.models import LAYERS, init_params, params_to_numpy  # noqa: E402
from brl_b200.optim import flatten_params  # noqa: E402
from oracle import ppo_ref  # noqa: E402

DEV = "cuda:0"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
g = torch.Generator().manual_seed(3)
obs = (torch.rand((B, 480), generator=g) < 0.04)
mask = torch.rand((B, 38), generator=g) < 0.5
mask[:, 0] = True
params = init_params(9, DEV)
pn = params_to_numpy(params)
ref = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in pn.items()}
lg, vl = ppo_ref.mlp_forward_torch(ref, obs.double())
with torch.no_grad():
    ml = torch.where(mask, lg, torch.tensor(float("-inf"), dtype=torch.float64))
    action = torch.distributions.Categorical(logits=ml).sample().to(torch.int32)
    lp = torch.log_softmax(ml, 1).gather(1, action[:, None].long())[:, 0]
    old_lp = lp + (torch.rand(B, generator=g, dtype=torch.float64) - 0.5) * 0.6
    adv = torch.randn(B, generator=g, dtype=torch.float64)
    old_v = vl.clone()
    tgt = vl + torch.randn(B, generator=g, dtype=torch.float64) * 0.2
cfgk = dict(clip_eps=0.2, ent_coef=0.001, vf_coef=0.5)
total, _ = ppo_ref.loss_fn(lg, vl, mask, action, old_lp, old_v, adv, tgt, **cfgk)
total.backward()
order = [f"{c}{i}" for i in range(6) for c in ("w", "b")]
want = np.concatenate([ref[k].grad.numpy().reshape(-1) for k in order])
f32 = lambda t: t.to(torch.float32).to(DEV).contiguous()  # noqa: E731
cfg = dict(illegal_l2_coef=0.0, value_clipping=True, reward_scaling=False, masked_policy=True, **cfgk)
# tensor-core path
flat_p, _ = flatten_params(params)
blob = ops.mlp_pack_train(flat_p)
grads = torch.empty_like(flat_p)
stats = torch.zeros(8, device=DEV)
acc = ops.ppo_scratch(DEV)
ops.ppo_grad(f32(obs), blob, ops.mlp_train_scratch(B, DEV), None, mask.to(torch.uint8).to(DEV), action.to(DEV), f32(old_lp), f32(old_v),
             f32(adv), f32(tgt), grads, stats, acc, **cfg)
got_tc = grads.cpu().numpy().astype(np.float64)
# library fp32 path
torch.backends.cuda.matmul.allow_tf32 = False
leaves = {name: {k: params[name][k].detach().clone().requires_grad_() for k in ("w", "b")} for name in LAYERS}
h = f32(obs)
for name in LAYERS[:4]:
    h = torch.relu(torch.addmm(leaves[name]["b"], h, leaves[name]["w"]))
logits = torch.addmm(leaves[LAYERS[4]]["b"], h, leaves[LAYERS[4]]["w"])
value = torch.addmm(leaves[LAYERS[5]]["b"], h, leaves[LAYERS[5]]["w"]).squeeze(-1)
t32, _ = ppo_ref.loss_fn(logits, value, mask.to(DEV), action.to(DEV), f32(old_lp), f32(old_v), f32(adv), f32(tgt), **cfgk)
t32.backward()
got_32 = np.concatenate([leaves[n][k].grad.cpu().numpy().reshape(-1) for n in LAYERS for k in ("w", "b")]).astype(np.float64)
for name, got in (("tc", got_tc), ("fp32", got_32)):
    err = np.abs(got - want)
    nz = np.abs(want) > 0
    rel = err[nz] / np.abs(want[nz])
    print(name, "rel L2", np.linalg.norm(got - want) / np.linalg.norm(want), "max abs", err.max(), "max |g|", np.abs(want).max(),
          "median |g|", np.median(np.abs(want[nz])), "rel-err quantiles 0.5/0.9/0.99/0.999",
          [float(np.quantile(rel, q)) for q in (0.5, 0.9, 0.99, 0.999)], "frac rel>1e-3", float((rel > 1e-3).mean()))
    off = 0
    for k in order:
        n = ref[k].numel()
        e, w = got[off:off + n] - want[off:off + n], want[off:off + n]
        print("   ", k, "rel L2 %.3e" % (np.linalg.norm(e) / max(np.linalg.norm(w), 1e-300)), "max abs %.3e" % np.abs(e).max(), "max |g| %.3e" % np.abs(w).max())
        off += n
