"""TEST INFRASTRUCTURE ONLY (never imported by brl_b200/): CPU restatement of the PPO update
arithmetic -- `_loss_fn` of src/update.py:91-162 in float64 PyTorch with autograd standing in
for `jax.value_and_grad` (src/update.py:164-167), and optax's
`chain(clip_by_global_norm, adam(eps=1e-5))` (ppo.py:195-211) in NumPy.

Parity status: the reference's loss needs jax / distrax / optax, none installable here, so this
restatement is pinned only by construction from the cited lines ("parity unpinned" for f1); the
CUDA kernels are checked against it, and it against closed-form cases in tests/test_ppo_ref.py.
"""
from __future__ import annotations

import numpy as np
import torch


def loss_fn(logits, value, mask, action, old_log_prob, old_value, gae, targets, *, clip_eps, ent_coef, vf_coef,
            illegal_l2_coef=0.0, value_clipping=True, reward_scaling=False, masked_policy=True):
    """Returns (total_loss, (value_loss, loss_actor, entropy, approx_kl, clipflacs, illegal_action_loss))."""
    neg_inf = torch.tensor(float("-inf"), dtype=logits.dtype)
    masked_logits = torch.where(mask, logits, neg_inf)                 # src/update.py:14 (select semantics)
    logp_masked = torch.log_softmax(masked_logits, dim=1)
    logp_policy = logp_masked if masked_policy else torch.log_softmax(logits, dim=1)   # :12-24
    log_prob = logp_policy.gather(1, action.long()[:, None])[:, 0]    # :98
    if value_clipping:                                                 # :49-60
        v_clipped = old_value + (value - old_value).clamp(-clip_eps, clip_eps)
        value_loss = 0.5 * torch.maximum((value - targets) ** 2, (v_clipped - targets) ** 2).mean()
    else:                                                              # :66-68
        value_loss = 0.5 * ((value - targets) ** 2).mean()
    logratio = log_prob - old_log_prob                                 # :116-117
    ratio = torch.exp(logratio)
    if reward_scaling:                                                 # :35-36
        gae = (gae - gae.mean()) / (gae.std(unbiased=False) + 1e-8)
    loss_actor = -torch.minimum(ratio * gae, ratio.clamp(1.0 - clip_eps, 1.0 + clip_eps) * gae).mean()   # :121-131
    p_masked = torch.exp(logp_masked)
    # 0 * log 0 := 0 (distrax); the select is applied BEFORE the product so autograd never sees 0 * inf
    entropy = -(p_masked * torch.where(mask, logp_masked, torch.zeros_like(logp_masked))).sum(1).mean()  # :133-137
    probs = torch.softmax(logits, dim=1)                               # :139-140
    # jnp.linalg.norm(x, ord=2) of a 2-D [minibatch, 38] array is the SPECTRAL norm (largest singular value), not Frobenius
    illegal_action_loss = torch.linalg.matrix_norm(probs * (~mask), ord=2) / 2                          # :141-142
    total = loss_actor + vf_coef * value_loss - ent_coef * entropy + illegal_l2_coef * illegal_action_loss  # :144-149
    approx_kl = ((ratio - 1) - logratio).mean()                        # :153
    clipflacs = ((ratio - 1.0).abs() > clip_eps).to(logits.dtype).mean()   # :154-156
    return total, (value_loss, loss_actor, entropy, approx_kl, clipflacs, illegal_action_loss)


def adam_clip_step(p, g, m, v, count, *, lr, max_grad_norm, b1=0.9, b2=0.999, eps=1e-5):
    """One optax.chain(clip_by_global_norm(c), adam(lr, eps)) step on flat float64 arrays.
    `count` = steps taken before this one.  Returns (p, m, v)."""
    g = np.asarray(g, np.float64)
    if max_grad_norm and max_grad_norm > 0:
        norm = np.sqrt((g * g).sum())
        if not norm < max_grad_norm:                                   # optax: where(norm < c, g, g / norm * c)
            g = g / norm * max_grad_norm
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    t = count + 1
    mhat, vhat = m / (1 - b1 ** t), v / (1 - b2 ** t)
    return p - lr * mhat / (np.sqrt(vhat) + eps), m, v


def mlp_forward_torch(params, x):
    """src/models.py:23-33 on a dict {'w0'..'w5','b0'..'b5'} of torch tensors (w [in,out])."""
    h = x
    for i in range(4):
        h = torch.relu(h @ params[f"w{i}"] + params[f"b{i}"])
    return h @ params["w4"] + params["b4"], (h @ params["w5"] + params["b5"])[:, 0]
