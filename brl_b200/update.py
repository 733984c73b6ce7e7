"""`make_update_step` of src/update.py:9-244: `update_epochs` passes over the rollout, each a
fresh permutation split into `num_minibatches` minibatches; per minibatch the net is re-run,
the PPO loss and its gradient are formed, and the optimizer steps.

Two back ends, selected by the forward pass's precision:
  "tc" / "tc-bf16" (default for ReLU nets)  `brl_ppo_grad`: minibatch take, forward with kept
      activations, loss head + its backward, and the backward GEMMs all in hand-written kernels
      on tcgen05 (csrc/brl_mlp_train.cu, three-term bf16 split = fp32-class gradients), then
      `brl_mlp_adam_step` (clip + Adam + refresh of the kernels' weight layout in one pass) -- no
      library GEMM, no autograd tape;
  "fp32" (and tanh nets)  the minibatch gather, loss head and clip + Adam step are the same
      kernels (csrc/brl_ppo.cu) around plain library GEMMs (cuBLAS fp32 through torch
      autograd) -- kept as the independent cross-check of the tensor-core path.
Functional semantics are kept: the caller's `params` / `opt_state` are not modified (ppo.py
keeps the pre-update params as `opp_params`), a new flat copy is updated and returned."""
from __future__ import annotations

import torch

from . import ops
from . import random as brandom
from .models import LAYERS
from .optim import OptState, flatten_params

_STAT_NAMES = ("total_loss", "value_loss", "loss_actor", "entropy", "approx_kl", "clipflacs", "illegal_action_loss")


class _LossHead(torch.autograd.Function):
    """total_loss(logits, value) with the gradient produced by the same kernel pass."""

    @staticmethod
    def forward(ctx, logits, value, call):
        dlogits, dvalue, stats = torch.empty_like(logits), torch.empty_like(value), call["stats"]
        ops.ppo_loss(logits, value, call["index"], call["mask"], call["action"], call["old_log_prob"], call["old_value"],
                     call["adv"], call["targets"], dlogits, dvalue, stats, call["scratch"], **call["cfg"])
        ctx.save_for_backward(dlogits, dvalue)
        return stats[0].clone()

    @staticmethod
    def backward(ctx, gout):
        dlogits, dvalue = ctx.saved_tensors
        return gout * dlogits, gout * dvalue, None


def _forward_autograd(p, x, act):
    h = x
    for name in LAYERS[:4]:
        h = act(torch.addmm(p[name]["b"], h, p[name]["w"]))
    logits = torch.addmm(p[LAYERS[4]]["b"], h, p[LAYERS[4]]["w"])
    value = torch.addmm(p[LAYERS[5]]["b"], h, p[LAYERS[5]]["w"]).squeeze(-1)
    return logits, value


def make_update_step(config, actor_forward_pass, optimizer, permutation_fn=None):
    """`permutation_fn(rng, batch_size) -> int64/32 tensor` replaces jax.random.permutation
    (src/update.py:193); default: torch.randperm seeded by the key (tests inject a fixed one)."""
    masked = bool(config["actor_illegal_action_mask"])
    cfg = dict(clip_eps=config["clip_eps"], ent_coef=config["ent_coef"], vf_coef=config["vf_coef"],
               illegal_l2_coef=config.get("illegal_action_l2norm_coef", 0.0),
               value_clipping=bool(config.get("value_clipping", True)),
               reward_scaling=bool(config.get("reward_scaling", False)), masked_policy=masked)
    act = getattr(actor_forward_pass, "_activation", torch.relu)
    tensor_core = getattr(actor_forward_pass, "precision", "fp32") in ("tc", "tc-bf16") and act is torch.relu
    tune = int(config.get("brl_train_tune", 0))

    def default_permutation(rng, batch_size, device):
        g = torch.Generator(device="cpu").manual_seed(rng & 0x7FFFFFFFFFFFFFFF)
        return torch.randperm(batch_size, generator=g).to(device=device, dtype=torch.int32)

    def update_step(runner_state, traj_batch, advantages, targets):
        params, opt_state, env_state, last_obs, terminated_count, rng = runner_state
        dev = advantages.device
        nmb, mbs = int(config["num_minibatches"]), int(config["minibatch_size"])
        batch_size = nmb * mbs
        T, n = advantages.shape
        assert batch_size == T * n, "batch size must be equal to number of steps * number of envs"  # src/update.py:189-192
        # flat views of the rollout (src/update.py:195-197)
        obs = traj_batch.obs.reshape(batch_size, -1)
        mask = traj_batch.legal_action_mask.reshape(batch_size, -1).view(torch.uint8).contiguous()
        action = traj_batch.action.reshape(batch_size).contiguous()
        old_lp = traj_batch.log_prob.reshape(batch_size).contiguous()
        old_v = traj_batch.value.reshape(batch_size).contiguous()
        adv, tgt = advantages.reshape(batch_size).contiguous(), targets.reshape(batch_size).contiguous()
        if obs.dtype == torch.bool:
            obs = obs.view(torch.uint8)
        obs = obs.contiguous()

        flat_p, new_params = flatten_params(params)
        if opt_state is None:
            opt_state = optimizer.init(params)
        state = OptState(opt_state.count, opt_state.mu.clone(), opt_state.nu.clone())
        n_epochs = int(config["update_epochs"])
        stats_all = torch.zeros((n_epochs, nmb, 8), dtype=torch.float32, device=dev)
        acc = torch.zeros(16, dtype=torch.float64, device=dev)

        def permutation(_rng):
            return (permutation_fn(_rng, batch_size) if permutation_fn is not None
                    else default_permutation(_rng, batch_size, dev)).to(device=dev, dtype=torch.int32).contiguous()

        def finish():
            # loss_info: (total_loss, (value_loss, loss_actor, entropy, approx_kl, clipflacs, illegal_action_loss)),
            # each [update_epochs, num_minibatches] (src/update.py:170-173, 226-229; read at ppo.py:487-506)
            cols = [stats_all[:, :, i] for i in range(7)]
            return (new_params, state, env_state, last_obs, terminated_count, rng), (cols[0], tuple(cols[1:]))

        if tensor_core:
            blob = ops.mlp_pack_train(flat_p)
            scratch = ops.mlp_train_scratch(mbs, dev)
            flat_g = torch.empty_like(flat_p)
            for epoch in range(n_epochs):
                rng, _rng = brandom.split(rng)                                        # src/update.py:187
                perm = permutation(_rng)
                for mb in range(nmb):                                                  # src/update.py:207-209
                    ops.ppo_grad(obs, blob, scratch, perm[mb * mbs:(mb + 1) * mbs], mask, action, old_lp, old_v, adv, tgt,
                                 flat_g, stats_all[epoch, mb], acc, tune=tune, **cfg)  # src/update.py:164-167
                    state = optimizer.update_mlp_(flat_p, flat_g, state, acc[14:15], blob)  # src/update.py:168-169
            return finish()

        leaves = {name: {k: new_params[name][k].detach().requires_grad_() for k in ("w", "b")} for name in LAYERS}
        flat_g = torch.zeros_like(flat_p)
        off = 0
        for name in LAYERS:  # gradients accumulate straight into the flat buffer the optimizer kernel reads
            for k in ("w", "b"):
                t = leaves[name][k]
                t.grad = flat_g[off: off + t.numel()].view(t.shape)
                off += t.numel()
        call = dict(mask=mask, action=action, old_log_prob=old_lp, old_value=old_v, adv=adv, targets=tgt, cfg=cfg, scratch=acc)
        x_mb = torch.empty((mbs, obs.shape[1]), dtype=obs.dtype, device=dev)
        prev_tf32 = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False  # the reference computes in fp32
        try:
            for epoch in range(n_epochs):
                rng, _rng = brandom.split(rng)                                        # src/update.py:187
                perm = permutation(_rng)
                for mb in range(nmb):                                                  # src/update.py:207-209
                    index = perm[mb * mbs:(mb + 1) * mbs]
                    ops.gather_rows(obs, index, x_mb)
                    x = x_mb.to(torch.float32)                                         # src/update.py:95
                    logits, value = _forward_autograd(leaves, x, act)
                    call["index"], call["stats"] = index, stats_all[epoch, mb]
                    loss = _LossHead.apply(logits.contiguous(), value.contiguous(), call)
                    flat_g.zero_()
                    loss.backward()
                    state = optimizer.update_(flat_p, flat_g, state)                  # src/update.py:168-169
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev_tf32
        return finish()

    return update_step
