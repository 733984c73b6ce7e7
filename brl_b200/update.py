"""`make_update_step` of src/update.py:9-244: `update_epochs` passes over the rollout, each a
fresh permutation split into `num_minibatches` minibatches; per minibatch the net is re-run,
the PPO loss and its gradient are formed, and the optimizer steps.

One back end: `brl_ppo_grad` -- minibatch take, forward with kept activations, loss head + its backward, and the
backward GEMMs all in hand-written kernels on tcgen05 (csrc/brl_mlp_train.cu, three-term bf16 split = fp32-class
gradients) -- then `brl_mlp_adam_step` (clip + Adam + refresh of the kernels' weight layout in one pass).  No library
GEMM, no autograd tape.  (The cuBLAS-fp32 + autograd cross-check of this path is scripts/torch_baseline.py.)
Functional semantics are kept: the caller's `params` / `opt_state` are not modified (ppo.py
keeps the pre-update params as `opp_params`), a new flat copy is updated and returned."""
from __future__ import annotations

import torch

from . import ops
from . import random as brandom
from .optim import OptState, flatten_params

_STAT_NAMES = ("total_loss", "value_loss", "loss_actor", "entropy", "approx_kl", "clipflacs", "illegal_action_loss")


def make_update_step(config, actor_forward_pass, optimizer, permutation_fn=None):
    """`permutation_fn(rng, batch_size) -> int64/32 tensor` replaces jax.random.permutation
    (src/update.py:193); default: torch.randperm seeded by the key (tests inject a fixed one)."""
    masked = bool(config["actor_illegal_action_mask"])
    cfg = dict(clip_eps=config["clip_eps"], ent_coef=config["ent_coef"], vf_coef=config["vf_coef"],
               illegal_l2_coef=config.get("illegal_action_l2norm_coef", 0.0),
               value_clipping=bool(config.get("value_clipping", True)),
               reward_scaling=bool(config.get("reward_scaling", False)), masked_policy=masked)
    # illegal_action_loss = ||probs * ~mask||_2 / 2 is the SPECTRAL norm of the [minibatch, 38] matrix (src/update.py:141-142): a
    # 38 x 38 eigen-problem per minibatch.  With the default illegal_action_l2norm_coef = 0 it is a logged statistic only,
    # and ppo.py logs just the LAST minibatch of the last epoch (ppo.py:505, `illegal_action_loss[-1][-1]`).  "last" (default)
    # computes it for the last minibatch of every epoch and leaves NaN elsewhere in loss_info -- never a wrong number;
    # "all" fills every entry as the reference does, at ~19 us (11 %) per optimizer step (scripts/exp_gram_cost.py).
    # A non-zero coefficient needs the norm for the gradient and always computes it.
    stat_mode = config.get("illegal_action_loss_stat", "last")
    if stat_mode not in ("last", "all"):
        raise ValueError("illegal_action_loss_stat must be 'last' or 'all'")
    if getattr(actor_forward_pass, "precision", None) not in ("tc", "tc-bf16"):
        raise TypeError("make_update_step needs a brl_b200.models.ForwardPass (tensor-core ReLU net); there is no library "
                        "back end on the product path")
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
        acc = ops.ppo_scratch(dev)

        def permutation(_rng):
            return (permutation_fn(_rng, batch_size) if permutation_fn is not None
                    else default_permutation(_rng, batch_size, dev)).to(device=dev, dtype=torch.int32).contiguous()

        def finish():
            # loss_info: (total_loss, (value_loss, loss_actor, entropy, approx_kl, clipflacs, illegal_action_loss)),
            # each [update_epochs, num_minibatches] (src/update.py:170-173, 226-229; read at ppo.py:487-506)
            cols = [stats_all[:, :, i] for i in range(7)]
            return (new_params, state, env_state, last_obs, terminated_count, rng), (cols[0], tuple(cols[1:]))

        blob = ops.mlp_pack_train(flat_p)
        scratch = ops.mlp_train_scratch(mbs, dev)
        flat_g = torch.empty_like(flat_p)
        for epoch in range(n_epochs):
            rng, _rng = brandom.split(rng)                                        # src/update.py:187
            perm = permutation(_rng)
            for mb in range(nmb):                                                  # src/update.py:207-209
                ops.ppo_grad(obs, blob, scratch, perm[mb * mbs:(mb + 1) * mbs], mask, action, old_lp, old_v, adv, tgt,
                             flat_g, stats_all[epoch, mb], acc, tune=tune, illegal_stat=stat_mode == "all" or mb == nmb - 1,
                             **cfg)                                                # src/update.py:164-167
                state = optimizer.update_mlp_(flat_p, flat_g, state, acc[14:15], blob)  # src/update.py:168-169
        return finish()

    return update_step
