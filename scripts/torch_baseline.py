"""Library-GEMM cross-checks of the tensor-core kernels -- NOT part of the product path.

`brl_b200` runs the policy net and the PPO update only through its own hand-written tcgen05 kernels (no multi-backend
dispatch).  What used to be the `precision="fp32" / "tf32" / "bf16"` branch of `brl_b200.models.ForwardPass` and the
autograd branch of `brl_b200.update.make_update_step` lives here, for tests / bench comparison legs / experiment scripts:

  TorchForwardPass         forward.apply / .act through cuBLAS (`torch.addmm`), any activation
  make_update_step_autograd  src/update.py:74-242 with cuBLAS fp32 GEMMs through torch autograd around the same
                           loss-head and clip + Adam kernels (csrc/brl_ppo.cu)
"""
from __future__ import annotations

import torch

from brl_b200 import ops
from brl_b200 import random as brandom
from brl_b200.models import LAYERS
from brl_b200.optim import OptState, flatten_params


class TorchForwardPass:
    def __init__(self, activation: str = "relu", precision: str = "fp32"):
        if precision not in ("fp32", "tf32", "bf16"):
            raise ValueError(f"unknown library precision {precision!r}")
        self._activation = torch.relu if activation == "relu" else torch.tanh
        self.precision = precision

    @property
    def input_dtype(self):
        return torch.float32

    def apply(self, params, x: torch.Tensor):
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = self.precision == "tf32"
        try:
            dt = torch.bfloat16 if self.precision == "bf16" else torch.float32
            h = x.to(dt)
            for name in LAYERS[:4]:
                h = self._activation(torch.addmm(params[name]["b"].to(dt), h, params[name]["w"].to(dt)))
            logits = torch.addmm(params[LAYERS[4]]["b"].to(dt), h, params[LAYERS[4]]["w"].to(dt)).float()
            value = torch.addmm(params[LAYERS[5]]["b"].to(dt), h, params[LAYERS[5]]["w"].to(dt)).float().squeeze(-1)
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev
        return logits, value

    def act(self, params, x, mask, action, log_prob=None, value=None, *, sample: bool, seed: int, env_offset: int = 0):
        logits, v = self.apply(params, x)
        ops.categorical(logits.contiguous(), mask, action, log_prob, sample=sample, seed=seed, env_offset=env_offset)
        if value is not None:
            value.copy_(v)


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


def make_update_step_autograd(config, actor_forward_pass, optimizer, permutation_fn=None):
    """Same contract as brl_b200.update.make_update_step; GEMMs = cuBLAS fp32 via torch autograd."""
    masked = bool(config["actor_illegal_action_mask"])
    cfg = dict(clip_eps=config["clip_eps"], ent_coef=config["ent_coef"], vf_coef=config["vf_coef"],
               illegal_l2_coef=config.get("illegal_action_l2norm_coef", 0.0),
               value_clipping=bool(config.get("value_clipping", True)),
               reward_scaling=bool(config.get("reward_scaling", False)), masked_policy=masked)
    act = getattr(actor_forward_pass, "_activation", torch.relu)

    def default_permutation(rng, batch_size, device):
        g = torch.Generator(device="cpu").manual_seed(rng & 0x7FFFFFFFFFFFFFFF)
        return torch.randperm(batch_size, generator=g).to(device=device, dtype=torch.int32)

    def update_step(runner_state, traj_batch, advantages, targets):
        params, opt_state, env_state, last_obs, terminated_count, rng = runner_state
        dev = advantages.device
        nmb, mbs = int(config["num_minibatches"]), int(config["minibatch_size"])
        batch_size = nmb * mbs
        T, n = advantages.shape
        assert batch_size == T * n
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
                rng, _rng = brandom.split(rng)
                perm = (permutation_fn(_rng, batch_size) if permutation_fn is not None
                        else default_permutation(_rng, batch_size, dev)).to(device=dev, dtype=torch.int32).contiguous()
                for mb in range(nmb):
                    index = perm[mb * mbs:(mb + 1) * mbs]
                    ops.gather_rows(obs, index, x_mb)
                    x = x_mb.to(torch.float32)
                    logits, value = _forward_autograd(leaves, x, act)
                    call["index"], call["stats"] = index, stats_all[epoch, mb]
                    loss = _LossHead.apply(logits.contiguous(), value.contiguous(), call)
                    flat_g.zero_()
                    loss.backward()
                    state = optimizer.update_(flat_p, flat_g, state)
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev_tf32
        cols = [stats_all[:, :, i] for i in range(7)]
        return (new_params, state, env_state, last_obs, terminated_count, rng), (cols[0], tuple(cols[1:]))

    return update_step
