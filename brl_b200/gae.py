"""`make_calc_gae` of src/gae.py:5-47 over the brl_gae reverse-scan kernel."""
from __future__ import annotations

import torch

from . import ops


def make_calc_gae(config, actor_forward_pass):
    def calc_gae(runner_state, traj_batch):
        params, opt_state, env_state, last_obs, terminated_count, rng = runner_state
        _, last_val = actor_forward_pass.apply(params, last_obs)          # src/gae.py:16-18
        done = traj_batch.done.view(torch.uint8).contiguous()
        value = traj_batch.value.contiguous()
        reward = traj_batch.reward.contiguous()
        adv = torch.empty_like(value)
        targets = torch.empty_like(value)
        ops.gae(done, value, reward, last_val.contiguous(), adv, targets, config["gamma"], config["gae_lambda"])
        return adv, targets                                                # src/gae.py:40-41

    return calc_gae
