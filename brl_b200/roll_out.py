"""`make_roll_out` of src/roll_out.py:23-110: T env steps of the learner against an
opponent net through the auto-reset quad step, recording `Transition`s.

Buffers: the trajectory is preallocated [T, N, ...]; the env kernel writes the
observation / mask that the NEXT step records directly into its trajectory slot, so
`Transition.obs` costs no extra copy."""
from __future__ import annotations

from typing import NamedTuple

import torch

from . import ops
from . import random as brandom
from .env import State
from .utils import auto_reset, single_play_step_free_run, single_play_step_two_policy_commpetitive


class Transition(NamedTuple):
    """src/roll_out.py:13-20"""
    done: torch.Tensor
    action: torch.Tensor
    value: torch.Tensor
    reward: torch.Tensor
    log_prob: torch.Tensor
    obs: torch.Tensor
    legal_action_mask: torch.Tensor


def make_roll_out(config, env, actor_forward_pass, opp_forward_pass):
    masked = bool(config["actor_illegal_action_mask"])
    if not masked and not config.get("actor_illegal_action_penalty", False):
        raise ValueError("one of actor_illegal_action_mask / actor_illegal_action_penalty must be set (src/roll_out.py:24-39)")
    if config["game_mode"] == "competitive":
        make_step_fn = single_play_step_two_policy_commpetitive
    elif config["game_mode"] == "free-run":
        make_step_fn = single_play_step_free_run
    else:
        raise ValueError(config["game_mode"])
    T = int(config["num_steps"])
    scale = float(config["reward_scale"])

    def roll_out(runner_state, opp_params, trace=None):
        params, opt_state, env_state, last_obs, terminated_count, rng = runner_state
        step_fn = make_step_fn(step_fn=auto_reset(env.step, env.init), actor_forward_pass=actor_forward_pass,
                               actor_params=params, opp_forward_pass=opp_forward_pass, opp_params=opp_params)
        step_fn.trace = trace
        n, dev = env_state.num_envs, env.device
        traj = Transition(
            done=torch.empty((T, n), dtype=torch.uint8, device=dev), action=torch.empty((T, n), dtype=torch.int32, device=dev),
            value=torch.empty((T, n), dtype=torch.float32, device=dev), reward=torch.empty((T, n), dtype=torch.float32, device=dev),
            log_prob=torch.empty((T, n), dtype=torch.float32, device=dev),
            obs=torch.empty((T, n, ops.OBS_DIM), dtype=last_obs.dtype, device=dev),
            legal_action_mask=torch.empty((T, n, ops.NUM_ACTIONS), dtype=torch.uint8, device=dev))
        traj.obs[0].copy_(last_obs)
        traj.legal_action_mask[0].copy_(env_state._mask_u8)
        # working state whose observation / mask outputs alias the NEXT trajectory slot
        packed = env_state._packed.clone()
        actor = env_state.current_player.clone()
        spare_obs = torch.empty_like(last_obs)
        spare_mask = torch.empty_like(env_state._mask_u8)
        out = ops.EnvOutputs(n, dev, last_obs.dtype)
        cur = State(env, packed, out)
        cur.observation, cur._mask_u8 = traj.obs[0], traj.legal_action_mask[0]
        cur.current_player.copy_(actor)
        for t in range(T):
            rng, _rng = brandom.split(rng)
            actor_forward_pass.act(params, traj.obs[t], traj.legal_action_mask[t] if masked else None, traj.action[t],
                                   traj.log_prob[t], traj.value[t], sample=True, seed=_rng,
                                   env_offset=getattr(env, "env_offset", 0))               # :73-81
            actor.copy_(cur.current_player)
            rng, _rng = brandom.split(rng)
            nxt = State(env, packed, out)
            nxt.observation = traj.obs[t + 1] if t + 1 < T else spare_obs
            nxt._mask_u8 = traj.legal_action_mask[t + 1] if t + 1 < T else spare_mask
            cur = step_fn(cur, traj.action[t], _rng, out_state=nxt)                        # :84
            terminated_count = terminated_count + cur._terminated_u8.sum()                 # :85
            traj.done[t].copy_(cur._terminated_u8)
            ops.gather_reward(cur.rewards, actor, traj.reward[t], scale)                   # :86-94
        runner_state = (params, opt_state, cur, cur.observation, terminated_count, rng)
        return runner_state, traj._replace(done=traj.done.view(torch.bool),
                                           legal_action_mask=traj.legal_action_mask.view(torch.bool))

    return roll_out
