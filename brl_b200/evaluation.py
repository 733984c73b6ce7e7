"""Evaluation loops of src/evaluation.py: `make_simple_duplicate_evaluate` (69-204, the
eval.py path and the league probe) and `make_simple_evaluate` (11-66)."""
from __future__ import annotations

import torch

from . import dist as bdist
from . import ops
from . import random as brandom
from .duplicate import Table_info, duplicate_step
from .models import load_params, make_forward_pass
from .utils import single_play_step_two_policy_commpetitive_deterministic

_CHECK_EVERY = 8  # stepping a finished env is a zero-reward no-op, so the all-done test needs no per-step sync


def make_simple_evaluate(eval_env, team1_activation, team1_model_type, team2_activation, team2_model_type,
                         team2_model_path, num_eval_envs, team2_params=None):
    """src/evaluation.py:11-66: deterministic quad steps vs a fixed opponent; mean raw score."""
    actor_forward_pass = make_forward_pass(activation=team1_activation, model_type=team1_model_type)
    opp_forward_pass = make_forward_pass(activation=team2_activation, model_type=team2_model_type)
    opp_params = team2_params if team2_params is not None else load_params(team2_model_path, eval_env.device)

    def simple_evaluate(actor_params, rng):
        step_fn = single_play_step_two_policy_commpetitive_deterministic(
            step_fn=eval_env.step, actor_params=actor_params, actor_forward_pass=actor_forward_pass,
            opp_params=opp_params, opp_forward_pass=opp_forward_pass)
        rng_key, sub_key = brandom.split(rng)
        state = eval_env.init(eval_env.make_keys(sub_key, num_eval_envs))
        R = torch.zeros(num_eval_envs, dtype=torch.float32, device=eval_env.device)
        action = torch.empty(num_eval_envs, dtype=torch.int32, device=eval_env.device)
        r_actor = torch.empty_like(R)
        it = 0
        while True:
            actor = state.current_player.clone()
            logits, _ = actor_forward_pass.apply(actor_params, state.observation)
            ops.categorical(logits.contiguous(), state._mask_u8, action, None, sample=False)
            rng_key, _rng = brandom.split(rng_key)
            state = step_fn(state, action, _rng, out_state=state)
            ops.gather_reward(state.rewards, actor, r_actor, 1.0)
            R += r_actor
            it += 1
            if it % _CHECK_EVERY == 0 and bool(state._terminated_u8.all()):
                break
        return R.mean()

    return simple_evaluate


def make_simple_duplicate_evaluate(eval_env, team1_activation, team1_model_type, team2_activation, team2_model_type,
                                   num_eval_envs, env_offset: int = 0):
    """src/evaluation.py:69-204.  `num_eval_envs` is THIS rank's shard; with
    torch.distributed initialised the statistics are all-reduced over ranks (one
    collective per match) so every rank returns the whole-match numbers."""
    team1_forward_pass = make_forward_pass(activation=team1_activation, model_type=team1_model_type)
    team2_forward_pass = make_forward_pass(activation=team2_activation, model_type=team2_model_type)

    def duplicate_evaluate(team1_params, team2_params, rng_key, trace=None):
        step_fn = duplicate_step(eval_env.step)
        rng_key, sub_key = brandom.split(rng_key)
        state = eval_env.init(eval_env.make_keys(sub_key, num_eval_envs, env_offset))      # :93-95
        table_a_info = Table_info.from_state(state)                                        # :97-112
        table_b_info = Table_info.from_state(state)
        dev = eval_env.device
        cum_return = torch.zeros(num_eval_envs, dtype=torch.float32, device=dev)
        a1 = torch.empty(num_eval_envs, dtype=torch.int32, device=dev)
        a2 = torch.empty_like(a1)
        count = 0
        while True:
            # :124-151 -- under vmap both nets run on every env; the team of current_player picks
            l1, _ = team1_forward_pass.apply(team1_params, state.observation)
            l2, _ = team2_forward_pass.apply(team2_params, state.observation)
            ops.categorical(l1.contiguous(), state._mask_u8, a1, None, sample=False)
            ops.categorical(l2.contiguous(), state._mask_u8, a2, None, sample=False)
            action = torch.where(state.current_player < 2, a1, a2)
            if trace is not None:  # tests replay the same action sequence on the oracle
                trace.append((action.clone(), l1.clone(), l2.clone()))
            state, table_a_info, table_b_info = step_fn(state, action, table_a_info, table_b_info)  # :164
            cum_return += state.rewards[:, 0]                                              # :167-169
            count += 1
            if count % _CHECK_EVERY == 0 and bool(state._terminated_u8.all()):
                break
        sums = torch.zeros(8, dtype=torch.float64, device=dev)
        ops.match_stats(cum_return, sums)
        bdist.allreduce_sums(sums)
        mean, std_error, win_rate = bdist.stats_from_sums(sums.cpu())                      # :199-201
        log_info = (mean, std_error, win_rate)
        return log_info, table_a_info, table_b_info, cum_return

    return duplicate_evaluate
