"""Env wrappers of src/utils.py: `auto_reset` and the four-call "quad" steps.

Same names, argument meaning and results as the reference; the difference is that the
env is natively batched and the wrappers map onto kernel flags instead of traced
`lax.cond`s (which, under vmap, execute `env.init` for every env on every sub-step):
  auto_reset               -> brl_step with BRL_F_AUTORESET      (src/utils.py:9-58)
  rewards1+..+4, OR of terminated -> BRL_F_ACCUMULATE / BRL_F_QUAD_LAST (src/utils.py:126-128)
"""
from __future__ import annotations

import torch

from . import ops
from . import random as brandom
from .env import BridgeBidding, State


class _AutoResetStep:
    def __init__(self, env: BridgeBidding):
        self.env = env

    def __call__(self, state: State, action: torch.Tensor) -> State:
        return self.env.step(state, action, autoreset=True)


def auto_reset(step_fn, init_fn):
    """src/utils.py:9-58.  `step_fn`/`init_fn` must be `env.step`/`env.init` of one env."""
    env = getattr(step_fn, "__self__", None)
    if not isinstance(env, BridgeBidding) or getattr(init_fn, "__self__", None) is not env:
        raise TypeError("auto_reset(step_fn, init_fn) expects env.step and env.init of a brl_b200.BridgeBidding")
    return _AutoResetStep(env)


def _env_of(step_fn):
    if isinstance(step_fn, _AutoResetStep):
        return step_fn.env, True
    env = getattr(step_fn, "__self__", None)
    if isinstance(env, BridgeBidding):
        return env, False
    raise TypeError("step_fn must be env.step or auto_reset(env.step, env.init)")


class _QuadStep:
    """Four consecutive env sub-steps (src/utils.py:69-128): the given action, then opp /
    actor / opp nets act on the fresh observation; returns the last state with
    rewards = sum of the four and terminated = OR of the four."""

    def __init__(self, step_fn, actor_forward_pass, actor_params, opp_forward_pass, opp_params, mode: str):
        self.env, self.autoreset = _env_of(step_fn)
        self.actor_fp, self.actor_params = actor_forward_pass, actor_params
        self.opp_fp, self.opp_params = opp_forward_pass, opp_params
        self.mode = mode  # "sample" | "deterministic" | "free-run"
        self._scratch = None
        self._mid_obs = None
        self.trace = None  # tests set a list here to receive the four sub-step action tensors

    def _buffers(self, n, device):
        if self._scratch is None or self._scratch[0].shape[0] != n:
            self._scratch = (torch.empty(n, dtype=torch.int32, device=device),)
        return self._scratch

    def _policy_action(self, fp, params, state: State, seed: int, sample: bool, act_buf):
        fp.act(params, state.observation, state._mask_u8, act_buf, sample=sample, seed=seed, env_offset=self.env.env_offset)
        return act_buf

    def __call__(self, state: State, action: torch.Tensor, rng: int, *, out_state: State = None) -> State:
        env = self.env
        n = state.num_envs
        (act_buf,) = self._buffers(n, env.device)
        # sub-step 1 writes fresh rewards/terminated, sub-steps 2-4 accumulate into them
        packed, out = (out_state._packed, out_state.outputs()) if out_state is not None else env._fresh(n)
        kw = dict(autoreset=self.autoreset, illegal_penalty=env.illegal_penalty, illegal_bonus=env.illegal_bonus)
        # the observations of sub-steps 1-3 are read by a policy net only (the caller sees the 4th): when the nets
        # take bf16 input, write those straight in bf16 into a private buffer (no cast launch, half the bytes)
        mid = out
        if self.mode != "free-run" and getattr(self.actor_fp, "input_dtype", None) == torch.bfloat16 and \
                getattr(self.opp_fp, "input_dtype", None) == torch.bfloat16 and out.observation.dtype != torch.bfloat16:
            if self._mid_obs is None or self._mid_obs.shape[0] != n:
                self._mid_obs = torch.empty((n, ops.OBS_DIM), dtype=torch.bfloat16, device=env.device)
            mid = State(env, packed, out).outputs()
            mid.observation = self._mid_obs
        ops.step(state._packed, action.to(torch.int32), env.table, packed, mid, **kw)
        if self.trace is not None:
            self.trace.append(action.clone())
        cur = State(env, packed, mid)
        plan = ((self.opp_fp, self.opp_params), (self.actor_fp, self.actor_params), (self.opp_fp, self.opp_params))
        for k, (fp, params) in enumerate(plan):
            rng, sub = brandom.split(rng)
            if self.mode == "free-run" and k != 1:
                act_buf.zero_()  # opponents always pass (src/utils.py:217,235)
            else:
                sample = self.mode == "sample"
                self._policy_action(fp, params, cur, sub, sample, act_buf)
            if self.trace is not None:
                self.trace.append(act_buf.clone())
            ops.step(packed, act_buf, env.table, packed, out if k == 2 else mid, accumulate=True, quad_last=(k == 2), **kw)
        return State(env, packed, out)


def single_play_step_two_policy_commpetitive(step_fn, actor_forward_pass, actor_params, opp_forward_pass, opp_params):
    """src/utils.py:61-130 (sampled opponents / partner)"""
    return _QuadStep(step_fn, actor_forward_pass, actor_params, opp_forward_pass, opp_params, "sample")


def single_play_step_two_policy_commpetitive_deterministic(step_fn, actor_forward_pass, actor_params,
                                                           opp_forward_pass, opp_params):
    """src/utils.py:133-202 (argmax opponents / partner)"""
    return _QuadStep(step_fn, actor_forward_pass, actor_params, opp_forward_pass, opp_params, "deterministic")


def single_play_step_free_run(step_fn, actor_forward_pass, actor_params, opp_forward_pass, opp_params):
    """src/utils.py:205-246 (opponents always pass, partner argmax)"""
    return _QuadStep(step_fn, actor_forward_pass, actor_params, opp_forward_pass, opp_params, "free-run")


def normal_step(step_fn):
    """src/utils.py:249-254"""
    def wrapped_step_fn(state, action, rng):
        return step_fn(state, action)
    return wrapped_step_fn
