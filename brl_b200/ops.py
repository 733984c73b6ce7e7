"""Torch-tensor front-end of the stream-first C-ABI ops (include/brl_b200.h).

PyTorch is plumbing here: it owns device memory and the current CUDA stream; all
arithmetic happens in the hand-written sm_100a kernels behind `_lib.call`.
Every function enqueues on `torch.cuda.current_stream()` and never synchronises.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from ._lib import (BrlParams, F_ACCUMULATE, F_AUTORESET, F_MLP_BF16, F_OBS_BF16, F_OBS_U8, F_QUAD_LAST,
                   F_RANDOM_ACTION, F_RESULT_I16, F_SAMPLE, F_UNIFORM_U16)

NUM_ACTIONS = 38
OBS_DIM = 480
STATE_PLANES = 5

_OBS_FLAG = {torch.float32: 0, torch.uint8: F_OBS_U8, torch.bool: F_OBS_U8, torch.bfloat16: F_OBS_BF16}


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise _lib.BrlError("brl_b200 ops need CUDA tensors (there is no CPU path)")
    if not t.is_contiguous():
        raise _lib.BrlError("brl_b200 ops need contiguous tensors")
    return t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _call(name: str, buffers, params) -> None:
    # buffers were validated (CUDA, contiguous) while the list was built; only then touch the stream
    _lib.call(name, _stream(), buffers, params)


def _params(n, *, flags=0, env_offset=0, stride=0, seed=0, n_deals=0, step=0, k_steps=0, illegal_penalty=-1.0,
            illegal_bonus=1.0, gamma=0.0, gae_lambda=0.0) -> BrlParams:
    return BrlParams(int(n), int(env_offset), int(stride or n), int(seed) & 0xFFFFFFFFFFFFFFFF, int(n_deals), int(flags),
                     int(step) & 0xFFFFFFFF, int(k_steps), float(illegal_penalty), float(illegal_bonus), float(gamma),
                     float(gae_lambda))


def obs_flag(dtype: torch.dtype) -> int:
    try:
        return _OBS_FLAG[dtype]
    except KeyError:
        raise _lib.BrlError(f"observation dtype {dtype} not supported (float32, uint8/bool, bfloat16)") from None


def new_state(n: int, device) -> torch.Tensor:
    """Packed env state: uint4[5][n] plane-major, seen from torch as int32[5, n, 4]."""
    return torch.empty((STATE_PLANES, n, 4), dtype=torch.int32, device=device)


class EnvOutputs:
    """The Env surface brl consumes (src/roll_out.py:72-100): caller-owned output buffers."""

    __slots__ = ("observation", "legal_action_mask", "rewards", "terminated", "current_player")

    def __init__(self, n: int, device, obs_dtype=torch.float32, rows: int = 0):
        lead = (rows, n) if rows else (n,)
        odt = torch.uint8 if obs_dtype == torch.bool else obs_dtype
        self.observation = torch.empty(lead + (OBS_DIM,), dtype=odt, device=device)
        self.legal_action_mask = torch.empty(lead + (NUM_ACTIONS,), dtype=torch.uint8, device=device)
        self.rewards = torch.empty(lead + (4,), dtype=torch.float32, device=device)
        self.terminated = torch.empty(lead, dtype=torch.uint8, device=device)
        self.current_player = torch.empty(lead, dtype=torch.int8, device=device)

    def ptrs(self):
        return [_ptr(self.observation), _ptr(self.legal_action_mask), _ptr(self.rewards), _ptr(self.terminated),
                _ptr(self.current_player)]

    def flag(self) -> int:
        return obs_flag(self.observation.dtype)


def make_keys(seed: int, n: int, device, env_offset: int = 0) -> torch.Tensor:
    keys = torch.empty(n, dtype=torch.int64, device=device)
    _call("brl_make_keys", [_ptr(keys)], _params(n, seed=seed, env_offset=env_offset))
    return keys


def init(keys: torch.Tensor, table: torch.Tensor, state: torch.Tensor, out: EnvOutputs, tune: int = 0) -> None:
    n = keys.shape[0]
    _call("brl_init", [_ptr(keys), _ptr(table), _ptr(state), *out.ptrs()],
              _params(n, n_deals=table.shape[0], stride=state.shape[1], flags=out.flag() | tune))


def reset_fields(deal, dealer, vul_ns, vul_ew, players, rng_key, table, state, out: EnvOutputs) -> None:
    n = deal.shape[0]
    _call("brl_reset_fields", [_ptr(deal), _ptr(dealer), _ptr(vul_ns), _ptr(vul_ew), _ptr(players), _ptr(rng_key), _ptr(table),
               _ptr(state), *out.ptrs()],
              _params(n, n_deals=table.shape[0], stride=state.shape[1], flags=out.flag()))


def step(state_in, action, table, state_out, out: EnvOutputs, *, autoreset=False, random_action=False,
         accumulate=False, quad_last=False, seed=0, step_index=0, env_offset=0, action_out=None,
         illegal_penalty=-1.0, illegal_bonus=1.0, tune: int = 0) -> None:
    n = state_in.shape[1]
    flags = out.flag() | tune
    flags |= F_AUTORESET if autoreset else 0
    flags |= F_RANDOM_ACTION if random_action else 0
    flags |= F_ACCUMULATE if accumulate else 0
    flags |= F_QUAD_LAST if quad_last else 0
    _call("brl_step", [_ptr(state_in), _ptr(action), _ptr(table), _ptr(state_out), *out.ptrs(), _ptr(action_out)],
              _params(n, flags=flags, n_deals=table.shape[0], stride=state_in.shape[1], seed=seed, step=step_index,
                      env_offset=env_offset, illegal_penalty=illegal_penalty, illegal_bonus=illegal_bonus))


class TableInfoBuffers:
    """src/duplicate.py:138-144 Table_info as SoA device buffers."""

    __slots__ = ("terminated", "rewards", "last_bid", "last_bidder", "call_x", "call_xx")

    def __init__(self, n: int, device):
        self.terminated = torch.zeros(n, dtype=torch.uint8, device=device)
        self.rewards = torch.zeros((n, 4), dtype=torch.float32, device=device)
        self.last_bid = torch.full((n,), -1, dtype=torch.int32, device=device)
        self.last_bidder = torch.full((n,), -1, dtype=torch.int32, device=device)
        self.call_x = torch.zeros(n, dtype=torch.uint8, device=device)
        self.call_xx = torch.zeros(n, dtype=torch.uint8, device=device)

    def ptrs(self):
        return [_ptr(getattr(self, k)) for k in self.__slots__]


def duplicate_step(state_in, action, table, info_a: TableInfoBuffers, info_b: TableInfoBuffers, state_out,
                   out: EnvOutputs, illegal_penalty=-1.0, illegal_bonus=1.0) -> None:
    n = state_in.shape[1]
    _call("brl_duplicate_step", [_ptr(state_in), _ptr(action), _ptr(table), *info_a.ptrs(), *info_b.ptrs(), _ptr(state_out), *out.ptrs()],
              _params(n, flags=out.flag(), n_deals=table.shape[0], stride=state_in.shape[1],
                      illegal_penalty=illegal_penalty, illegal_bonus=illegal_bonus))


def duplicate_init(state_in, table, state_out, out: EnvOutputs) -> None:
    n = state_in.shape[1]
    _call("brl_duplicate_init", [_ptr(state_in), _ptr(table), _ptr(state_out), *out.ptrs()],
              _params(n, flags=out.flag(), n_deals=table.shape[0], stride=state_in.shape[1]))


def observe(state, table, obs: torch.Tensor, player_id: Optional[torch.Tensor] = None, tune: int = 0) -> None:
    n = state.shape[1]
    _call("brl_observe", [_ptr(state), _ptr(player_id), _ptr(table), _ptr(obs)],
              _params(n, flags=obs_flag(obs.dtype) | tune, n_deals=table.shape[0], stride=state.shape[1]))


def legal_mask(state, mask: torch.Tensor, tune: int = 0) -> None:
    n = state.shape[1]
    _call("brl_legal_mask", [_ptr(state), _ptr(mask)], _params(n, stride=state.shape[1], flags=tune))


def rollout_random(state, table, k_steps: int, out: Optional[EnvOutputs], *, seed=0, step0=0, env_offset=0,
                   action_out=None, stats=None, obs_only: Optional[torch.Tensor] = None, tune: int = 0,
                   uniforms: Optional[torch.Tensor] = None, result16: Optional[torch.Tensor] = None) -> None:
    """K auto-reset random-legal steps in ONE launch; `out` holds [K, n, ...] trajectories.
    `uniforms`: caller-owned randomness, int32/uint32 [K, n], or int16/uint16 [K, n] (stands for u16 << 16);
    `result16`: int16 [K, n] compact result = 2 * rewards[player 0] + terminated (decode_result16)."""
    n = state.shape[1]
    extra = 0
    if uniforms is not None and uniforms.element_size() == 2:
        extra |= F_UNIFORM_U16
    if result16 is not None:
        if result16.dtype != torch.int16:
            raise TypeError("result16 must be an int16 tensor")
        extra |= F_RESULT_I16
    if out is not None:
        ptrs, flag = out.ptrs(), out.flag()
    else:
        ptrs = [_ptr(obs_only), None, None, None, None]
        flag = obs_flag(obs_only.dtype) if obs_only is not None else 0
    _call("brl_rollout_random", [_ptr(state), _ptr(table), *ptrs, _ptr(action_out), _ptr(stats), _ptr(uniforms),
                                 _ptr(result16)],
              _params(n, flags=flag | tune | extra, n_deals=table.shape[0], stride=state.shape[1], seed=seed, step=step0,
                      env_offset=env_offset, k_steps=k_steps))


def decode_result16(result16: torch.Tensor):
    """compact rollout result -> (rewards f32[..., 4], terminated u8[...]) exactly as the Env surface gives them"""
    w = result16.to(torch.int32)
    s = (w >> 1).to(torch.float32)
    return torch.stack((s, s, -s, -s), dim=-1), (w & 1).to(torch.uint8)


def imp_reward(a_rewards, b_rewards, out) -> None:
    _call("brl_imp_reward", [_ptr(a_rewards), _ptr(b_rewards), _ptr(out)], _params(a_rewards.shape[0]))


def gae(done, value, reward, last_val, adv, targets, gamma: float, gae_lambda: float) -> None:
    t, n = done.shape
    _call("brl_gae", [_ptr(done), _ptr(value), _ptr(reward), _ptr(last_val), _ptr(adv), _ptr(targets)],
              _params(n, k_steps=t, gamma=gamma, gae_lambda=gae_lambda))


def categorical(logits, mask, action, log_prob, *, sample=False, seed=0, env_offset=0, step_index=0) -> None:
    n = logits.shape[0]
    _call("brl_categorical", [_ptr(logits), _ptr(mask), _ptr(action), _ptr(log_prob)],
              _params(n, flags=F_SAMPLE if sample else 0, seed=seed, env_offset=env_offset, step=step_index))


def match_stats(x: torch.Tensor, sums: torch.Tensor) -> None:
    _call("brl_match_stats", [_ptr(x), _ptr(sums)], _params(x.shape[0]))


def gather_reward(rewards, actor, out, scale: float, done: Optional[torch.Tensor] = None,
                  count: Optional[torch.Tensor] = None) -> None:
    """reward = rewards[actor] / scale (src/roll_out.py:86-94); with `done` (u8) and `count` (int64[1]): count += sum(done)
    in the same launch (src/roll_out.py:85)."""
    counting = done is not None and count is not None
    _call("brl_gather_reward", [_ptr(rewards), _ptr(actor), _ptr(out), _ptr(done) if counting else None,
                                _ptr(count) if counting else None],
          _params(rewards.shape[0], gamma=scale, flags=_lib.F_COUNT_DONE if counting else 0))


def mlp_pack(weights, biases, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Six haiku-layout fp32 (w[in,out], b[out]) pairs -> the packed bf16 hi/lo parameter blob (`out`: re-pack in place)."""
    dev = weights[0].device
    blob = out if out is not None else torch.empty(_lib.load().brl_mlp_packed_bytes(), dtype=torch.uint8, device=dev)
    ws = [w.detach().to(torch.float32).contiguous() for w in weights]
    bs = [b.detach().to(torch.float32).contiguous() for b in biases]
    _call("brl_mlp_pack", [_ptr(t) for t in ws] + [_ptr(t) for t in bs] + [_ptr(blob)], _params(0))
    return blob


def obs_to_bf16(obs: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """0/1 observation in a reference dtype (f32 / u8 / bool) -> bf16 for the tensor-core forward."""
    if obs.dtype == torch.bfloat16:
        return obs
    n = obs.shape[0]
    if out is None:
        out = torch.empty((n, OBS_DIM), dtype=torch.bfloat16, device=obs.device)
    src = obs.view(torch.uint8) if obs.dtype == torch.bool else obs
    if src.dtype not in (torch.float32, torch.uint8):
        raise _lib.BrlError(f"observation dtype {obs.dtype} not supported")
    _call("brl_obs_to_bf16", [_ptr(src), _ptr(out)], _params(n, flags=F_OBS_U8 if src.dtype == torch.uint8 else 0))
    return out


def mlp_scratch(n: int, device) -> torch.Tensor:
    return torch.empty(_lib.load().brl_mlp_scratch_bytes(n), dtype=torch.uint8, device=device)


def mlp_forward(obs_bf16: torch.Tensor, packed: torch.Tensor, scratch: torch.Tensor, logits: torch.Tensor,
                value: torch.Tensor, single_bf16: bool = False, tune: int = 0) -> None:
    """logits f32[n,38], value f32[n] = DeepMind MLP(obs) on tcgen05 (3-term bf16 split unless single_bf16)."""
    n = obs_bf16.shape[0]
    if obs_bf16.dtype != torch.bfloat16:
        raise _lib.BrlError("mlp_forward needs a bf16 observation (ops.obs_to_bf16)")
    if scratch.numel() < _lib.load().brl_mlp_scratch_bytes(n):
        raise _lib.BrlError("mlp_forward: scratch too small (ops.mlp_scratch)")
    _call("brl_mlp_forward", [_ptr(obs_bf16), _ptr(packed), _ptr(scratch), _ptr(logits), _ptr(value)],
          _params(n, flags=(F_MLP_BF16 if single_bf16 else 0) | tune))


def policy_act(obs_bf16: torch.Tensor, packed: torch.Tensor, scratch: torch.Tensor, mask: Optional[torch.Tensor],
               action: torch.Tensor, log_prob: Optional[torch.Tensor] = None, value: Optional[torch.Tensor] = None,
               logits: Optional[torch.Tensor] = None, *, sample: bool = False, seed: int = 0, env_offset: int = 0,
               step_index: int = 0, single_bf16: bool = False, tune: int = 0, seed_salt: Optional[torch.Tensor] = None) -> None:
    """forward.apply + masked Categorical sample / mode (+ log_prob, value, logits on request) in one call
    (src/roll_out.py:73-81); same noise stream as `categorical`, so it equals mlp_forward followed by categorical."""
    n = obs_bf16.shape[0]
    if obs_bf16.dtype != torch.bfloat16:
        raise _lib.BrlError("policy_act needs a bf16 observation (ops.obs_to_bf16)")
    if scratch.numel() < _lib.load().brl_mlp_scratch_bytes(n):
        raise _lib.BrlError("policy_act: scratch too small (ops.mlp_scratch)")
    _call("brl_policy_act", [_ptr(obs_bf16), _ptr(packed), _ptr(scratch), _ptr(mask), _ptr(action), _ptr(log_prob), _ptr(value),
                             _ptr(logits), _ptr(seed_salt)],
          _params(n, flags=(F_MLP_BF16 if single_bf16 else 0) | (F_SAMPLE if sample else 0) | tune |
                  (_lib.F_SEED_SALT if seed_salt is not None else 0), seed=seed,
                  env_offset=env_offset, step=step_index))


def mlp_rows_scratch(n_rows: int, device) -> torch.Tensor:
    return torch.empty(_lib.load().brl_mlp_rows_scratch_bytes(n_rows), dtype=torch.uint8, device=device)


def policy_act_rows(obs_bf16: torch.Tensor, packed: torch.Tensor, scratch: torch.Tensor, mask: Optional[torch.Tensor],
                    action: torch.Tensor, rows: torch.Tensor, log_prob: Optional[torch.Tensor] = None,
                    logits: Optional[torch.Tensor] = None, *, sample: bool = False,
                    seed: int = 0, env_offset: int = 0, step_index: int = 0, single_bf16: bool = False, tune: int = 0) -> None:
    """`policy_act` restricted to the envs listed in `rows` (int32): only their observations are read and only their
    action / log_prob entries written (src/evaluation.py:124-151 without forwarding envs a net does not decide)."""
    n = rows.shape[0]
    if n == 0:
        return
    if obs_bf16.dtype != torch.bfloat16 or rows.dtype != torch.int32:
        raise _lib.BrlError("policy_act_rows needs a bf16 observation and int32 rows")
    if scratch.numel() < _lib.load().brl_mlp_rows_scratch_bytes(n):
        raise _lib.BrlError("policy_act_rows: scratch too small (ops.mlp_rows_scratch)")
    _call("brl_policy_act_rows", [_ptr(obs_bf16), _ptr(packed), _ptr(scratch), _ptr(mask), _ptr(action), _ptr(rows), _ptr(log_prob),
                                  _ptr(logits)],
          _params(n, flags=(F_MLP_BF16 if single_bf16 else 0) | (F_SAMPLE if sample else 0) | tune, seed=seed,
                  env_offset=env_offset, step=step_index))


def team_rows(current_player: torch.Tensor, done: Optional[torch.Tensor], rows1: torch.Tensor, rows2: torch.Tensor,
              counts: torch.Tensor) -> None:
    """rows of the live envs (done == 0) split by acting team: players 0/1 -> rows1, players 2/3 -> rows2; counts i32[2]."""
    _call("brl_team_rows", [_ptr(current_player), _ptr(done), _ptr(rows1), _ptr(rows2), _ptr(counts)],
          _params(current_player.shape[0]))


def ppo_scratch(device) -> torch.Tensor:
    """scratch of ppo_loss / ppo_grad as a float64 tensor (element 14 = sum of squares of the gradient after ppo_grad)"""
    return torch.zeros(_lib.PPO_SCRATCH_BYTES // 8, dtype=torch.float64, device=device)


def ppo_loss(logits, value, index, mask, action, old_log_prob, old_value, adv, targets, dlogits, dvalue, stats, scratch,
             *, clip_eps, ent_coef, vf_coef, illegal_l2_coef=0.0, value_clipping=True, reward_scaling=False,
             masked_policy=True) -> None:
    """_loss_fn of src/update.py:91-162 + its gradient w.r.t. (logits, value) for one minibatch."""
    flags = (_lib.PPO_VALUE_CLIPPING if value_clipping else 0) | (_lib.PPO_REWARD_SCALING if reward_scaling else 0) | \
            (0 if masked_policy else _lib.PPO_UNMASKED_POLICY)
    p = _lib.BrlPpoParams(int(logits.shape[0]), int(action.numel()), float(clip_eps), float(ent_coef), float(vf_coef),
                          float(illegal_l2_coef), flags, 0)
    _call("brl_ppo_loss", [_ptr(logits), _ptr(value), _ptr(index), _ptr(mask), _ptr(action), _ptr(old_log_prob),
                           _ptr(old_value), _ptr(adv), _ptr(targets), _ptr(dlogits), _ptr(dvalue), _ptr(stats),
                           _ptr(scratch)], p)


def adam_clip(params, grads, m, v, scratch, *, step: int, lr: float, beta1=0.9, beta2=0.999, eps=1e-5,
              max_grad_norm=0.0, sumsq: Optional[torch.Tensor] = None) -> None:
    """optax.chain(clip_by_global_norm, adam) on flat fp32 buffers, in place (ppo.py:195-211).  `sumsq` (f64[1], e.g.
    `acc[14:15]` after `ppo_grad`): the gradient's sum of squares is already known, skip the pass that forms it."""
    p = _lib.BrlAdamParams(int(params.numel()), int(step), float(lr), float(beta1), float(beta2), float(eps),
                           float(max_grad_norm))
    if sumsq is not None:
        if sumsq.dtype != torch.float64:
            raise _lib.BrlError("adam_clip: sumsq must be float64")
        _call("brl_adam_apply", [_ptr(params), _ptr(grads), _ptr(m), _ptr(v), _ptr(sumsq)], p)
    else:
        _call("brl_adam_clip", [_ptr(params), _ptr(grads), _ptr(m), _ptr(v), _ptr(scratch)], p)


def mlp_num_params() -> int:
    return int(_lib.load().brl_mlp_num_params())


def mlp_pack_train(flat_params: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Flat fp32 parameters (optim.flatten_params order) -> the training blob of `ppo_grad`: W[in, out] as bf16 hi / lo in
    haiku's orientation + fp32 biases.  Once per parameter set; `mlp_adam_step` keeps it current inside the update loop."""
    L = _lib.load()
    if flat_params.dtype != torch.float32 or flat_params.numel() != L.brl_mlp_num_params():
        raise _lib.BrlError("mlp_pack_train needs the flat fp32 parameter buffer of the DeepMind MLP")
    if out is None:
        out = torch.empty(L.brl_mlp_train_blob_bytes(), dtype=torch.uint8, device=flat_params.device)
    _call("brl_mlp_pack_train", [_ptr(flat_params), _ptr(out)], _params(0))
    return out


def mlp_adam_step(params, grads, m, v, sumsq, blob, *, step: int, lr: float, beta1=0.9, beta2=0.999, eps=1e-5,
                  max_grad_norm=0.0) -> None:
    """optax.chain(clip_by_global_norm, adam) on the flat MLP parameters, in place, and the training blob refreshed in
    the same pass (ppo.py:195-211, src/update.py:168-169).  `sumsq`: f64[1] = `acc[14:15]` after `ppo_grad`."""
    if sumsq.dtype != torch.float64:
        raise _lib.BrlError("mlp_adam_step: sumsq must be float64")
    p = _lib.BrlAdamParams(int(params.numel()), int(step), float(lr), float(beta1), float(beta2), float(eps),
                           float(max_grad_norm))
    _call("brl_mlp_adam_step", [_ptr(params), _ptr(grads), _ptr(m), _ptr(v), _ptr(sumsq), _ptr(blob)], p)


def mlp_train_scratch(batch: int, device) -> torch.Tensor:
    return torch.empty(_lib.load().brl_mlp_train_scratch_bytes(batch), dtype=torch.uint8, device=device)


def ppo_grad(obs, blob, scratch, index, mask, action, old_log_prob, old_value, adv, targets, grads, stats, acc, *, clip_eps,
             ent_coef, vf_coef, illegal_l2_coef=0.0, value_clipping=True, reward_scaling=False, masked_policy=True,
             illegal_stat: bool = True, tune: int = 0) -> None:
    """One minibatch of jax.value_and_grad(_loss_fn) (src/update.py:91-167): take(index), forward, loss, backward
    through the MLP on tcgen05 -> flat fp32 gradients (+ the loss statistics of `ppo_loss`).  `illegal_stat=False` (only
    honoured with illegal_l2_coef == 0): skip the logged illegal_action_loss (stats[6] = NaN), ~19 us per call."""
    L = _lib.load()
    B = int(index.shape[0]) if index is not None else int(obs.shape[0])
    if obs.dtype == torch.bool:
        obs = obs.view(torch.uint8)
    try:
        oflag = {torch.float32: 0, torch.uint8: _lib.PPO_OBS_U8, torch.bfloat16: _lib.PPO_OBS_BF16}[obs.dtype]
    except KeyError:
        raise _lib.BrlError(f"ppo_grad: observation dtype {obs.dtype} not supported") from None
    if scratch.numel() < L.brl_mlp_train_scratch_bytes(B) or blob.numel() < L.brl_mlp_train_blob_bytes():
        raise _lib.BrlError("ppo_grad: scratch / blob too small (ops.mlp_train_scratch, ops.mlp_pack_train)")
    if grads.dtype != torch.float32 or grads.numel() != L.brl_mlp_num_params():
        raise _lib.BrlError("ppo_grad: grads must be the flat fp32 gradient buffer")
    flags = (_lib.PPO_VALUE_CLIPPING if value_clipping else 0) | (_lib.PPO_REWARD_SCALING if reward_scaling else 0) | \
            (0 if masked_policy else _lib.PPO_UNMASKED_POLICY) | (_lib.PPO_ILLEGAL_STAT if illegal_stat else 0) | oflag
    p = _lib.BrlPpoParams(B, int(action.numel()), float(clip_eps), float(ent_coef), float(vf_coef), float(illegal_l2_coef),
                          flags, int(tune))
    _call("brl_ppo_grad", [_ptr(obs), _ptr(blob), _ptr(scratch), _ptr(index), _ptr(mask), _ptr(action), _ptr(old_log_prob),
                           _ptr(old_value), _ptr(adv), _ptr(targets), _ptr(grads), _ptr(stats), _ptr(acc)], p)


def gather_rows(src: torch.Tensor, index: torch.Tensor, dst: torch.Tensor) -> None:
    """dst[b] = src[index[b]] over the leading axis (minibatch take, src/update.py:194-199)."""
    row_bytes = src[0].numel() * src.element_size()
    _call("brl_gather_rows", [_ptr(src), _ptr(index), _ptr(dst)], _params(index.shape[0], k_steps=row_bytes))


def eval_act_log(logits_team1, logits_team2, mask, current_player, terminated, action, acc, indicator_bids=False) -> None:
    """masked argmax of the acting team's logits + update_log_info (src/evaluation.py:236-385, 650-745)."""
    n = logits_team1.shape[0]
    _call("brl_eval_act_log", [_ptr(logits_team1), _ptr(logits_team2), _ptr(mask), _ptr(current_player), _ptr(terminated),
                               _ptr(action), _ptr(acc)],
          _params(n, flags=_lib.F_EVAL_INDICATOR_BIDS if indicator_bids else 0))


def eval_summary(acc, cum_return, step_count, table_a, table_b, sums) -> None:
    """table_* = (last_bid i32, last_bidder i32, call_x u8, call_xx u8, rewards f32[n,4] | None, pass_num i32 | None);
    table_b may be None (single-table evaluate).  sums f64[brl_eval_num_sums()] += partial sums."""
    tb = table_b if table_b is not None else (None,) * 6
    _call("brl_eval_summary", [_ptr(acc), _ptr(cum_return), _ptr(step_count)] + [_ptr(t) for t in table_a] +
          [_ptr(t) for t in tb] + [_ptr(sums)], _params(cum_return.shape[0]))


_FIELD_SPECS = (("deal", torch.int32, ()), ("dealer", torch.int32, ()), ("shuffled_players", torch.int8, (4,)),
                ("vul", torch.uint8, (2,)), ("last_bid", torch.int32, ()), ("last_bidder", torch.int32, ()),
                ("call_x", torch.uint8, ()), ("call_xx", torch.uint8, ()), ("pass_num", torch.int32, ()),
                ("step_count", torch.int32, ()), ("rng_key", torch.int64, ()))


def state_fields(state: torch.Tensor) -> dict:
    """Unpack the private pgx-style fields brl reads (src/evaluation.py:97-112, 465-495)."""
    n = state.shape[1]
    out = {name: torch.empty((n,) + shape, dtype=dt, device=state.device) for name, dt, shape in _FIELD_SPECS}
    _call("brl_state_fields", [_ptr(state)] + [_ptr(out[name]) for name, _, _ in _FIELD_SPECS],
              _params(n, stride=state.shape[1]))
    return out
