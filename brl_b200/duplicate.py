"""Duplicate-table wrappers of src/duplicate.py: `_imp_reward`, `duplicate_init`,
`Table_info`, `duplicate_step` -- same names and results, executed by
brl_duplicate_step / brl_duplicate_init / brl_imp_reward."""
from __future__ import annotations

from typing import NamedTuple

import torch

from . import ops
from .env import BridgeBidding, State

PASS_ACTION_NUM = 0
DOUBLE_ACTION_NUM = 1
REDOUBLE_ACTION_NUM = 2
BID_OFFSET_NUM = 3


def _imp_reward(table_a_reward: torch.Tensor, table_b_reward: torch.Tensor) -> torch.Tensor:
    """src/duplicate.py:15-70, batched: f32[n,4] x f32[n,4] -> f32[n,4]."""
    a = table_a_reward.to(torch.float32).reshape(-1, 4).contiguous()
    b = table_b_reward.to(torch.float32).reshape(-1, 4).contiguous()
    out = torch.empty_like(a)
    ops.imp_reward(a, b, out)
    return out.reshape(table_a_reward.shape)


class Table_info(NamedTuple):
    """src/duplicate.py:138-144"""
    terminated: torch.Tensor
    rewards: torch.Tensor
    last_bid: torch.Tensor
    last_bidder: torch.Tensor
    call_x: torch.Tensor
    call_xx: torch.Tensor

    @classmethod
    def from_state(cls, state: State) -> "Table_info":
        """The snapshot src/evaluation.py:97-112 takes of the init state."""
        return cls(state._terminated_u8.clone(), state.rewards.clone(), state._last_bid.clone().to(torch.int32),
                   state._last_bidder.clone().to(torch.int32), state._f("call_x").clone(), state._f("call_xx").clone())

    def _buffers(self) -> ops.TableInfoBuffers:
        b = ops.TableInfoBuffers.__new__(ops.TableInfoBuffers)
        b.terminated, b.rewards, b.last_bid = self.terminated.view(torch.uint8), self.rewards, self.last_bid
        b.last_bidder, b.call_x, b.call_xx = self.last_bidder, self.call_x.view(torch.uint8), self.call_xx.view(torch.uint8)
        return b


def duplicate_init(state: State) -> State:
    """src/duplicate.py:132-135: table B start -- seats handed to the other team, same deal."""
    env: BridgeBidding = state.env
    packed, out = env._fresh(state.num_envs)
    ops.duplicate_init(state._packed, env.table, packed, out)
    return State(env, packed, out)


def duplicate_step(step_fn):
    """src/duplicate.py:147-192.  Table infos are updated IN PLACE and returned."""
    env = getattr(step_fn, "__self__", None)
    if not isinstance(env, BridgeBidding):
        raise TypeError("duplicate_step(step_fn) expects env.step of a brl_b200.BridgeBidding")

    def wrapped_step(state: State, action: torch.Tensor, table_a_info: Table_info, table_b_info: Table_info,
                     *, inplace: bool = True):
        n = state.num_envs
        packed, out = (state._packed, state.outputs()) if inplace else env._fresh(n)
        ops.duplicate_step(state._packed, action.to(torch.int32), env.table, table_a_info._buffers(),
                           table_b_info._buffers(), packed, out, env.illegal_penalty, env.illegal_bonus)
        return State(env, packed, out), table_a_info, table_b_info

    return wrapped_step
