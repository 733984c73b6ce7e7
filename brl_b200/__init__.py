"""brl_b200 -- B200-native bridge-bidding environment + rollout path of harukaki/brl.

Host-side mirror of the reference's interface for this path (same names / argument
meaning): `BridgeBidding`, `State`, `auto_reset`, quad steps, `duplicate_step`,
`Table_info`, `make_roll_out`, `make_calc_gae`, `make_simple_duplicate_evaluate`,
`make_simple_evaluate`, `make_forward_pass`.  All arithmetic of the path runs in the
hand-written sm_100a kernels behind the C ABI of include/brl_b200.h; there is no CPU
fallback (importing works without a GPU, calling an op does not).
"""
from .env import BridgeBidding, State, act_randomly  # noqa: F401

__all__ = ["BridgeBidding", "State", "act_randomly"]
