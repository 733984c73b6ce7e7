"""Where bench.py's 94 us per rollout launch goes beyond the 90 us of the bare kernel: per-launch event records,
the action trajectory, the episode statistics."""
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from brl_b200 import ops  # noqa: E402
from brl_b200.deals import synthetic_deal_table  # noqa: E402

dev = "cuda:0"
table = torch.as_tensor(synthetic_deal_table(100000, 0), device=dev)
n, k, steps = 8192, 32, 50
state, out0 = ops.new_state(n, dev), ops.EnvOutputs(n, dev)
ops.init(ops.make_keys(1, n, dev), table, state, out0)
traj = ops.EnvOutputs(n, dev, rows=k)
actions = torch.empty((k, n), dtype=torch.int32, device=dev)
stats = torch.zeros(4, dtype=torch.int64, device=dev)
step = [0]


def run(per_launch_events, with_actions, with_stats):
    for _ in range(5):
        ops.rollout_random(state, table, k, traj, seed=1, step0=step[0], action_out=actions if with_actions else None,
                           stats=stats if with_stats else None)
        step[0] += k
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    evs[0].record()
    for i in range(steps):
        ops.rollout_random(state, table, k, traj, seed=1, step0=step[0], action_out=actions if with_actions else None,
                           stats=stats if with_stats else None)
        step[0] += k
        if per_launch_events:
            evs[i + 1].record()
    if not per_launch_events:
        evs[steps].record()
    torch.cuda.synchronize()
    return evs[0].elapsed_time(evs[steps]) / steps * 1e3


cfgs = [(ev, a, s) for ev in (True, False) for a in (True, False) for s in (True, False)]
res = {c: [] for c in cfgs}
for r in range(7):
    for c in cfgs:
        res[c].append(run(*c))
for c, v in sorted(res.items(), key=lambda kv: statistics.median(kv[1])):
    print(f"per-launch events={c[0]!s:5}  action_out={c[1]!s:5}  stats={c[2]!s:5}  median {statistics.median(v):7.2f} us  min {min(v):7.2f}")
