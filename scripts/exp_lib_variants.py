"""Times the bench-shape rollout for a few launch shapes with whatever library BRL_B200_LIB selects (one process per library
build; scripts/build_variant.py makes them)."""
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from brl_b200 import _lib, ops  # noqa: E402
from brl_b200.deals import synthetic_deal_table  # noqa: E402

dev = "cuda:0"
table = torch.as_tensor(synthetic_deal_table(100000, 0), device=dev)
n, k, reps, rounds = 8192, 32, 20, 7
state, out0 = ops.new_state(n, dev), ops.EnvOutputs(n, dev)
ops.init(ops.make_keys(1, n, dev), table, state, out0)
traj = ops.EnvOutputs(n, dev, rows=k)
cfgs = {"default": 0, "epb32 w4 bal": _lib.tune(epw=32, writers=4, balanced=True), "epb32 w3 plain": _lib.tune(epw=32, writers=3, balanced=False),
        "epb16 w2 bal": _lib.tune(epw=16, writers=2, balanced=True), "epb32 w3 bal": _lib.tune(epw=32, writers=3, balanced=True),
        "epb32 w5 bal": _lib.tune(epw=32, writers=5, balanced=True), "epb16 w3 bal": _lib.tune(epw=16, writers=3, balanced=True)}
res = {name: [] for name in cfgs}
step = 0
for r in range(rounds):
    for name, tune in cfgs.items():
        for _ in range(3):
            ops.rollout_random(state, table, k, traj, seed=1, step0=step, tune=tune); step += k
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            ops.rollout_random(state, table, k, traj, seed=1, step0=step, tune=tune); step += k
        e1.record()
        torch.cuda.synchronize()
        res[name].append(e0.elapsed_time(e1) / reps)
tag = os.path.basename(os.environ.get("BRL_B200_LIB", "default-lib"))
print(tag, " | ".join(f"{name}: {statistics.median(v)*1e3:.2f} us" for name, v in res.items()))
