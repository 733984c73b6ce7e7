"""Kernel-only timing of the rollout with caller-supplied uniforms, u32 against u16, with and without the i16 result
(device-resident buffers; interleaved, median over rounds)."""
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from brl_b200 import ops  # noqa: E402
from brl_b200.deals import synthetic_deal_table  # noqa: E402

dev = "cuda:0"
table = torch.as_tensor(synthetic_deal_table(100000, 0), device=dev)
n, k, reps, rounds = 8192, 32, 20, 9
state, out0 = ops.new_state(n, dev), ops.EnvOutputs(n, dev)
ops.init(ops.make_keys(1, n, dev), table, state, out0)
traj = ops.EnvOutputs(n, dev, rows=k)
import numpy as np  # noqa: E402
_u = np.random.default_rng(0).integers(0, 2 ** 32, size=(k, n), dtype=np.uint32)  # the same draws at both widths
u32 = torch.from_numpy(_u.view(np.int32)).to(dev)
u16 = torch.from_numpy((_u >> 16).astype(np.uint16).view(np.int16)).to(dev)
res16 = torch.empty((k, n), dtype=torch.int16, device=dev)
stats = torch.zeros(4, dtype=torch.int64, device=dev)
cfgs = {"philox": {}, "u32": dict(uniforms=u32), "u16": dict(uniforms=u16), "u32 + result16 + stats": dict(uniforms=u32, result16=res16, stats=stats),
        "u16 + result16 + stats": dict(uniforms=u16, result16=res16, stats=stats)}
res = {name: [] for name in cfgs}
step = 0
for r in range(rounds):
    for name, kw in cfgs.items():
        for _ in range(3):
            ops.rollout_random(state, table, k, traj, seed=1, step0=step, **kw); step += k
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            ops.rollout_random(state, table, k, traj, seed=1, step0=step, **kw); step += k
        e1.record()
        torch.cuda.synchronize()
        res[name].append(e0.elapsed_time(e1) / reps)
print(" | ".join(f"{name}: {statistics.median(v)*1e3:.2f} us" for name, v in res.items()))
