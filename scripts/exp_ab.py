"""Interleaved A/B timing of rollout-kernel launch shapes (median of rounds; run-to-run noise on the
shared boxes is ~10 %, so configurations are alternated inside one process instead of run back to back)."""
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from brl_b200 import _lib, ops  # noqa: E402
from brl_b200.deals import synthetic_deal_table  # noqa: E402

dev = "cuda:0"
table = torch.as_tensor(synthetic_deal_table(100000, 0), device=dev)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    k, reps, rounds = 32, 20, 7
    state, out0 = ops.new_state(n, dev), ops.EnvOutputs(n, dev)
    ops.init(ops.make_keys(1, n, dev), table, state, out0)
    traj = ops.EnvOutputs(n, dev, rows=k)
    cfgs = {}
    for epb in (32, 16):
        for w in ((2, 3, 4, 5) if epb == 32 else (1, 2, 3, 4)):
            for bal in (False, True):
                cfgs[f"epb{epb} w{w} {'bal' if bal else 'plain'}"] = _lib.tune(epw=epb, writers=w, balanced=bal)
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
    for name, v in sorted(res.items(), key=lambda kv: statistics.median(kv[1])):
        med = statistics.median(v)
        print(f"{name:18s} n={n} median={med*1e3:7.2f} us  min={min(v)*1e3:7.2f}  max={max(v)*1e3:7.2f}  "
              f"frac(median)={1980*n*k/med/1e6/6537.3:.3f}")


if __name__ == "__main__":
    main()
