"""Interleaved A/B of the rollout kernel's observation path: per-warp shared staging + one cp.async.bulk per writer warp and
step (default) against direct 128-bit stores, by launch shape and observation dtype.  Median over rounds."""
import json
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from brl_b200 import _lib, ops  # noqa: E402
from brl_b200.deals import synthetic_deal_table  # noqa: E402

dev = "cuda:0"
table = torch.as_tensor(synthetic_deal_table(100000, 0), device=dev)
peak = 6443.2
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def rollout(n, k, rounds=7):
    state, out0 = ops.new_state(n, dev), ops.EnvOutputs(n, dev)
    ops.init(ops.make_keys(1, n, dev), table, state, out0)
    cfgs = {}
    for dt, row in ((torch.float32, 1920), (torch.bfloat16, 960), (torch.uint8, 480)):
        traj = ops.EnvOutputs(n, dev, rows=k, obs_dtype=dt)
        for direct in (False, True):
            shapes = ((32, 2), (32, 3), (32, 4), (32, 5), (32, 6), (16, 2), (16, 3)) if dt == torch.float32 else ((0, 0),)
            for epb, w in shapes:
                cfgs[f"{str(dt)[6:]:8s} epb{epb} w{w} {'direct' if direct else 'bulk  '}"] = \
                    (traj, _lib.tune(epw=epb, writers=w, direct_stores=direct), row + 60)
            if dt == torch.float32:  # floors: no outputs at all (env warp + Philox only), and observation rows only
                for w in (1, 4):
                    cfgs[f"env only (no outputs) w{w} {'direct' if direct else 'bulk  '}"] = (None, _lib.tune(epw=32, writers=w, direct_stores=direct), 1980)
                cfgs[f"obs rows only epb32 w4 {'direct' if direct else 'bulk  '}"] = (traj.observation, _lib.tune(epw=32, writers=4, direct_stores=direct), 1920)
    res = {name: [] for name in cfgs}
    step = [0]
    for r in range(rounds):
        for name, (traj, tune, _) in cfgs.items():
            def go():
                if traj is None or isinstance(traj, torch.Tensor):
                    ops.rollout_random(state, table, k, None, obs_only=traj, seed=1, step0=step[0], tune=tune); step[0] += k
                    return
                ops.rollout_random(state, table, k, traj, seed=1, step0=step[0], tune=tune); step[0] += k
            res[name].append(timed(go, 20))
    for name, v in sorted(res.items(), key=lambda kv: statistics.median(kv[1])):
        med = statistics.median(v)
        b = cfgs[name][2]
        print(f"{name:34s} n={n} median={med*1e3:8.2f} us  min={min(v)*1e3:8.2f}  GB/s={b*n*k/med/1e6:7.0f}  frac={b*n*k/med/1e6/peak:.3f}")


if __name__ == "__main__":
    rollout(8192, 32)
    rollout(65536, 8, rounds=5)
