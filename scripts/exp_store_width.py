"""Interleaved A/B of 256-bit (st.global.v8, default when the buffer is 32-byte aligned) against 128-bit observation-row stores:
the warp-specialised rollout at the bench shape by writer count and observation dtype, and the one-launch step / observe
kernels at 1M envs.  Median over rounds; configurations alternate inside one process."""
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


def rollout(n, k, rounds=9):
    state, out0 = ops.new_state(n, dev), ops.EnvOutputs(n, dev)
    ops.init(ops.make_keys(1, n, dev), table, state, out0)
    cfgs = {}
    for dt, row in ((torch.float32, 1920), (torch.bfloat16, 960), (torch.uint8, 480)):
        traj = ops.EnvOutputs(n, dev, rows=k, obs_dtype=dt)
        for narrow in (False, True):
            for w in ((3, 4, 5) if dt == torch.float32 else (0,)):
                cfgs[f"rollout {str(dt)[6:]:8s} w{w} {'128-bit' if narrow else '256-bit'}"] = (traj, _lib.tune(writers=w, narrow_stores=narrow), row + 60)
    res = {name: [] for name in cfgs}
    step = [0]
    for r in range(rounds):
        for name, (traj, tune, _) in cfgs.items():
            def go():
                ops.rollout_random(state, table, k, traj, seed=1, step0=step[0], tune=tune); step[0] += k
            res[name].append(timed(go, 20))
    for name, v in sorted(res.items()):
        med = statistics.median(v)
        b = cfgs[name][2]
        print(f"{name:36s} n={n} median={med*1e3:8.2f} us  min={min(v)*1e3:8.2f}  GB/s={b*n*k/med/1e6:7.0f}  frac={b*n*k/med/1e6/peak:.3f}")


def single(n, rounds=7):
    state, out = ops.new_state(n, dev), ops.EnvOutputs(n, dev)
    ops.init(ops.make_keys(1, n, dev), table, state, out)
    cfgs = {}
    for narrow in (False, True):
        t = _lib.tune(narrow_stores=narrow)
        tag = '128-bit' if narrow else '256-bit'
        cfgs[f"step+autoreset random f32 {tag}"] = (lambda t=t: ops.step(state, None, table, state, out, autoreset=True, random_action=True, seed=3, tune=t), 80 + 80 + 1920 + 60)
        cfgs[f"observe f32 {tag}"] = (lambda t=t: ops.observe(state, table, out.observation, tune=t), 80 + 1920)
    res = {name: [] for name in cfgs}
    for r in range(rounds):
        for name, (fn, _) in cfgs.items():
            res[name].append(timed(fn, 10))
    for name, v in sorted(res.items()):
        med = statistics.median(v)
        b = cfgs[name][1]
        print(f"{name:36s} n={n} median={med*1e3:8.2f} us  min={min(v)*1e3:8.2f}  GB/s={b*n/med/1e6:7.0f}  frac={b*n/med/1e6/peak:.3f}")


if __name__ == "__main__":
    rollout(8192, 32)
    rollout(65536, 8, rounds=5)
    single(1 << 20)
