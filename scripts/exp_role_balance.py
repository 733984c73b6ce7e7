"""Interleaved A/B of role-balance variants of the rollout kernel (bench shape: 8192 envs x 32 steps, f32 observations):
who draws the Philox uniforms (writer warp 0 / the env warp) and who stores the per-env scalars (last writer warp / the env
warp), by writer-warp count; Philox mode and caller-supplied uniforms.  Median over rounds; configurations alternate."""
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


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    peak = 6443.2
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    k, reps, rounds = 32, 20, 9
    state, out0 = ops.new_state(n, dev), ops.EnvOutputs(n, dev)
    ops.init(ops.make_keys(1, n, dev), table, state, out0)
    traj = ops.EnvOutputs(n, dev, rows=k)
    u = torch.randint(0, 2 ** 31 - 1, (k, n), dtype=torch.int32, device=dev)
    cfgs = {}
    for ep in (False, True):
        for es in (False, True):
            for w in (3, 4, 5):
                for uni in (False, True):
                    if uni and (ep or w != 4):
                        continue
                    cfgs[f"philox@{'env' if ep else 'w0 '} scalars@{'env ' if es else 'last'} w{w} {'caller-u' if uni else 'philox'}"] = \
                        (_lib.tune(writers=w, env_philox=ep, env_scalars=es), uni)
    res = {name: [] for name in cfgs}
    step = 0
    for r in range(rounds):
        for name, (tune, uni) in cfgs.items():
            kw = dict(uniforms=u) if uni else {}
            for _ in range(3):
                ops.rollout_random(state, table, k, traj, seed=1, step0=step, tune=tune, **kw); step += k
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                ops.rollout_random(state, table, k, traj, seed=1, step0=step, tune=tune, **kw); step += k
            e1.record()
            torch.cuda.synchronize()
            res[name].append(e0.elapsed_time(e1) / reps)
    for name, v in sorted(res.items(), key=lambda kv: statistics.median(kv[1])):
        med = statistics.median(v)
        print(f"{name:44s} n={n} median={med*1e3:7.2f} us  min={min(v)*1e3:7.2f}  max={max(v)*1e3:7.2f}  "
              f"frac(median)={1980*n*k/med/1e6/peak:.3f}")


if __name__ == "__main__":
    main()
