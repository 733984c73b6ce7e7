"""Stand-alone timing of every kernel on the env path (CUDA events), one JSON line each.

    python scripts/prof_kernels.py [--only rollout,step,observe,mask,gae,dup,categorical] [--reps R]

Run plain for the numbers; run under `ncu --set full -k regex:<kernel>` for the DRAM-traffic
evidence committed in profiles/ (numbers printed under the profiler are never bench values).
Sizes are chosen so every launch's footprint exceeds the 126 MB L2 (1,048,576 envs for the
one-launch-per-step kernels), i.e. the stores are real DRAM traffic.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from brl_b200 import ops  # noqa: E402
from brl_b200.deals import synthetic_deal_table  # noqa: E402

dev = "cuda:0"


def peak():
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"])
    except Exception:
        return 6650.0


def timed(fn, reps, warm=3, graph=False):
    """CUDA-event time per launch.  graph=True: the `reps` launches are captured into one CUDA graph and replayed,
    so a kernel shorter than the Python/ctypes launch path (~15-20 us) is timed back to back on the device."""
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    if graph:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(reps):
                fn(warm + i)
        g.replay()
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / reps)
        ts.sort()
        return sum(ts) / len(ts), ts[len(ts) // 2], ts[0]
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    evs[0].record()
    for i in range(reps):
        fn(warm + i)
        evs[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(reps))
    return sum(ts) / len(ts), ts[len(ts) // 2], ts[0]


def report(name, n, units, bytes_per_unit, ms_avg, ms_med, ms_min, note=""):
    gbs = bytes_per_unit * units / (ms_avg * 1e-3) / 1e9
    print(json.dumps({"kernel": name, "n_envs": n, "units": units, "algorithmic_bytes_per_unit": bytes_per_unit,
                      "algorithmic_MB_per_launch": round(bytes_per_unit * units / 1e6, 2),
                      "ms_avg": round(ms_avg, 5), "ms_median": round(ms_med, 5), "ms_min": round(ms_min, 5),
                      "GBps": round(gbs, 1), "frac_of_measured_hbm": round(gbs / peak(), 4),
                      "units_per_sec": units / (ms_avg * 1e-3), "note": note}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="rollout,step,observe,mask,gae,dup,categorical")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--big", type=int, default=1 << 20)
    ap.add_argument("--graph", action="store_true", help="time CUDA-graph replays (no host launch gaps)")
    a = ap.parse_args()
    global timed
    if a.graph:
        _timed = timed
        timed = lambda fn, reps, warm=3: _timed(fn, reps, warm, graph=True)  # noqa: E731
    only = set(a.only.split(","))
    table = torch.as_tensor(synthetic_deal_table(100000, 0), device=dev)
    R = a.reps

    if "rollout" in only:
        for n, k in ((8192, 32), (65536, 8)):
            state, out0 = ops.new_state(n, dev), ops.EnvOutputs(n, dev)
            ops.init(ops.make_keys(1, n, dev), table, state, out0)
            traj = ops.EnvOutputs(n, dev, rows=k)
            act = torch.empty((k, n), dtype=torch.int32, device=dev)
            t = timed(lambda i: ops.rollout_random(state, table, k, traj, seed=1, step0=i * k, action_out=act), R)
            report("k_rollout_ws<f32>", n, n * k, 1980, *t, note=f"{k} auto-reset sub-steps per launch")
            del state, out0, traj, act

    n = a.big
    if only & {"step", "observe", "mask", "dup"}:
        state, out = ops.new_state(n, dev), ops.EnvOutputs(n, dev)
        ops.init(ops.make_keys(1, n, dev), table, state, out)
        # advance a few random steps so histories / masks are non-trivial
        for i in range(6):
            ops.step(state, None, table, state, out, autoreset=True, random_action=True, seed=1, step_index=i)
    if "step" in only:
        act = torch.empty(n, dtype=torch.int32, device=dev)
        t = timed(lambda i: ops.step(state, None, table, state, out, autoreset=True, random_action=True, seed=1,
                                     step_index=10 + i, action_out=act), R)
        report("k_step<32,f32> (auto-reset, in-kernel random action)", n, n, 1980, *t)
    if "observe" in only:
        t = timed(lambda i: ops.observe(state, table, out.observation), R)
        report("k_produce observe f32", n, n, 1920, *t, note="+80 B/env packed-state read not counted")
        obs8 = torch.empty((n, 480), dtype=torch.uint8, device=dev)
        t = timed(lambda i: ops.observe(state, table, obs8), R)
        report("k_produce observe u8 (pgx bool dtype)", n, n, 480, *t, note="+80 B/env packed-state read not counted")
        del obs8
    if "mask" in only:
        t = timed(lambda i: ops.legal_mask(state, out.legal_action_mask), R)
        report("k_legal_mask", n, n, 38, *t, note="reads one 16 B state plane per env for 38 B/env out: 54 B/env of real "
               "traffic, so 0.70 is the ceiling of this algorithmic fraction")
    if "dup" in only:
        ia, ib = ops.TableInfoBuffers(n, dev), ops.TableInfoBuffers(n, dev)
        st2 = state.clone()
        actions = torch.zeros(n, dtype=torch.int32, device=dev)
        t = timed(lambda i: ops.duplicate_step(st2, actions, table, ia, ib, st2, out), R)
        report("k_dup_step<32,f32> (all-pass actions)", n, n, 1980, *t)
        del st2, ia, ib
    if "gae" in only:
        for T, m in ((32, 8192), (32, 1 << 20)):
            done = (torch.rand((T, m), device=dev) < 0.1).to(torch.uint8)
            value, reward = torch.randn((T, m), device=dev), torch.randn((T, m), device=dev)
            last = torch.randn(m, device=dev)
            adv, tgt = torch.empty_like(value), torch.empty_like(value)
            t = timed(lambda i: ops.gae(done, value, reward, last, adv, tgt, 1.0, 0.95), R)
            report("k_gae", m, T * m, 17, *t, note=f"T={T}")
            del done, value, reward, last, adv, tgt
    if "categorical" in only:
        m = 1 << 20
        logits = torch.randn((m, 38), device=dev)
        mask = (torch.rand((m, 38), device=dev) < 0.6).to(torch.uint8)
        mask[:, 0] = 1
        act, lp = torch.empty(m, dtype=torch.int32, device=dev), torch.empty(m, device=dev)
        t = timed(lambda i: ops.categorical(logits, mask, act, lp, sample=True, seed=3, step_index=i), R)
        report("k_categorical (Gumbel sample + log-prob)", m, m, 38 * 4 + 38 + 8, *t)


if __name__ == "__main__":
    main()
