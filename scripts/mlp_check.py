"""Numerics + timing of the tcgen05 policy forward against torch (fp64 reference, cuBLAS fp32/tf32/bf16).

    python scripts/mlp_check.py [n_envs ...]
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from brl_b200 import ops  # noqa: E402
from brl_b200.deals import synthetic_deal_table  # noqa: E402
from scripts.torch_baseline import TorchForwardPass  # noqa: E402
from brl_b200.models import LAYERS, init_params, make_forward_pass  # noqa: E402

dev = "cuda:0"


def real_obs(n):
    table = torch.as_tensor(synthetic_deal_table(1000, 1), device=dev)
    state, out = ops.new_state(n, dev), ops.EnvOutputs(n, dev)
    ops.init(ops.make_keys(3, n, dev), table, state, out)
    for i in range(7):
        ops.step(state, None, table, state, out, autoreset=True, random_action=True, seed=3, step_index=i)
    return out.observation


def ref64(params, x):
    h = x.double()
    for name in LAYERS[:4]:
        h = torch.relu(h @ params[name]["w"].double() + params[name]["b"].double())
    return h @ params[LAYERS[4]]["w"].double() + params[LAYERS[4]]["b"].double(), \
        (h @ params[LAYERS[5]]["w"].double() + params[LAYERS[5]]["b"].double()).squeeze(-1)


def timeit(fn, reps=20):
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


def main():
    modes = ("tc", "tc-bf16", "fp32", "tf32", "bf16")
    args = sys.argv[1:]
    if args and args[0].startswith("--modes="):
        modes = tuple(args.pop(0).split("=")[1].split(","))
    sizes = [int(a) for a in args] or [100, 1000, 8192]
    params = init_params(5, dev)
    for name in LAYERS:  # non-zero biases so the bias path is exercised
        params[name]["b"] = torch.randn_like(params[name]["b"]) * 0.1
    for n in sizes:
        x = real_obs(n)
        l64, v64 = ref64(params, x)
        scale_l, scale_v = float(l64.abs().max()), float(v64.abs().max())
        row = {"n_envs": n}
        for prec in modes:
            fp = make_forward_pass(precision=prec) if prec.startswith("tc") else TorchForwardPass("relu", prec)
            logits, value = fp.apply(params, x)
            torch.cuda.synchronize()
            el = float((logits.double() - l64).abs().max()) / scale_l
            ev = float((value.double() - v64).abs().max()) / scale_v
            agree = float((logits.argmax(1) == l64.argmax(1)).float().mean())
            ms = timeit(lambda: fp.apply(params, x))
            row[prec] = {"logit_err_rel_max": el, "value_err_rel_max": ev, "argmax_agree": agree, "ms": round(ms, 4),
                         "TFLOPs": round(7354368 * n / ms / 1e9, 1)}
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
