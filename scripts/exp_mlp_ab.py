"""A/B of MLP forward tile shapes (interleaved, CUDA events)."""
import os, sys, statistics, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from brl_b200 import ops
from brl_b200.models import LAYERS, init_params
dev = "cuda:0"
for n in (8192, 65536):
    params = init_params(1, dev)
    blob = ops.mlp_pack([params[k]["w"] for k in LAYERS], [params[k]["b"] for k in LAYERS])
    x = (torch.rand((n, 480), device=dev) < 0.05).to(torch.bfloat16)
    scratch = ops.mlp_scratch(n, dev)
    lg, v = torch.empty((n, 38), device=dev), torch.empty(n, device=dev)
    ref = None
    res = {}
    for rnd in range(5):
        for name, kw in (("bn128 x3", {}), ("bn256 x3", {"tune": 1 << 26}), ("bn128 bf16", {"single_bf16": True}),
                         ("bn256 bf16", {"single_bf16": True, "tune": 1 << 26})):
            for _ in range(3): ops.mlp_forward(x, blob, scratch, lg, v, **kw)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20): ops.mlp_forward(x, blob, scratch, lg, v, **kw)
            e1.record(); torch.cuda.synchronize()
            res.setdefault(name, []).append(e0.elapsed_time(e1) / 20)
            if name == "bn128 x3": ref = lg.clone()
            if name == "bn256 x3": assert torch.equal(lg, ref), "bn256 differs from bn128"
    for name, t in res.items():
        m = statistics.median(t)
        print(f"n={n} {name:12s} median {m*1e3:8.1f} us  model TFLOP/s {7354368*n/m/1e9:7.1f}")
