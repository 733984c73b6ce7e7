"""Per-tile timeline of the fused policy forward (k_mlp_fused) at 8192 envs, from a -DBRL_MLP_TRACE build:
    python scripts/build_variant.py mlptrace -DBRL_MLP_TRACE
    BRL_B200_LIB=brl_b200/lib/libbrl_mlptrace.so python scripts/exp_mlp_trace.py
Prints, per layer: when its tiles start / end, the dependency-wait time, the main-loop time, the epilogue lag; and the
idle time of every CTA pair."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brl_b200 import _lib, ops  # noqa: E402
from brl_b200.models import init_params, make_forward_pass  # noqa: E402

dev, n = "cuda:0", int(sys.argv[1]) if len(sys.argv) > 1 else 8192
g = torch.Generator().manual_seed(0)
obs = ops.obs_to_bf16((torch.rand((n, 480), generator=g) < 0.05).to(torch.float32).to(dev))
mask = torch.ones((n, 38), dtype=torch.uint8, device=dev)
action = torch.empty(n, dtype=torch.int32, device=dev)
fp = make_forward_pass(precision=sys.argv[2] if len(sys.argv) > 2 else "tc")
params = init_params(1, dev)
for i in range(5):
    fp.act(params, obs, mask, action, None, None, sample=False, seed=i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(20):
    fp.act(params, obs, mask, action, None, None, sample=False, seed=i)
e1.record()
torch.cuda.synchronize()
print("forward ms", e0.elapsed_time(e1) / 20)
L = _lib.load()
epi = np.zeros(4, np.uint64)
L.brl_debug_mlp_epi(epi.ctypes.data_as(C.c_void_p), 1)
fp.act(params, obs, mask, action, None, None, sample=False, seed=0)
torch.cuda.synchronize()
L.brl_debug_mlp_epi(epi.ctypes.data_as(C.c_void_p), 0)
print("epilogue of CTA 0 warp 2 over one forward (SM clocks): tcgen05.ld+wait %d, arithmetic %d, stores %d" % tuple(int(x) for x in epi[:3]))
buf = np.zeros((1024, 8), np.uint64)
assert L.brl_debug_mlp_trace(buf.ctypes.data_as(C.c_void_p)) == 0
nmb = (n + 255) // 256
tpl = nmb * 4
n_tiles = 4 * tpl + nmb
t = buf[:n_tiles].astype(np.int64)
t0 = t[:, 0].min()
rel = (t[:, :7] - t0) / 1e3
print(f"tiles {n_tiles}, span {rel[:, 6].max():.1f} us")
for layer in range(5):
    sl = slice(layer * tpl, (layer + 1) * tpl) if layer < 4 else slice(4 * tpl, n_tiles)
    r = rel[sl]
    print(f"layer {layer}: first start {r[:,0].min():6.1f} last end {r[:,6].max():6.1f} | dep wait mean {np.mean(r[:,1]-r[:,0]):5.2f} max {np.max(r[:,1]-r[:,0]):5.2f}"
          f" | mma span mean {np.mean(r[:,4]-r[:,3]):5.2f} | acc-free wait (3-1) mean {np.mean(r[:,3]-r[:,1]):5.2f}"
          f" | epilogue (6-5) mean {np.mean(r[:,6]-r[:,5]):5.2f} | tile total (6-0) mean {np.mean(r[:,6]-r[:,0]):5.2f}")
pairs = t[:, 7]
busy = {}
for p in np.unique(pairs):
    sel = pairs == p
    busy[p] = (rel[sel, 4] - rel[sel, 3]).sum()
b = np.array(list(busy.values()))
print(f"MMA-busy per pair: mean {b.mean():.1f} us min {b.min():.1f} max {b.max():.1f}  (span {rel[:,6].max():.1f})")
# pair 0's timeline
sel = np.nonzero(pairs == 0)[0]
for i in sel:
    print("pair0 tile", i, "layer", min(i // tpl, 4), " ".join(f"{x:7.2f}" for x in rel[i]))
