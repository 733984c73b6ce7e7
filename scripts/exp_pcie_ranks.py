"""Where does the host-buffer rollout (`e2e`) lose time as ranks are added?  Run under torchrun at N = 1/2/4/8.

Per transfer size of the e2e payloads (f32 payload: 4.46 MB out + 1.05 MB in per step; compact: 0.52 MB out + 0.52 / 1.05 MB in)
each rank times pinned-memory cudaMemcpyAsync D2H / H2D / both (two streams)
  (a) alone  -- one rank at a time, the others idle at a barrier;
  (b) all ranks at once.
(a) == (b) per rank means private links end to end; (b) slower means a shared hop (PCIe switch uplink, root complex or host
DRAM).  Also times the bare submit loop of the C-ABI (uniforms=NULL, result copies only) to bound the host-side call cost.
Prints one JSON line per rank-0 measurement.
"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
    dist.barrier()


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def time_copies(pairs, reps=200):
    """pairs: list of (dst, src, stream); per-iteration time (us) of issuing all pairs back to back, `reps` times"""
    for dst, src, st in pairs:
        with torch.cuda.stream(st):
            for _ in range(5):
                dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        for dst, src, st in pairs:
            with torch.cuda.stream(st):
                dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e6


s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
cases = {"f32_payload": (int(4.46e6), int(1.05e6)), "compact_u32": (int(0.524e6), int(1.05e6)), "compact_u16": (int(0.524e6), int(0.524e6))}
for name, (out_b, in_b) in cases.items():
    d_o = torch.empty(out_b, dtype=torch.uint8, device=dev); h_o = torch.empty(out_b, dtype=torch.uint8).pin_memory()
    d_i = torch.empty(in_b, dtype=torch.uint8, device=dev); h_i = torch.empty(in_b, dtype=torch.uint8).pin_memory()
    legs = {"d2h": [(h_o, d_o, s1)], "h2d": [(d_i, h_i, s2)], "duplex": [(h_o, d_o, s1), (d_i, h_i, s2)]}
    for leg, pairs in legs.items():
        alone = []
        for r in range(world):  # one rank at a time
            barrier()
            if r == rank:
                alone.append(time_copies(pairs))
            barrier()
        barrier()
        together = time_copies(pairs)
        barrier()
        t = torch.tensor([alone[0], together], dtype=torch.float64, device=dev)
        if world > 1:
            g = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(g, t)
        else:
            g = [t]
        if rank == 0:
            a = [float(x[0]) for x in g]; b = [float(x[1]) for x in g]
            nbytes = sum(p[0].numel() for p in pairs)
            print(json.dumps({"case": name, "leg": leg, "bytes": nbytes, "world": world,
                              "alone_us": [round(x, 1) for x in a], "together_us": [round(x, 1) for x in b],
                              "alone_GBps": round(nbytes / (sum(a) / len(a)) / 1e3, 1),
                              "together_GBps_per_rank": round(nbytes / (sum(b) / len(b)) / 1e3, 1),
                              "together_GBps_total": round(world * nbytes / max(b) / 1e3, 1)}), flush=True)

# host-side cost of the C-ABI submit loop (no uniforms: nothing but launch + result copy per call)
import ctypes as C  # noqa: E402

import numpy as np  # noqa: E402

from brl_b200 import _lib  # noqa: E402
from brl_b200.deals import synthetic_deal_table  # noqa: E402

L = _lib.load()
n, k = 8192, 32
tbl = np.ascontiguousarray(synthetic_deal_table(100_000, seed=0))
h = L.brl_env_create(n, rank * n, tbl.ctypes.data, tbl.shape[0], 1, _lib.F_AUTORESET)
assert L.brl_env_init_host(h, None, None, None, None, None) == 0
res = [torch.zeros((k, n), dtype=torch.int16).pin_memory() for _ in range(3)]
st = [torch.zeros(4, dtype=torch.int64).pin_memory() for _ in range(3)]
vp = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
host_us = []
for i in range(60):
    t0 = time.perf_counter()
    t = L.brl_env_rollout_host_compact_async(h, k, None, vp(res[i % 3]), vp(st[i % 3]))
    host_us.append((time.perf_counter() - t0) * 1e6)
    if i >= 2:
        L.brl_env_wait(h, t - 2)
torch.cuda.synchronize()
L.brl_env_destroy(h)
if rank == 0:
    host_us.sort()
    print(json.dumps({"case": "submit_call_host_us", "median": round(host_us[len(host_us) // 2], 1), "p90": round(host_us[int(len(host_us) * 0.9)], 1)}))
if world > 1:
    dist.destroy_process_group()
