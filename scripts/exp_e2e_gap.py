"""Where do the ~14 us per step between `value` (0.0915 ms) and `e2e` (0.1056 ms) go at N=1?
(1) the kernel alone, back to back on one stream: Philox vs caller-supplied u32 / u16 uniforms, with / without result16;
(2) the pipelined host API with uniforms=NULL (pipeline overhead only) and with uniforms, depth 2/3/4."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brl_b200 import _lib, ops  # noqa: E402
from brl_b200.deals import synthetic_deal_table  # noqa: E402

dev = "cuda:0"
n, k = 8192, 32
table_np = synthetic_deal_table(100_000, seed=0)
table = torch.as_tensor(table_np, device=dev)
state, out0 = ops.new_state(n, dev), ops.EnvOutputs(n, dev)
ops.init(ops.make_keys(1, n, dev), table, state, out0)
traj = ops.EnvOutputs(n, dev, rows=k)
rng = np.random.default_rng(0)
pool32 = [torch.as_tensor(rng.integers(0, 2 ** 32, size=(k, n), dtype=np.uint32).view(np.int32), device=dev) for _ in range(8)]
pool16 = [torch.as_tensor(rng.integers(0, 2 ** 16, size=(k, n), dtype=np.uint16).view(np.int16), device=dev) for _ in range(8)]
res = torch.zeros((k, n), dtype=torch.int16, device=dev)
stats = torch.zeros(4, dtype=torch.int64, device=dev)


def kern(name, **kw):
    def f(i):
        u = kw.get("pool")
        ops.rollout_random(state, table, k, traj, seed=1, step0=i * k, stats=stats, uniforms=u[i % 8] if u else None,
                           result16=res if kw.get("res") else None)
    for i in range(5):
        f(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(50):
        f(5 + i)
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({"kernel_only": name, "ms_per_step": e0.elapsed_time(e1) / 50}), flush=True)


kern("philox")
kern("philox+result16", res=True)
kern("u32", pool=pool32)
kern("u16", pool=pool16)
kern("u16+result16", pool=pool16, res=True)

L = _lib.load()
tbl = np.ascontiguousarray(table_np)
vp = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
hp16 = torch.from_numpy(rng.integers(0, 2 ** 16, size=(8, k, n), dtype=np.uint16).view(np.int16)).pin_memory()
for depth in (2, 3, 4):
    for with_u in (False, True):
        h = L.brl_env_create(n, 0, tbl.ctypes.data, tbl.shape[0], 1, _lib.F_AUTORESET | _lib.F_UNIFORM_U16)
        assert L.brl_env_init_host(h, None, None, None, None, None) == 0
        r = [torch.zeros((k, n), dtype=torch.int16).pin_memory() for _ in range(depth)]
        st = [torch.zeros(4, dtype=torch.int64).pin_memory() for _ in range(depth)]

        def loop(count):
            infl = []
            for i in range(count):
                j = i % depth
                if len(infl) == depth:
                    L.brl_env_wait(h, infl.pop(0))
                infl.append(L.brl_env_rollout_host_compact_async(h, k, vp(hp16[i % 8]) if with_u else None, vp(r[j]), vp(st[j])))
            for t in infl:
                L.brl_env_wait(h, t)
        loop(10)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        loop(200)
        dt = time.perf_counter() - t0
        L.brl_env_destroy(h)
        print(json.dumps({"host_api": "compact_async", "depth": depth, "uniforms_from_host": with_u, "ms_per_step": 1e3 * dt / 200}), flush=True)
