"""Debug experiment: per-role cycle accounting of the warp-specialised rollout kernel.
Builds a private copy of the library with -DBRL_ROLE_TIMING (never the shipped one)."""
import ctypes as C, os, subprocess, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
so = "/tmp/libbrl_timing.so"
src = [os.path.join(ROOT, "brl_b200", "csrc", f) for f in ("brl_env.cu", "brl_algo.cu", "brl_host.cu", "xla_ffi_shim.cc")]
subprocess.run(["nvcc", "-std=c++17", "-O3", "-lineinfo", "-gencode", "arch=compute_100a,code=sm_100a", "-DBRL_ROLE_TIMING",
                "-Xcompiler", "-fPIC", "-shared", "-cudart", "static", "-o", so, *src], check=True)
from brl_b200 import _lib, build
build.LIB = so
_lib._LIB = None
import brl_b200.build as b
b._stale = lambda: False
L = _lib.load()
from brl_b200 import ops
from brl_b200.deals import synthetic_deal_table
dev = "cuda:0"
table = torch.as_tensor(synthetic_deal_table(100000, 0), device=dev)
n, k = 8192, 32
for kw in ({"epw": 32, "writers": 3}, {"epw": 32, "writers": 7}, {"epw": 16, "writers": 2}, {"epw": 8, "writers": 1}, {"epw": 8, "writers": 2}):
    state, out0 = ops.new_state(n, dev), ops.EnvOutputs(n, dev)
    ops.init(ops.make_keys(1, n, dev), table, state, out0)
    traj = ops.EnvOutputs(n, dev, rows=k)
    tune = _lib.tune(**kw)
    for i in range(3):
        ops.rollout_random(state, table, k, traj, seed=1, step0=i * k, tune=tune)
    torch.cuda.synchronize()
    out = (C.c_ulonglong * 8)()
    L.brl_debug_role_cycles(out, 1)
    reps = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        ops.rollout_random(state, table, k, traj, seed=1, step0=(3 + i) * k, tune=tune)
    e1.record(); torch.cuda.synchronize()
    L.brl_debug_role_cycles(out, 1)
    blocks = (n + kw["epw"] - 1) // kw["epw"]
    per = lambda v, warps: v / (reps * blocks * warps * k)  # cycles per step per warp
    w_other = max(1, kw["writers"] - 1)
    print(kw, f"ms={e0.elapsed_time(e1)/reps:.4f}",
          f"env work={per(out[0],1):.0f} wait={per(out[1],1):.0f} | writer0 work={per(out[2],1):.0f} wait={per(out[3],1):.0f}"
          f" | other writers work={per(out[4],w_other):.0f} wait={per(out[5],w_other):.0f}  (cycles per step)")
