"""Per-role SM clocks per step of k_rollout_ws (bench shape, default tuning) from a -DBRL_ROLE_TIMING build:
slots {0,1} env warp {work, barrier wait}, {2,3} warp 0, {4,5} the other non-env warps (summed)."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from brl_b200 import _lib, ops
from brl_b200.deals import synthetic_deal_table
L = _lib.load()
dev = "cuda:0"
table = torch.as_tensor(synthetic_deal_table(100000, 0), device=dev)
n, k = 8192, 32
state, out0 = ops.new_state(n, dev), ops.EnvOutputs(n, dev)
ops.init(ops.make_keys(1, n, dev), table, state, out0)
traj = ops.EnvOutputs(n, dev, rows=k)
for i in range(3):
    ops.rollout_random(state, table, k, traj, seed=1, step0=i * k)
torch.cuda.synchronize()
out = (C.c_ulonglong * 8)()
L.brl_debug_role_cycles(out, 1)
reps = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(reps):
    ops.rollout_random(state, table, k, traj, seed=1, step0=(3 + i) * k)
e1.record(); torch.cuda.synchronize()
L.brl_debug_role_cycles(out, 1)
blocks = 296
per = lambda v: v / (reps * blocks * (k + 1))
print("ms per launch (timing build)", e0.elapsed_time(e1) / reps)
print("env warp: work %.0f  barrier %.0f | warp 0: work %.0f barrier %.0f | other warps (sum): work %.0f barrier %.0f  [SM clocks per step]" %
      tuple(per(out[i]) for i in range(6)))
