"""Quick timing of the rollout kernel variants at the bench shape (8192 envs x 32 steps)."""
import sys, os, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from brl_b200 import ops, _lib
from brl_b200.deals import synthetic_deal_table
dev = "cuda:0"
table = torch.as_tensor(synthetic_deal_table(100000, 0), device=dev)
def run(n, k, tune=0, reps=20, label="", obs_dtype=torch.float32):
    state, out0 = ops.new_state(n, dev), ops.EnvOutputs(n, dev, obs_dtype)
    ops.init(ops.make_keys(1, n, dev), table, state, out0)
    traj = ops.EnvOutputs(n, dev, obs_dtype, rows=k)
    for i in range(3): ops.rollout_random(state, table, k, traj, seed=1, step0=i * k, tune=tune)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps): ops.rollout_random(state, table, k, traj, seed=1, step0=(3 + i) * k, tune=tune)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    bpe = {torch.float32: 1980, torch.uint8: 540, torch.bfloat16: 1020}[obs_dtype]
    gbs = bpe * n * k / ms / 1e6
    print(f"{label:30s} n={n:8d} k={k:3d} ms={ms:.4f} us/step={1e3*ms/k:.3f} GB/s={gbs:7.1f} frac={gbs/6555.2:.3f}")
if __name__ == "__main__":
    for n in (8192, 16384, 4096, 12000):
        run(n, 32, _lib.tune(balanced=False), label="ws auto, plain grid")
        run(n, 32, _lib.tune(balanced=True), label="ws auto, balanced grid")
        run(n, 32, _lib.tune(balanced=True, writers=4), label="ws4, balanced grid")
        run(n, 32, _lib.tune(balanced=True, writers=5), label="ws5, balanced grid")
    for epb, ws in ((32, (3, 5, 7)), (16, (2, 3))):
        for w in ws:
            run(8192, 32, _lib.tune(epw=epb, writers=w), label=f"ws epb={epb} writers={w}")
    run(8192, 32, _lib.tune(classic_rollout=True, epw=8), label="tile8")
    run(4096, 32, 0, label="ws auto"); run(16384, 32, 0, label="ws auto"); run(65536, 32, 0, label="ws auto")
    run(8192, 32, 0, label="ws auto u8 obs", obs_dtype=torch.uint8)
    run(8192, 32, 0, label="ws auto bf16 obs", obs_dtype=torch.bfloat16)
