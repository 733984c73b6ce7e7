import sys, torch, time
sys.path.insert(0, "/root/repo")
from brl_b200 import ops, _lib
from brl_b200.deals import synthetic_deal_table
dev = "cuda:0"
table = torch.as_tensor(synthetic_deal_table(100000, 0), device=dev)
def run(n, k, tune=0, obs=True, reps=20, label=""):
    state, out0 = ops.new_state(n, dev), ops.EnvOutputs(n, dev)
    ops.init(ops.make_keys(1, n, dev), table, state, out0)
    traj = ops.EnvOutputs(n, dev, rows=k)
    if not obs:
        class O: pass
        o = ops.EnvOutputs.__new__(ops.EnvOutputs)
        o.observation = traj.observation; o.legal_action_mask = traj.legal_action_mask; o.rewards = traj.rewards
        o.terminated = traj.terminated; o.current_player = traj.current_player
        # call with obs pointer None via obs_only=None path
        def call(i):
            ops._call("brl_rollout_random", [ops._ptr(state), ops._ptr(table), None, ops._ptr(traj.legal_action_mask), ops._ptr(traj.rewards), ops._ptr(traj.terminated), ops._ptr(traj.current_player), None, None, None],
                      ops._params(n, flags=tune, n_deals=table.shape[0], stride=n, seed=1, step=i * k, k_steps=k))
    else:
        def call(i):
            ops.rollout_random(state, table, k, traj, seed=1, step0=i * k, tune=tune)
    for i in range(3): call(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps): call(3 + i)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    gbs = 1980 * n * k / ms / 1e6
    print(f"{label:28s} n={n:8d} k={k:3d} ms={ms:.4f} us/step={1e3*ms/k:.3f} GB/s={gbs:7.1f} frac={gbs/6555.2:.3f}")
for epb, ws in ((8, (1, 2, 3)), (16, (1, 2, 3, 5)), (32, (3, 5, 7))):
    for w in ws:
        run(8192, 32, _lib.tune(epw=epb, writers=w), True, label=f"ws epb={epb} writers={w}")
run(8192, 32, _lib.tune(epw=8, writers=1), False, label="ws epb=8 w=1 no-obs")
for n in (4096, 16384, 32768, 65536, 262144):
    run(n, 32 if n <= 65536 else 8, 0, True, label="ws auto")
for n in (65536, 1048576):
    run(n, 8, _lib.tune(classic_rollout=True, epw=8, wpb=4), True, label="tile8")
    run(n, 2, 0, True, label="ws auto k=2")
    run(n, 2, _lib.tune(classic_rollout=True, epw=8, wpb=4), True, label="tile8 k=2")
