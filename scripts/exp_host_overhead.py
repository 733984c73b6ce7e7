"""Host enqueue cost vs GPU time of the policy-loop calls (is the rollout CPU-launch-bound?)."""
import os, sys, time, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from brl_b200 import ops, BridgeBidding
from brl_b200 import random as brandom
from brl_b200.deals import synthetic_deal_table
from scripts.torch_baseline import TorchForwardPass  # noqa: E402
from brl_b200.models import LAYERS, init_params, make_forward_pass
from brl_b200.roll_out import make_roll_out
dev = "cuda:0"
n = 8192
table_np = synthetic_deal_table(100000, 1)
env = BridgeBidding(table=table_np, device=dev)
params = init_params(1, dev)
state = env.init(env.make_keys(1, n))


def host_and_gpu(fn, reps=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(reps): fn()
    e1.record(); t1 = time.perf_counter()
    torch.cuda.synchronize()
    return (t1 - t0) / reps * 1e6, e0.elapsed_time(e1) / reps * 1e3


for prec in ("tc", "tc-bf16"):
    fp = make_forward_pass(precision=prec) if prec.startswith("tc") else TorchForwardPass("relu", prec)
    a = torch.empty(n, dtype=torch.int32, device=dev)
    xb = ops.obs_to_bf16(state.observation)
    print(prec, "fp.act(bf16 obs)      host %.1f us  gpu %.1f us" % host_and_gpu(lambda: fp.act(params, xb, state._mask_u8, a, sample=True, seed=3)))
    print(prec, "fp.act(f32 obs)       host %.1f us  gpu %.1f us" % host_and_gpu(lambda: fp.act(params, state.observation, state._mask_u8, a, sample=True, seed=3)))
    print(prec, "fp.apply(f32 obs)     host %.1f us  gpu %.1f us" % host_and_gpu(lambda: fp.apply(params, state.observation)))
st2 = env.init(env.make_keys(2, n))
print("env.step autoreset        host %.1f us  gpu %.1f us" % host_and_gpu(lambda: env.step(st2, a, autoreset=True)))
config = dict(actor_illegal_action_mask=True, actor_illegal_action_penalty=False, game_mode="competitive",
              num_steps=32, reward_scale=7600.0, gamma=1.0, gae_lambda=0.95)
for prec in ("tc", "tc-bf16"):
    fp = make_forward_pass(precision=prec) if prec.startswith("tc") else TorchForwardPass("relu", prec)
    opp = init_params(2, dev)
    roll = make_roll_out(config, env, fp, fp)
    runner = [(params, None, state, state.observation, torch.zeros((), dtype=torch.int64, device=dev), brandom.PRNGKey(3))]
    def one():
        runner[0], _ = roll(runner[0], opp)
    h, g = host_and_gpu(one, reps=3)
    print(prec, "roll_out 32 x 8192: host %.2f ms  gpu %.2f ms  (per sub-step host %.1f us gpu %.1f us)" % (h / 1e3, g / 1e3, h / 128, g / 128))
