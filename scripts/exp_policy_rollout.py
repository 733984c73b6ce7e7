"""Where does the configs[2] rollout spend its time?  GPU-bound vs host-bound check: CUDA-event time of the
rollout vs the host time to ENQUEUE it (no sync inside), at 8192 and 16384 envs."""
import os, sys, time, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from brl_b200 import BridgeBidding, random as brandom
from brl_b200.deals import synthetic_deal_table
from brl_b200.gae import make_calc_gae
from brl_b200.models import init_params, make_forward_pass
from brl_b200.roll_out import make_roll_out
dev = "cuda:0"
table = synthetic_deal_table(100000, 0)
for n in (8192, 16384):
    env = BridgeBidding(table=table, device=dev)
    config = dict(actor_illegal_action_mask=True, actor_illegal_action_penalty=False, game_mode="competitive",
                  num_steps=32, reward_scale=7600.0, gamma=1.0, gae_lambda=0.95)
    fp = make_forward_pass("relu", "DeepMind")
    params, opp = init_params(1, dev), init_params(2, dev)
    state = env.init(env.make_keys(1, n))
    runner = (params, None, state, state.observation, torch.zeros((), dtype=torch.int64, device=dev), brandom.PRNGKey(3))
    roll_out = make_roll_out(config, env, fp, fp)
    runner, traj = roll_out(runner, opp)
    torch.cuda.synchronize()
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        runner, traj = roll_out(runner, opp)
        e1.record()
        t_enq = time.perf_counter() - t0
        torch.cuda.synchronize()
        print(f"n={n} rollout: gpu {e0.elapsed_time(e1):7.2f} ms   host enqueue {1e3*t_enq:7.2f} ms")
