"""Wall-clock of whole ppo.py iterations at the reference's default shape (num_envs=8192, num_steps=32, minibatch 1024,
update_epochs=10 -> 2560 optimizer steps per iteration), through brl_b200.ppo.train: rollout / GAE / update times as
ppo.py logs them ("time/rollout", "time/calc_gae", "time/update"), tensor-core path vs the library-GEMM path.

    python scripts/ppo_iteration.py [--iters 3] [--fp32-iters 1]
"""
import argparse
import json
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--fp32-iters", type=int, default=1)
    args = ap.parse_args()
    from scripts.torch_baseline import TorchForwardPass, make_update_step_autograd
    from brl_b200 import ppo, random as brandom
    from brl_b200.deals import synthetic_deal_table
    tables = [synthetic_deal_table(100_000, seed=k) for k in range(2)]
    eval_table = synthetic_deal_table(10_000, seed=99)
    out = {"shape": "num_envs 8192, num_steps 32, minibatch 1024, update_epochs 10 (2560 optimizer steps), lr 1e-6 (ppo.py defaults)"}
    for prec, iters in (("tc", args.iters), ("fp32", args.fp32_iters)):
        if iters <= 0:
            continue
        with tempfile.TemporaryDirectory() as tmp:
            cfg = dict(total_timesteps=8192 * 32 * (iters + 2), num_eval_envs=1000, num_eval_step=1000, num_prioritized_envs=100,
                       save_model=False, log_path=tmp)
            hooks = {} if prec == "tc" else dict(forward_pass_factory=lambda act, mt: TorchForwardPass(act, "fp32"),
                                                 update_step_factory=make_update_step_autograd)
            _, logs = ppo.train(cfg, brandom.PRNGKey(0), tables=tables, eval_table=eval_table, device="cuda:0", **hooks)
        logs = logs[2:]  # the first two iterations pay lazy initialisation (library load, the second deal table's buffers)
        mean = lambda k: sum(log[k] for log in logs) / len(logs)  # noqa: E731
        out[prec] = {"iterations_timed": len(logs), "rollout_s": mean("time/rollout"), "calc_gae_s": mean("time/calc_gae"),
                     "update_s": mean("time/update"), "rollout_s_each": [round(log["time/rollout"], 4) for log in logs],
                     "table_rotated_before": ["table" in log for log in logs], "rollout+gae+update_s": mean("time/rollout") + mean("time/calc_gae") + mean("time/update"),
                     "agent_steps_per_s": 8192 * 32 / (mean("time/rollout") + mean("time/calc_gae") + mean("time/update")),
                     "last_total_loss": logs[-1]["train/total_loss"], "last_entropy": logs[-1]["train/policy_entropy"]}
    if "tc" in out and "fp32" in out:
        out["speedup"] = out["fp32"]["rollout+gae+update_s"] / out["tc"]["rollout+gae+update_s"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
