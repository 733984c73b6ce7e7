"""BASELINE.json configs[4]: FSP/PFSP league probe -- the learner against a pool of 5 models, 1,048,576 envs in
total, sharded by global env index over the ranks, ONE all-reduce of f64[5, 8].

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port P \
        scripts/league_rollout.py --envs 1048576

The five bundled `bridge_models/*.pkl` cannot travel to the GPU box; pass --models <dir> to use them, otherwise
random-init nets of the same architecture stand in (the arithmetic and traffic are identical)."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from brl_b200 import BridgeBidding, random as brandom  # noqa: E402
from brl_b200.deals import synthetic_deal_table  # noqa: E402
from brl_b200.evaluation import make_league_evaluate  # noqa: E402
from brl_b200.models import init_params, load_params  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1 << 20)
    ap.add_argument("--models", default=None)
    ap.add_argument("--reps", type=int, default=2)
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    env = BridgeBidding(table=synthetic_deal_table(100_000, seed=0), device=dev)
    if a.models:
        files = sorted(f for f in os.listdir(a.models) if f.endswith(".pkl"))[:5]
        pool = [load_params(os.path.join(a.models, f), dev) for f in files]
    else:
        pool = [init_params(100 + m, dev) for m in range(5)]
    actor = init_params(1, dev)
    league = make_league_evaluate(env, "relu", "DeepMind", a.envs, len(pool))
    league(actor, pool, brandom.PRNGKey(0))  # warm-up
    best = None
    for r in range(a.reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = league(actor, pool, brandom.PRNGKey(1))
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best = float(t) if best is None else min(best, float(t))
    if rank == 0:
        print(json.dumps({"workload": "configs[4]: league probe vs a pool of 5 models", "n_envs_total": a.envs, "n_gpus": world,
                          "ms_per_league": best, "boards_per_sec": 2 * a.envs / (best * 1e-3),
                          "imp_mean": [r[0] for r in res], "imp_se": [r[1] for r in res], "win_rate": [r[2] for r in res],
                          "collective": "1 all-reduce of f64[5,8]" if world > 1 else "none (1 rank)"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
