"""BASELINE.json configs[3]: duplicate self-play match, envs sharded by GLOBAL index across ranks,
IMP statistics all-reduced once over NCCL.

    python scripts/dup_selfplay.py [--envs 65536] [--reps 3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        scripts/dup_selfplay.py --envs 65536

Prints one JSON line on rank 0: match statistics (identical for every rank count, because keys and
action RNG use global env indices), env-steps/s (2 tables x auction length x envs / max-over-ranks
CUDA-event time) and boards/s."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from brl_b200 import BridgeBidding, dist as bdist, random as brandom  # noqa: E402
from brl_b200.deals import synthetic_deal_table  # noqa: E402
from brl_b200.evaluation import make_simple_duplicate_evaluate  # noqa: E402
from brl_b200.models import init_params  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--deals", type=int, default=100_000)
    a = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lo, hi = bdist.shard_range(a.envs, rank, world)
    env = BridgeBidding(table=synthetic_deal_table(a.deals, seed=0), device=dev)
    p1, p2 = init_params(1, dev), init_params(2, dev)   # self-play pool stand-ins (random init; no checkpoints offline)
    evaluate = make_simple_duplicate_evaluate(env, "relu", "DeepMind", "relu", "DeepMind", hi - lo, env_offset=lo)
    evaluate(p1, p2, brandom.PRNGKey(0))               # warm-up (packs weights, loads kernels, NCCL init)
    times, res = [], None
    for r in range(a.reps):
        trace = []
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = evaluate(p1, p2, brandom.PRNGKey(1 + r), trace=trace)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        steps = torch.tensor([float(len(trace))], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(steps, op=dist.ReduceOp.MAX)
        times.append((float(t), int(steps)))
    (mean, se, win), info_a, info_b, cum = res
    # env-steps actually taken = live (env, step) pairs; count them from the table infos' auction lengths is not
    # recorded, so report loop iterations x envs as the upper bound and boards/s as the exact figure
    ms, iters = times[-1]
    if rank == 0:
        print(json.dumps({"workload": "configs[3]: duplicate self-play, DeepMind MLP x2 (random init), argmax play",
                          "n_envs_total": a.envs, "n_gpus": world, "envs_per_rank": hi - lo,
                          "imp_mean": mean, "imp_se": se, "win_rate": win, "ms_per_match": ms, "loop_iterations": iters,
                          "boards_per_sec": 2 * a.envs / (ms * 1e-3), "env_steps_per_sec_upper": iters * a.envs / (ms * 1e-3),
                          "all_ms": [t for t, _ in times],
                          "collective": "1 all-reduce of f64[8] per match (NCCL)" if world > 1 else "none (1 rank)"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
