"""CPU, world_size 2 over gloo: the N>1 host logic -- envs shard by GLOBAL index, no
collective in the data path, ONE all-reduce of the f64[8] statistics vector; the
whole-job result does not depend on the rank count."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from brl_b200 import deals, dist as bdist
from oracle import oracle as orc

N_TOTAL, K, SEED = 600, 10, 77


def _shard_returns(lo, hi, table):
    """per-env sum of player-0 rewards over K random-legal auto-reset steps (oracle compute)."""
    env = orc.OracleEnv(table, hi - lo)
    env.init(orc.make_keys(SEED, hi - lo, lo))           # keys by GLOBAL env index
    out = env.rollout_random(SEED, 0, K, env_offset=lo)  # action RNG by GLOBAL env index
    return out["rewards"][:, :, 0].sum(axis=0).astype(np.float64), out["n_terminated"]


def _sums(x):
    return torch.tensor([len(x), x.sum(), (x * x).sum(), (x > 0).sum(), 0, 0, 0, 0], dtype=torch.float64)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    table = deals.synthetic_deal_table(400, seed=1)
    lo, hi = bdist.shard_range(N_TOTAL, rank, world)
    x, _ = _shard_returns(lo, hi, table)
    sums = bdist.allreduce_sums(_sums(x))
    q.put((rank, lo, hi, sums.tolist(), bdist.stats_from_sums(sums)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_reproduce_the_single_rank_match_statistics():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, sums0, st0), (r1, lo1, hi1, sums1, st1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 300, 300, 600)
    assert sums0 == sums1 and st0 == st1                    # every rank holds the whole-match numbers
    table = deals.synthetic_deal_table(400, seed=1)
    x, _ = _shard_returns(0, N_TOTAL, table)                # one rank, all envs
    want = _sums(x)
    assert sums0[:4] == want[:4].tolist()
    np.testing.assert_allclose(st0, orc.match_stats(x), rtol=1e-12)
