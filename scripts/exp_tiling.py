"""Tile-shape sweep of the one-launch-per-step kernels at 1,048,576 envs: envs per block (EPW) x warps per block.
The stores of a block sweep EPW*WPB consecutive 1920-byte rows; scripts/exp_store_paths.cu shows a pure-write
kernel goes from 6.5 TB/s (32 rows/warp) to 7.5 TB/s (1 row/warp) as the window of addresses in flight shrinks."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from brl_b200 import ops  # noqa: E402
from brl_b200.deals import synthetic_deal_table  # noqa: E402
from scripts.prof_kernels import timed, peak  # noqa: E402

dev = "cuda:0"
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
table = torch.as_tensor(synthetic_deal_table(100000, 0), device=dev)
state, out = ops.new_state(n, dev), ops.EnvOutputs(n, dev)
ops.init(ops.make_keys(1, n, dev), table, state, out)
for i in range(6):
    ops.step(state, None, table, state, out, autoreset=True, random_action=True, seed=1, step_index=i)
obs8 = torch.empty((n, 480), dtype=torch.uint8, device=dev)
act = torch.empty(n, dtype=torch.int32, device=dev)
from brl_b200 import _lib  # noqa: E402
for epw in (0, 8, 16, 32):
    for wpb in (0, 1, 2, 4, 8):
        if (epw == 0) != (wpb == 0):
            continue
        tune = _lib.tune(epw=epw, wpb=wpb)
        r = {"epw": epw, "wpb": wpb}
        t = timed(lambda i: ops.observe(state, table, out.observation, tune=tune), 20)
        r["observe_f32_GBps"] = round(1920 * n / t[0] / 1e6, 1)
        t = timed(lambda i: ops.observe(state, table, obs8, tune=tune), 20)
        r["observe_u8_GBps"] = round(480 * n / t[0] / 1e6, 1)
        t = timed(lambda i: ops.legal_mask(state, out.legal_action_mask, tune=tune), 20)
        r["mask_GBps"] = round(38 * n / t[0] / 1e6, 1)
        t = timed(lambda i: ops.step(state, None, table, state, out, autoreset=True, random_action=True, seed=1,
                                     step_index=10 + i, action_out=act, tune=tune), 20)
        r["step_GBps"] = round(1980 * n / t[0] / 1e6, 1)
        print(json.dumps(r), flush=True)
print("n", n, "peak", peak())
