"""Per-tile timeline of the fused PPO-update launches (tune bit 3 makes k_train_fused stamp %globaltimer)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brl_b200 import _lib, ops  # noqa: E402
from brl_b200.models import init_params  # noqa: E402
from brl_b200.optim import flatten_params  # noqa: E402

tune = (int(sys.argv[1]) if len(sys.argv) > 1 else 0) | 8
dev, B, total = "cuda:0", 1024, 8192
g = torch.Generator().manual_seed(0)
obs = (torch.rand((total, 480), generator=g) < 0.05).to(torch.bfloat16).to(dev)
mask = torch.ones((total, 38), dtype=torch.uint8, device=dev)
action = torch.zeros(total, dtype=torch.int32, device=dev)
z = torch.zeros(total, device=dev)
perm = torch.randperm(total, generator=g).to(torch.int32).to(dev)
flat_p, _ = flatten_params(init_params(1, dev))
blob = ops.mlp_pack_train(flat_p)
scratch = ops.mlp_train_scratch(B, dev)
flat_g = torch.empty_like(flat_p)
stats = torch.zeros(8, dtype=torch.float32, device=dev)
acc = ops.ppo_scratch(dev)
cfg = dict(clip_eps=0.2, ent_coef=0.01, vf_coef=0.5)
for i in range(4):
    ops.ppo_grad(obs, blob, scratch, perm[i * B:(i + 1) * B], mask, action, z - 1.0, z, z + 0.5, z + 0.1, flat_g, stats, acc, tune=tune, **cfg)
torch.cuda.synchronize()
off = _lib.load().brl_mlp_train_trace_offset(B)
tr = scratch[off: off + 2 * 4096 * 8 * 8].view(torch.int64).cpu().numpy().reshape(2, 4096, 8)
names = ["dep_wait", "load_issue", "load->mma0", "mma_issue", "acc_ready", "epilogue"]
for ph, label, tiles_per_op in ((0, "forward", None), (1, "backward", None)):
    t = tr[ph]
    used = t[:, 6] > 0
    n = int(used.sum())
    t = t[:n].astype(np.float64)
    t0 = t[:, 0].min()
    print(f"== {label}: {n} tiles, span {(t[:, 6].max() - t0) / 1e3:.1f} us")
    d = {"dep_wait": t[:, 1] - t[:, 0], "loads_issued": t[:, 2] - t[:, 1], "first_mma_after_dep": t[:, 3] - t[:, 1],
         "mma_span": t[:, 4] - t[:, 3], "acc_seen_after_last_issue": t[:, 5] - t[:, 4], "epilogue+publish": t[:, 6] - t[:, 5],
         "tile_total": t[:, 6] - t[:, 0]}
    for k, v in d.items():
        print(f"   {k:28s} mean {v.mean() / 1e3:7.2f} us  median {np.median(v) / 1e3:7.2f}  max {v.max() / 1e3:7.2f}")
    # per-op summary (ops are contiguous ranges of the list): start / end of each block of 128 tiles
    step = 128
    for s0 in range(0, n, step):
        blk = t[s0:s0 + step]
        print(f"   tiles {s0:4d}..{min(s0 + step, n) - 1:4d}: start {(blk[:, 0].min() - t0) / 1e3:6.1f}  dep_ok {(np.median(blk[:, 1]) - t0) / 1e3:6.1f}  "
              f"mma0 {(np.median(blk[:, 3]) - t0) / 1e3:6.1f}  last_issue {(np.median(blk[:, 4]) - t0) / 1e3:6.1f}  acc {(np.median(blk[:, 5]) - t0) / 1e3:6.1f}  done {(np.median(blk[:, 6]) - t0) / 1e3:6.1f} (max {(blk[:, 6].max() - t0) / 1e3:6.1f})")
