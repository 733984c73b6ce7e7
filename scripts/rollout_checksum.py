"""One small warp-specialised rollout; prints a checksum of every output (for comparing library builds)."""
import hashlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brl_b200 import ops  # noqa: E402
from brl_b200.deals import synthetic_deal_table  # noqa: E402

dev, n, k = "cuda:0", 1024, 24
table = torch.as_tensor(synthetic_deal_table(1000, seed=1), device=dev)
state, out0 = ops.new_state(n, dev), ops.EnvOutputs(n, dev)
ops.init(ops.make_keys(7, n, dev), table, state, out0)
traj = ops.EnvOutputs(n, dev, rows=k)
act = torch.empty((k, n), dtype=torch.int32, device=dev)
stats = torch.zeros(4, dtype=torch.int64, device=dev)
ops.rollout_random(state, table, k, traj, seed=7, step0=0, action_out=act, stats=stats)
torch.cuda.synchronize()
h = hashlib.sha256()
for t in (traj.observation, traj.legal_action_mask, traj.rewards, traj.terminated, traj.current_player, act, state, stats):
    h.update(t.cpu().numpy().tobytes())
print("checksum", h.hexdigest()[:32], "finished_auctions", int(stats[0]))
