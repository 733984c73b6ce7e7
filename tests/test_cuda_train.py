"""GPU: the ppo.py training loop mirror (brl_b200/ppo.py) end to end on a tiny configuration --
rollout -> GAE -> update, strength probes, PFSP league, DDS-table rotation, model save/load."""
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_train_loop_runs_and_rotates_tables(tmp_path):
    from brl_b200 import ppo, random as brandom
    from brl_b200.deals import synthetic_deal_table
    from brl_b200.models import load_params, params_to_numpy
    tables = [synthetic_deal_table(300, seed=k) for k in range(3)]
    config = dict(num_envs=128, num_steps=4, minibatch_size=128, update_epochs=2, total_timesteps=128 * 4 * 3,
                  hash_size=20, num_eval_envs=64, num_prioritized_envs=32, num_eval_step=2, ratio_model_zoo=1.0,
                  prioritized_fictitious=True, lr=1e-4, save_model=True, log_path=str(tmp_path), exp_name="t")
    seen = []
    runner, logs = ppo.train(config, brandom.PRNGKey(0), tables=tables, eval_table=synthetic_deal_table(200, seed=9),
                             device=DEV, log_fn=seen.append)
    assert len(logs) == 3 and seen == logs
    for log in logs:
        for k in ("train/score", "train/total_loss", "train/value_loss", "train/policy_entropy", "train/approx_kl",
                  "train/imp_opp_before", "train/imp_opp_after"):
            assert math.isfinite(log[k]), (k, log[k])
        assert log["train/lr"] == pytest.approx(1e-4)
    assert "eval/IMP_reward" in logs[0] and "eval/IMP_reward" not in logs[1] and "eval/actor_bid_probs/1C" in logs[2]
    assert logs[-1]["steps"] == 128 * 4 * 3 and logs[-1]["board_num"] > 20
    assert any("table" in log for log in logs), "hash_size=20 boards must trigger a table rotation"
    assert runner[1].count == 3 * 2 * 4   # iterations x epochs x minibatches optimizer steps
    # saved models round-trip through the reference's pickle layout
    saved = sorted(os.listdir(os.path.join(tmp_path, "t", "rl_params")))
    assert saved == ["opt_state-00000003.pkl", "params-00000001.pkl", "params-00000002.pkl", "params-00000003.pkl"]
    final = load_params(os.path.join(tmp_path, "t", "rl_params", "params-00000003.pkl"), DEV)
    a, b = params_to_numpy(final), params_to_numpy(runner[0])
    assert all((a[k] == b[k]).all() for k in a)
    first = params_to_numpy(load_params(os.path.join(tmp_path, "t", "rl_params", "params-00000001.pkl"), DEV))
    assert any((first[k] != b[k]).any() for k in first), "training must move the parameters"


def test_pfsp_probabilities():
    from brl_b200.ppo import pfsp_probabilities
    p = pfsp_probabilities(np.array([1.0, 0.0, -2.0]), 0.1)
    assert p.argmax() == 2 and abs(p.sum() - 1) < 1e-12   # the opponent we lose most IMPs to is sampled most
