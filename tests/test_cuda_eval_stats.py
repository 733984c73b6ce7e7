"""GPU parity of the full evaluation statistics (src/evaluation.py:207-1032) against the NumPy
restatement in oracle/eval_ref.py: the GPU loop's traced logits drive the oracle env and the
restated log arithmetic, so every entry of `log_info` is compared (counts exactly, the
illegal-probability means to 1e-5 -- fp32 softmax on both sides, different summation order)."""
import numpy as np
import pytest
import torch

from tests import helpers as H

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _assert_log_info(got, want):
    assert len(got) == len(want)
    for k, (g, w) in enumerate(zip(got, want)):
        np.testing.assert_allclose(np.asarray(g, np.float64), np.asarray(w, np.float64), rtol=1e-5, atol=1e-7,
                                   err_msg=f"log_info[{k}]", equal_nan=True)


@pytest.mark.parametrize("game_mode", ["competitive", "free-run"])
def test_duplicate_evaluate_full_statistics(game_mode):
    from brl_b200 import BridgeBidding
    from brl_b200 import random as brandom
    from brl_b200.evaluation import make_evaluate, make_evaluate_log
    from brl_b200.models import init_params
    from oracle import eval_ref
    from oracle import oracle as orc
    boards = H.load_boards()
    env = BridgeBidding(table=boards["table"], device=DEV)
    n = 300
    p1, p2 = init_params(21, DEV), init_params(22, DEV)
    ev = make_evaluate(env, "relu", "DeepMind", "relu", "DeepMind", None, n, game_mode, duplicate=True, team2_params=p2)
    trace = []
    rng = brandom.PRNGKey(4)
    log_info, info_a, info_b = ev(p1, rng, trace=trace)
    _, sub = brandom.split(rng)
    ref = orc.OracleEnv(boards["table"], n)
    ref.init(env.make_keys(sub, n).cpu().numpy().view(np.uint64))
    ref.duplicate_tables_from_state()
    log = eval_ref.EvalLog(n, duplicate=True)
    cum = np.zeros(n, np.float32)
    for action, l1, l2 in trace:
        e = ref.export()
        want_a, probs = log.make_action(l1.cpu().numpy(), None if l2 is None else l2.cpu().numpy(),
                                        e["legal_action_mask"], e["current_player"])
        got_a = action.cpu().numpy()
        live = e["terminated"] == 0
        assert (want_a[live] == got_a[live]).all()
        log.step_log(probs, e["legal_action_mask"], e["current_player"], got_a, e["terminated"])
        ref.duplicate_step(got_a)
        cum += ref.export()["rewards"][:, 0]
    assert ref.export()["terminated"].all()
    priv = ref.export_private()
    ta = eval_ref.table_logs(ref.info_a["last_bid"], ref.info_a["last_bidder"], ref.info_a["call_x"], ref.info_a["call_xx"],
                             ref.info_a["rewards"][:, 0])
    tb = eval_ref.table_logs(ref.info_b["last_bid"], ref.info_b["last_bidder"], ref.info_b["call_x"], ref.info_b["call_xx"],
                             ref.info_b["rewards"][:, 0])
    want = eval_ref.log_info_duplicate(log, cum, priv["step_count"].astype(np.float64), ta, tb,
                                       ref.info_a["rewards"][:, 0], ref.info_b["rewards"][:, 0])
    _assert_log_info(log_info, want)
    d = make_evaluate_log(log_info)
    assert d["eval/IMP_reward"] == log_info[0] and "eval/actor_bid_probs/7NT" in d and len(d) == 19 + 4 * 35
    if game_mode == "free-run":
        assert log_info[4] == 0.0 and log_info[22] == 1.0      # the opponent always passes, legally


def test_single_table_evaluate_full_statistics():
    from brl_b200 import BridgeBidding
    from brl_b200 import random as brandom
    from brl_b200.evaluation import make_evaluate
    from brl_b200.models import init_params
    from oracle import eval_ref
    from oracle import oracle as orc
    boards = H.load_boards()
    env = BridgeBidding(table=boards["table"], device=DEV)
    n = 257
    p1, p2 = init_params(31, DEV), init_params(32, DEV)
    ev = make_evaluate(env, "relu", "DeepMind", "relu", "DeepMind", None, n, "competitive", duplicate=False, team2_params=p2)
    trace = []
    rng = brandom.PRNGKey(6)
    state, log_info = ev(p1, rng, trace=trace)
    _, sub = brandom.split(rng)
    ref = orc.OracleEnv(boards["table"], n)
    ref.init(env.make_keys(sub, n).cpu().numpy().view(np.uint64))
    log = eval_ref.EvalLog(n, duplicate=False)
    cum = np.zeros(n, np.float32)
    total_rewards = np.zeros((n, 4), np.float32)
    for action, l1, l2 in trace:
        e = ref.export()
        want_a, probs = log.make_action(l1.cpu().numpy(), l2.cpu().numpy(), e["legal_action_mask"], e["current_player"])
        got_a = action.cpu().numpy()
        live = e["terminated"] == 0
        assert (want_a[live] == got_a[live]).all()
        log.step_log(probs, e["legal_action_mask"], e["current_player"], got_a, e["terminated"])
        ref.step(got_a)
        r = ref.export()["rewards"]
        cum += r[:, 0]
        total_rewards += r
    priv = ref.export_private()
    t = eval_ref.table_logs(priv["last_bid"], priv["last_bidder"], priv["call_x"], priv["call_xx"], cum, priv["pass_num"])
    want = eval_ref.log_info_single(log, cum, priv["step_count"].astype(np.float64), t)
    _assert_log_info(log_info, want)
    assert (state.rewards.cpu().numpy() == total_rewards).all()
