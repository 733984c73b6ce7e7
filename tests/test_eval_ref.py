"""CPU: known-answer checks of the evaluation-statistics restatement (oracle/eval_ref.py) that the CUDA statistics kernels
are compared with, worked by hand from src/evaluation.py:236-596 -- the restatement needs jax to be run against the
original, so these hand cases are what pins it."""
import numpy as np

from oracle import eval_ref


def test_terminated_and_contract_log_by_hand():
    # src/evaluation.py:463-567.  env0: passed out; env1: actor team (player 1) plays bid 12 doubled and cum_return > 0;
    # env2: opponents (player 3) play bid 30 redoubled and cum_return > 0; env3: actor team (player 0) plays bid 0, cum_return < 0
    last_bid = np.array([-1, 12, 30, 0])
    last_bidder = np.array([-1, 1, 3, 0])
    call_x = np.array([0, 1, 1, 0], np.uint8)
    call_xx = np.array([0, 0, 1, 0], np.uint8)
    cum = np.array([0.0, 500.0, 200.0, -50.0], np.float32)
    pass_num = np.array([4, 3, 3, 3])
    t = eval_ref.table_logs(last_bid, last_bidder, call_x, call_xx, cum, pass_num)
    assert t["pass_out"].tolist() == [True, False, False, False]
    assert t["actor_contract"].sum() == 2 and t["actor_contract"][1, 12] == 1 and t["actor_contract"][3, 0] == 1
    assert t["opp_contract"].sum() == 1 and t["opp_contract"][2, 30] == 1
    assert t["actor_doubled"].tolist() == [False, True, False, False] and not t["actor_redoubled"].any()
    assert t["opp_doubled"].tolist() == [False, False, True, False] and t["opp_redoubled"].tolist() == [False, False, True, False]
    # :520-547 keys "make" on cum_return >= 0 for WHOEVER declared (cum_return is team 1's): env2 is logged as opp_make
    assert t["actor_make"].tolist() == [False, True, False, False] and t["opp_make"].tolist() == [False, False, True, False]
    assert t["actor_down"].tolist() == [False, False, False, True] and not t["opp_down"].any()
    # a state with last_bid == -1 but fewer than four passes is not a pass-out (:465-467)
    t2 = eval_ref.table_logs(np.array([-1]), np.array([-1]), np.zeros(1, np.uint8), np.zeros(1, np.uint8), np.zeros(1), np.array([2]))
    assert not t2["pass_out"][0]

    log = eval_ref.EvalLog(4, duplicate=False)
    log.steps[:] = 1.0   # avoid 0 / 0 in the ratios of this hand case
    info = eval_ref.log_info_single(log, cum, np.array([4.0, 9.0, 11.0, 6.0]), t)
    assert info[0] == cum.mean() and info[3] == 7.5
    assert info[8] == 0.5 and info[9] == 0.25            # declarer ratios: contracts / n
    assert info[10] == 0.25 and info[11] == 0.0 and info[12] == 0.25 and info[13] == 0.25
    assert info[14] == 0.25 and info[15] == 0.25 and info[16] == 0.25 and info[17] == 0.0 and info[18] == 0.25


def test_step_log_by_hand():
    # src/evaluation.py:311-385 / 673-745: two envs, three steps.  probs put 0.25 on one illegal action in env 0 at step 0.
    n = 2
    mask = np.ones((n, 38), bool)
    mask[0, 5] = False
    probs = np.full((n, 38), 0.75 / 37)
    probs[0, 5] = 0.25
    probs[1] = 1.0 / 38
    for duplicate in (False, True):
        log = eval_ref.EvalLog(n, duplicate)
        # step 0: both envs, team 1 to act (players 0 and 1); env 0 bids action 10 (bid 7), env 1 passes
        log.step_log(probs, mask, np.array([0, 1]), np.array([10, 0]), np.zeros(n, np.uint8))
        # step 1: team 2 to act (players 2, 3); env 0 passes, env 1 bids action 3 (bid 0)
        log.step_log(probs, mask, np.array([2, 3]), np.array([0, 3]), np.zeros(n, np.uint8))
        # step 2: team 1 again; env 0 repeats bid 7 -- env 1 has terminated and must not be logged
        log.step_log(probs, mask, np.array([0, 1]), np.array([10, 10]), np.array([0, 1], np.uint8))
        np.testing.assert_allclose(log.ill[0], [0.5, 0.0])      # team 1: env 0 twice 0.25; env 1 has no illegal action
        np.testing.assert_allclose(log.ill[1], [0.25, 0.0])
        assert log.steps[0].tolist() == [2, 1] and log.steps[1].tolist() == [1, 1]
        assert log.passes[0].tolist() == [0, 1] and log.passes[1].tolist() == [1, 0]
        assert log.bid[0][0, 7] == (2 if duplicate else 1)       # duplicate: += 1 (:699-706); single table: .set(1) (:345-354)
        assert log.bid[1][1, 0] == 1 and log.bid[0].sum() == (2 if duplicate else 1) and log.bid[1].sum() == 1


def test_make_action_is_masked_argmax_of_the_acting_team():
    # src/evaluation.py:236-252: logits of the team of current_player, masked argmax, probs of the UNMASKED logits
    log = eval_ref.EvalLog(3, duplicate=True)
    l1 = np.zeros((3, 38), np.float32); l2 = np.zeros((3, 38), np.float32)
    l1[:, 7] = 5.0; l2[:, 9] = 5.0
    mask = np.ones((3, 38), np.uint8); mask[2, 9] = 0
    action, probs = log.make_action(l1, l2, mask, np.array([0, 2, 3]))
    assert action.tolist() == [7, 9, 0]                          # env 2: team 2's favourite is illegal -> first of the rest
    assert probs[1].argmax() == 9 and abs(probs.sum(1) - 1).max() < 1e-6
    # free-run opponent: always Pass with a one-hot distribution (:248-252)
    action, probs = log.make_action(l1, None, mask, np.array([0, 2, 3]))
    assert action.tolist() == [7, 0, 0] and probs[1, 0] == 1.0
