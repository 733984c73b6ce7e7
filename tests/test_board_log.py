"""CPU: board logs of a duplicate match (brl_b200/board_log.py) in the reference's JSON format.
A match is played on the oracle env with random-legal actions; the written files are then read
back (a) structurally, against the golden score table, and (b) -- when the reference tree is
mounted (this container, not the GPU box) -- by the reference's OWN `JsonParser`, `calc_score`
and `score_to_imp` following wb5/analyze_log.py:63-155, whose IMPs must equal the env's."""
import io
import json
import os
import sys

import numpy as np
import pytest

from brl_b200 import board_log
from brl_b200 import deals
from oracle import oracle as orc
from tests import helpers as H

REF_ENV = "/root/reference/submodule/bridge_env"


def _play_match(n=200, seed=5):
    boards = H.load_boards()
    env = orc.OracleEnv(boards["table"], n)
    env.init(orc.make_keys(seed, n))
    priv0 = env.export_private()
    env.duplicate_tables_from_state()
    record, cum = [], np.zeros(n)
    for step in range(400):
        a_term, b_term = env.info_a["terminated"].copy(), env.info_b["terminated"].copy()
        if env.export()["terminated"].all():
            break
        action = env.random_legal_actions(seed, step)
        record.append((action, a_term, b_term))
        env.duplicate_step(action)
        cum += env.export()["rewards"][:, 0]
    assert env.export()["terminated"].all()
    return boards, env, priv0, record, cum


def test_score_closed_form_matches_golden_table():
    tab = np.load(H.GOLDEN + "/score_table.npy")
    for b in range(35):
        for d, (x, xx) in enumerate(((0, 0), (1, 0), (1, 1))):
            for v in (0, 1):
                for t in range(14):
                    assert board_log.contract_score(b, bool(x), bool(xx), bool(v), t) == tab[b, d, v, t]


def test_known_auction_contract_and_strings():
    known = H.load_known()
    # wb5/utils.py:60-68 auction + 'Pass': 6C by E (tests/test_bidding_phase.py:28-68 shape)
    calls = [0, 9, 11, 20, 1, 0, 22, 1, 2, 0, 0, 28, 0, 0, 0]
    last_bid, x, xx, declarer = board_log.resolve_contract(calls, dealer=2)
    assert (last_bid, x, xx) == (25, False, False) and board_log.ACTION_STR[28] == "6C"
    assert board_log.SEATS[declarer] == "E"                      # 6C by E
    assert board_log.resolve_contract([0, 0, 0, 0], 1) == (None, False, False, None)
    assert [board_log.ACTION_STR[a] for a in (0, 1, 2, 3, 7, 37)] == ["Pass", "X", "XX", "1C", "1NT", "7NT"]
    assert known is not None


def test_match_logs_roundtrip_and_scores():
    boards, env, priv0, record, cum = _play_match()
    t1, t2 = board_log.match_to_board_logs(boards["table"], priv0["deal"], priv0["dealer"], priv0["vul"],
                                           priv0["shuffled_players"], record, team_names=("actor", "opp"))
    buf = io.StringIO()
    board_log.write_logs(buf, t1)
    back = json.loads(buf.getvalue())["logs"]
    assert back == t1 and len(t1) == len(t2) == 200
    owners, dd = deals.unpack_deal_table(boards["table"])
    for i, (e1, e2) in enumerate(zip(t1, t2)):
        # same board at both tables, seats swapped between the teams
        assert e1["deal"] == e2["deal"] and e1["dealer"] == e2["dealer"] and e1["vulnerability"] == e2["vulnerability"]
        assert all(e1["players"][s] != e2["players"][s] for s in "NESW")
        assert e1["players"]["N"] == e1["players"]["S"] != e1["players"]["E"] == e1["players"]["W"]
        assert sorted(sum(e1["deal"].values(), [])) == sorted(board_log._card_str(c) for c in range(52))
        # the env's own terminal info agrees with the log text
        for e, info in ((e1, env.info_a[i]), (e2, env.info_b[i])):
            if e["contract"] == "Passed_out":
                assert info["last_bid"] == -1 and e["declarer"] is None and e["taken_trick"] is None
            else:
                assert e["contract"].rstrip("X") == board_log.ACTION_STR[info["last_bid"] + 3]
                assert e["contract"].endswith("XX") == bool(info["call_xx"])
                ns_team = 0 if e["players"]["N"] == "actor" else 1
                want_ns = info["rewards"][0] if ns_team == 0 else info["rewards"][2]
                assert e["scores"]["NS"] == int(want_ns)
        # IMPs of team "actor" from the two logs == the env's duplicate reward (src/duplicate.py:46-70)
        s1 = e1["scores"]["NS"] if e1["players"]["N"] == "actor" else e1["scores"]["EW"]
        s2 = e2["scores"]["NS"] if e2["players"]["N"] == "actor" else e2["scores"]["EW"]
        assert orc.imp(s1 + s2) == int(cum[i])


@pytest.mark.skipif(not os.path.isdir(REF_ENV), reason="reference tree not mounted (GPU box)")
def test_reference_parser_and_analysis_accept_the_logs(tmp_path):
    boards, env, priv0, record, cum = _play_match(n=120, seed=8)
    t1, t2 = board_log.match_to_board_logs(boards["table"], priv0["deal"], priv0["dealer"], priv0["vul"],
                                           priv0["shuffled_players"], record, team_names=("actor", "opp"),
                                           board_ids=[str(i) for i in range(120)])
    for name, entries in (("t1.json", t1), ("t2.json", t2)):
        with open(tmp_path / name, "w") as fh:
            board_log.write_logs(fh, entries)
    sys.path.insert(0, REF_ENV)
    try:
        from bridge_env.data_handler.json_handler.parser import JsonParser
        from bridge_env.score import calc_score, score_to_imp
        from bridge_env import Pair
    finally:
        sys.path.remove(REF_ENV)
    parser = JsonParser()
    with open(tmp_path / "t1.json") as fh:
        data1 = parser.parse_board_logs(fh)
    with open(tmp_path / "t2.json") as fh:
        data2 = parser.parse_board_logs(fh)

    def dda_score(d):                                  # wb5/analyze_log.py:131-147
        if d.contract.is_passed_out():
            return 0
        return calc_score(d.contract, d.dda[d.declarer][d.contract.trump])

    for i, (d1, d2) in enumerate(zip(data1, data2)):
        s1, s2 = dda_score(d1), dda_score(d2)
        # wb5/analyze_log.py:86-95 orients both scores to table 1's N/S team; here to team "actor"
        from bridge_env import Player
        actor_pair_1 = Pair.NS if d1.players[Player.N] == "actor" else Pair.EW
        actor_pair_2 = Pair.NS if d2.players[Player.N] == "actor" else Pair.EW
        if d1.declarer is not None and d1.declarer.pair is not actor_pair_1:
            s1 = -s1
        if d2.declarer is not None and d2.declarer.pair is not actor_pair_2:
            s2 = -s2
        assert score_to_imp(s1, s2) == int(cum[i]), f"board {i}"
        assert [str(b) for b in d1.bid_history] == t1[i]["bid_history"]


@pytest.mark.gpu
def test_gpu_match_record_to_board_logs_and_reference_parser(tmp_path):
    """SURVEY 8f-4 on the GPU path: a duplicate match played by the CUDA kernels (`record=` of
    make_simple_duplicate_evaluate: per step the action and both tables' terminated flags), written as the reference's
    JSON board logs.  The logs' own scores must give the GPU match's IMPs board by board; where the reference tree is
    mounted its `JsonParser` + `calc_score` + `score_to_imp` (wb5/analyze_log.py:63-155) must agree too."""
    import torch
    from brl_b200 import BridgeBidding
    from brl_b200 import random as brandom
    from brl_b200.evaluation import make_simple_duplicate_evaluate
    from brl_b200.models import init_params
    dev = "cuda:0"
    boards = H.load_boards()
    env = BridgeBidding(table=boards["table"], device=dev)
    n = 150
    rng = brandom.PRNGKey(21)
    _, sub = brandom.split(rng)
    state0 = env.init(env.make_keys(sub, n))          # the same draws the evaluator makes: private fields of the start state
    priv = {k: getattr(state0, "_" + k).cpu().numpy() for k in ("deal", "dealer", "shuffled_players")}
    vul = np.stack([state0._vul_NS.cpu().numpy(), state0._vul_EW.cpu().numpy()], axis=1).astype(np.uint8)
    evaluate = make_simple_duplicate_evaluate(env, "relu", "DeepMind", "relu", "DeepMind", n)
    record = []
    (mean, se, win), info_a, info_b, cum = evaluate(init_params(5, dev), init_params(6, dev), rng, record=record)
    cum = cum.cpu().numpy()
    t1, t2 = board_log.match_to_board_logs(boards["table"], priv["deal"], priv["dealer"], vul, priv["shuffled_players"], record,
                                           team_names=("actor", "opp"), board_ids=[str(i) for i in range(n)])
    assert len(t1) == len(t2) == n
    for i, (e1, e2) in enumerate(zip(t1, t2)):
        s1 = e1["scores"]["NS"] if e1["players"]["N"] == "actor" else e1["scores"]["EW"]
        s2 = e2["scores"]["NS"] if e2["players"]["N"] == "actor" else e2["scores"]["EW"]
        assert orc.imp(s1 + s2) == int(cum[i]), f"board {i}"
        for e, info_bid in ((e1, int(info_a.last_bid[i])), (e2, int(info_b.last_bid[i]))):
            if e["contract"] == "Passed_out":
                assert info_bid == -1
            else:
                assert e["contract"].rstrip("X") == board_log.ACTION_STR[info_bid + 3]
    assert abs(mean - float(cum.astype(np.float64).mean())) < 1e-9
    if not os.path.isdir(REF_ENV):
        return  # GPU box: the structural check above is all that can run
    for name, entries in (("t1.json", t1), ("t2.json", t2)):
        with open(tmp_path / name, "w") as fh:
            board_log.write_logs(fh, entries)
    sys.path.insert(0, REF_ENV)
    try:
        from bridge_env import Pair, Player
        from bridge_env.data_handler.json_handler.parser import JsonParser
        from bridge_env.score import calc_score, score_to_imp
    finally:
        sys.path.remove(REF_ENV)
    parsed = []
    for name in ("t1.json", "t2.json"):
        with open(tmp_path / name) as fh:
            parsed.append(JsonParser().parse_board_logs(fh))

    def actor_score(d):
        if d.contract.is_passed_out():
            return 0
        s = calc_score(d.contract, d.dda[d.declarer][d.contract.trump])
        return s if d.declarer.pair is (Pair.NS if d.players[Player.N] == "actor" else Pair.EW) else -s

    for i, (d1, d2) in enumerate(zip(*parsed)):
        assert score_to_imp(actor_score(d1), actor_score(d2)) == int(cum[i]), f"board {i}"
