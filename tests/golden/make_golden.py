"""Generate the golden fixtures in this directory FROM THE REFERENCE'S OWN PYTHON.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):   python tests/golden/make_golden.py

What is executed is reference code, unmodified, imported from where it lies:
  * submodule/bridge_env/bridge_env  BiddingPhase.take_bid / available_bid /
    contract (bidding_phase.py:119-206), calc_score (score.py:109-125),
    score_to_imp (score.py:140-146)
  * wb5/utils.py  convert_obs / convert_leagal_action_mask (lines 11-52); its
    module top imports jax + pgx, which are absent, so empty stub modules are
    registered for those two names before the import (no reference code changed).
Outputs (committed, small):
  score_table.npy        int32[35,3,2,14]   calc_bid_score for every (bid, none/X/XX, vul, tricks)
  imp_table.npy          int32[1601]        score_to_imp(d,0) for d = -8000..8000 step 10
  auctions.npz           2000 auctions over the 1000 real boards: calls, per-call
                         obs bits + mask of the player to act, declarer, score
  boards_wb5_1000.npz    the 1000 boards of wb5/dataset_for_vs_wb5.json as a packed
                         deal table (brl_b200/deals.py row format) + dealer/vul
  known_auctions.json    the two fixed auctions the reference tests use
                         (wb5/utils.py:60-68; tests/test_bidding_phase.py:28-68)
"""
import json
import os
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

# -- stubs for the two absent imports at the top of wb5/utils.py ---------------
for name in ("jax", "jax.numpy", "pgx", "pgx.bridge_bidding"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["jax"].numpy = sys.modules["jax.numpy"]
sys.modules["pgx.bridge_bidding"].BridgeBidding = object
sys.modules["pgx.bridge_bidding"]._state_to_pbn = None
sys.path.insert(0, REF)

from submodule.bridge_env.bridge_env import Bid, BiddingPhase, Hands, Player, Vul  # noqa: E402
from submodule.bridge_env.bridge_env.bidding_phase import BiddingPhaseState  # noqa: E402
from submodule.bridge_env.bridge_env.card import Card  # noqa: E402
from submodule.bridge_env.bridge_env.score import calc_bid_score, calc_score, score_to_imp  # noqa: E402
from wb5.utils import convert_act_pgx2be, convert_leagal_action_mask, convert_obs  # noqa: E402

from brl_b200.deals import boards_from_json  # noqa: E402  (format conversion only)

SEATS = "NESW"
STRAINS = ("C", "D", "H", "S", "NT")


def score_table():
    tab = np.zeros((35, 3, 2, 14), dtype=np.int32)
    for b in range(35):
        for d, (x, xx) in enumerate(((False, False), (True, False), (True, True))):
            for v in (0, 1):
                for t in range(14):
                    tab[b, d, v, t] = calc_bid_score(Bid(b + 1), x=x, xx=xx, vul=bool(v), taken_trick_num=t)
    return tab


def imp_table():
    return np.array([score_to_imp(d, 0) for d in range(-8000, 8001, 10)], dtype=np.int32)


def vul_of(ns, ew):
    return Vul(1 + int(ns) + 2 * int(ew))  # NONE=1, NS=2, EW=3, BOTH=4 (vul.py:9-12)


def hands_of(board):
    hs = {}
    for name in SEATS:
        hs[name] = {Card.str_to_card(c) for c in board["deal"][name]}
    return Hands(north_hand=hs["N"], east_hand=hs["E"], south_hand=hs["S"], west_hand=hs["W"])


def play_auction(board, dealer_seat, vul_ns, vul_ew, rng, pass_bias, forced=None):
    """Random-legal auction through the reference BiddingPhase; records what the
    player to act sees before every call."""
    bp = BiddingPhase(dealer=Player(dealer_seat + 1), vul=vul_of(vul_ns, vul_ew))
    hands = hands_of(board).to_binary()
    calls, obs_rows, mask_rows = [], [], []
    k = 0
    while not bp.has_done():
        mask = convert_leagal_action_mask(bp.available_bid).astype(np.uint8)
        obs = convert_obs(dealer=bp.dealer, vul=bp.vul, active_player=bp.active_player,
                          bid_history=bp.bid_history, hand=hands[bp.active_player])
        if forced is not None:
            a = forced[k]
        elif pass_bias > 0 and rng.random() < pass_bias:
            a = 0
        else:
            legal = np.flatnonzero(mask)
            a = int(legal[rng.integers(len(legal))])
            if pass_bias > 0 and a >= 3:  # keep realistic levels: prefer the cheapest few bids
                bids = legal[legal >= 3]
                a = int(bids[min(len(bids) - 1, int(rng.integers(0, 6)))])
        st = bp.take_bid(convert_act_pgx2be(a))
        assert st is not BiddingPhaseState.ILLEGAL
        calls.append(a)
        obs_rows.append(np.packbits(obs.astype(np.uint8), bitorder="little"))
        mask_rows.append(mask)
        k += 1
    contract = bp.contract()
    if contract.is_passed_out():
        return calls, obs_rows, mask_rows, (-1, -1, 0, 0, 0, 0)
    declarer = contract.declarer
    tricks = board["dda"][str(declarer)][str(contract.final_bid.suit)]
    sc = calc_score(contract, tricks)
    return calls, obs_rows, mask_rows, (declarer.value - 1, contract.final_bid.idx, int(contract.x),
                                        int(contract.xx), int(contract.is_vul()), sc)


def main():
    np.save(os.path.join(HERE, "score_table.npy"), score_table())
    np.save(os.path.join(HERE, "imp_table.npy"), imp_table())

    with open(os.path.join(REF, "wb5", "dataset_for_vs_wb5.json")) as fh:
        data = json.load(fh)
    table, dealer, vul_ns, vul_ew, board_id = boards_from_json(data)
    np.savez_compressed(os.path.join(HERE, "boards_wb5_1000.npz"), table=table, dealer=dealer,
                        vul_ns=vul_ns, vul_ew=vul_ew, board_id=board_id)

    rng = np.random.default_rng(20241017)
    rec = dict(board=[], dealer=[], vul_ns=[], vul_ew=[], offsets=[0], calls=[], obs_bits=[], mask=[], final=[])
    for i, board in enumerate(data["logs"]):
        for variant in range(2):
            if variant == 0:  # the board's own dealer / vulnerability, uniform random-legal
                d, vn, ve, bias = int(dealer[i]), int(vul_ns[i]), int(vul_ew[i]), 0.0
            else:             # random dealer / vulnerability, pass-biased (realistic contracts, pass-outs)
                d, vn, ve, bias = int(rng.integers(4)), int(rng.integers(2)), int(rng.integers(2)), 0.62
            calls, obs_rows, mask_rows, final = play_auction(board, d, vn, ve, rng, bias)
            rec["board"].append(i); rec["dealer"].append(d); rec["vul_ns"].append(vn); rec["vul_ew"].append(ve)
            rec["calls"].extend(calls); rec["obs_bits"].extend(obs_rows); rec["mask"].extend(mask_rows)
            rec["offsets"].append(len(rec["calls"])); rec["final"].append(final)
    np.savez_compressed(
        os.path.join(HERE, "auctions.npz"),
        board=np.array(rec["board"], np.int32), dealer=np.array(rec["dealer"], np.int8),
        vul_ns=np.array(rec["vul_ns"], np.uint8), vul_ew=np.array(rec["vul_ew"], np.uint8),
        offsets=np.array(rec["offsets"], np.int32), calls=np.array(rec["calls"], np.int8),
        obs_bits=np.stack(rec["obs_bits"]).astype(np.uint8), mask=np.stack(rec["mask"]).astype(np.uint8),
        final=np.array(rec["final"], np.int32))  # declarer seat, bid idx, x, xx, vul, declarer score

    # fixed auctions the reference itself uses
    known = {}
    b0 = data["logs"][0]
    calls, obs_rows, mask_rows, final = play_auction(  # wb5/utils.py:60-68: dealer E, vul None
        b0, 1, 0, 0, rng, 0.0, forced=[0, 9, 11, 20, 1, 0, 22, 1, 2, 0, 0, 28, 0, 0, 0])
    known["wb5_utils_14call_plus_final_pass"] = dict(
        board=0, dealer=1, vul_ns=0, vul_ew=0, calls=calls, final=list(final),
        set_bits=[np.flatnonzero(np.unpackbits(r, bitorder="little")[:480]).tolist() for r in obs_rows],
        n_legal=[int(m.sum()) for m in mask_rows])
    # tests/test_bidding_phase.py:28-68: same calls, dealer S, vul NS; 5S and XX are illegal
    # before the final pass; contract 6C by E, not vulnerable.
    calls, obs_rows, mask_rows, final = play_auction(
        b0, 2, 1, 0, rng, 0.0, forced=[0, 9, 11, 20, 1, 0, 22, 1, 2, 0, 0, 28, 0, 0, 0])
    assert mask_rows[14][26] == 0 and mask_rows[14][2] == 0 and final[0] == 1 and final[1] == 25 and final[4] == 0
    known["test_bidding_phase1"] = dict(
        board=0, dealer=2, vul_ns=1, vul_ew=0, calls=calls, final=list(final),
        illegal_before_final_pass=[26, 2], n_legal=[int(m.sum()) for m in mask_rows])
    with open(os.path.join(HERE, "known_auctions.json"), "w") as fh:
        json.dump(known, fh, indent=1)
    print("wrote golden fixtures:", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
