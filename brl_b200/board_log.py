"""Board logs of a GPU duplicate match in the JSON format the reference's tooling reads
(SURVEY 8f-4): `bridge_env/data_handler/json_handler` (writer.py:85-152, parser.py:78-126,
README.md) -- the format `wb5/analyze_log.py:30-155` analyses and
`wb5/dataset_for_vs_wb5.json` uses for board settings.  One file per table:
table 1 = "table A" of `duplicate_step` (team 1 seated as dealt), table 2 = "table B"
(seats handed to the other team, src/duplicate.py:113-129).

Host-side format code only: the auction itself ran on the GPU; `record` is the per-step
(action, table_a.terminated, table_b.terminated) trace `make_simple_duplicate_evaluate(...,
record=[])` collects.
"""
from __future__ import annotations

import json
from typing import Dict, IO, List, Optional, Sequence

import numpy as np

from . import deals as _deals

SEATS = "NESW"
STRAINS = ("C", "D", "H", "S", "NT")
_RANKS = "23456789TJQKA"
_SUITS = "CDHS"
# action -> bid string (bid.py:101-104, 185-194): 0 Pass, 1 X, 2 XX, 3.. = 1C..7NT
ACTION_STR = ["Pass", "X", "XX"] + [f"{lvl}{s}" for lvl in range(1, 8) for s in STRAINS]


def split_tables(record) -> tuple:
    """record: per step (action[n], a_terminated_before[n], b_terminated_before[n]) as array-likes.
    -> (calls of table A, calls of table B): lists of n int lists.  A step belongs to table A while A has
    not finished, to table B afterwards until B finishes; later steps are no-ops (src/duplicate.py:147-192)."""
    n = len(np.asarray(record[0][0]))
    hist_a: List[List[int]] = [[] for _ in range(n)]
    hist_b: List[List[int]] = [[] for _ in range(n)]
    for action, a_term, b_term in record:
        action, a_term, b_term = (np.asarray(x.cpu() if hasattr(x, "cpu") else x) for x in (action, a_term, b_term))
        for i in np.nonzero(a_term == 0)[0]:
            hist_a[i].append(int(action[i]))
        for i in np.nonzero((a_term != 0) & (b_term == 0))[0]:
            hist_b[i].append(int(action[i]))
    return hist_a, hist_b


def resolve_contract(calls: Sequence[int], dealer: int):
    """(last_bid or None, x, xx, declarer seat or None) of a finished auction
    (bidding_phase.py:119-206: declarer = first of the declaring side to name the strain)."""
    last_bid, x, xx, bidder = None, False, False, None
    first_namer: Dict[tuple, int] = {}
    seat = dealer
    for a in calls:
        if a >= 3:
            last_bid, x, xx, bidder = a - 3, False, False, seat
            first_namer.setdefault((seat & 1, last_bid % 5), seat)
        elif a == 1:
            x = True
        elif a == 2:
            xx = True
        seat = (seat + 1) & 3
    if last_bid is None:
        return None, False, False, None
    return last_bid, x, xx, first_namer[(bidder & 1, last_bid % 5)]


def _card_str(c: int) -> str:
    return _SUITS[c % 4] + _RANKS[c // 4]


def _deal_json(owners_row: np.ndarray) -> Dict[str, List[str]]:
    """hands as sorted card strings "C2".."SA" (json_handler/writer.py:155-165, card.py order: suit, then rank)"""
    out = {}
    for seat, name in enumerate(SEATS):
        cards = np.nonzero(owners_row == seat)[0]
        out[name] = [_card_str(int(c)) for c in sorted(cards, key=lambda c: (c % 4, c // 4))]
    return out


def _vul_str(vul_ns: int, vul_ew: int) -> str:
    return ("None", "NS", "EW", "Both")[int(bool(vul_ns)) + 2 * int(bool(vul_ew))]   # vul.py:16-26


def board_log_entry(board_id, owners_row, dd_row, dealer: int, vul_ns: int, vul_ew: int, seat_names: Sequence[str],
                    calls: Sequence[int], score_fn) -> dict:
    """One item of the "logs" list (json_handler/writer.py:128-152).  `score_fn(bid, x, xx, vul, tricks)` is the
    duplicate score of a contract (score.py:50-106); taken_trick is the double-dummy trick count, as
    wb5/analyze_log.py:131-147 scores bidding-only logs."""
    last_bid, x, xx, declarer = resolve_contract(calls, dealer)
    if last_bid is None:
        contract, tricks, score_ns = "Passed_out", None, 0
    else:
        contract = ACTION_STR[last_bid + 3] + ("XX" if xx else ("X" if x else ""))
        tricks = int(dd_row[declarer][last_bid % 5])
        vul = bool(vul_ew) if declarer & 1 else bool(vul_ns)
        s = int(score_fn(last_bid, x, xx, vul, tricks))
        score_ns = s if declarer % 2 == 0 else -s
    return {
        "players": {name: seat_names[seat] for seat, name in enumerate(SEATS)},
        "board_id": board_id,
        "dealer": SEATS[dealer],
        "deal": _deal_json(owners_row),
        "vulnerability": _vul_str(vul_ns, vul_ew),
        "bid_history": [ACTION_STR[a] for a in calls],
        "contract": contract,
        "declarer": None if declarer is None else SEATS[declarer],
        "play_history": None,
        "taken_trick": tricks,
        "score_type": "IMP",
        "scores": {"NS": score_ns, "EW": -score_ns},
        "dda": {name: {s: int(dd_row[seat][k]) for k, s in enumerate(STRAINS)} for seat, name in enumerate(SEATS)},
    }


def write_logs(fp: IO[str], entries: Sequence[dict]) -> None:
    """{"logs": [ ... ]} with one board per line (json_handler/writer.py:27-47)."""
    fp.write('{"logs": [\n')
    fp.write(",\n".join(json.dumps(e, indent=None) for e in entries))
    fp.write("\n]}" if entries else "]}")


def match_to_board_logs(table: np.ndarray, deal, dealer, vul, shuffled_players, record, team_names=("team1", "team2"),
                        board_ids: Optional[Sequence] = None, score_fn=None):
    """-> (entries of table 1, entries of table 2) for a recorded duplicate match.
    `deal`, `dealer`, `vul[n,2]`, `shuffled_players[n,4]` are the private fields of the INITIAL state
    (ops.state_fields); player ids 0/1 are team 1, 2/3 team 2 (src/evaluation.py:148)."""
    if score_fn is None:
        score_fn = contract_score
    owners, dd = _deals.unpack_deal_table(table)
    hist_a, hist_b = split_tables(record)
    deal, dealer, vul, players = (np.asarray(x.cpu() if hasattr(x, "cpu") else x) for x in (deal, dealer, vul, shuffled_players))
    t1, t2 = [], []
    for i in range(len(hist_a)):
        row = int(deal[i])
        bid = board_ids[i] if board_ids is not None else i
        names_a = [team_names[int(players[i][s]) // 2] for s in range(4)]
        names_b = [team_names[1 - int(players[i][s]) // 2] for s in range(4)]     # seats handed to the other team
        args = (bid, owners[row], dd[row], int(dealer[i]), int(vul[i][0]), int(vul[i][1]))
        t1.append(board_log_entry(*args, names_a, hist_a[i], score_fn))
        t2.append(board_log_entry(*args, names_b, hist_b[i], score_fn))
    return t1, t2


def contract_score(bid: int, x: bool, xx: bool, vul: bool, tricks: int) -> int:
    """Duplicate score of `bid` (0..34) for the declaring side (score.py:5-106), host-side copy of the
    closed form the kernels use (csrc/env_device.cuh contract_score) for writing log files."""
    level, strain = bid // 5 + 1, bid % 5
    need = level + 6
    if need > tricks:
        n = need - tricks
        if not x and not xx:
            s = (100 if vul else 50) * n
        else:
            s = 300 * n - 100 if vul else (200 * n - 100 if n <= 3 else 300 * n - 400)
            if xx:
                s *= 2
        return -s
    over = tricks - need
    per = 20 if strain <= 1 else 30
    score = per * level + (10 if strain == 4 else 0)
    score *= 4 if xx else (2 if x else 1)
    if score >= 100:
        score += 450 if vul else 250
        if level >= 6:
            score += 750 if vul else 500
            if level == 7:
                score += 750 if vul else 500
    score += 50
    if xx:
        score += 100
        per = 400 if vul else 200
    elif x:
        score += 50
        per = 200 if vul else 100
    return score + per * over
