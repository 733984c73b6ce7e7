"""Deal / double-dummy (DD) tables in the packed row format the kernels read.

The reference draws every episode's deal from a table of pre-solved boards
(`BridgeBidding("dds_results/test_000.npy")`, eval.py:43; one 100k-row file per
env, ppo.py:127-128,297-308) and resolves the terminal score from the table's
DD tricks.  The DDS dataset cannot be downloaded offline, so tables here are
synthetic or converted from board JSON in the format of
`wb5/dataset_for_vs_wb5.json`.

Row layout (48 B, 16 B aligned; `include/brl_b200.h` BRL_DEAL_ROW_BYTES):
    bytes  0..31  u64 hand_mask[seat N,E,S,W]; bit c = card in OpenSpiel order
                  c = rank*4 + suit, rank 0='2'..12='A', suit 0=C,1=D,2=H,3=S
                  (the order of obs[428:480], wb5/utils.py:18-26)
    bytes 32..41  20 DD-trick nibbles, index seat*5 + strain (C,D,H,S,NT),
                  even index in the low nibble
    bytes 42..47  zero
"""
from __future__ import annotations

import json

import numpy as np

DEAL_ROW_BYTES = 48
SEATS = "NESW"
STRAINS = ("C", "D", "H", "S", "NT")
_RANKS = "23456789TJQKA"
_SUITS = "CDHS"


def pack_deal_table(owners: np.ndarray, dd: np.ndarray) -> np.ndarray:
    """owners: int[n,52] seat (0..3) of every card in OpenSpiel order;
    dd: int[n,4,5] tricks (0..13) for declarer seat x strain."""
    owners = np.asarray(owners)
    dd = np.asarray(dd)
    n = owners.shape[0]
    assert owners.shape == (n, 52) and dd.shape == (n, 4, 5)
    assert ((owners >= 0) & (owners < 4)).all() and ((dd >= 0) & (dd <= 13)).all()
    assert ((owners[:, :, None] == np.arange(4)).sum(axis=1) == 13).all(), "every seat holds 13 cards"
    table = np.zeros((n, DEAL_ROW_BYTES), dtype=np.uint8)
    bit = np.uint64(1) << np.arange(52, dtype=np.uint64)
    masks = np.zeros((n, 4), dtype=np.uint64)
    for seat in range(4):
        masks[:, seat] = ((owners == seat).astype(np.uint64) * bit).sum(axis=1, dtype=np.uint64)
    table[:, :32] = masks.view(np.uint8).reshape(n, 32)
    flat = dd.reshape(n, 20).astype(np.uint8)
    table[:, 32:42] = flat[:, 0::2] | (flat[:, 1::2] << 4)
    return table


def unpack_deal_table(table: np.ndarray):
    table = np.ascontiguousarray(table, dtype=np.uint8).reshape(-1, DEAL_ROW_BYTES)
    n = table.shape[0]
    masks = table[:, :32].copy().view(np.uint64).reshape(n, 4)
    owners = np.zeros((n, 52), dtype=np.int8)
    for seat in range(4):
        bits = (masks[:, seat, None] >> np.arange(52, dtype=np.uint64)) & np.uint64(1)
        owners[bits.astype(bool)] = seat
    nib = table[:, 32:42]
    dd = np.zeros((n, 20), dtype=np.int8)
    dd[:, 0::2] = nib & 15
    dd[:, 1::2] = nib >> 4
    return owners, dd.reshape(n, 4, 5)


def synthetic_deal_table(n_deals: int = 100_000, seed: int = 0) -> np.ndarray:
    """Uniform random deals + a plausible DD table: per strain
    t_N ~ clip(round(N(6.5,2.5))), t_S = t_N+eps, t_E = 13-max(t_N,t_S)-eta,
    t_W = t_E+eps' -- declarer-asymmetric on ~40% of boards like the real fixture."""
    rng = np.random.default_rng(seed)
    base = np.repeat(np.arange(4, dtype=np.int8), 13)
    owners = rng.permuted(np.tile(base, (n_deals, 1)), axis=1)
    t_n = np.clip(np.rint(rng.normal(6.5, 2.5, size=(n_deals, 5))), 0, 13).astype(np.int64)
    eps = rng.choice(np.array([-1, 0, 0, 0, 1]), size=(n_deals, 5))
    eps2 = rng.choice(np.array([-1, 0, 0, 0, 1]), size=(n_deals, 5))
    eta = rng.choice(np.array([0, 0, 1]), size=(n_deals, 5))
    t_s = np.clip(t_n + eps, 0, 13)
    t_e = np.clip(13 - np.maximum(t_n, t_s) - eta, 0, 13)
    t_w = np.clip(t_e + eps2, 0, 13)
    dd = np.stack([t_n, t_e, t_s, t_w], axis=1)
    return pack_deal_table(owners, dd)


def card_to_index(card: str) -> int:
    """'C6' / 'SA' (suit then rank, bridge_env card.py:100-112) -> OpenSpiel index."""
    return _RANKS.index(card[1]) * 4 + _SUITS.index(card[0])


def boards_from_json(path_or_obj):
    """Convert board logs in the `wb5/dataset_for_vs_wb5.json` schema
    ({"logs":[{board_id, dealer, deal{N,E,S,W}, vulnerability, dda{seat{strain}}}]})
    into (table u8[n,48], dealer i32[n], vul_ns u8[n], vul_ew u8[n], board_id i64[n])."""
    obj = path_or_obj
    if isinstance(path_or_obj, (str, bytes)):
        with open(path_or_obj, "r") as fh:
            obj = json.load(fh)
    logs = obj["logs"]
    n = len(logs)
    owners = np.zeros((n, 52), dtype=np.int8)
    dd = np.zeros((n, 4, 5), dtype=np.int8)
    dealer = np.zeros(n, dtype=np.int32)
    vul_ns = np.zeros(n, dtype=np.uint8)
    vul_ew = np.zeros(n, dtype=np.uint8)
    board_id = np.zeros(n, dtype=np.int64)
    for i, b in enumerate(logs):
        for seat, name in enumerate(SEATS):
            for card in b["deal"][name]:
                owners[i, card_to_index(card)] = seat
            for k, strain in enumerate(STRAINS):
                dd[i, seat, k] = b["dda"][name][strain]
        dealer[i] = SEATS.index(b["dealer"])
        v = b["vulnerability"]
        vul_ns[i] = v in ("NS", "Both", "All")
        vul_ew[i] = v in ("EW", "Both", "All")
        board_id[i] = b.get("board_id", i)
    return pack_deal_table(owners, dd), dealer, vul_ns, vul_ew, board_id
