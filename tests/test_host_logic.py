"""CPU: host-side logic -- deal-table packing / formats, sharding, statistics, PRNG keys,
oracle self-consistency (auto-reset, duplicate, GAE, categorical) at small sizes."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from brl_b200 import deals, dist as bdist, random as brandom
from oracle import oracle as orc
from tests import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_deal_table_pack_roundtrip_and_layout():
    t = deals.synthetic_deal_table(500, seed=4)
    assert t.shape == (500, 48) and t.dtype == np.uint8 and (t[:, 42:] == 0).all()
    owners, dd = deals.unpack_deal_table(t)
    assert (deals.pack_deal_table(owners, dd) == t).all()
    masks = t[:, :32].copy().view(np.uint64).reshape(-1, 4)
    assert (np.bitwise_or.reduce(masks, axis=1) == (1 << 52) - 1).all()       # 52 cards, each owned once
    assert all(bin(int(m)).count("1") == 13 for m in masks[:50].ravel())
    assert 0.3 < ((dd[:, 0] != dd[:, 2]) | (dd[:, 1] != dd[:, 3])).any(axis=1).mean() < 0.6  # declarer-asymmetric boards


def test_pgx_dds_roundtrip_and_value_docstring_example():
    t = deals.synthetic_deal_table(200, seed=9)
    keys, values = deals.table_to_pgx_dds(t)
    assert keys.shape == (200, 4) and values.shape == (200, 4)
    assert (deals.pgx_dds_to_table(keys, values) == t).all()
    dd = deals.unpack_deal_table(deals.pgx_dds_to_table(keys[:1], np.array([[4160, 904605, 4160, 904605]])))[1][0]
    assert dd.tolist() == [[0, 1, 0, 4, 0], [13, 12, 13, 9, 13]] * 2


def test_pgx_key_matches_reference_pbn_to_key_convention(tmp_path):
    """wb5/vis_pgx.py:13-24: key digit = owner, suits S,H,D,C, ranks A,2..K, most significant first."""
    owners = np.zeros((1, 52), np.int8)
    # North: all spades; East: all hearts; South: all diamonds; West: all clubs
    for rank in range(13):
        owners[0, rank * 4 + 3], owners[0, rank * 4 + 2], owners[0, rank * 4 + 1], owners[0, rank * 4 + 0] = 0, 1, 2, 3
    t = deals.pack_deal_table(owners, np.zeros((1, 4, 5), np.int8))
    keys, _ = deals.table_to_pgx_dds(t)
    digits = lambda d: sum(d << (2 * k) for k in range(13))  # noqa: E731
    assert keys[0].tolist() == [digits(0), digits(1), digits(2), digits(3)]
    np.save(tmp_path / "dds.npy", np.stack(deals.table_to_pgx_dds(t)))
    assert (deals.load_table(str(tmp_path / "dds.npy")) == t).all()


def test_board_json_schema_conversion():
    boards = H.load_boards()
    owners, dd = deals.unpack_deal_table(boards["table"])
    # wb5/dataset_for_vs_wb5.json row 0: N holds C6 (rank 4, suit 0); dda N/C = 11, E/S = 8
    assert owners[0, 4 * 4 + 0] == 0 and dd[0, 0, 0] == 11 and dd[0, 1, 3] == 8
    assert boards["dealer"][0] == 2 and boards["vul_ns"][0] == 0 and boards["vul_ew"][0] == 1
    obj = {"logs": [{"board_id": 7, "dealer": "W", "vulnerability": "Both",
                     "deal": {"N": [s + r for s in "C" for r in "23456789TJQKA"], "E": [s + r for s in "D" for r in "23456789TJQKA"],
                              "S": [s + r for s in "H" for r in "23456789TJQKA"], "W": [s + r for s in "S" for r in "23456789TJQKA"]},
                     "dda": {p: {"C": 1, "D": 2, "H": 3, "S": 4, "NT": 5} for p in "NESW"}}]}
    t, dealer, vns, vew, bid = deals.boards_from_json(obj)
    assert dealer[0] == 3 and vns[0] == 1 and vew[0] == 1 and bid[0] == 7
    assert deals.unpack_deal_table(t)[1][0, 2].tolist() == [1, 2, 3, 4, 5]


def test_shard_ranges_cover_everything_once():
    for n, w in ((65536, 8), (100, 3), (7, 8), (1048576, 4)):
        spans = [bdist.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def test_stats_from_sums_equals_reference_formula():
    rng = np.random.default_rng(0)
    x = rng.integers(-24, 25, 1000).astype(np.float64)
    sums = [len(x), x.sum(), (x * x).sum(), (x > 0).sum(), 0, 0, 0, 0]
    mean, se, win = bdist.stats_from_sums(sums)
    assert mean == pytest.approx(x.mean()) and win == pytest.approx((x > 0).mean())
    assert se == pytest.approx(x.std(ddof=1) / np.sqrt(len(x)))  # src/evaluation.py:199
    assert orc.match_stats(x).tolist() == pytest.approx([mean, se, win])


def test_host_keys_split_is_deterministic_and_distinct():
    k = brandom.PRNGKey(0)
    a, b = brandom.split(k)
    assert (a, b) == brandom.split(k) and a != b and 0 <= a < 2 ** 64
    assert len(set(brandom.split(k, 64))) == 64


def test_philox_known_answer():
    # Random123 kat_vectors: philox4x32-10, counter = key = 0 and all-ones
    assert orc.philox([0, 0, 0, 0], [0, 0]).tolist() == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert orc.philox([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2).tolist() == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]


def test_oracle_episode_draws_are_valid_and_cover_all_seatings():
    seen = set()
    key = 12345
    for _ in range(400):
        key, (deal, dealer, vns, vew, seating) = orc.draw_episode(key, 1000)
        assert 0 <= deal < 1000 and 0 <= dealer < 4 and vns in (0, 1) and vew in (0, 1)
        p = orc.seating_to_players(seating).tolist()
        assert sorted(p) == [0, 1, 2, 3] and p[0] // 2 == p[2] // 2 and p[1] // 2 == p[3] // 2 and p[0] // 2 != p[1] // 2
        seen.add(tuple(p))
    assert seen == {tuple(r) for r in H.SEATINGS.tolist()}


def test_oracle_autoreset_keeps_flag_and_reward_but_swaps_episode():
    # src/utils.py:33-56
    boards = H.load_boards()
    env = orc.OracleEnv(boards["table"], 64)
    env.init(orc.make_keys(3, 64))
    first = env.export_private()
    seen_terminal = 0
    for s in range(40):
        env.step(env.random_legal_actions(1, s), autoreset=True)
        out, priv = env.export(), env.export_private()
        t = out["terminated"] == 1
        seen_terminal += int(t.sum())
        assert (priv["step_count"][t] == 0).all() and (priv["last_bid"][t] == -1).all()     # fresh episode
        assert (out["legal_action_mask"][t].sum(1) == 36).all()                              # init mask, not all-True
        assert (out["rewards"][~t] == 0).all()
    assert seen_terminal > 64 and (env.export_private()["rng_key"] != first["rng_key"]).any()


def test_oracle_duplicate_flow_known_board():
    """board 5000000, dealer S, vul EW: table A 6C by E doubled-history auction
    (-1000 E/W), table B 3NT by S (-150 N/S) => +15 IMPs for the team N/S at table A."""
    boards = H.load_boards()
    env = orc.OracleEnv(boards["table"], 1)
    env.reset_fields([0], [2], [0], [1], H.SEATINGS[[0]])  # N=0,E=2,S=1,W=3: team {0,1} sits N/S at table A
    env.duplicate_tables_from_state()
    total = 0.0
    for a in [0, 9, 11, 20, 1, 0, 22, 1, 2, 0, 0, 28, 0, 0, 0]:
        env.duplicate_step(np.array([a], np.int32))
        total += env.export()["rewards"][0, 0]
    assert env.info_a["terminated"][0] == 1 and env.info_a["rewards"][0].tolist() == [1000.0, 1000.0, -1000.0, -1000.0]
    assert env.export()["terminated"][0] == 0 and total == 0                     # table B started, rewards forced 0
    assert env.export_private()["shuffled_players"][0].tolist() == [2, 0, 3, 1]  # seats handed over
    for a in [7, 0, 17, 0, 0, 0]:  # S 1NT, W P, N 3NT, P P P  (dealer S)
        env.duplicate_step(np.array([a], np.int32))
        total += env.export()["rewards"][0, 0]
    assert env.info_b["terminated"][0] == 1 and env.info_b["rewards"][0, 0] == 150.0  # team {0,1} now E/W: +150
    assert total == 15.0 and env.export()["rewards"][0].tolist() == [15.0, 15.0, -15.0, -15.0]


def test_oracle_gae_against_float64_formula():
    rng = np.random.default_rng(5)
    T, n = 32, 50
    done = (rng.random((T, n)) < 0.2).astype(np.uint8)
    value = rng.normal(size=(T, n)).astype(np.float32)
    reward = rng.normal(size=(T, n)).astype(np.float32)
    last = rng.normal(size=n).astype(np.float32)
    adv, tgt = orc.gae(done, value, reward, last, 0.99, 0.95)
    g, nv = np.zeros(n), last.astype(np.float64)
    for t in range(T - 1, -1, -1):  # src/gae.py:20-39
        nd = 1.0 - done[t]
        g = reward[t] + 0.99 * nv * nd - value[t] + 0.99 * 0.95 * nd * g
        nv = value[t].astype(np.float64)
        np.testing.assert_allclose(adv[t], g, rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(tgt, adv + value, rtol=1e-6)


def test_oracle_threads_give_identical_results():
    t = deals.synthetic_deal_table(500, seed=1)
    outs = []
    for nt in (1, 4):
        env = orc.OracleEnv(t, 300, n_threads=nt)
        env.init(orc.make_keys(8, 300))
        outs.append(env.rollout_random(8, 0, 12))
    for k in ("observation", "legal_action_mask", "rewards", "terminated", "action"):
        assert (outs[0][k] == outs[1][k]).all()
    assert outs[0]["n_terminated"] == outs[1]["n_terminated"] > 0


def test_bench_reference_arm_prints_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "env_steps_per_sec" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
