"""GPU parity: the CUDA env path (through the C ABI) against the CPU oracle and the
golden vectors from the reference's own Python.  Bit-exact: observations, legal masks,
terminal flags, rewards (integer scores / IMPs), current player, private fields."""
import numpy as np
import pytest
import torch

from tests import helpers as H

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _orc():
    from oracle import oracle as orc
    return orc


def _ops():
    from brl_b200 import ops
    return ops


def _t(a, dtype=None):
    return torch.as_tensor(np.ascontiguousarray(a), device=DEV) if dtype is None else torch.as_tensor(
        np.ascontiguousarray(a).astype(dtype), device=DEV)


def _cmp_outputs(out, ref, where=None, tag=""):
    sel = slice(None) if where is None else where
    got_obs = out.observation.float().cpu().numpy()
    assert (got_obs[sel] == ref["observation"][sel]).all(), f"{tag}: observation"
    assert (out.legal_action_mask.cpu().numpy()[sel] == ref["legal_action_mask"][sel]).all(), f"{tag}: mask"
    assert (out.rewards.cpu().numpy()[sel] == ref["rewards"][sel]).all(), f"{tag}: rewards"
    assert (out.terminated.cpu().numpy()[sel] == ref["terminated"][sel]).all(), f"{tag}: terminated"
    assert (out.current_player.cpu().numpy()[sel] == ref["current_player"][sel]).all(), f"{tag}: current_player"


def _reset(ops, table_t, deal, dealer, vns, vew, players, obs_dtype=torch.float32):
    n = len(deal)
    state = ops.new_state(n, DEV)
    out = ops.EnvOutputs(n, DEV, obs_dtype)
    ops.reset_fields(_t(deal, np.int32), _t(dealer, np.int32), _t(vns, np.uint8), _t(vew, np.uint8),
                     _t(players, np.int8), None, table_t, state, out)
    return state, out


def test_golden_auctions_through_cuda():
    """19,174 (obs, mask) pairs + 2000 final contracts produced by the reference's own
    BiddingPhase / convert_obs / calc_score (tests/golden/make_golden.py)."""
    ops = _ops()
    boards, gold = H.load_boards(), H.load_auctions()
    calls, lens = H.auction_matrix(gold)
    n = len(lens)
    players = H.SEATINGS[np.arange(n) % 8]
    table_t = _t(boards["table"])
    state, out = _reset(ops, table_t, gold["board"], gold["dealer"], gold["vul_ns"], gold["vul_ew"], players, torch.uint8)
    want_obs = H.unpack_obs_bits(gold["obs_bits"])
    off = gold["offsets"]
    total_r = np.zeros((n, 4), np.float32)
    for k in range(calls.shape[1]):
        live = np.flatnonzero(lens > k)
        rows = off[live] + k
        assert (out.observation.cpu().numpy()[live] == want_obs[rows]).all(), f"obs at call {k}"
        assert (out.legal_action_mask.cpu().numpy()[live] == gold["mask"][rows]).all(), f"mask at call {k}"
        seat = (gold["dealer"][live].astype(int) + k) % 4
        assert (out.current_player.cpu().numpy()[live] == players[live, seat]).all()
        act = _t(np.where(lens > k, calls[:, k], 0), np.int32)
        ops.step(state, act, table_t, state, out)
        assert (out.terminated.cpu().numpy() == (lens <= k + 1)).all()
        total_r += out.rewards.cpu().numpy()
    want_r = np.stack([H.expected_rewards(gold["final"][i], players[i]) for i in range(n)])
    assert (total_r == want_r).all()
    f = ops.state_fields(state)
    assert (f["last_bid"].cpu().numpy() == gold["final"][:, 1]).all()
    has = gold["final"][:, 1] >= 0
    assert (f["call_x"].cpu().numpy()[has] == gold["final"][has, 2]).all()
    assert (f["call_xx"].cpu().numpy()[has] == gold["final"][has, 3]).all()
    assert (f["step_count"].cpu().numpy() == lens).all()


def test_score_table_exhaustive_through_step_kernel():
    """Every (bid, -/X/XX, vul, tricks) of submodule/bridge_env/tests/test_score.py:130-163,
    resolved by the terminal branch of the step kernel from a crafted DD table."""
    ops = _ops()
    from brl_b200.deals import pack_deal_table
    tab = np.load(H.GOLDEN + "/score_table.npy")
    owners = np.tile(np.repeat(np.arange(4), 13), (14, 1))
    dd = np.stack([np.full((4, 5), t) for t in range(14)])
    table_t = _t(pack_deal_table(owners, dd))
    cases = [(b, d, v, t) for b in range(35) for d in range(3) for v in range(2) for t in range(14)]
    n = len(cases)
    c = np.array(cases)
    players = np.tile(H.SEATINGS[0], (n, 1))
    # dealer N (player 0, N/S) bids; vul flag on N/S
    state, out = _reset(ops, table_t, c[:, 3], np.zeros(n), c[:, 2], np.zeros(n), players)
    seq = [c[:, 0] + 3, np.where(c[:, 1] >= 1, 1, 0), np.where(c[:, 1] == 2, 2, 0)]
    total = np.zeros((n, 4), np.float32)
    # N bids; E doubles (or passes); S redoubles (or passes); then passes until everything ended
    for k in range(7):
        a = seq[k] if k < 3 else np.zeros(n)
        ops.step(state, _t(a, np.int32), table_t, state, out)
        total += out.rewards.cpu().numpy()
    assert out.terminated.cpu().numpy().all()
    want = tab[c[:, 0], c[:, 1], c[:, 2], c[:, 3]].astype(np.float32)
    assert (total[:, 0] == want).all() and (total[:, 1] == want).all()
    assert (total[:, 2] == -want).all() and (total[:, 3] == -want).all()


@pytest.mark.parametrize("n,obs_dtype,epw", [(4096, torch.float32, 0), (1000, torch.uint8, 8), (777, torch.bfloat16, 16),
                                             (33, torch.float32, 32), (1, torch.float32, 8)])
def test_random_play_autoreset_matches_oracle(n, obs_dtype, epw):
    """init -> many auto-reset steps with random-legal actions chosen by the ORACLE,
    every Env-surface output compared bit-exactly after every step."""
    ops, orc = _ops(), _orc()
    from brl_b200 import _lib
    from brl_b200.deals import synthetic_deal_table
    table = synthetic_deal_table(5000, seed=3)
    table_t = _t(table)
    seed = 1234 + n
    keys = orc.make_keys(seed, n)
    env = orc.OracleEnv(table, n)
    env.init(keys)
    state = ops.new_state(n, DEV)
    out = ops.EnvOutputs(n, DEV, obs_dtype)
    keys_t = ops.make_keys(seed, n, DEV)
    assert (keys_t.cpu().numpy().view(np.uint64) == keys).all()
    tune = _lib.tune(epw=epw)
    ops.init(keys_t, table_t, state, out, tune=tune)
    _cmp_outputs(out, env.export(), tag="init")
    steps = 60 if n >= 1000 else 120
    n_term = 0
    for s in range(steps):
        act = env.random_legal_actions(seed, s)
        if s % 17 == 5:  # sprinkle illegal calls: X / XX where not allowed, out-of-range
            act = act.copy()
            act[::7] = 2
        env.step(act, autoreset=True)
        ops.step(state, _t(act), table_t, state, out, autoreset=True, tune=tune)
        ref = env.export()
        _cmp_outputs(out, ref, tag=f"step {s}")
        n_term += int(ref["terminated"].sum())
    assert n_term > 0
    f, p = ops.state_fields(state), env.export_private()
    for name in ("deal", "dealer", "shuffled_players", "vul", "last_bid", "last_bidder", "call_x", "call_xx", "pass_num",
                 "step_count"):
        assert (f[name].cpu().numpy() == p[name]).all(), name
    assert (f["rng_key"].cpu().numpy().view(np.uint64) == p["rng_key"]).all()


def test_plain_step_terminal_is_absorbing_and_noop():
    ops, orc = _ops(), _orc()
    from brl_b200.deals import synthetic_deal_table
    n = 512
    table = synthetic_deal_table(300, seed=5)
    table_t = _t(table)
    keys = orc.make_keys(9, n)
    env = orc.OracleEnv(table, n)
    env.init(keys)
    state, out = ops.new_state(n, DEV), ops.EnvOutputs(n, DEV)
    ops.init(_t(keys.view(np.int64)), table_t, state, out)
    for s in range(45):  # no auto-reset: every env ends and then no-ops with zero rewards
        act = env.random_legal_actions(77, s)
        env.step(act)
        ops.step(state, _t(act), table_t, state, out)
        _cmp_outputs(out, env.export(), tag=f"step {s}")
    assert out.terminated.all()
    assert (out.legal_action_mask == 1).all()  # pgx: all-True mask at a terminal
    assert (out.rewards == 0).all()


@pytest.mark.parametrize("n,k,tune_kw", [
    (2048, 24, {}), (100, 40, {"writers": 1}), (1000, 9, {"writers": 3, "epw": 32}), (33, 5, {"writers": 5, "epw": 16}),
    (1, 3, {}), (70000, 3, {}), (777, 6, {"epw": 8, "writers": 2}),
    (2048, 24, {"classic_rollout": True}), (100, 40, {"classic_rollout": True, "epw": 32}),
    (8, 7, {"classic_rollout": True, "epw": 8}),
    # balanced env split (grid = multiple of 148 SMs, ragged 27/28-env blocks, unaligned mask runs)
    (8192, 6, {}), (8192, 4, {"balanced": False}), (5000, 7, {"balanced": True}), (1001, 5, {"balanced": True, "epw": 16}),
    (37, 4, {"balanced": True}), (8192, 40, {"writers": 7}), (70000, 3, {"writers": 5, "epw": 16})])
def test_fused_rollout_kernel_matches_oracle(n, k, tune_kw):
    """brl_rollout_random (K steps in one launch, in-kernel random-legal policy) against
    the oracle's rollout: whole [K, n, ...] trajectories bit-exact."""
    ops, orc = _ops(), _orc()
    from brl_b200 import _lib
    from brl_b200.deals import synthetic_deal_table
    table = synthetic_deal_table(2000, seed=11)
    table_t = _t(table)
    seed, offset = 4242, 1000
    keys = orc.make_keys(seed, n, offset)
    env = orc.OracleEnv(table, n)
    env.init(keys)
    ref = env.rollout_random(seed, 5, k, env_offset=offset)
    state, out0 = ops.new_state(n, DEV), ops.EnvOutputs(n, DEV)
    ops.init(ops.make_keys(seed, n, DEV, env_offset=offset), table_t, state, out0)
    traj = ops.EnvOutputs(n, DEV, rows=k)
    actions = torch.empty((k, n), dtype=torch.int32, device=DEV)
    stats = torch.zeros(4, dtype=torch.int64, device=DEV)
    ops.rollout_random(state, table_t, k, traj, seed=seed, step0=5, env_offset=offset, action_out=actions, stats=stats,
                       tune=_lib.tune(**tune_kw))
    assert (actions.cpu().numpy() == ref["action"]).all()
    _cmp_outputs(traj, ref, tag="trajectory")
    st = stats.cpu().numpy()
    assert st[0] == ref["n_terminated"] and st[2] == n * k
    assert st[1] == int(ref["rewards"][:, :, 0].astype(np.int64).sum())
    # final packed state equals the oracle's
    f, p = ops.state_fields(state), env.export_private()
    for name in ("deal", "dealer", "last_bid", "pass_num", "step_count"):
        assert (f[name].cpu().numpy() == p[name]).all(), name


def test_duplicate_step_matches_oracle():
    """src/duplicate.py:147-192 -- table A, seat swap, table B, IMP rewards, Table_info."""
    ops, orc = _ops(), _orc()
    boards = H.load_boards()
    n = 1000
    table_t = _t(boards["table"])
    players = H.SEATINGS[np.arange(n) % 8]
    deal = np.arange(n)
    env = orc.OracleEnv(boards["table"], n)
    env.reset_fields(deal, boards["dealer"], boards["vul_ns"], boards["vul_ew"], players)
    env.duplicate_tables_from_state()
    state, out = _reset(ops, table_t, deal, boards["dealer"], boards["vul_ns"], boards["vul_ew"], players)
    ia, ib = ops.TableInfoBuffers(n, DEV), ops.TableInfoBuffers(n, DEV)
    cum = np.zeros(n)
    for s in range(64):
        act = env.random_legal_actions(31337, s)
        env.duplicate_step(act)
        ops.duplicate_step(state, _t(act), table_t, ia, ib, state, out)
        ref = env.export()
        _cmp_outputs(out, ref, tag=f"dup step {s}")
        for buf, info in ((ia, env.info_a), (ib, env.info_b)):
            assert (buf.terminated.cpu().numpy() == info["terminated"]).all()
            assert (buf.rewards.cpu().numpy() == info["rewards"]).all()
            done = info["terminated"] == 1
            for name in ("last_bid", "last_bidder", "call_x", "call_xx"):
                assert (getattr(buf, name).cpu().numpy()[done] == info[name][done]).all(), name
        cum += ref["rewards"][:, 0]
        if ref["terminated"].all():
            break
    assert ref["terminated"].all() and (np.abs(cum) <= 24).all() and (cum != 0).any()


def test_duplicate_init_and_observe_any_player():
    ops, orc = _ops(), _orc()
    boards = H.load_boards()
    n = 256
    table_t = _t(boards["table"])
    players = H.SEATINGS[np.arange(n) % 8]
    deal = np.arange(n) + 100
    env = orc.OracleEnv(boards["table"], n)
    env.reset_fields(deal, boards["dealer"][:n], boards["vul_ns"][:n], boards["vul_ew"][:n], players)
    state, out = _reset(ops, table_t, deal, boards["dealer"][:n], boards["vul_ns"][:n], boards["vul_ew"][:n], players)
    for s in range(6):
        act = env.random_legal_actions(5, s)
        env.step(act)
        ops.step(state, _t(act), table_t, state, out)
    for pid in range(4):  # _observe(state, player) for every player id
        obs = torch.empty((n, 480), dtype=torch.uint8, device=DEV)
        ops.observe(state, table_t, obs, _t(np.full(n, pid), np.int8))
        assert (obs.cpu().numpy() == env.observe(np.full(n, pid))).all()
    mask = torch.empty((n, 38), dtype=torch.uint8, device=DEV)
    ops.legal_mask(state, mask)
    assert (mask.cpu().numpy() == env.export()["legal_action_mask"]).all()
    env.duplicate_init()
    ops.duplicate_init(state, table_t, state, out)
    _cmp_outputs(out, env.export(), tag="duplicate_init")
    f, p = ops.state_fields(state), env.export_private()
    assert (f["shuffled_players"].cpu().numpy() == p["shuffled_players"]).all()
    assert (f["shuffled_players"].cpu().numpy() == players[:, [1, 0, 3, 2]]).all()  # src/duplicate.py:113-114


def test_abi_errors_are_reported_not_swallowed():
    from brl_b200 import _lib
    ops = _ops()
    with pytest.raises(_lib.BrlError, match="NULL"):
        _lib.call("brl_step", 0, [None] * 10, ops._params(4))
    with pytest.raises(_lib.BrlError):
        ops.step(torch.zeros((5, 4, 4), dtype=torch.int32), None, None, None, ops.EnvOutputs(4, DEV))  # CPU tensor


@pytest.mark.parametrize("obs_dtype", [torch.uint8, torch.bfloat16])
@pytest.mark.parametrize("n,tune_kw", [(3000, {}), (8192, {}), (100, {"writers": 2})])
def test_fused_rollout_other_observation_dtypes(obs_dtype, n, tune_kw):
    """pgx's own bool/u8 observation dtype and the bf16 form that feeds the tensor-core policy: same bits as f32."""
    ops, orc = _ops(), _orc()
    from brl_b200 import _lib
    from brl_b200.deals import synthetic_deal_table
    table = synthetic_deal_table(1500, seed=3)
    table_t = _t(table)
    k, seed = 7, 99
    env = orc.OracleEnv(table, n)
    env.init(orc.make_keys(seed, n))
    ref = env.rollout_random(seed, 0, k)
    state, out0 = ops.new_state(n, DEV), ops.EnvOutputs(n, DEV, obs_dtype)
    ops.init(ops.make_keys(seed, n, DEV), table_t, state, out0)
    traj = ops.EnvOutputs(n, DEV, obs_dtype, rows=k)
    ops.rollout_random(state, table_t, k, traj, seed=seed, step0=0, tune=_lib.tune(**tune_kw))
    assert (traj.observation.float().cpu().numpy() == ref["observation"]).all()
    assert (traj.legal_action_mask.cpu().numpy() == ref["legal_action_mask"]).all()
    assert (traj.rewards.cpu().numpy() == ref["rewards"]).all()
