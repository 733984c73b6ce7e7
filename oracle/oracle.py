"""ctypes front-end of the CPU oracle (``oracle/brl_oracle.c``).

TEST INFRASTRUCTURE ONLY.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this
module, and only as the checker / the timed CPU baseline.  The product package
``brl_b200`` never imports it (tests/test_layout.py enforces that).

See ``brl_oracle.h`` for which reference file:line every function restates.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libbrl_oracle.so")
_SO_V3 = os.path.join(_HERE, "libbrl_oracle_v3.so")   # same source, -march=x86-64-v3 (AVX2 / BMI2 / FMA)

NUM_ACTIONS = 38
OBS_DIM = 480
DEAL_ROW_BYTES = 48


def build(force: bool = False) -> str:
    """Compile the C restatement with the committed Makefile (gcc only)."""
    src = os.path.join(_HERE, "brl_oracle.c")
    hdr = os.path.join(_HERE, "brl_oracle.h")
    stale = any((not os.path.exists(so)) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in (src, hdr))
                for so in (_SO, _SO_V3))
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "all"], check=True, capture_output=True)
    return _SO


def _host_has_avx2() -> bool:
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith("flags"):
                    f = line.split()
                    return "avx2" in f and "bmi2" in f and "fma" in f
    except OSError:
        pass
    return False


def so_path() -> str:
    """the build the host CPU can run fastest (x86-64-v3 when AVX2 / BMI2 / FMA are present)"""
    return _SO_V3 if (_host_has_avx2() and os.path.exists(_SO_V3)) else _SO


class _Params(C.Structure):
    _fields_ = [
        ("deal_table", C.c_void_p),
        ("n_deals", C.c_int32),
        ("illegal_penalty", C.c_float),
        ("illegal_bonus", C.c_float),
    ]


class _TableInfo(C.Structure):
    _fields_ = [
        ("terminated", C.c_uint8),
        ("rewards", C.c_float * 4),
        ("last_bid", C.c_int32),
        ("last_bidder", C.c_int32),
        ("call_x", C.c_uint8),
        ("call_xx", C.c_uint8),
    ]


TABLE_INFO_DTYPE = np.dtype(
    [
        ("terminated", np.uint8),
        ("rewards", np.float32, (4,)),
        ("last_bid", np.int32),
        ("last_bidder", np.int32),
        ("call_x", np.uint8),
        ("call_xx", np.uint8),
    ],
    align=True,
)

_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(so_path())
        L.orc_state_size.restype = C.c_size_t
        L.orc_score.restype = C.c_int32
        L.orc_score.argtypes = [C.c_int32] + [C.c_int] * 4
        L.orc_imp.restype = C.c_int32
        L.orc_imp.argtypes = [C.c_int32]
        L.orc_make_key.restype = C.c_uint64
        L.orc_make_key.argtypes = [C.c_uint64, C.c_uint64]
        L.orc_rollout_random.restype = C.c_int64
        L.orc_random_legal_action.restype = C.c_int32
        L.orc_random_legal_action.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32]
        assert C.sizeof(_TableInfo) == TABLE_INFO_DTYPE.itemsize
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def philox(ctr, key):
    ctr = np.asarray(ctr, dtype=np.uint32)
    key = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().orc_philox4x32(_p(ctr), _p(key), _p(out))
    return out


def score(bid: int, x: bool, xx: bool, vul: bool, tricks: int) -> int:
    return int(lib().orc_score(int(bid), int(x), int(xx), int(vul), int(tricks)))


def imp(diff: int) -> int:
    return int(lib().orc_imp(int(diff)))


def imp_reward(a, b):
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    out = np.zeros(4, dtype=np.float32)
    lib().orc_imp_reward(_p(a), _p(b), _p(out))
    return out


def make_keys(seed: int, n: int, offset: int = 0) -> np.ndarray:
    L = lib()
    return np.array([L.orc_make_key(C.c_uint64(seed), C.c_uint64(offset + i)) for i in range(n)], dtype=np.uint64)


def draw_episode(key: int, n_deals: int):
    new_key = C.c_uint64()
    vals = [C.c_int32() for _ in range(5)]
    lib().orc_draw_episode(C.c_uint64(int(key)), C.c_int32(n_deals), C.byref(new_key), *[C.byref(v) for v in vals])
    return int(new_key.value), [int(v.value) for v in vals]


def seating_to_players(seating: int) -> np.ndarray:
    out = np.zeros(4, dtype=np.int8)
    lib().orc_seating_to_players(C.c_int32(seating), _p(out))
    return out


class OracleEnv:
    """Batch of ``n`` oracle envs over one deal table (rows of 48 bytes)."""

    def __init__(self, deal_table: np.ndarray, n: int, illegal_penalty: float = -1.0,
                 illegal_bonus: float = 1.0, n_threads: int = 1):
        self.L = lib()
        self.table = np.ascontiguousarray(deal_table, dtype=np.uint8).reshape(-1, DEAL_ROW_BYTES)
        self.n = int(n)
        self.n_threads = int(n_threads)
        self.params = _Params(self.table.ctypes.data, self.table.shape[0], illegal_penalty, illegal_bonus)
        self.ssize = int(self.L.orc_state_size())
        self.buf = np.zeros(self.n * self.ssize, dtype=np.uint8)
        self.info_a = np.zeros(self.n, dtype=TABLE_INFO_DTYPE)
        self.info_b = np.zeros(self.n, dtype=TABLE_INFO_DTYPE)

    # -- env surface -----------------------------------------------------
    def init(self, keys: np.ndarray):
        keys = np.ascontiguousarray(keys, dtype=np.uint64)
        assert keys.shape == (self.n,)
        self.L.orc_init_batch(_p(self.buf), C.byref(self.params), _p(keys), C.c_int64(self.n), self.n_threads)

    def reset_fields(self, deal, dealer, vul_ns, vul_ew, players, rng_key=None):
        """Place every env on a GIVEN (deal, dealer, vul, seating) -- parity is
        conditional on identical episode draws (north star)."""
        players = np.ascontiguousarray(players, dtype=np.int8).reshape(self.n, 4)
        for i in range(self.n):
            self.L.orc_reset_fields(
                C.c_void_p(self.buf.ctypes.data + i * self.ssize), C.byref(self.params), C.c_int32(int(deal[i])),
                C.c_int32(int(dealer[i])), C.c_int(int(vul_ns[i])), C.c_int(int(vul_ew[i])),
                _p(players[i]), C.c_uint64(0 if rng_key is None else int(rng_key[i])))

    def step(self, actions: np.ndarray, autoreset: bool = False):
        actions = np.ascontiguousarray(actions, dtype=np.int32)
        assert actions.shape == (self.n,)
        self.L.orc_step_batch(_p(self.buf), C.byref(self.params), _p(actions), C.c_int64(self.n),
                              int(autoreset), self.n_threads)

    def duplicate_tables_from_state(self):
        for i in range(self.n):
            base = C.c_void_p(self.buf.ctypes.data + i * self.ssize)
            self.L.orc_table_info_from_state(base, C.c_void_p(self.info_a.ctypes.data + i * TABLE_INFO_DTYPE.itemsize))
            self.L.orc_table_info_from_state(base, C.c_void_p(self.info_b.ctypes.data + i * TABLE_INFO_DTYPE.itemsize))

    def duplicate_step(self, actions: np.ndarray):
        actions = np.ascontiguousarray(actions, dtype=np.int32)
        self.L.orc_duplicate_step_batch(_p(self.buf), C.byref(self.params), _p(actions), _p(self.info_a),
                                        _p(self.info_b), C.c_int64(self.n), self.n_threads)

    def duplicate_init(self):
        for i in range(self.n):
            self.L.orc_duplicate_init(C.c_void_p(self.buf.ctypes.data + i * self.ssize), C.byref(self.params))

    def export(self, obs_dtype=np.float32):
        n = self.n
        obs = np.zeros((n, OBS_DIM), dtype=obs_dtype)
        mask = np.zeros((n, NUM_ACTIONS), dtype=np.uint8)
        rewards = np.zeros((n, 4), dtype=np.float32)
        term = np.zeros(n, dtype=np.uint8)
        cur = np.zeros(n, dtype=np.int8)
        f32 = obs if obs_dtype == np.float32 else None
        u8 = obs if obs_dtype == np.uint8 else None
        self.L.orc_export(_p(self.buf), C.c_int64(n), _p(f32), _p(u8), _p(mask), _p(rewards), _p(term), _p(cur))
        return dict(observation=obs, legal_action_mask=mask, rewards=rewards, terminated=term, current_player=cur)

    def export_private(self):
        n = self.n
        out = dict(
            deal=np.zeros(n, np.int32), dealer=np.zeros(n, np.int32), shuffled_players=np.zeros((n, 4), np.int8),
            vul=np.zeros((n, 2), np.uint8), last_bid=np.zeros(n, np.int32), last_bidder=np.zeros(n, np.int32),
            call_x=np.zeros(n, np.uint8), call_xx=np.zeros(n, np.uint8), pass_num=np.zeros(n, np.int32),
            step_count=np.zeros(n, np.int32), rng_key=np.zeros(n, np.uint64))
        self.L.orc_export_private(_p(self.buf), C.c_int64(n), *[_p(out[k]) for k in (
            "deal", "dealer", "shuffled_players", "vul", "last_bid", "last_bidder", "call_x", "call_xx",
            "pass_num", "step_count", "rng_key")])
        return out

    def observe(self, player_ids: np.ndarray) -> np.ndarray:
        out = np.zeros((self.n, OBS_DIM), dtype=np.uint8)
        for i in range(self.n):
            self.L.orc_observe(C.c_void_p(self.buf.ctypes.data + i * self.ssize), C.byref(self.params),
                               C.c_int(int(player_ids[i])), _p(out[i]))
        return out

    def random_legal_actions(self, seed: int, step: int, env_offset: int = 0) -> np.ndarray:
        mask = self.export()["legal_action_mask"]
        return np.array([self.L.orc_random_legal_action(_p(mask[i]), C.c_uint64(seed), C.c_uint64(env_offset + i),
                                                        C.c_uint32(step)) for i in range(self.n)], dtype=np.int32)

    def rollout_random(self, seed: int, step0: int, k_steps: int, env_offset: int = 0, want_outputs: bool = True,
                       n_threads: int | None = None):
        """K auto-reset steps with random-legal actions (the timed CPU baseline)."""
        n, k = self.n, int(k_steps)
        out = {}
        if want_outputs:
            out = dict(observation=np.zeros((k, n, OBS_DIM), np.float32), legal_action_mask=np.zeros((k, n, NUM_ACTIONS), np.uint8),
                       rewards=np.zeros((k, n, 4), np.float32), terminated=np.zeros((k, n), np.uint8),
                       current_player=np.zeros((k, n), np.int8), action=np.zeros((k, n), np.int32))
        g = out.get
        nt = self.n_threads if n_threads is None else n_threads
        n_term = self.L.orc_rollout_random(
            _p(self.buf), C.byref(self.params), C.c_int64(n), C.c_int64(env_offset), C.c_uint64(seed), C.c_uint32(step0),
            C.c_int32(k), _p(g("observation")), _p(g("legal_action_mask")), _p(g("rewards")), _p(g("terminated")),
            _p(g("current_player")), _p(g("action")), C.c_int(nt))
        out["n_terminated"] = int(n_term)
        return out


def gae(done, value, reward, last_val, gamma: float, lam: float):
    done = np.ascontiguousarray(done, dtype=np.uint8)
    value = np.ascontiguousarray(value, dtype=np.float32)
    reward = np.ascontiguousarray(reward, dtype=np.float32)
    last_val = np.ascontiguousarray(last_val, dtype=np.float32)
    t, n = done.shape
    adv = np.zeros((t, n), np.float32)
    tgt = np.zeros((t, n), np.float32)
    lib().orc_gae(_p(done), _p(value), _p(reward), _p(last_val), C.c_int32(t), C.c_int64(n),
                  C.c_float(gamma), C.c_float(lam), _p(adv), _p(tgt))
    return adv, tgt


def categorical(logits, mask, sample: bool, seed: int = 0, env_offset: int = 0, step: int = 0):
    logits = np.ascontiguousarray(logits, dtype=np.float32)
    mask = np.ascontiguousarray(mask, dtype=np.uint8)
    n = logits.shape[0]
    action = np.zeros(n, np.int32)
    logp = np.zeros(n, np.float32)
    lib().orc_categorical(_p(logits), _p(mask), C.c_int64(n), int(sample), C.c_uint64(seed), C.c_uint64(env_offset),
                          C.c_uint32(step), _p(action), _p(logp))
    return action, logp


def match_stats(cum_return):
    x = np.ascontiguousarray(cum_return, dtype=np.float64)
    out = np.zeros(3, np.float64)
    lib().orc_match_stats(_p(x), C.c_int64(x.shape[0]), _p(out))
    return out


def mlp_forward(params: dict, x: np.ndarray):
    """DeepMind 4x1024 ReLU actor-critic (src/models.py:23-33), fp32 NumPy.
    ``params`` maps 'w0'..'w5','b0'..'b5' with w [in,out] (haiku: y = x@w + b)."""
    h = np.asarray(x, dtype=np.float32)
    for i in range(4):
        h = np.maximum(h @ params[f"w{i}"] + params[f"b{i}"], 0.0)
    logits = h @ params["w4"] + params["b4"]
    value = (h @ params["w5"] + params["b5"])[:, 0]
    return logits, value
