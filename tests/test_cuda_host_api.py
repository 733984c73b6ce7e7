"""GPU: the host-buffer C-ABI layer (brl_env_*) and caller-supplied action randomness."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TAG_ACT = 0x41435430


def _uniforms(orc, seed, n, k, offset=0, step0=0):
    u = np.zeros((k, n), np.uint32)
    key = [seed & 0xFFFFFFFF, seed >> 32]
    for s in range(k):
        for i in range(n):
            g = offset + i
            u[s, i] = orc.philox([g & 0xFFFFFFFF, g >> 32, TAG_ACT, step0 + s], key)[0]
    return u


@pytest.mark.parametrize("tune_kw,k", [({}, 12), ({"classic_rollout": True, "epw": 8}, 12), ({}, 100), ({"writers": 1}, 70)])
def test_rollout_with_caller_supplied_uniforms_equals_philox_mode(tune_kw, k):
    from brl_b200 import _lib, ops
    from brl_b200.deals import synthetic_deal_table
    from oracle import oracle as orc
    n, seed = 200, 99
    table = synthetic_deal_table(700, seed=2)
    table_t = torch.as_tensor(table, device=DEV)
    env = orc.OracleEnv(table, n)
    env.init(orc.make_keys(seed, n))
    ref = env.rollout_random(seed, 0, k)
    u = torch.as_tensor(_uniforms(orc, seed, n, k).view(np.int32), device=DEV)
    state, out0 = ops.new_state(n, DEV), ops.EnvOutputs(n, DEV)
    ops.init(ops.make_keys(seed, n, DEV), table_t, state, out0)
    traj = ops.EnvOutputs(n, DEV, rows=k)
    actions = torch.empty((k, n), dtype=torch.int32, device=DEV)
    ops.rollout_random(state, table_t, k, traj, seed=12345, step0=777, action_out=actions, uniforms=u, tune=_lib.tune(**tune_kw))
    assert (actions.cpu().numpy() == ref["action"]).all()
    assert (traj.observation.cpu().numpy() == ref["observation"]).all()
    assert (traj.rewards.cpu().numpy() == ref["rewards"]).all()


def test_host_buffer_api_step_and_rollout_match_oracle():
    from brl_b200 import _lib
    from brl_b200.deals import synthetic_deal_table
    from oracle import oracle as orc
    L = _lib.load()
    n, seed, offset = 300, 5, 40
    table = synthetic_deal_table(900, seed=3)
    h = L.brl_env_create(n, offset, table.ctypes.data, table.shape[0], seed, _lib.F_AUTORESET | _lib.F_OBS_U8)
    assert h, L.brl_last_error()
    obs, mask = np.zeros((n, 480), np.uint8), np.zeros((n, 38), np.uint8)
    rew, term, cur = np.zeros((n, 4), np.float32), np.zeros(n, np.uint8), np.zeros(n, np.int8)
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    assert L.brl_env_init_host(h, p(obs), p(mask), p(rew), p(term), p(cur)) == 0
    env = orc.OracleEnv(table, n)
    env.init(orc.make_keys(seed, n, offset))
    e = env.export(np.uint8)
    assert (obs == e["observation"]).all() and (mask == e["legal_action_mask"]).all() and (cur == e["current_player"]).all()
    for s in range(25):
        act = env.random_legal_actions(seed, s, env_offset=offset)
        env.step(act, autoreset=True)
        assert L.brl_env_step_host(h, p(act), p(obs), p(mask), p(rew), p(term), p(cur)) == 0
        e = env.export(np.uint8)
        assert (obs == e["observation"]).all() and (mask == e["legal_action_mask"]).all()
        assert (rew == e["rewards"]).all() and (term == e["terminated"]).all() and (cur == e["current_player"]).all()
    # rollout with host-supplied randomness: rewards / terminated / stats come back to the host
    k = 10
    u = _uniforms(orc, seed, n, k, offset=offset, step0=1000)
    ref = env.rollout_random(seed, 1000, k, env_offset=offset)
    rr, tt, st = np.zeros((k, n, 4), np.float32), np.zeros((k, n), np.uint8), np.zeros(4, np.uint64)
    assert L.brl_env_rollout_host(h, k, p(u), p(rr), p(tt), p(st)) == 0, L.brl_last_error()
    assert (rr == ref["rewards"]).all() and (tt == ref["terminated"]).all()
    assert int(st[0]) == ref["n_terminated"] and int(st[2]) == n * k
    ptrs = (C.c_void_p * 6)()
    assert L.brl_env_trajectory(h, ptrs) == 0 and all(ptrs)
    assert L.brl_env_step_host(None, p(act), None, None, None, None, None) == -4  # bad handle is an error, not a crash
    L.brl_env_destroy(h)


@pytest.mark.parametrize("staged", [False, True])
def test_host_rollout_into_pinned_buffers_zero_copy(staged):
    """pinned result buffers are written by the kernel itself over PCIe (no copy phase); same numbers as
    the staged path and the oracle, for a chunkable (k=32) and an odd (k=5) horizon."""
    from brl_b200 import _lib
    from brl_b200.deals import synthetic_deal_table
    from oracle import oracle as orc
    L = _lib.load()
    n, seed = 8192, 17
    table = synthetic_deal_table(1200, seed=4)
    flags = _lib.F_AUTORESET | (_lib.F_HOST_STAGED if staged else 0)
    h = L.brl_env_create(n, 0, table.ctypes.data, table.shape[0], seed, flags)
    assert h, L.brl_last_error()
    assert L.brl_env_init_host(h, None, None, None, None, None) == 0
    env = orc.OracleEnv(table, n)
    env.init(orc.make_keys(seed, n))
    vp = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    step0 = 0
    for k in (32, 5):
        rng = np.random.default_rng(k)
        u_np = rng.integers(0, 2 ** 32, size=(k, n), dtype=np.uint32)
        u = torch.from_numpy(u_np.view(np.int32)).pin_memory()
        rew = torch.full((k, n, 4), -1.0).pin_memory()
        term = torch.full((k, n), 7, dtype=torch.uint8).pin_memory()
        stats = torch.zeros(4, dtype=torch.int64).pin_memory()
        assert L.brl_env_rollout_host(h, k, vp(u), vp(rew), vp(term), vp(stats)) == 0, L.brl_last_error()
        # oracle with the same caller-supplied uniforms: replay action by action
        n_term = 0
        for s in range(k):
            e = env.export()
            m = e["legal_action_mask"].astype(bool)
            n_legal = m.sum(1)
            kth = ((u_np[s].astype(np.uint64) * n_legal.astype(np.uint64)) >> np.uint64(32)).astype(np.int64)
            act = np.array([np.nonzero(m[i])[0][kth[i]] for i in range(n)], np.int32)
            env.step(act, autoreset=True)
            e = env.export()
            assert (rew[s].numpy() == e["rewards"]).all(), (k, s)
            assert (term[s].numpy() == e["terminated"]).all(), (k, s)
            n_term += int(e["terminated"].sum())
        assert int(stats[0]) == n_term and int(stats[2]) == n * k
        step0 += k
    L.brl_env_destroy(h)


def test_pipelined_host_rollout_equals_blocking_calls():
    from brl_b200 import _lib
    from brl_b200.deals import synthetic_deal_table
    L = _lib.load()
    n, seed, k, calls = 2048, 23, 8, 9  # 9 calls through 4 staging slots: every slot is reused at least once
    table = synthetic_deal_table(800, seed=6)
    vp = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    rng = np.random.default_rng(0)
    us = [torch.from_numpy(rng.integers(0, 2 ** 32, size=(k, n), dtype=np.uint32).view(np.int32)).pin_memory() for _ in range(calls)]

    def run(pipelined):
        h = L.brl_env_create(n, 0, table.ctypes.data, table.shape[0], seed, _lib.F_AUTORESET)
        assert L.brl_env_init_host(h, None, None, None, None, None) == 0
        outs = [(torch.zeros((k, n, 4)).pin_memory(), torch.zeros((k, n), dtype=torch.uint8).pin_memory(),
                 torch.zeros(4, dtype=torch.int64).pin_memory()) for _ in range(calls)]
        tickets = []
        for c in range(calls):
            r, t, s = outs[c]
            if pipelined:
                tk = L.brl_env_rollout_host_async(h, k, vp(us[c]), vp(r), vp(t), vp(s))
                assert tk == c + 1, L.brl_last_error()
                tickets.append(tk)
                if c >= 3:  # BRL_ENV_PIPELINE_DEPTH - 1 older calls stay in flight
                    assert L.brl_env_wait(h, tickets[c - 3]) == 0
            else:
                assert L.brl_env_rollout_host(h, k, vp(us[c]), vp(r), vp(t), vp(s)) == 0
        if pipelined:
            assert L.brl_env_wait(h, tickets[-1]) == 0 and L.brl_env_wait(h, 1) == 0
            assert L.brl_env_wait(h, calls + 1) == -1
        L.brl_env_destroy(h)
        return outs

    a, b = run(False), run(True)
    for (r0, t0, s0), (r1, t1, s1) in zip(a, b):
        assert torch.equal(r0, r1) and torch.equal(t0, t1) and torch.equal(s0, s1)
        assert int(s0[2]) == n * k and int(s0[0]) > 0


def _replay_with_uniforms(env, u32, n):
    """oracle replay of a rollout whose action uniforms are caller-supplied: action = mulhi(u, #legal)-th legal one"""
    rews, terms, n_term = [], [], 0
    for s in range(u32.shape[0]):
        m = env.export()["legal_action_mask"].astype(bool)
        n_legal = m.sum(1)
        kth = ((u32[s].astype(np.uint64) * n_legal.astype(np.uint64)) >> np.uint64(32)).astype(np.int64)
        act = np.array([np.nonzero(m[i])[0][kth[i]] for i in range(n)], np.int32)
        env.step(act, autoreset=True)
        e = env.export()
        rews.append(e["rewards"].copy()); terms.append(e["terminated"].copy())
        n_term += int(e["terminated"].sum())
    return np.stack(rews), np.stack(terms), n_term


@pytest.mark.parametrize("tune_kw", [{}, {"classic_rollout": True, "epw": 8}])
@pytest.mark.parametrize("u16", [False, True])
def test_compact_result_and_u16_uniforms_decode_to_the_oracle_outputs(tune_kw, u16):
    """brl_rollout_random with BRL_F_RESULT_I16 (+ BRL_F_UNIFORM_U16): the 2-byte result decodes to exactly the f32[4]
    rewards / u8 terminated the same launch writes and the oracle gives (what src/roll_out.py:86-94 consumes)."""
    from brl_b200 import _lib, ops
    from brl_b200.deals import synthetic_deal_table
    from oracle import oracle as orc
    n, k, seed = 333, 40, 31
    table = synthetic_deal_table(900, seed=8)
    table_t = torch.as_tensor(table, device=DEV)
    rng = np.random.default_rng(5)
    if u16:
        u_np = rng.integers(0, 2 ** 16, size=(k, n), dtype=np.uint16)
        u32 = u_np.astype(np.uint32) << np.uint32(16)
        u = torch.as_tensor(u_np.view(np.int16), device=DEV)
    else:
        u32 = rng.integers(0, 2 ** 32, size=(k, n), dtype=np.uint32)
        u = torch.as_tensor(u32.view(np.int32), device=DEV)
    env = orc.OracleEnv(table, n)
    env.init(orc.make_keys(seed, n))
    ref_rew, ref_term, _ = _replay_with_uniforms(env, u32, n)
    state, out0 = ops.new_state(n, DEV), ops.EnvOutputs(n, DEV)
    ops.init(ops.make_keys(seed, n, DEV), table_t, state, out0)
    traj = ops.EnvOutputs(n, DEV, rows=k)
    res = torch.full((k, n), -7, dtype=torch.int16, device=DEV)
    ops.rollout_random(state, table_t, k, traj, uniforms=u, result16=res, tune=_lib.tune(**tune_kw))
    rew, term = ops.decode_result16(res)
    assert (traj.rewards.cpu().numpy() == ref_rew).all() and (traj.terminated.cpu().numpy() == ref_term).all()
    assert (rew.cpu().numpy() == ref_rew).all() and (term.cpu().numpy() == ref_term).all()
    assert ref_term.sum() > 100 and (ref_rew != 0).sum() > 100
    with pytest.raises(_lib.BrlError):  # the flag without the buffer is a usage error, not a crash
        _lib.call("brl_rollout_random", 0, [state.data_ptr(), table_t.data_ptr()] + [None] * 9,
                  ops._params(n, flags=_lib.F_RESULT_I16, n_deals=table.shape[0], stride=n, k_steps=k))


@pytest.mark.parametrize("u16", [False, True])
def test_compact_host_rollout_pipelined_equals_f32_payload_and_oracle(u16):
    """brl_env_rollout_host_compact[_async]: same rollout as the f32-payload calls, 2 B per env-step each way"""
    from brl_b200 import _lib
    from brl_b200.deals import synthetic_deal_table
    from oracle import oracle as orc
    L = _lib.load()
    n, seed, k, calls = 1500, 29, 8, 7
    table = synthetic_deal_table(800, seed=9)
    vp = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    rng = np.random.default_rng(1)
    if u16:
        us_np = [rng.integers(0, 2 ** 16, size=(k, n), dtype=np.uint16) for _ in range(calls)]
        us32 = [x.astype(np.uint32) << np.uint32(16) for x in us_np]
        us = [torch.from_numpy(x.view(np.int16)).pin_memory() for x in us_np]
    else:
        us32 = [rng.integers(0, 2 ** 32, size=(k, n), dtype=np.uint32) for _ in range(calls)]
        us = [torch.from_numpy(x.view(np.int32)).pin_memory() for x in us32]
    flags = _lib.F_AUTORESET | (_lib.F_UNIFORM_U16 if u16 else 0)
    h = L.brl_env_create(n, 0, table.ctypes.data, table.shape[0], seed, flags)
    assert h, L.brl_last_error()
    assert L.brl_env_init_host(h, None, None, None, None, None) == 0
    res = [torch.zeros((k, n), dtype=torch.int16).pin_memory() for _ in range(calls)]
    st = [torch.zeros(4, dtype=torch.int64).pin_memory() for _ in range(calls)]
    tickets = []
    for c in range(calls):
        t = L.brl_env_rollout_host_compact_async(h, k, vp(us[c]), vp(res[c]), vp(st[c]))
        assert t == c + 1, L.brl_last_error()
        tickets.append(t)
        if c >= 3:
            assert L.brl_env_wait(h, tickets[c - 3]) == 0
    assert L.brl_env_wait(h, tickets[-1]) == 0
    # the device-resident trajectory of the last call still holds the full f32 rewards / u8 terminated
    ptrs = (C.c_void_p * 6)()
    assert L.brl_env_trajectory(h, ptrs) == 0
    env = orc.OracleEnv(table, n)
    env.init(orc.make_keys(seed, n))
    for c in range(calls):
        ref_rew, ref_term, n_term = _replay_with_uniforms(env, us32[c], n)
        rew = np.zeros((k, n, 4), np.float32); term = np.zeros((k, n), np.uint8)
        L.brl_result16_decode(vp(res[c]), k * n, rew.ctypes.data_as(C.c_void_p), term.ctypes.data_as(C.c_void_p))
        assert (rew == ref_rew).all() and (term == ref_term).all(), c
        assert int(st[c][0]) == n_term and int(st[c][2]) == n * k
    # blocking form continues the same env
    r1 = torch.zeros((k, n), dtype=torch.int16).pin_memory()
    assert L.brl_env_rollout_host_compact(h, k, vp(us[0]), vp(r1), vp(st[0])) == 0, L.brl_last_error()
    ref_rew, ref_term, _ = _replay_with_uniforms(env, us32[0], n)
    rew = np.zeros((k, n, 4), np.float32); term = np.zeros((k, n), np.uint8)
    L.brl_result16_decode(vp(r1), k * n, rew.ctypes.data_as(C.c_void_p), term.ctypes.data_as(C.c_void_p))
    assert (rew == ref_rew).all() and (term == ref_term).all()
    assert L.brl_env_rollout_host_compact(h, k, vp(us[0]), None, None) == -2  # NULL result is a usage error
    L.brl_env_destroy(h)
