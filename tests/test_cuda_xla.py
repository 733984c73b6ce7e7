"""GPU: the legacy XLA custom-call leg (csrc/xla_ffi_shim.cc).  Each `<op>_xla` symbol is called through ctypes with
exactly the signature XLA uses -- `void f(cudaStream_t, void** buffers, const char* opaque, size_t len,
XlaCustomCallStatus*)`, buffers = operands then results -- and must produce what the direct op produces.  Call sites the
ops stand in for: src/roll_out.py:51 (env.step), src/duplicate.py:134,149 (_observe, duplicate_step), src/gae.py:32-38."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _xla_order(layout: str, op_buffers):
    """op-ordered buffer list -> XLA's (operands, then results); in-place buffers appear in both"""
    ins = [b for c, b in zip(layout, op_buffers) if c in "ix"]
    outs = [b for c, b in zip(layout, op_buffers) if c in "osx"]
    return ins + outs


def _call_xla(name, op_buffers, params, status=None):
    from brl_b200 import _lib
    L = _lib.load()
    layout = L.brl_xla_layout(name.encode()).decode()
    assert len(layout) == len(op_buffers), (name, layout, len(op_buffers))
    bufs = _xla_order(layout, op_buffers)
    arr = (C.c_void_p * len(bufs))(*[C.c_void_p(b.data_ptr()) if b is not None else None for b in bufs])
    stream = torch.cuda.current_stream().cuda_stream
    getattr(L, name + "_xla")(C.c_void_p(stream), arr, C.cast(C.byref(params), C.c_char_p), C.sizeof(params), status)
    return layout


def _env(n, seed=3, rows=700):
    from brl_b200 import ops
    from brl_b200.deals import synthetic_deal_table
    table = torch.as_tensor(synthetic_deal_table(rows, seed=2), device=DEV)
    state, out = ops.new_state(n, DEV), ops.EnvOutputs(n, DEV)
    ops.init(ops.make_keys(seed, n, DEV), table, state, out)
    for i in range(5):
        ops.step(state, None, table, state, out, autoreset=True, random_action=True, seed=seed, step_index=i)
    return table, state, out


def test_layout_table_covers_every_op_and_matches_the_header():
    from brl_b200 import _lib
    L = _lib.load()
    for op in _lib.OPS:
        layout = L.brl_xla_layout(op.encode())
        assert layout is not None and set(layout.decode()) <= set("iosx"), op
    assert L.brl_xla_layout(b"brl_nope") is None
    assert L.brl_xla_layout(b"brl_step").decode() == "iiiooooooo"
    assert L.brl_xla_layout(b"brl_gae").decode() == "iiiioo"


def test_step_and_observe_xla_equal_direct_ops():
    from brl_b200 import _lib, ops
    n = 500
    table, state, out = _env(n)
    action = torch.zeros(n, dtype=torch.int32, device=DEV)
    ops.step(state, None, table, state.clone(), ops.EnvOutputs(n, DEV), autoreset=True, random_action=True, seed=9, step_index=7,
             action_out=action)
    # direct
    s1, o1 = state.clone(), ops.EnvOutputs(n, DEV)
    ops.step(state, action, table, s1, o1, autoreset=True)
    # through the XLA custom-call symbol (out of place: state_in != state_out, XLA's functional form)
    s2, o2 = torch.zeros_like(state), ops.EnvOutputs(n, DEV)
    taken = torch.full((n,), -1, dtype=torch.int32, device=DEV)
    p = ops._params(n, flags=_lib.F_AUTORESET, n_deals=table.shape[0], stride=n)
    _call_xla("brl_step", [state, action, table, s2, o2.observation, o2.legal_action_mask, o2.rewards, o2.terminated,
                           o2.current_player, taken], p)
    torch.cuda.synchronize()
    assert torch.equal(s1, s2) and torch.equal(o1.observation, o2.observation) and torch.equal(o1.rewards, o2.rewards)
    assert torch.equal(o1.legal_action_mask, o2.legal_action_mask) and torch.equal(o1.terminated, o2.terminated)
    assert torch.equal(o1.current_player, o2.current_player) and torch.equal(taken, action)
    # _observe(state, player) (src/duplicate.py:134)
    pid = (o1.current_player + 1) % 4
    want = torch.empty((n, 480), dtype=torch.float32, device=DEV)
    ops.observe(s1, table, want, pid.to(torch.int8).contiguous())
    got = torch.zeros_like(want)
    _call_xla("brl_observe", [s1, pid.to(torch.int8).contiguous(), table, got], ops._params(n, stride=n, n_deals=table.shape[0]))
    torch.cuda.synchronize()
    assert torch.equal(want, got)


def test_gae_xla_equals_direct_op():
    from brl_b200 import ops
    T, n = 32, 777
    g = torch.Generator(device=DEV).manual_seed(1)
    done = (torch.rand((T, n), generator=g, device=DEV) < 0.1).to(torch.uint8)
    value = torch.randn((T, n), generator=g, device=DEV)
    reward = torch.randn((T, n), generator=g, device=DEV) * done
    last = torch.randn(n, generator=g, device=DEV)
    a1, t1 = torch.empty_like(value), torch.empty_like(value)
    ops.gae(done, value, reward, last, a1, t1, 1.0, 0.95)
    a2, t2 = torch.zeros_like(value), torch.zeros_like(value)
    _call_xla("brl_gae", [done, value, reward, last, a2, t2], ops._params(n, k_steps=T, gamma=1.0, gae_lambda=0.95))
    torch.cuda.synchronize()
    assert torch.equal(a1, a2) and torch.equal(t1, t2)


def test_rollout_random_xla_in_place_buffers_are_aliased_results():
    """buffers the op updates in place (state, stats) are an XLA operand AND the result aliased to it"""
    from brl_b200 import _lib, ops
    n, k = 300, 9
    table, state, _ = _env(n)
    s1, s2 = state.clone(), state.clone()
    t1, t2 = ops.EnvOutputs(n, DEV, rows=k), ops.EnvOutputs(n, DEV, rows=k)
    a1 = torch.empty((k, n), dtype=torch.int32, device=DEV)
    a2 = torch.empty_like(a1)
    st1, st2 = torch.zeros(4, dtype=torch.int64, device=DEV), torch.zeros(4, dtype=torch.int64, device=DEV)
    u = torch.randint(-2 ** 31, 2 ** 31 - 1, (k, n), dtype=torch.int32, device=DEV)
    r1 = torch.zeros((k, n), dtype=torch.int16, device=DEV)
    r2 = torch.zeros_like(r1)
    ops.rollout_random(s1, table, k, t1, action_out=a1, stats=st1, uniforms=u, result16=r1)
    p = ops._params(n, flags=_lib.F_RESULT_I16, n_deals=table.shape[0], stride=n, k_steps=k)
    layout = _call_xla("brl_rollout_random", [s2, table, t2.observation, t2.legal_action_mask, t2.rewards, t2.terminated,
                                              t2.current_player, a2, st2, u, r2], p)
    torch.cuda.synchronize()
    assert layout.count("x") == 2
    assert torch.equal(s1, s2) and torch.equal(t1.observation, t2.observation) and torch.equal(a1, a2)
    assert torch.equal(st1, st2) and torch.equal(r1, r2) and int(st1[2]) == n * k


def test_duplicate_step_xla_equals_direct_op():
    from brl_b200 import BridgeBidding, ops
    from brl_b200.deals import synthetic_deal_table
    from brl_b200.duplicate import Table_info
    n = 400
    env = BridgeBidding(table=synthetic_deal_table(600, seed=5), device=DEV)
    state = env.init(env.make_keys(11, n))
    infos = [[Table_info.from_state(state), Table_info.from_state(state)] for _ in range(2)]
    packed = [state._packed.clone(), state._packed.clone()]
    outs = [ops.EnvOutputs(n, DEV), ops.EnvOutputs(n, DEV)]
    mask = state._mask_u8.clone()
    for it in range(40):
        # any legal action: the lowest-numbered legal bid every fifth call, Pass otherwise (both tables finish in ~20 calls)
        m = mask.to(torch.int32)
        first_bid = (m[:, 3:].argmax(1) + 3).to(torch.int32)
        action = torch.where((m[:, 3:].sum(1) > 0) & (torch.arange(n, device=DEV) + it) .remainder(5).eq(0), first_bid,
                             torch.zeros_like(first_bid)).contiguous()
        ops.duplicate_step(packed[0], action, env.table, infos[0][0]._buffers(), infos[0][1]._buffers(), packed[0], outs[0],
                           env.illegal_penalty, env.illegal_bonus)
        tb = [infos[1][0]._buffers(), infos[1][1]._buffers()]
        flat = lambda t: [t.terminated, t.rewards, t.last_bid, t.last_bidder, t.call_x, t.call_xx]  # noqa: E731
        o = outs[1]
        new_state = torch.zeros_like(packed[1])
        p = ops._params(n, n_deals=env.n_deals, stride=n, illegal_penalty=env.illegal_penalty, illegal_bonus=env.illegal_bonus)
        _call_xla("brl_duplicate_step", [packed[1], action, env.table, *flat(tb[0]), *flat(tb[1]), new_state, o.observation,
                                         o.legal_action_mask, o.rewards, o.terminated, o.current_player], p)
        packed[1] = new_state
        mask = outs[0].legal_action_mask
    torch.cuda.synchronize()
    assert torch.equal(packed[0], packed[1])
    for k in ("observation", "legal_action_mask", "rewards", "terminated", "current_player"):
        assert torch.equal(getattr(outs[0], k), getattr(outs[1], k)), k
    for a, b in zip(infos[0], infos[1]):
        for x, y in zip(a, b):
            assert torch.equal(x, y)
    assert bool(outs[0].terminated.any())


_FAKE_STATUS_C = r"""
#include <string.h>
static char last[512];
static int calls;
void XlaCustomCallStatusSetFailure(void* status, const char* msg, size_t len) {
    if (len > 511) len = 511;
    memcpy(last, msg, len); last[len] = 0; ++calls;
    if (status) *(int*)status = 1;
}
const char* fake_last(void) { return last; }
int fake_calls(void) { return calls; }
"""


def test_failing_op_reports_through_the_xla_status_symbol(tmp_path):
    """Without XLA in the process the failure is counted, not dropped silently; once a process-wide
    XlaCustomCallStatusSetFailure exists (here: a stand-in .so loaded RTLD_GLOBAL, as XLA's runtime would provide it), the
    op's message reaches it."""
    from brl_b200 import _lib, ops
    L = _lib.load()
    n = 16
    bad = ops._params(n, k_steps=0)  # brl_gae: k_steps must be > 0
    t = torch.zeros((1, n), device=DEV)
    before = L.brl_xla_unreported_failures()
    probe = C.CDLL(None)
    have_symbol = hasattr(probe, "XlaCustomCallStatusSetFailure")
    if not have_symbol:
        _call_xla("brl_gae", [t.to(torch.uint8), t, t, t[0], t.clone(), t.clone()], bad)
        assert L.brl_xla_unreported_failures() == before + 1
        assert b"k_steps" in L.brl_last_error()
    src = tmp_path / "fake_xla_status.c"
    src.write_text(_FAKE_STATUS_C)
    so = tmp_path / "libfake_xla_status.so"
    subprocess.run(["gcc", "-shared", "-fPIC", "-O1", "-o", str(so), str(src)], check=True)
    fake = C.CDLL(str(so), mode=os.RTLD_GLOBAL if hasattr(os, "RTLD_GLOBAL") else C.RTLD_GLOBAL)
    fake.fake_last.restype = C.c_char_p
    status = C.c_int(0)
    _call_xla("brl_gae", [t.to(torch.uint8), t, t, t[0], t.clone(), t.clone()], bad, C.byref(status))
    assert fake.fake_calls() == 1 and status.value == 1
    assert b"k_steps" in fake.fake_last()
    # an in-place buffer that is NOT aliased is a usage error reported the same way
    s = ops.new_state(8, DEV)
    stats, stats2 = torch.zeros(4, dtype=torch.int64, device=DEV), torch.zeros(4, dtype=torch.int64, device=DEV)
    layout = L.brl_xla_layout(b"brl_match_stats").decode()
    assert layout == "ix"
    x = torch.zeros(8, device=DEV)
    arr = (C.c_void_p * 3)(C.c_void_p(x.data_ptr()), C.c_void_p(stats.data_ptr()), C.c_void_p(stats2.data_ptr()))
    p = ops._params(8)
    L.brl_match_stats_xla(None, arr, C.cast(C.byref(p), C.c_char_p), C.sizeof(p), C.byref(status))
    assert fake.fake_calls() == 2 and b"alias" in fake.fake_last()
    del s
