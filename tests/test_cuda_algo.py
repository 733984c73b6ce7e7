"""GPU parity of the per-env algorithms around the env: GAE (<=1e-6 relative, in
practice bit-exact), masked categorical, IMP reward, match statistics."""
import numpy as np
import pytest
import torch

from tests import helpers as H

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _t(a):
    return torch.as_tensor(np.ascontiguousarray(a), device=DEV)


@pytest.mark.parametrize("t_steps,n", [(32, 8192), (5, 77), (1, 1), (19, 1000)])
def test_gae_matches_oracle(t_steps, n):
    from brl_b200 import ops
    from oracle import oracle as orc
    rng = np.random.default_rng(t_steps * 1000 + n)
    done = (rng.random((t_steps, n)) < 0.1).astype(np.uint8)
    value = rng.normal(0, 0.3, (t_steps, n)).astype(np.float32)
    reward = (rng.integers(-7600, 7601, (t_steps, n)) / 7600.0 * done).astype(np.float32)
    last_val = rng.normal(0, 0.3, n).astype(np.float32)
    adv_ref, tgt_ref = orc.gae(done, value, reward, last_val, 1.0, 0.95)  # ppo.py:162-163
    adv = torch.empty((t_steps, n), dtype=torch.float32, device=DEV)
    tgt = torch.empty_like(adv)
    ops.gae(_t(done), _t(value), _t(reward), _t(last_val), adv, tgt, 1.0, 0.95)
    # north-star tolerance: 1e-6 relative (the kernel keeps the reference's operation order, so it is exact)
    np.testing.assert_allclose(adv.cpu().numpy(), adv_ref, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(tgt.cpu().numpy(), tgt_ref, rtol=1e-6, atol=1e-7)
    # float64 restatement of src/gae.py:20-39 as an independent check
    g, nv, ref64 = np.zeros(n), last_val.astype(np.float64), np.zeros((t_steps, n))
    for t in range(t_steps - 1, -1, -1):
        nd = 1.0 - done[t]
        delta = reward[t] + 1.0 * nv * nd - value[t]
        g = delta + 1.0 * 0.95 * nd * g
        nv = value[t].astype(np.float64)
        ref64[t] = g
    np.testing.assert_allclose(adv.cpu().numpy(), ref64, rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize("n", [4096, 37, 1])
def test_categorical_mode_and_logprob(n):
    from brl_b200 import ops
    from oracle import oracle as orc
    rng = np.random.default_rng(n)
    logits = rng.normal(0, 3, (n, 38)).astype(np.float32)
    mask = (rng.random((n, 38)) < 0.5).astype(np.uint8)
    mask[:, 0] = 1  # Pass is always legal
    logits[::5, 7] = logits[::5, 3]  # exact ties -> lowest index wins
    a_ref, lp_ref = orc.categorical(logits, mask, sample=False)
    act = torch.empty(n, dtype=torch.int32, device=DEV)
    lp = torch.empty(n, dtype=torch.float32, device=DEV)
    ops.categorical(_t(logits), _t(mask), act, lp, sample=False)
    assert (act.cpu().numpy() == a_ref).all()
    np.testing.assert_allclose(lp.cpu().numpy(), lp_ref, rtol=1e-5, atol=1e-5)
    assert mask[np.arange(n), act.cpu().numpy()].all()


def test_categorical_sample_follows_gumbel_argmax_and_distribution():
    from brl_b200 import ops
    from oracle import oracle as orc
    n = 20000
    rng = np.random.default_rng(0)
    base = rng.normal(0, 1.5, 38).astype(np.float32)
    logits = np.tile(base, (n, 1))
    mask = np.ones((n, 38), np.uint8)
    mask[:, [1, 2, 30, 31]] = 0
    a_ref, _ = orc.categorical(logits, mask, sample=True, seed=99, env_offset=5, step=3)
    act = torch.empty(n, dtype=torch.int32, device=DEV)
    lp = torch.empty(n, dtype=torch.float32, device=DEV)
    ops.categorical(_t(logits), _t(mask), act, lp, sample=True, seed=99, env_offset=5, step_index=3)
    a = act.cpu().numpy()
    # same counter-based noise on both sides: only float-rounding near-ties may differ
    assert (a == a_ref).mean() > 0.999
    assert mask[np.arange(n), a].all()
    p = np.exp(base - base.max()) * mask[0]
    p /= p.sum()
    freq = np.bincount(a, minlength=38) / n
    assert np.abs(freq - p).max() < 0.02
    np.testing.assert_allclose(lp.cpu().numpy(), np.log(p[a]), rtol=1e-4, atol=1e-4)


def test_imp_reward_table_and_doctests():
    from brl_b200 import ops
    tab = np.load(H.GOLDEN + "/imp_table.npy")
    d = np.arange(-8000, 8001, 10).astype(np.float32)
    a = np.stack([d, d, -d, -d], 1)
    b = np.zeros_like(a)
    out = torch.empty((len(d), 4), dtype=torch.float32, device=DEV)
    ops.imp_reward(_t(a), _t(b), out)
    o = out.cpu().numpy()
    assert (o[:, 0] == tab).all() and (o[:, 1] == tab).all() and (o[:, 2] == -tab).all() and (o[:, 3] == -tab).all()
    # src/duplicate.py:20-43
    a = np.array([[0, 0, 0, 0], [0, 0, 0, 0], [-100, -100, 100, 100], [-100, -100, 100, 100], [-3500, -3500, 3500, 3500],
                  [2000, 2000, -2000, -2000]], np.float32)
    b = np.array([[0, 0, 0, 0], [100, 100, -100, -100], [0, 0, 0, 0], [100, 100, -100, -100], [0, 0, 0, 0],
                  [2000, 2000, -2000, -2000]], np.float32)
    out = torch.empty((6, 4), dtype=torch.float32, device=DEV)
    ops.imp_reward(_t(a), _t(b), out)
    assert out.cpu().numpy()[:, 0].tolist() == [0, 3, -3, 0, -23, 24]


def test_match_stats_sums():
    from brl_b200 import ops
    from oracle import oracle as orc
    rng = np.random.default_rng(1)
    x = rng.integers(-24, 25, 100000).astype(np.float32)
    sums = torch.zeros(8, dtype=torch.float64, device=DEV)
    ops.match_stats(_t(x), sums)
    s = sums.cpu().numpy()
    n, s1, s2, w = s[:4]
    mean = s1 / n
    se = np.sqrt((s2 - s1 * s1 / n) / (n - 1)) / np.sqrt(n)
    ref = orc.match_stats(x)
    np.testing.assert_allclose([mean, se, w / n], ref, rtol=1e-12, atol=1e-12)


def test_gather_reward():
    from brl_b200 import ops
    rng = np.random.default_rng(2)
    r = rng.integers(-7600, 7601, (5000, 4)).astype(np.float32)
    actor = rng.integers(0, 4, 5000).astype(np.int8)
    out = torch.empty(5000, dtype=torch.float32, device=DEV)
    ops.gather_reward(_t(r), _t(actor), out, 7600.0)
    assert (out.cpu().numpy() == (r[np.arange(5000), actor] / np.float32(7600.0))).all()
