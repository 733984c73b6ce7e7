"""GPU parity of the PPO update kernels (csrc/brl_ppo.cu) against the float64 restatement of
src/update.py / optax in oracle/ppo_ref.py.  Floating point: kernels compute in fp32 like the
reference; tolerances are written at each comparison."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _case(B, total, seed):
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn((B, 38), generator=g, dtype=torch.float64) * 2
    value = torch.randn(B, generator=g, dtype=torch.float64) * 0.3
    index = torch.randperm(total, generator=g)[:B].to(torch.int32)
    mask = torch.rand((total, 38), generator=g) < 0.5
    mask[:, 0] = True
    # recorded actions are legal; old log-probs near the current ones so ratios straddle the clip range
    action = torch.zeros(total, dtype=torch.int32)
    for i in range(total):
        legal = torch.nonzero(mask[i])[:, 0]
        action[i] = legal[torch.randint(len(legal), (1,), generator=g)]
    old_lp = -torch.rand(total, generator=g, dtype=torch.float64) * 3
    old_v = torch.randn(total, generator=g, dtype=torch.float64) * 0.3
    adv = torch.randn(total, generator=g, dtype=torch.float64)
    tgt = torch.randn(total, generator=g, dtype=torch.float64) * 0.3
    return logits, value, index, mask, action, old_lp, old_v, adv, tgt


@pytest.mark.parametrize("value_clipping,reward_scaling,masked,ill", [
    (True, False, True, 0.0),      # ppo.py defaults
    (False, True, True, 0.0),
    (True, True, False, 0.3),      # penalty mode + illegal-action L2 term
    (True, False, True, 0.7),
])
def test_ppo_loss_and_gradient_match_autograd(value_clipping, reward_scaling, masked, ill):
    from brl_b200 import ops
    from oracle import ppo_ref
    B, total = 333, 1000
    logits, value, index, mask, action, old_lp, old_v, adv, tgt = _case(B, total, 7)
    # make the policy's log-prob of the taken action close to old_lp for a spread of ratios around 1
    cfg = dict(clip_eps=0.2, ent_coef=0.01, vf_coef=0.5, illegal_l2_coef=ill, value_clipping=value_clipping,
               reward_scaling=reward_scaling, masked_policy=masked)
    idx = index.long()
    with torch.no_grad():
        ml = torch.where(mask[idx], logits, torch.tensor(float("-inf"), dtype=torch.float64))
        lp_now = (torch.log_softmax(ml if masked else logits, 1)).gather(1, action[idx].long()[:, None])[:, 0]
        old_lp[idx] = lp_now + (torch.rand(B, dtype=torch.float64) - 0.5) * 0.8
    lr, vr = logits.clone().requires_grad_(), value.clone().requires_grad_()
    total_ref, aux = ppo_ref.loss_fn(lr, vr, mask[idx], action[idx], old_lp[idx], old_v[idx], adv[idx], tgt[idx], **cfg)
    total_ref.backward()
    f32 = lambda t: t.to(torch.float32).to(DEV).contiguous()  # noqa: E731
    dlogits = torch.empty((B, 38), dtype=torch.float32, device=DEV)
    dvalue = torch.empty(B, dtype=torch.float32, device=DEV)
    stats = torch.zeros(8, dtype=torch.float32, device=DEV)
    scratch = ops.ppo_scratch(DEV)
    ops.ppo_loss(f32(logits), f32(value), index.to(DEV), mask.to(torch.uint8).to(DEV).contiguous(), action.to(DEV), f32(old_lp),
                 f32(old_v), f32(adv), f32(tgt), dlogits, dvalue, stats, scratch, **cfg)
    got = stats.cpu().numpy()
    want = np.array([float(total_ref.detach())] + [float(a.detach()) for a in aux])
    np.testing.assert_allclose(got[:7], want, rtol=2e-5, atol=2e-6)       # fp32 kernel vs float64 reference
    np.testing.assert_allclose(dlogits.cpu().numpy(), lr.grad.numpy(), rtol=2e-4, atol=2e-7)
    np.testing.assert_allclose(dvalue.cpu().numpy(), vr.grad.numpy(), rtol=2e-4, atol=2e-7)
    # illegal actions of a masked policy receive no actor / entropy gradient
    if masked and ill == 0.0:
        assert float(dlogits[~mask[idx].to(DEV)].abs().max()) == 0.0


def test_adam_clip_matches_optax_restatement():
    from brl_b200 import ops
    from oracle import ppo_ref
    rng = np.random.default_rng(0)
    n = 100_003
    p = rng.normal(0, 0.1, n)
    m, v = np.zeros(n), np.zeros(n)
    tp = torch.as_tensor(p, dtype=torch.float32, device=DEV)
    tm, tv = torch.zeros(n, dtype=torch.float32, device=DEV), torch.zeros(n, dtype=torch.float32, device=DEV)
    scratch = torch.zeros(1, dtype=torch.float64, device=DEV)
    p = tp.cpu().numpy().astype(np.float64)
    for step, gscale in enumerate([1.0, 1e-4, 0.3, 1e-3]):   # global norm above and below max_grad_norm = 0.5
        g = (rng.normal(0, 1, n) * gscale).astype(np.float32)
        p, m, v = ppo_ref.adam_clip_step(p, g, m, v, step, lr=1e-3, max_grad_norm=0.5)
        ops.adam_clip(tp, torch.as_tensor(g, device=DEV), tm, tv, scratch, step=step + 1, lr=1e-3, max_grad_norm=0.5)
        np.testing.assert_allclose(tm.cpu().numpy(), m, rtol=1e-5, atol=1e-9)
        np.testing.assert_allclose(tv.cpu().numpy(), v, rtol=1e-5, atol=1e-12)
        np.testing.assert_allclose(tp.cpu().numpy(), p, rtol=0, atol=2e-6)  # lr = 1e-3: 2e-6 = 0.2 % of one step


def test_gather_rows():
    from brl_b200 import ops
    for dtype, width in ((torch.float32, 480), (torch.uint8, 480), (torch.uint8, 38), (torch.bfloat16, 480), (torch.uint8, 7)):
        src = (torch.rand((500, width), device=DEV) * 200).to(dtype)
        index = torch.randperm(500, device=DEV)[:123].to(torch.int32)
        dst = torch.empty((123, width), dtype=dtype, device=DEV)
        ops.gather_rows(src, index, dst)
        assert torch.equal(dst, src[index.long()])


def _flat_order():
    return [f"{c}{i}" for i in range(6) for c in ("w", "b")]


@pytest.mark.parametrize("B,total,obs_dtype,tune", [
    (5, 60, torch.float32, 0),         # a handful of samples: one ragged row block everywhere
    (64, 200, torch.float32, 0),
    (333, 1000, torch.uint8, 0),       # ragged batch: row / K tails of every GEMM
    (333, 1000, torch.bfloat16, 4),    # tune 4 / 16: the other tile width in the fused forward / backward launch
    (1024, 3000, torch.bfloat16, 0),   # ppo.py's minibatch size; tune 0: the two fused persistent launches (default)
    (1024, 2500, torch.float32, 16),
    (2048, 5000, torch.bfloat16, 0),   # 16 row blocks: more tiles than SMs in every op
    (4096, 9000, torch.uint8, 0),      # 32 row blocks: several tiles per SM in every op, > 4096 tiles in the launch
    (777, 2500, torch.uint8, 4 | 16),  # tune 4 / 16: the other tile width in the fused forward / backward launch
])
def test_ppo_grad_matches_float64_autograd(B, total, obs_dtype, tune):
    """brl_ppo_grad (take + forward + loss + backward on tcgen05) vs float64 autograd of the restated
    _loss_fn over the restated MLP (src/update.py:91-167, src/models.py:23-33)."""
    from brl_b200 import ops
    from brl_b200.models import init_params, params_to_numpy
    from brl_b200.optim import flatten_params
    from oracle import ppo_ref
    _, _, _, mask, action, old_lp, old_v, adv, tgt = _case(B, total, 11)
    g = torch.Generator().manual_seed(5)
    obs = (torch.rand((total, 480), generator=g) < 0.05)
    params = init_params(4, DEV)
    # non-zero biases so that every bias path is exercised
    for name in params:
        params[name]["b"] = torch.randn(params[name]["b"].shape, generator=g).to(DEV) * 0.05
    # The gradient is discontinuous where a hidden pre-activation crosses zero (ReLU kink): there a 1e-6 difference
    # between the split-bf16 and the float64 forward flips one mask bit and changes that sample's gradient by O(1).
    # The minibatch is therefore drawn from rows whose pre-activations all stay 2e-5 away from zero in float64
    # (about ten times the forward's error); ~60 % of the rows qualify.
    with torch.no_grad():
        pn = {k: torch.tensor(v, dtype=torch.float64) for k, v in params_to_numpy(params).items()}
        h, margin = obs.double(), torch.full((total,), float("inf"), dtype=torch.float64)
        for i in range(4):
            z = h @ pn[f"w{i}"] + pn[f"b{i}"]
            margin = torch.minimum(margin, z.abs().min(dim=1).values)
            h = torch.relu(z)
    safe = torch.nonzero(margin > 2e-5)[:, 0]
    assert len(safe) >= B
    index = safe[torch.randperm(len(safe), generator=g)[:B]].to(torch.int32)
    cfg = dict(clip_eps=0.2, ent_coef=0.01, vf_coef=0.5, illegal_l2_coef=0.0, value_clipping=True, reward_scaling=False,
               masked_policy=True)
    idx = index.long()
    ref = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in params_to_numpy(params).items()}
    # the float64 net (as ppo_ref.mlp_forward_torch), keeping every layer's input and pre-activation
    hs, zs = [obs[idx].double()], []
    for i in range(4):
        z = hs[-1] @ ref[f"w{i}"] + ref[f"b{i}"]
        z.retain_grad()
        zs.append(z)
        hs.append(torch.relu(z))
    lg = hs[-1] @ ref["w4"] + ref["b4"]
    vl = (hs[-1] @ ref["w5"] + ref["b5"])[:, 0]
    lg.retain_grad()
    vl.retain_grad()
    with torch.no_grad():  # old log-probs near the current policy: ratios straddle the clip range
        ml = torch.where(mask[idx], lg, torch.tensor(float("-inf"), dtype=torch.float64))
        lp_now = torch.log_softmax(ml, 1).gather(1, action[idx].long()[:, None])[:, 0]
        old_lp[idx] = lp_now + (torch.rand(B, generator=g, dtype=torch.float64) - 0.5) * 0.8
        old_v[idx] = vl + torch.randn(B, generator=g, dtype=torch.float64) * 0.1
        tgt[idx] = vl + torch.randn(B, generator=g, dtype=torch.float64) * 0.2
    total_ref, aux = ppo_ref.loss_fn(lg, vl, mask[idx], action[idx], old_lp[idx], old_v[idx], adv[idx], tgt[idx],
                                     clip_eps=0.2, ent_coef=0.01, vf_coef=0.5)
    total_ref.backward()
    # Every gradient element is a sum over the batch of signed per-sample terms; a split product carries 2^-16 relative
    # error per TERM, so next to the bar on the tensor's scale there is one on sum |term| (matters where the terms
    # cancel: the value head's column, the head biases, large batches).
    dzs = [z.grad for z in zs] + [lg.grad, vl.grad[:, None]]
    ins = hs[:4] + [hs[4], hs[4]]
    bound = {}
    for i in range(6):
        bound[f"w{i}"] = (ins[i].detach().abs().T @ dzs[i].abs()).numpy().reshape(-1)
        bound[f"b{i}"] = dzs[i].abs().sum(0).numpy().reshape(-1)
    f32 = lambda t: t.to(torch.float32).to(DEV).contiguous()  # noqa: E731
    flat_p, _ = flatten_params(params)
    blob = ops.mlp_pack_train(flat_p)
    scratch = ops.mlp_train_scratch(B, DEV)
    grads = torch.full_like(flat_p, float("nan"))
    stats = torch.zeros(8, dtype=torch.float32, device=DEV)
    acc = ops.ppo_scratch(DEV)
    ops.ppo_grad(obs.to(obs_dtype).to(DEV).contiguous(), blob, scratch, index.to(DEV), mask.to(torch.uint8).to(DEV).contiguous(),
                 action.to(DEV), f32(old_lp), f32(old_v), f32(adv), f32(tgt), grads, stats, acc, tune=tune, **cfg)
    got = stats.cpu().numpy()
    want = np.array([float(total_ref.detach())] + [float(a.detach()) for a in aux])
    np.testing.assert_allclose(got[:7], want, rtol=5e-5, atol=5e-6)   # three-term bf16 split forward vs float64
    gg = grads.cpu().numpy()
    assert np.isfinite(gg).all()                                         # every gradient element was written
    # |grads|^2 for clip_by_global_norm comes out of the gradient epilogues (scratch element 14, input of brl_adam_apply)
    np.testing.assert_allclose(float(acc[14]), float((gg.astype(np.float64) ** 2).sum()), rtol=1e-5)
    off = 0
    for k in _flat_order():
        want_g = ref[k].grad.numpy().reshape(-1)
        got_g = gg[off:off + want_g.size]
        off += want_g.size
        scale = np.abs(want_g).max()
        assert scale > 0
        # fp32-class: 2e-4 of the tensor's largest gradient (the library-GEMM path's bar is 2e-4 relative)
        tol = 2e-4 * scale + 2e-5 * bound[k]
        bad = np.abs(got_g - want_g) > tol
        assert not bad.any(), (k, int(bad.sum()), float(np.abs(got_g - want_g).max()), scale)
    assert off == gg.size


def _check_train_blob(blob, params):
    """W[in, out_pad] as bf16 hi / lo (hi = bf16(w), hi + lo = w to 2^-16) and the fp32 bias, per layer."""
    from brl_b200.models import LAYERS
    off = 0
    for li, (k_in, n_pad) in enumerate(((480, 1024), (1024, 1024), (1024, 1024), (1024, 1024), (1024, 64))):
        nbytes = k_in * n_pad * 2
        hi = blob[off:off + nbytes].view(torch.bfloat16).view(k_in, n_pad).float()
        lo = blob[off + nbytes:off + 2 * nbytes].view(torch.bfloat16).view(k_in, n_pad).float()
        bias = blob[off + 2 * nbytes:off + 2 * nbytes + 4 * n_pad].view(torch.float32)
        off = (off + 2 * nbytes + 4 * n_pad + 255) & ~255
        if li < 4:
            w, b = params[LAYERS[li]]["w"], params[LAYERS[li]]["b"]
        else:  # 38 policy columns, the value column, zeros
            w = torch.zeros((1024, 64), device=DEV)
            w[:, :38] = params[LAYERS[4]]["w"]
            w[:, 38] = params[LAYERS[5]]["w"][:, 0]
            b = torch.zeros(64, device=DEV)
            b[:38] = params[LAYERS[4]]["b"]
            b[38] = params[LAYERS[5]]["b"][0]
        assert torch.equal(hi, w.to(torch.bfloat16).float()), li
        assert float((hi + lo - w).abs().max()) <= 2.0 ** -16 * float(w.abs().max()), li
        assert torch.equal(bias, b), li
    assert off == blob.numel()


def test_pack_train_and_adam_step_keep_the_blob_current():
    """brl_mlp_pack_train lays the parameters out for brl_ppo_grad; brl_mlp_adam_step = brl_adam_clip on the flat buffer
    (bit-identical parameters and moments) with the blob refreshed in the same pass."""
    from brl_b200 import ops
    from brl_b200.models import init_params
    from brl_b200.optim import flatten_params
    params = init_params(21, DEV)
    g = torch.Generator().manual_seed(1)
    for name in params:
        params[name]["b"] = torch.randn(params[name]["b"].shape, generator=g).to(DEV)
    flat_p, views = flatten_params(params)
    assert flat_p.numel() == ops.mlp_num_params() == 3681319
    blob = ops.mlp_pack_train(flat_p)
    _check_train_blob(blob, views)
    p2, m2, v2 = flat_p.clone(), torch.zeros_like(flat_p), torch.zeros_like(flat_p)
    m1, v1 = torch.zeros_like(flat_p), torch.zeros_like(flat_p)
    scratch = torch.zeros(1, dtype=torch.float64, device=DEV)
    for step, gscale in enumerate([1.0, 1e-4, 0.3]):  # global norm above and below max_grad_norm
        grads = torch.randn(flat_p.shape, generator=g).to(DEV) * gscale
        sumsq = (grads.double() ** 2).sum().reshape(1)
        ops.mlp_adam_step(flat_p, grads, m1, v1, sumsq, blob, step=step + 1, lr=1e-3, max_grad_norm=0.5)
        ops.adam_clip(p2, grads, m2, v2, scratch, step=step + 1, lr=1e-3, max_grad_norm=0.5)
        # the clip scale comes from a differently-ordered f64 sum: identical to the last bit or one ulp of the norm apart
        assert float((flat_p - p2).abs().max()) <= 1e-9 and float((m1 - m2).abs().max()) <= 1e-9 * gscale
        _check_train_blob(blob, views)   # `views` are views into flat_p: the refreshed parameters


@pytest.mark.parametrize("precision", ["fp32", "tc"])
def test_update_step_matches_float64_reference(precision):
    """src/update.py:74-242 end to end on a small rollout with an injected permutation, through the library-GEMM
    back end and through the all-kernel tensor-core back end."""
    from brl_b200.models import LAYERS, init_params, make_forward_pass, params_to_numpy
    from brl_b200.optim import AdamWithClip
    from brl_b200.roll_out import Transition
    from brl_b200.update import make_update_step
    from brl_b200 import random as brandom
    from oracle import ppo_ref
    from scripts.torch_baseline import TorchForwardPass, make_update_step_autograd
    T, n, mbs, epochs = 4, 64, 64, 2
    nmb = T * n // mbs
    config = dict(actor_illegal_action_mask=True, actor_illegal_action_penalty=False, clip_eps=0.2, ent_coef=0.001,
                  vf_coef=0.5, illegal_action_l2norm_coef=0.0, value_clipping=True, reward_scaling=False,
                  num_minibatches=nmb, minibatch_size=mbs, update_epochs=epochs, num_steps=T, num_envs=n)
    g = torch.Generator().manual_seed(3)
    obs = (torch.rand((T, n, 480), generator=g) < 0.04).to(torch.float32)
    mask = torch.rand((T, n, 38), generator=g) < 0.5
    mask[..., 0] = True
    action = torch.zeros((T, n), dtype=torch.int32)
    params = init_params(9, DEV)
    fp = TorchForwardPass("relu", "fp32")
    torch.manual_seed(3)  # the action sample below draws from the global CUDA generator
    with torch.no_grad():
        logits, value = fp.apply(params, obs.reshape(-1, 480).to(DEV))
        ml = torch.where(mask.reshape(-1, 38).to(DEV), logits, torch.tensor(float("-inf"), device=DEV))
        action = torch.distributions.Categorical(logits=ml).sample().to(torch.int32).reshape(T, n).cpu()
        lp = torch.log_softmax(ml, 1).gather(1, action.reshape(-1, 1).long().to(DEV))[:, 0].reshape(T, n).cpu()
        value = value.reshape(T, n).cpu()
    old_lp = lp + (torch.rand((T, n), generator=g) - 0.5) * 0.6
    adv = torch.randn((T, n), generator=g)
    tgt = value + torch.randn((T, n), generator=g) * 0.2
    traj = Transition(done=torch.zeros((T, n), dtype=torch.bool, device=DEV), action=action.to(DEV), value=value.to(DEV),
                      reward=torch.zeros((T, n), device=DEV), log_prob=old_lp.to(DEV), obs=obs.to(DEV),
                      legal_action_mask=mask.to(DEV))
    perms = [torch.randperm(T * n, generator=g) for _ in range(epochs)]
    it = iter(perms)
    lr = 1e-3
    opt = AdamWithClip(lr, eps=1e-5, max_grad_norm=0.5)
    if precision == "fp32":   # cuBLAS + autograd cross-check (scripts/torch_baseline.py), not a product back end
        update_step = make_update_step_autograd(config, fp, opt, permutation_fn=lambda rng, bs: next(it))
        with pytest.raises(TypeError):
            make_update_step(config, fp, opt)
    else:
        update_step = make_update_step(config, make_forward_pass("relu", "DeepMind", precision=precision), opt,
                                       permutation_fn=lambda rng, bs: next(it))
    before = {k: v.copy() for k, v in params_to_numpy(params).items()}
    runner = (params, opt.init(params), None, None, 0, brandom.PRNGKey(0))
    runner2, (total_loss, aux) = update_step(runner, traj, adv.to(DEV), tgt.to(DEV))
    # the caller's params are untouched (ppo.py keeps them as opp_params)
    for k, v in params_to_numpy(params).items():
        assert (v == before[k]).all()
    assert runner2[1].count == epochs * nmb and total_loss.shape == (epochs, nmb)
    # float64 reference of the whole update with the same permutations
    ref = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in before.items()}
    order = [f"{c}{i}" for i in range(6) for c in ("w", "b")]
    flat = lambda d: np.concatenate([d[k].detach().numpy().reshape(-1) for k in order])  # noqa: E731
    p = flat(ref)
    m, v, count = np.zeros_like(p), np.zeros_like(p), 0
    fo, fm, fa, fl, fv, fadv, ftgt = (t.reshape(T * n, *t.shape[2:]) for t in (obs, mask, action, old_lp, value, adv, tgt))
    losses = []
    for e in range(epochs):
        for mb in range(nmb):
            idx = perms[e][mb * mbs:(mb + 1) * mbs]
            for t in ref.values():
                t.grad = None
            lg, vl = ppo_ref.mlp_forward_torch(ref, fo[idx].double())
            total, _ = ppo_ref.loss_fn(lg, vl, fm[idx], fa[idx], fl[idx].double(), fv[idx].double(), fadv[idx].double(),
                                       ftgt[idx].double(), clip_eps=0.2, ent_coef=0.001, vf_coef=0.5)
            total.backward()
            grad = np.concatenate([ref[k].grad.numpy().reshape(-1) for k in order])
            p, m, v = ppo_ref.adam_clip_step(p, grad, m, v, count, lr=lr, max_grad_norm=0.5)
            count += 1
            off = 0
            with torch.no_grad():
                for k in order:
                    sz = ref[k].numel()
                    ref[k].copy_(torch.as_tensor(p[off:off + sz]).reshape(ref[k].shape))
                    off += sz
            losses.append(float(total))
    got_losses = total_loss.cpu().numpy().reshape(-1)
    # first minibatch: same parameters on both sides, fp32-vs-float64 evaluation only
    np.testing.assert_allclose(got_losses[0], losses[0], rtol=5e-4, atol=5e-5)
    # later minibatches see parameters that went through Adam, whose normalised step turns fp32 gradient
    # noise on near-zero gradients into O(lr) parameter differences: the loss may drift by O(lr * steps)
    np.testing.assert_allclose(got_losses, losses, rtol=5e-4, atol=0.1 * lr * len(losses))
    got = params_to_numpy(runner2[0])
    got_flat = np.concatenate([got[k].reshape(-1) for k in order])
    delta_ref, delta_got = p - flat({k: torch.tensor(before[k]) for k in order}), got_flat - flat({k: torch.tensor(before[k]) for k in order})
    # Adam's step is ~lr per element per step; fp32-vs-float64 gradient noise only matters where |g| ~ eps
    err = np.abs(delta_got - delta_ref)
    if precision == "fp32":
        assert np.quantile(err, 0.999) <= 0.05 * lr * count
    else:
        # Split-bf16 products carry 2^-17 relative error (fp32: 2^-24), so ~1 hidden unit per 64-sample minibatch lands on
        # the other side of its ReLU kink than in float64 (scripts/exp_grad_error.py).  Either side is a valid
        # subgradient, but the flip changes that sample's back-propagated signal through every earlier layer by ~1 %,
        # i.e. the minibatch gradient by 1e-4 .. 1e-3 relative, which Adam's normalised step passes on: the bar is 1 % of
        # the total movement at the median and 10 % at the 99th percentile (the fp32 bar above: 5 % at the 99.9th).
        assert np.median(err) <= 0.01 * lr * count, float(np.median(err))
        assert np.quantile(err, 0.99) <= 0.1 * lr * count, [float(np.quantile(err, q)) for q in (0.5, 0.9, 0.99, 0.999, 1.0)]
    assert np.abs(delta_ref).max() > 0.5 * lr   # the update actually moved the parameters


def test_illegal_action_statistic_is_opt_in_and_never_wrong():
    """brl_ppo_grad with the default coefficient 0: the logged illegal_action_loss (spectral norm) is formed when asked for
    (bit-identical to a run that asks for it) and is NaN -- not stale, not an approximation -- when not; nothing else moves."""
    from brl_b200 import ops
    from brl_b200.models import init_params
    from brl_b200.optim import flatten_params
    B, total = 512, 2000
    logits, value, index, mask, action, old_lp, old_v, adv, tgt = _case(B, total, 11)
    g = torch.Generator().manual_seed(2)
    obs = (torch.rand((total, 480), generator=g) < 0.05).to(torch.float32)
    f32 = lambda t: t.to(torch.float32).to(DEV).contiguous()  # noqa: E731
    flat_p, _ = flatten_params(init_params(3, DEV))
    blob = ops.mlp_pack_train(flat_p)
    scratch = ops.mlp_train_scratch(B, DEV)
    cfg = dict(clip_eps=0.2, ent_coef=0.01, vf_coef=0.5, illegal_l2_coef=0.0)
    out = {}
    for want in (True, False):
        grads = torch.empty_like(flat_p)
        stats = torch.full((8,), -3.0, dtype=torch.float32, device=DEV)
        acc = ops.ppo_scratch(DEV)
        ops.ppo_grad(obs.to(DEV), blob, scratch, index.to(DEV), mask.to(torch.uint8).to(DEV).contiguous(), action.to(DEV),
                     f32(old_lp), f32(old_v), f32(adv), f32(tgt), grads, stats, acc, illegal_stat=want, **cfg)
        torch.cuda.synchronize()
        out[want] = (grads.clone(), stats.cpu().numpy())
    assert torch.equal(out[True][0], out[False][0])
    assert (out[True][1][:6] == out[False][1][:6]).all()
    assert np.isnan(out[False][1][6]) and np.isfinite(out[True][1][6]) and out[True][1][6] > 0
    # and the value is the matrix 2-norm of the illegal probabilities of THIS forward's logits
    from scripts.torch_baseline import TorchForwardPass  # noqa: F401  (float64 check below needs no GEMM library)
    from brl_b200.models import LAYERS, init_params as ip
    p = ip(3, "cpu")
    h = obs[index.long()].double()
    for name in LAYERS[:4]:
        h = torch.relu(h @ p[name]["w"].double() + p[name]["b"].double())
    lg = h @ p[LAYERS[4]]["w"].double() + p[LAYERS[4]]["b"].double()
    X = torch.softmax(lg, 1) * (~mask[index.long()])
    np.testing.assert_allclose(out[True][1][6], float(torch.linalg.matrix_norm(X, 2)) / 2, rtol=5e-5)
