"""CPU: closed-form checks of the PPO restatement the CUDA kernels are compared with
(oracle/ppo_ref.py; src/update.py:91-162, ppo.py:195-211)."""
import numpy as np
import torch

from oracle import ppo_ref


def test_loss_closed_form_single_sample():
    # two legal actions with equal logits, third illegal: p = (.5, .5), entropy = ln 2
    logits = torch.tensor([[1.0, 1.0, 5.0] + [0.0] * 35], dtype=torch.float64)
    mask = torch.zeros((1, 38), dtype=torch.bool)
    mask[0, :2] = True
    old_lp = torch.tensor([np.log(0.25)], dtype=torch.float64)      # ratio = 0.5 / 0.25 = 2 -> clipped at 1.2 for gae > 0
    total, (vl, la, ent, kl, cf, ill) = ppo_ref.loss_fn(
        logits, torch.tensor([0.3], dtype=torch.float64), mask, torch.tensor([0]), old_lp,
        torch.tensor([0.0], dtype=torch.float64), torch.tensor([2.0], dtype=torch.float64),
        torch.tensor([1.0], dtype=torch.float64), clip_eps=0.2, ent_coef=0.001, vf_coef=0.5)
    assert abs(float(ent) - np.log(2)) < 1e-12
    assert abs(float(la) - (-1.2 * 2.0)) < 1e-12
    # value clipping: v = .3, old = 0 -> clipped to .2; max((.3-1)^2, (.2-1)^2) = .64
    assert abs(float(vl) - 0.5 * 0.64) < 1e-12
    assert float(cf) == 1.0 and abs(float(kl) - (1.0 - np.log(2.0))) < 1e-12
    assert abs(float(total) - (float(la) + 0.5 * float(vl) - 0.001 * float(ent))) < 1e-12
    # illegal mass: softmax over all 38 of [1,1,5,0...]: everything except the two legal entries
    q = torch.softmax(logits[0], 0)
    assert abs(float(ill) - float(torch.sqrt((q[2:] ** 2).sum()) / 2)) < 1e-12


def test_adam_clip_first_step_is_sign_step():
    g = np.array([3.0, -4.0, 0.0])                                    # norm 5 > 0.5 -> scaled to norm 0.5
    p, m, v = ppo_ref.adam_clip_step(np.zeros(3), g, np.zeros(3), np.zeros(3), 0, lr=0.1, max_grad_norm=0.5)
    gc = g / 5 * 0.5
    np.testing.assert_allclose(m, 0.1 * gc)
    np.testing.assert_allclose(v, 0.001 * gc * gc)
    np.testing.assert_allclose(p, -0.1 * gc / (np.abs(gc) + 1e-5), rtol=1e-12)   # bias-corrected first step
    # below the threshold the gradient is untouched
    p2, m2, _ = ppo_ref.adam_clip_step(np.zeros(3), g / 100, np.zeros(3), np.zeros(3), 0, lr=0.1, max_grad_norm=0.5)
    np.testing.assert_allclose(m2, 0.1 * g / 100)


def test_illegal_action_loss_is_the_spectral_norm_of_the_minibatch_matrix():
    """src/update.py:141-142: `jnp.linalg.norm(X, ord=2)` of the 2-D [minibatch, 38] matrix X = probs * ~mask is its
    largest singular value, NOT the Frobenius norm of the flattened matrix.  Known answer: two samples whose illegal mass
    sits on DIFFERENT single actions give orthogonal rows, so sigma_max = max row norm while Frobenius = sqrt(sum)."""
    B = 2
    logits = torch.zeros((B, 38), dtype=torch.float64)
    logits[0, 5], logits[1, 9] = 3.0, 1.0
    mask = torch.ones((B, 38), dtype=torch.bool)
    mask[0, 5] = False          # sample 0: action 5 illegal, sample 1: action 9 illegal
    mask[1, 9] = False
    z = torch.zeros(B, dtype=torch.float64)
    _, aux = ppo_ref.loss_fn(logits, z, mask, torch.zeros(B, dtype=torch.long), z, z, z, z, clip_eps=0.2, ent_coef=0.0,
                             vf_coef=0.5)
    q0 = float(torch.softmax(logits[0], 0)[5])
    q1 = float(torch.softmax(logits[1], 0)[9])
    assert abs(float(aux[5]) - max(q0, q1) / 2) < 1e-12
    assert abs(float(aux[5]) - np.sqrt(q0 * q0 + q1 * q1) / 2) > 1e-3      # and it is not the Frobenius value
    # general case against NumPy's own matrix 2-norm (numpy.linalg.norm(x, 2) == what jnp.linalg.norm mirrors)
    g = torch.Generator().manual_seed(3)
    logits = torch.randn((64, 38), generator=g, dtype=torch.float64) * 2
    mask = torch.rand((64, 38), generator=g) < 0.5
    mask[:, 0] = True
    z = torch.zeros(64, dtype=torch.float64)
    _, aux = ppo_ref.loss_fn(logits, z, mask, torch.zeros(64, dtype=torch.long), z, z, z, z, clip_eps=0.2, ent_coef=0.0,
                             vf_coef=0.5)
    X = (torch.softmax(logits, 1) * (~mask)).numpy()
    assert abs(float(aux[5]) - np.linalg.norm(X, 2) / 2) < 1e-12
    assert abs(np.linalg.norm(X, 2) - np.linalg.svd(X, compute_uv=False)[0]) < 1e-12


def _loss(logits, value, mask, action, old_lp, old_v, gae, tgt, **kw):
    f = lambda x: torch.tensor(x, dtype=torch.float64)  # noqa: E731
    cfg = dict(clip_eps=0.2, ent_coef=0.0, vf_coef=0.5)
    cfg.update(kw)
    return ppo_ref.loss_fn(f(logits), f(value), torch.tensor(mask, dtype=torch.bool), torch.tensor(action), f(old_lp), f(old_v),
                           f(gae), f(tgt), **cfg)


def test_actor_loss_clipping_by_hand_both_signs_of_the_advantage():
    """src/update.py:116-131: -min(ratio * gae, clip(ratio, 1 - eps, 1 + eps) * gae), averaged over the minibatch.
    Four samples, all with p(action) = 0.5 under the new policy (two legal actions, equal logits):
      ratio 2,   gae +1 -> min(2, 1.2)    = 1.2   (clipped: no more reward for moving further)
      ratio 2,   gae -1 -> min(-2, -1.2)  = -2    (NOT clipped: the pessimistic bound)
      ratio 0.5, gae +1 -> min(.5, .8)    = 0.5   (not clipped)
      ratio 0.5, gae -1 -> min(-.5, -.8)  = -0.8  (clipped)"""
    logits = [[0.0, 0.0] + [9.0] * 36] * 4
    mask = [[True, True] + [False] * 36] * 4
    old_lp = [np.log(0.25), np.log(0.25), np.log(1.0), np.log(1.0)]
    total, (vl, la, ent, kl, cf, _) = _loss(logits, [0.0] * 4, mask, [0, 1, 0, 1], old_lp, [0.0] * 4, [1.0, -1.0, 1.0, -1.0], [0.0] * 4)
    assert abs(float(la) - (-(1.2 - 2.0 + 0.5 - 0.8) / 4)) < 1e-12
    assert float(cf) == 1.0                                            # |ratio - 1| = 1 and .5, both > .2
    r = np.array([2.0, 2.0, 0.5, 0.5])
    assert abs(float(kl) - float(((r - 1) - np.log(r)).mean())) < 1e-12   # :153
    assert abs(float(ent) - np.log(2)) < 1e-12 and float(vl) == 0.0
    assert abs(float(total) - float(la)) < 1e-12


def test_value_loss_with_and_without_clipping_by_hand():
    """src/update.py:49-68.  old value 0, eps .2.  Sample A: v = .5, target 1: clipped prediction .2 is FURTHER from the
    target, max picks it: (.8)^2.  Sample B: v = .5, target 0: clipped prediction .2 is closer, max picks the unclipped
    (.5)^2.  value_loss = 0.5 * mean."""
    logits = [[0.0] * 38] * 2
    mask = [[True] * 38] * 2
    lp = [np.log(1 / 38)] * 2
    _, aux = _loss(logits, [0.5, 0.5], mask, [0, 0], lp, [0.0, 0.0], [0.0, 0.0], [1.0, 0.0])
    assert abs(float(aux[0]) - 0.5 * (0.64 + 0.25) / 2) < 1e-12
    _, aux = _loss(logits, [0.5, 0.5], mask, [0, 0], lp, [0.0, 0.0], [0.0, 0.0], [1.0, 0.0], value_clipping=False)
    assert abs(float(aux[0]) - 0.5 * (0.25 + 0.25) / 2) < 1e-12


def test_reward_scaling_uses_the_population_std_plus_eps():
    """src/update.py:35-36: (gae - mean) / (std + 1e-8) with jnp's default ddof = 0.  gae = (1, 3): mean 2, std 1 ->
    (-1, +1); with ratio 1 everywhere the actor loss is -mean(scaled gae) = 0, and with ratios (1.1, 1) it is
    -(1.1 * -1 + 1 * 1) / 2 = 0.05."""
    logits = [[0.0, 0.0] + [0.0] * 36] * 2
    mask = [[True, True] + [False] * 36] * 2
    _, aux = _loss(logits, [0.0] * 2, mask, [0, 0], [np.log(0.5 / 1.1), np.log(0.5)], [0.0] * 2, [1.0, 3.0], [0.0] * 2,
                   reward_scaling=True)
    assert abs(float(aux[1]) - 0.05) < 1e-7                              # 1e-8 in the denominator
    _, aux = _loss(logits, [0.0] * 2, mask, [0, 0], [np.log(0.5 / 1.1), np.log(0.5)], [0.0] * 2, [1.0, 3.0], [0.0] * 2)
    assert abs(float(aux[1]) - (-(1.1 * 1.0 + 1.0 * 3.0) / 2)) < 1e-12  # unscaled


def test_unmasked_policy_takes_log_prob_from_all_38_logits_but_entropy_from_the_legal_ones():
    """src/update.py:12-24 (no_masked_policy) vs :133-137 (the entropy always masks)."""
    logits = [[np.log(2.0), 0.0, 0.0] + [-50.0] * 35]
    mask = [[True, True, False] + [False] * 35]
    # unmasked p(action 0) = 2 / 4; masked distribution over the two legal actions = (2/3, 1/3)
    _, aux = _loss(logits, [0.0], mask, [0], [np.log(0.5)], [0.0], [1.0], [0.0], masked_policy=False)
    assert abs(float(aux[1]) - (-1.0)) < 1e-9                            # ratio 1 -> -gae
    h = -(2 / 3 * np.log(2 / 3) + 1 / 3 * np.log(1 / 3))
    assert abs(float(aux[2]) - h) < 1e-9
    _, aux = _loss(logits, [0.0], mask, [0], [np.log(0.5)], [0.0], [1.0], [0.0], masked_policy=True)
    assert abs(float(aux[1]) - (-min(4 / 3, 1.2))) < 1e-9                # masked p = 2/3 -> ratio 4/3, clipped at 1.2


def test_adam_second_step_by_hand():
    """optax.adam(lr, eps=1e-5): m, v moments, bias correction by 1 - b^t with t counted from 1, eps OUTSIDE the root
    (eps_root = 0).  Two steps with g = 1 then g = -1 on one parameter, no clipping."""
    p, m, v = ppo_ref.adam_clip_step(np.zeros(1), np.array([1.0]), np.zeros(1), np.zeros(1), 0, lr=0.1, max_grad_norm=0)
    assert abs(p[0] - (-0.1 / (1 + 1e-5))) < 1e-15
    p, m, v = ppo_ref.adam_clip_step(p, np.array([-1.0]), m, v, 1, lr=0.1, max_grad_norm=0)
    m2 = 0.9 * 0.1 - 0.1                                                 # = -0.01
    v2 = 0.999 * 0.001 + 0.001
    step = 0.1 * (m2 / (1 - 0.9 ** 2)) / (np.sqrt(v2 / (1 - 0.999 ** 2)) + 1e-5)
    assert abs(m[0] - m2) < 1e-15 and abs(v[0] - v2) < 1e-15
    assert abs(p[0] - (-0.1 / (1 + 1e-5) - step)) < 1e-15


def test_linear_learning_rate_schedule_matches_ppo_py():
    """ppo.py:186-192: lr * (1 - (count // (num_minibatches * update_epochs)) / num_updates), evaluated by optax at the
    count BEFORE the increment; and global_gradient_clipping = False drops the clip (ppo.py:203, 210)."""
    from brl_b200.optim import make_optimizer
    cfg = dict(lr=1e-3, anneal_lr=True, num_minibatches=4, update_epochs=2, num_updates=10, max_grad_norm=0.5)
    opt = make_optimizer(cfg)
    assert opt.lr_at(0) == 1e-3 and opt.lr_at(7) == 1e-3               # the whole first update (8 optimizer steps)
    assert abs(opt.lr_at(8) - 1e-3 * 0.9) < 1e-18 and abs(opt.lr_at(79) - 1e-3 * 0.1) < 1e-18
    assert opt.max_grad_norm == 0.5
    opt = make_optimizer(dict(cfg, anneal_lr=False, global_gradient_clipping=False))
    assert opt.lr_at(123) == 1e-3 and opt.max_grad_norm == 0.0
