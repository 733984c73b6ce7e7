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
