"""TEST INFRASTRUCTURE ONLY: NumPy restatement of the evaluation statistics of
src/evaluation.py:207-1032 (`make_evaluate`), written the way the reference computes them
(per-step `update_log_info`, per-table `make_terminated_log` / `make_contract_log`, then the
`log_info` tuple), vectorised over envs.  Pinned only by construction from the cited lines
(needs jax/distrax to run the original): "parity unpinned" for this row."""
from __future__ import annotations

import numpy as np


class EvalLog:
    """the 8 accumulators carried through the while_loop (src/evaluation.py:262-271 / 631-639)"""

    def __init__(self, n: int, duplicate: bool):
        self.duplicate = duplicate
        self.ill = np.zeros((2, n), np.float32)     # actor_total_illegal_action_probs, opp_...
        self.steps = np.zeros((2, n), np.float32)
        self.bid = np.zeros((2, n, 35), np.float32)
        self.passes = np.zeros((2, n), np.float32)

    def make_action(self, l_actor, l_opp, mask, current_player):
        """masked argmax of the acting team's logits + pi.probs of the UNMASKED logits (:236-246, 650-663);
        free-run opponent: Pass with probs one-hot at Pass (:248-252)."""
        team1 = current_player < 2
        mask = mask.astype(bool)
        if l_opp is None:
            l_opp = np.full_like(l_actor, -np.inf)
            l_opp[:, 0] = 0.0
        logits = np.where(team1[:, None], l_actor, l_opp).astype(np.float32)
        masked = np.where(mask, logits, -np.inf)
        action = masked.argmax(1).astype(np.int32)
        e = np.exp(logits - logits.max(1, keepdims=True))
        probs = e / e.sum(1, keepdims=True)
        return action, probs

    def step_log(self, probs, mask, current_player, action, terminated):
        """make_step_log / update_log_info (:311-385, 673-745)"""
        live = ~terminated.astype(bool)
        illegal = (probs * (~mask.astype(bool))).sum(1).astype(np.float32)
        for t, sel in enumerate((current_player < 2, current_player >= 2)):
            s = sel & live
            self.ill[t][s] += illegal[s]
            self.steps[t][s] += 1
            self.passes[t][s & (action == 0)] += 1
            b = s & (action >= 3)
            idx = np.nonzero(b)[0]
            if self.duplicate:
                self.bid[t][idx, action[idx] - 3] += 1                 # jnp.zeros(35).at[a-3].set(1) + bid (:699-706)
            else:
                self.bid[t][idx, action[idx] - 3] = 1                  # bid.at[a-3].set(1) (:345-354)


def table_logs(last_bid, last_bidder, call_x, call_xx, sign_source, pass_num=None):
    """make_terminated_log + make_contract_log for one table (:448-563 / :839-925).
    sign_source = table_info.rewards[:, 0] (duplicate) or cum_return (single table)."""
    n = last_bid.shape[0]
    pass_out = (last_bidder == -1) & (last_bid == -1)
    if pass_num is not None:
        pass_out &= pass_num == 4
    live = ~pass_out
    actor = live & (last_bidder < 2)
    opp = live & (last_bidder >= 2)
    actor_contract, opp_contract = np.zeros((n, 35)), np.zeros((n, 35))
    actor_contract[np.nonzero(actor)[0], last_bid[actor]] = 1
    opp_contract[np.nonzero(opp)[0], last_bid[opp]] = 1
    x, xx = call_x.astype(bool), call_xx.astype(bool)
    made = sign_source >= 0
    return dict(pass_out=pass_out, actor_contract=actor_contract, opp_contract=opp_contract,
                actor_doubled=actor & x, actor_redoubled=actor & xx, opp_doubled=opp & x, opp_redoubled=opp & xx,
                actor_make=actor & made, opp_make=opp & made, actor_down=actor & ~made, opp_down=opp & ~made)


def log_info_duplicate(log: EvalLog, cum_return, step_count, ta, tb, score_a, score_b):
    """the 23-tuple of src/evaluation.py:986-1027"""
    n = cum_return.shape[0]
    with np.errstate(divide="ignore", invalid="ignore"):
        ill = log.ill / log.steps
        pr = log.passes / log.steps
    m = lambda a, b, k: (a[k].mean(axis=0) + b[k].mean(axis=0)) / 2  # noqa: E731
    return (cum_return.mean(), cum_return.std(ddof=1) / np.sqrt(n), (score_a.mean() + score_b.mean()) / 2,
            ill[0].mean(), ill[1].mean(), step_count.mean(), log.bid[0].mean(axis=0) / 2, log.bid[1].mean(axis=0) / 2,
            m(ta, tb, "actor_contract"), m(ta, tb, "opp_contract"),
            (ta["actor_contract"].sum() / n + tb["actor_contract"].sum() / n) / 2,
            (ta["opp_contract"].sum() / n + tb["opp_contract"].sum() / n) / 2,
            m(ta, tb, "actor_doubled"), m(ta, tb, "actor_redoubled"), m(ta, tb, "opp_doubled"), m(ta, tb, "opp_redoubled"),
            m(ta, tb, "actor_make"), m(ta, tb, "opp_make"), m(ta, tb, "actor_down"), m(ta, tb, "opp_down"),
            (ta["pass_out"].sum() / n + tb["pass_out"].sum() / n) / 2, pr[0].mean(), pr[1].mean())


def log_info_single(log: EvalLog, cum_return, step_count, t):
    """the 19-tuple of src/evaluation.py:575-596"""
    n = cum_return.shape[0]
    with np.errstate(divide="ignore", invalid="ignore"):
        ill = log.ill / log.steps
    return (cum_return.mean(), ill[0].mean(), ill[1].mean(), step_count.mean(), log.bid[0].mean(axis=0),
            log.bid[1].mean(axis=0), t["actor_contract"].mean(axis=0), t["opp_contract"].mean(axis=0),
            t["actor_contract"].sum() / n, t["opp_contract"].sum() / n, t["actor_doubled"].mean(),
            t["actor_redoubled"].mean(), t["opp_doubled"].mean(), t["opp_redoubled"].mean(), t["actor_make"].mean(),
            t["opp_make"].mean(), t["actor_down"].mean(), t["opp_down"].mean(), t["pass_out"].sum() / n)
