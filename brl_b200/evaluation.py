"""Evaluation loops of src/evaluation.py: `make_simple_duplicate_evaluate` (69-204, the
eval.py path and the league probe) and `make_simple_evaluate` (11-66)."""
from __future__ import annotations

import torch

from . import dist as bdist
from . import ops
from . import random as brandom
from .duplicate import Table_info, duplicate_step
from .models import load_params, make_forward_pass
from .utils import single_play_step_two_policy_commpetitive_deterministic

_CHECK_EVERY = 8  # stepping a finished env is a zero-reward no-op, so the all-done test needs no per-step sync


class _TeamRows:
    """Per-iteration lists of the envs still playing, split by the team of the player to act (brl_team_rows).  The
    reference's vmapped loop (src/evaluation.py:124-151) runs BOTH nets on every env and keeps forwarding finished envs
    until the slowest auction ends; here each net runs on exactly the envs it decides.  The two counts come back to the
    host once per iteration (8 bytes): they size the two forward launches and are the loop's termination test."""

    def __init__(self, n: int, device):
        self.rows1 = torch.empty(n, dtype=torch.int32, device=device)
        self.rows2 = torch.empty(n, dtype=torch.int32, device=device)
        self.counts = torch.zeros(2, dtype=torch.int32, device=device)
        self.host = torch.zeros(2, dtype=torch.int32).pin_memory()
        self.forwarded = 0  # rows sent through a net so far (tests / bench: compare with iterations * 2 * n)

    def split(self, current_player, done_u8):
        ops.team_rows(current_player, done_u8, self.rows1, self.rows2, self.counts)
        self.host.copy_(self.counts, non_blocking=True)
        torch.cuda.current_stream(self.counts.device).synchronize()
        c1, c2 = int(self.host[0]), int(self.host[1])
        self.forwarded += c1 + c2
        return self.rows1[:c1], self.rows2[:c2]


def _bf16_state(env, keys):
    """env.init with the observation written directly in bf16 (the tensor-core nets' input dtype; the observation of
    these loops is read by the nets only)"""
    n = keys.shape[0]
    from .env import State
    packed, out = ops.new_state(n, env.device), ops.EnvOutputs(n, env.device, torch.bfloat16)
    ops.init(keys, env.table, packed, out)
    return State(env, packed, out)


def make_simple_evaluate(eval_env, team1_activation, team1_model_type, team2_activation, team2_model_type,
                         team2_model_path, num_eval_envs, team2_params=None):
    """src/evaluation.py:11-66: deterministic quad steps vs a fixed opponent; mean raw score."""
    actor_forward_pass = make_forward_pass(activation=team1_activation, model_type=team1_model_type)
    opp_forward_pass = make_forward_pass(activation=team2_activation, model_type=team2_model_type)
    opp_params = team2_params if team2_params is not None else load_params(team2_model_path, eval_env.device)

    def simple_evaluate(actor_params, rng, trace=None):
        """`trace` (a list, tests only) receives the four sub-step action tensors of every quad step and finally
        ("R", per-env return) -- the inputs of an oracle replay."""
        step_fn = single_play_step_two_policy_commpetitive_deterministic(
            step_fn=eval_env.step, actor_params=actor_params, actor_forward_pass=actor_forward_pass,
            opp_params=opp_params, opp_forward_pass=opp_forward_pass)
        step_fn.trace = trace
        rng_key, sub_key = brandom.split(rng)
        state = eval_env.init(eval_env.make_keys(sub_key, num_eval_envs))
        R = torch.zeros(num_eval_envs, dtype=torch.float32, device=eval_env.device)
        action = torch.empty(num_eval_envs, dtype=torch.int32, device=eval_env.device)
        r_actor = torch.empty_like(R)
        it = 0
        while True:
            actor = state.current_player.clone()
            logits, _ = actor_forward_pass.apply(actor_params, state.observation)
            ops.categorical(logits.contiguous(), state._mask_u8, action, None, sample=False)
            rng_key, _rng = brandom.split(rng_key)
            state = step_fn(state, action, _rng, out_state=state)
            ops.gather_reward(state.rewards, actor, r_actor, 1.0)
            R += r_actor
            it += 1
            if it % _CHECK_EVERY == 0 and bool(state._terminated_u8.all()):
                break
        if trace is not None:
            trace.append(("R", R.clone()))
        return R.mean()

    return simple_evaluate


def make_simple_duplicate_evaluate(eval_env, team1_activation, team1_model_type, team2_activation, team2_model_type,
                                   num_eval_envs, env_offset: int = 0):
    """src/evaluation.py:69-204.  `num_eval_envs` is THIS rank's shard; with
    torch.distributed initialised the statistics are all-reduced over ranks (one
    collective per match) so every rank returns the whole-match numbers."""
    team1_forward_pass = make_forward_pass(activation=team1_activation, model_type=team1_model_type)
    team2_forward_pass = make_forward_pass(activation=team2_activation, model_type=team2_model_type)

    def duplicate_evaluate(team1_params, team2_params, rng_key, trace=None, record=None, local_sums=None):
        """`record` (a list) receives per step (action, table_a.terminated, table_b.terminated) BEFORE the step --
        the input of board_log.match_to_board_logs.  `local_sums` (f64[8] tensor): accumulate this rank's partial
        statistics there and skip the all-reduce (the league evaluator reduces all its matches at once)."""
        step_fn = duplicate_step(eval_env.step)
        rng_key, sub_key = brandom.split(rng_key)
        state = _bf16_state(eval_env, eval_env.make_keys(sub_key, num_eval_envs, env_offset))  # :93-95
        table_a_info = Table_info.from_state(state)                                        # :97-112
        table_b_info = Table_info.from_state(state)
        dev = eval_env.device
        cum_return = torch.zeros(num_eval_envs, dtype=torch.float32, device=dev)
        action = torch.zeros(num_eval_envs, dtype=torch.int32, device=dev)  # finished envs keep a (no-op) Pass
        teams = _TeamRows(num_eval_envs, dev)
        while True:
            # :124-151 -- team 1's net decides for players 0/1, team 2's for players 2/3; finished envs are skipped
            rows1, rows2 = teams.split(state.current_player, state._terminated_u8)
            if rows1.shape[0] + rows2.shape[0] == 0:                                       # :120-122 ~terminated.all()
                break
            team1_forward_pass.act_rows(team1_params, state.observation, state._mask_u8, action, rows1)
            team2_forward_pass.act_rows(team2_params, state.observation, state._mask_u8, action, rows2)
            if trace is not None:  # tests replay the same action sequence on the oracle
                trace.append((action.clone(), None, None))
            if record is not None:
                record.append((action.cpu(), table_a_info.terminated.view(torch.uint8).cpu(),
                               table_b_info.terminated.view(torch.uint8).cpu()))
            state, table_a_info, table_b_info = step_fn(state, action, table_a_info, table_b_info)  # :164
            cum_return += state.rewards[:, 0]                                              # :167-169
        duplicate_evaluate.rows_forwarded = teams.forwarded
        if local_sums is not None:
            ops.match_stats(cum_return, local_sums)
            return None, table_a_info, table_b_info, cum_return
        sums = torch.zeros(8, dtype=torch.float64, device=dev)
        ops.match_stats(cum_return, sums)
        bdist.allreduce_sums(sums)
        mean, std_error, win_rate = bdist.stats_from_sums(sums.cpu())                      # :199-201
        log_info = (mean, std_error, win_rate)
        return log_info, table_a_info, table_b_info, cum_return

    return duplicate_evaluate


def make_league_evaluate(eval_env, activation, model_type, num_eval_envs_total: int, num_models: int):
    """The PFSP league probe of ppo.py:412-436 as ONE sharded job (BASELINE configs[4]): the learner plays a
    duplicate match against every model of the pool; the `num_eval_envs_total` envs are split into `num_models`
    equal blocks by GLOBAL env index (block m plays pool model m) and every block is sharded across the ranks, so
    the result does not depend on the rank count.  One all-reduce of f64[num_models, 8] for the whole league.
    Returns per-model (mean IMP, SE, win rate) lists -- `imp_list` / `win_rate_list` of ppo.py:414-435."""
    rank, world = bdist.rank_world()
    per_model = num_eval_envs_total // num_models
    lo, hi = bdist.shard_range(per_model, rank, world)
    matches = [make_simple_duplicate_evaluate(eval_env, activation, model_type, activation, model_type, hi - lo,
                                              env_offset=m * per_model + lo) for m in range(num_models)]

    def league_evaluate(actor_params, pool_params, rng_key):
        assert len(pool_params) == num_models
        sums = torch.zeros((num_models, 8), dtype=torch.float64, device=eval_env.device)
        for m, (match, opp) in enumerate(zip(matches, pool_params)):
            if hi > lo:
                match(actor_params, opp, rng_key, local_sums=sums[m])
        bdist.allreduce_sums(sums)
        host = sums.cpu()
        return [bdist.stats_from_sums(host[m]) for m in range(num_models)]

    return league_evaluate


# ---- full evaluation statistics (src/evaluation.py:207-1115) --------------------------------------
_S_TABLE, _S_BIDS, _S_CONTRACTS = 11, 29, 99


def _log_info_from_sums(sums, duplicate: bool):
    """The `log_info` tuple of src/evaluation.py:575-596 (evaluate) / :986-1027 (duplicate_evaluate)
    from the all-reduced partial sums of brl_eval_summary."""
    import numpy as np
    s = np.asarray(sums, dtype=np.float64)
    n = s[0]
    mean_cum = s[1] / n
    bids = s[_S_BIDS:_S_BIDS + 70] / n
    ca, cb = s[_S_CONTRACTS:_S_CONTRACTS + 70] / n, s[_S_CONTRACTS + 70:_S_CONTRACTS + 140] / n
    ta, tb = s[_S_TABLE:_S_TABLE + 9] / n, s[_S_TABLE + 9:_S_TABLE + 18] / n
    if not duplicate:
        return (mean_cum, s[4] / n, s[5] / n, s[6] / n, bids[:35], bids[35:], ca[:35], ca[35:],
                ca[:35].sum(), ca[35:].sum(), ta[1], ta[2], ta[3], ta[4], ta[5], ta[6], ta[7], ta[8], ta[0])
    var = max(0.0, (s[2] - s[1] * s[1] / n) / (n - 1)) if n > 1 else float("nan")
    std_error = np.sqrt(var) / np.sqrt(n)
    half = lambda x, y: (x + y) / 2  # noqa: E731
    return (mean_cum, std_error, half(s[9] / n, s[10] / n), s[4] / n, s[5] / n, s[6] / n, bids[:35] / 2, bids[35:] / 2,
            half(ca[:35], cb[:35]), half(ca[35:], cb[35:]), half(ca[:35].sum(), cb[:35].sum()),
            half(ca[35:].sum(), cb[35:].sum()), half(ta[1], tb[1]), half(ta[2], tb[2]), half(ta[3], tb[3]),
            half(ta[4], tb[4]), half(ta[5], tb[5]), half(ta[6], tb[6]), half(ta[7], tb[7]), half(ta[8], tb[8]),
            half(ta[0], tb[0]), s[7] / n, s[8] / n)


def make_evaluate(eval_env, team1_activation, team1_model_type, team2_activation, team2_model_type, team2_model_path,
                  num_eval_envs, game_mode, duplicate=False, team2_params=None, env_offset: int = 0):
    """src/evaluation.py:207-1032.  Team 1 (`actor_params`, players 0/1) against a fixed team 2 (the pickle at
    `team2_model_path`, or always-pass in "free-run"); every decision is the masked argmax.  Besides the score it
    logs, per team, the unmasked probability mass on illegal actions, bid / contract histograms and declarer /
    double / make / down / pass-out ratios.  `num_eval_envs` is this rank's shard; the statistics are all-reduced
    (one collective) when torch.distributed is initialised."""
    actor_forward_pass = make_forward_pass(activation=team1_activation, model_type=team1_model_type)
    opp_forward_pass = make_forward_pass(activation=team2_activation, model_type=team2_model_type)
    if game_mode == "competitive":
        opp_params = team2_params if team2_params is not None else load_params(team2_model_path, eval_env.device)
    elif game_mode == "free-run":
        opp_params = None
    else:
        raise ValueError(game_mode)
    n, dev = num_eval_envs, eval_env.device
    n_sums = ops._lib.load().brl_eval_num_sums()

    def _loop(actor_params, rng_key, step, state, trace):
        acc = torch.zeros((n, ops._lib.EVAL_ACC_COLS), dtype=torch.float32, device=dev)
        cum_return = torch.zeros(n, dtype=torch.float32, device=dev)
        rewards = torch.zeros((n, 4), dtype=torch.float32, device=dev)
        action = torch.zeros(n, dtype=torch.int32, device=dev)
        # each live env's row holds the logits of the team that acts there (scattered by the listed forward)
        logits = torch.zeros((n, ops.NUM_ACTIONS), dtype=torch.float32, device=dev)
        teams = _TeamRows(n, dev)
        while True:
            rows1, rows2 = teams.split(state.current_player, state._terminated_u8)
            if rows1.shape[0] + rows2.shape[0] == 0:
                return state, acc, cum_return, rewards
            actor_forward_pass.act_rows(actor_params, state.observation, state._mask_u8, action, rows1, logits=logits)
            if opp_params is not None:
                opp_forward_pass.act_rows(opp_params, state.observation, state._mask_u8, action, rows2, logits=logits)
            # the step's decision + update_log_info (unmasked-softmax illegal mass, bid histograms) from the acting
            # team's logits; free-run: team 2 always passes
            ops.eval_act_log(logits, logits if opp_params is not None else None, state._mask_u8, state.current_player,
                             state._terminated_u8, action, acc, indicator_bids=not duplicate)
            if trace is not None:
                lg = logits.clone()  # row i = the logits of the team acting in env i
                trace.append((action.clone(), lg, lg if opp_params is not None else None))
            state = step(state, action)
            rewards += state.rewards
            cum_return += state.rewards[:, 0]

    def evaluate(actor_params, rng_key, trace=None):
        rng_key, sub_key = brandom.split(rng_key)
        state = _bf16_state(eval_env, eval_env.make_keys(sub_key, n, env_offset))
        state, acc, cum_return, rewards = _loop(actor_params, rng_key, lambda s, a: eval_env.step(s, a, inplace=True),
                                                state, trace)
        f = ops.state_fields(state._packed)
        sums = torch.zeros(n_sums, dtype=torch.float64, device=dev)
        ops.eval_summary(acc, cum_return, f["step_count"],
                         (f["last_bid"], f["last_bidder"], f["call_x"], f["call_xx"], None, f["pass_num"]), None, sums)
        bdist.allreduce_sums(sums)
        state.rewards.copy_(rewards)                                                   # state.replace(rewards=rewards), :574
        return state, _log_info_from_sums(sums.cpu().numpy(), duplicate=False)

    def duplicate_evaluate(actor_params, rng_key, trace=None):
        step_fn = duplicate_step(eval_env.step)
        rng_key, sub_key = brandom.split(rng_key)
        state = _bf16_state(eval_env, eval_env.make_keys(sub_key, n, env_offset))
        infos = [Table_info.from_state(state), Table_info.from_state(state)]

        def step(s, a):
            s, infos[0], infos[1] = step_fn(s, a, infos[0], infos[1])
            return s

        state, acc, cum_return, _ = _loop(actor_params, rng_key, step, state, trace)
        f = ops.state_fields(state._packed)
        sums = torch.zeros(n_sums, dtype=torch.float64, device=dev)
        tv = lambda t: (t.last_bid, t.last_bidder, t.call_x.view(torch.uint8), t.call_xx.view(torch.uint8), t.rewards, None)  # noqa: E731
        ops.eval_summary(acc, cum_return, f["step_count"], tv(infos[0]), tv(infos[1]), sums)
        bdist.allreduce_sums(sums)
        return _log_info_from_sums(sums.cpu().numpy(), duplicate=True), infos[0], infos[1]

    return duplicate_evaluate if duplicate else evaluate


_LOG_KEYS_DUP = ("eval/IMP_reward", "eval/IMP_SE", "eval/score_reward", "eval/actor_illegal_action_probs",
                 "eval/opp_illegal_action_probs", "eval/step count", None, None, None, None, "eval/actor_declarer_ratio",
                 "eval/opp_declarer_ratio", "eval/actor_doubled_ratio", "eval/actor_redoubled_ratio", "eval/opp_doubled_ratio",
                 "eval/opp_redoubled_ratio", "eval/actor_make_contract_ratio", "eval/opp_make_contract_ratio",
                 "eval/actor_down_contract_ratio", "eval/opp_down_contract_ratio", "eval/pass_out_ratio",
                 "eval/actor_pass_ratio", "eval/opp_pass_ratio")


def make_evaluate_log(log_info):
    """src/evaluation.py:1035-1115: the wandb dict of a duplicate evaluation (same keys)."""
    if len(log_info) != len(_LOG_KEYS_DUP):
        raise ValueError("make_evaluate_log expects the 23-entry log_info of duplicate_evaluate")
    log = {k: v for k, v in zip(_LOG_KEYS_DUP, log_info) if k is not None}
    actor_bid, opp_bid, actor_contract, opp_contract = log_info[6:10]
    groups = (("actor_bid_probs", actor_bid), ("actor_contract_probs", actor_contract), ("opp_bid_probs", opp_bid),
              ("opp_contract_probs", opp_contract))
    for name, vec in groups:
        index = 0
        for number in range(1, 8):
            for suit in ("C", "D", "H", "S", "NT"):
                log[f"eval/{name}/{number}{suit}"] = vec[index]
                index += 1
    return log
