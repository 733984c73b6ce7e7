/*
 * brl_b200.h -- C ABI of the B200-native bridge-bidding environment path.
 *
 * The reference (harukaki/brl) has NO native/FFI layer: its boundary for this
 * path is the pgx 1.4.0 `bridge_bidding` Env Python surface used under
 * jax.jit / vmap / lax.scan.  A jitted JAX program can only reach
 * CUDA through an XLA custom call, so every entry point below has exactly the
 * legacy XLA GPU custom-call shape
 *
 *     int32_t op(cudaStream_t stream, void **buffers,
 *                const void *opaque, size_t opaque_len);
 *
 * `buffers` = device pointers, inputs first then outputs, in the order listed
 * per op; `opaque` = one POD `BrlParams`.  The typed XLA FFI
 * (XLA_FFI_DEFINE_HANDLER) wraps the same symbols (csrc/xla_ffi_shim.cc,
 * INTEGRATION.md).  Rules for every op:
 *   - the caller owns every buffer; nothing is allocated, no stream is created,
 *     no host sync happens: the call only enqueues kernels on `stream` and is
 *     capturable in a CUDA graph / XLA command buffer;
 *   - optional outputs may be NULL (skipped); `state_in` may alias `state_out`
 *     (XLA input_output_aliases) -- every env is read before it is written;
 *   - re-entrant, no global mutable state;
 *   - returns 0 on success, a negative BRL_E_* on a usage error (bad opaque
 *     size, NULL required buffer, launch failure); `brl_last_error()` gives the
 *     message (thread-local).  Illegal ACTIONS are data, not errors: they
 *     terminate the env.
 * There is no CPU fallback anywhere behind this ABI.
 *
 * Reference call sites each op replaces (paths relative to /root/reference):
 *   brl_init              env.init(key)                src/evaluation.py:95, ppo.py:305
 *   brl_step              env.step(state, action)      src/roll_out.py:51, src/duplicate.py:149
 *                         (+ BRL_F_AUTORESET: auto_reset wrapper src/utils.py:9-58)
 *   brl_duplicate_step    duplicate_step(env.step)     src/duplicate.py:147-192
 *   brl_duplicate_init    duplicate_init(state)        src/duplicate.py:132-135
 *   brl_observe           _observe(state, player)      src/duplicate.py:6,134
 *   brl_legal_mask        state.legal_action_mask      src/roll_out.py:78
 *   brl_rollout_random    act_randomly + auto_reset step, K steps per launch
 *                                                      src/duplicate.py:7,234 ; src/utils.py:9-58
 *   brl_imp_reward        _imp_reward                  src/duplicate.py:15-70
 *   brl_gae               calc_gae reverse scan        src/gae.py:20-39
 *   brl_categorical       masked distrax.Categorical   src/roll_out.py:27-30,79-81 ; src/evaluation.py:128-133
 *   brl_match_stats       mean / SE / win-rate sums    src/evaluation.py:199-201
 *   brl_state_fields      reads of State._dealer, ._last_bid, ... src/evaluation.py:97-112
 *   brl_reset_fields      State(...) construction / state.replace(...) src/duplicate.py:120-128
 *   brl_mlp_forward       forward.apply(params, obs) -> (logits, value)  src/models.py:23-33,
 *                         src/roll_out.py:73-76, src/utils.py:78-82, src/evaluation.py:124-127
 *   brl_policy_act        forward.apply + masked Categorical sample / mode + log_prob, fused   src/roll_out.py:73-81
 *   brl_ppo_loss          _loss_fn + jax.value_and_grad w.r.t. the net outputs   src/update.py:91-167
 *   brl_adam_clip         optimizer.update + optax.apply_updates                  src/update.py:168-169, ppo.py:195-211
 *   brl_adam_apply        the same with |grads|^2 taken from brl_ppo_grad (no pass over the gradient)
 *   brl_gather_rows       minibatch take(permutation)                             src/update.py:194-199
 *   brl_ppo_grad          jax.value_and_grad(_loss_fn)(params, traj, gae, targets) -> (loss_info, grads): minibatch
 *                         take + forward + loss + backward through the MLP on tcgen05   src/update.py:91-167,194-199
 *   brl_mlp_pack_train    flat fp32 params -> device layout for brl_ppo_grad
 *   brl_mlp_adam_step     optimizer.update + apply_updates + refresh of that layout, one pass   src/update.py:168-169
 *   brl_eval_act_log      make_action + make_step_log of evaluate / duplicate_evaluate   src/evaluation.py:236-385, 650-745
 *   brl_eval_summary      make_terminated_log + the log_info means                        src/evaluation.py:448-596, 839-1027
 *   brl_mlp_pack          the params pytree (bridge_models/<name>.pkl, ppo.py:351-362) -> device layout
 */
#ifndef BRL_B200_H
#define BRL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef BRL_CUDA_STREAM_T
#define BRL_CUDA_STREAM_T
typedef struct CUstream_st *brl_stream_t; /* == cudaStream_t */
#endif

#define BRL_ABI_VERSION 2
#define BRL_NUM_ACTIONS 38      /* 0 Pass, 1 X, 2 XX, 3..37 = 1C..7NT (src/duplicate.py:9-12) */
#define BRL_OBS_DIM 480         /* ppo.py:241, wb5/utils.py:48-52 */
#define BRL_NUM_PLAYERS 4
#define BRL_DEAL_ROW_BYTES 48   /* brl_b200/deals.py */
#define BRL_STATE_PLANES 5      /* packed state = uint4[BRL_STATE_PLANES][state_stride] */
#define BRL_STATE_BYTES_PER_ENV 80

/* flags */
#define BRL_F_AUTORESET     0x0001 /* src/utils.py:9-58 around the step */
#define BRL_F_RANDOM_ACTION 0x0002 /* draw a uniform random-legal action in-kernel (act_randomly) */
#define BRL_F_ACCUMULATE    0x0004 /* rewards += , terminated |= (quad step, src/utils.py:126-128) */
#define BRL_F_OBS_U8        0x0010 /* observation as bool/u8 (pgx dtype) instead of f32 */
#define BRL_F_OBS_BF16      0x0020 /* observation as bf16 (feeds a bf16 policy GEMM) */
#define BRL_F_SAMPLE        0x0040 /* brl_categorical: Gumbel-argmax sample instead of mode */
#define BRL_F_QUAD_LAST     0x0100 /* with ACCUMULATE: write the OR-ed terminated flag back into the state (src/utils.py:128) */
#define BRL_F_MLP_BF16      0x0200 /* brl_mlp_forward: one bf16 product per term instead of the 3-term split */
#define BRL_F_RESULT_I16    0x1000 /* brl_rollout_random: also write the compact result i16[K,n] into buffers[10] */
#define BRL_F_UNIFORM_U16   0x2000 /* brl_rollout_random / brl_env_rollout_host_compact*: caller-supplied uniforms are u16[K,n] */
#define BRL_F_SEED_SALT     0x4000 /* brl_policy_act: buffers[8] = device u64 XORed into `seed` (replayed CUDA graphs draw fresh noise) */
#define BRL_F_COUNT_DONE    0x8000 /* brl_gather_reward: buffers[3] u8 done[n], [4] inout u64 count += sum(done) (src/roll_out.py:85) */
#define BRL_F_HOST_STAGED   0x0800 /* brl_env_create: never write results straight into pinned host buffers (always stage + copy) */

/* tuning (0 = automatic): bits 16-17 envs per block of the one-launch kernels (1->8, 2->16, 3->32), bits 18-19 warps
 * per block (1->1, 2->2, 3->4; 3 with bit 26: 8) */
#define BRL_F_TUNE_EPW(code) ((code) << 16)
#define BRL_F_TUNE_WPB(code) ((code) << 18)
#define BRL_F_TUNE_WPB8 ((3 << 18) | (1 << 26))
#define BRL_F_TUNE_CLASSIC_ROLLOUT (1 << 20)     /* tile-per-warp rollout kernel instead of the warp-specialised one */
#define BRL_F_TUNE_WRITERS(n) ((n) << 21)       /* writer warps (1..7) of the warp-specialised rollout; EPW bits = its envs per block */

/* errors */
#define BRL_OK 0
#define BRL_E_OPAQUE (-1)   /* opaque_len != sizeof(BrlParams) or bad field */
#define BRL_E_BUFFER (-2)   /* required buffer is NULL / misaligned */
#define BRL_E_LAUNCH (-3)   /* CUDA launch error (message in brl_last_error) */
#define BRL_E_HANDLE (-4)   /* bad BrlEnv handle */

typedef struct BrlParams {
    int64_t n_envs;        /* envs in this call (this rank's shard); 0 = an empty batch: every batched op returns BRL_OK
                              without touching (or requiring) its buffers -- brl_team_rows still zeroes its counts */
    int64_t env_offset;    /* global index of env 0 (RNG counters use global indices)  */
    int64_t state_stride;  /* envs per plane of the packed state buffers (>= n_envs)   */
    uint64_t seed;         /* key of the counter-based action / Gumbel RNG             */
    int32_t n_deals;       /* rows in the deal table                                   */
    int32_t flags;         /* BRL_F_*                                                  */
    uint32_t step;         /* step counter of the action RNG (rollout: first step)     */
    int32_t k_steps;       /* rollout length K / GAE horizon T                         */
    float illegal_penalty; /* reward of a player making an illegal call (pgx: -1)      */
    float illegal_bonus;   /* reward of the three others                               */
    float gamma;           /* GAE                                                      */
    float gae_lambda;      /* GAE                                                      */
} BrlParams;

typedef int32_t (*brl_op_fn)(brl_stream_t, void **, const void *, size_t);

/* thread-local message of the last failing call on this thread ("" if none) */
const char *brl_last_error(void);
int32_t brl_abi_version(void);

/* buffers: [0] out u64 keys[n]
 * key_i = philox(seed, env_offset + i): mirrors jax.random.split(key, n). */
int32_t brl_make_keys(brl_stream_t, void **buffers, const void *opaque, size_t opaque_len);

/* buffers: [0] in u64 keys[n]  [1] in deal_table
 *          [2] out state  [3] out obs[n,480]  [4] out u8 mask[n,38]  [5] out f32 rewards[n,4]
 *          [6] out u8 terminated[n]  [7] out i8 current_player[n] */
int32_t brl_init(brl_stream_t, void **buffers, const void *opaque, size_t opaque_len);

/* Place envs on GIVEN episode draws.
 * buffers: [0] in i32 deal[n]  [1] in i32 dealer[n]  [2] in u8 vul_ns[n]  [3] in u8 vul_ew[n]
 *          [4] in i8 shuffled_players[n,4]  [5] in u64 rng_key[n] (NULL -> 0)  [6] in deal_table
 *          [7..12] out as brl_init [2..7] */
int32_t brl_reset_fields(brl_stream_t, void **buffers, const void *opaque, size_t opaque_len);

/* buffers: [0] in state  [1] in i32 action[n] (ignored with BRL_F_RANDOM_ACTION)  [2] in deal_table
 *          [3] out state  [4] out obs  [5] out mask  [6] out rewards  [7] out terminated
 *          [8] out current_player  [9] out i32 action_taken[n] (optional) */
int32_t brl_step(brl_stream_t, void **buffers, const void *opaque, size_t opaque_len);

/* Table_info = SoA {u8 terminated[n], f32 rewards[n,4], i32 last_bid[n], i32 last_bidder[n],
 *                   u8 call_x[n], u8 call_xx[n]}  (src/duplicate.py:138-144), updated in place.
 * buffers: [0] in state  [1] in action  [2] in deal_table
 *          [3..8] inout table A info  [9..14] inout table B info
 *          [15] out state  [16] out obs  [17] out mask  [18] out rewards  [19] out terminated
 *          [20] out current_player */
int32_t brl_duplicate_step(brl_stream_t, void **buffers, const void *opaque, size_t opaque_len);

/* buffers: [0] in state  [1] in deal_table  [2..7] out as brl_init [2..7] */
int32_t brl_duplicate_init(brl_stream_t, void **buffers, const void *opaque, size_t opaque_len);

/* buffers: [0] in state  [1] in i8 player_id[n] (NULL -> state.current_player)  [2] in deal_table
 *          [3] out obs */
int32_t brl_observe(brl_stream_t, void **buffers, const void *opaque, size_t opaque_len);

/* buffers: [0] in state  [1] out u8 mask[n,38] */
int32_t brl_legal_mask(brl_stream_t, void **buffers, const void *opaque, size_t opaque_len);

/* K auto-reset steps with in-kernel random-legal actions; every output is a
 * trajectory [K, n, ...] (any may be NULL); state updated in place.
 * buffers: [0] inout state  [1] in deal_table
 *          [2] out obs[K,n,480]  [3] out mask[K,n,38]  [4] out rewards[K,n,4]  [5] out terminated[K,n]
 *          [6] out current_player[K,n]  [7] out i32 action[K,n]
 *          [8] inout u64 stats[4] = {terminal steps, sum reward[player 0] (two's complement), calls, 0}
 *          [9] in u32 uniforms[K,n] (optional): caller-owned randomness; action = the
 *              mulhi(u, #legal)-th legal action.  NULL -> Philox(seed, env_offset+i, step+s).
 *              With BRL_F_UNIFORM_U16 the buffer is u16[K,n] and stands for the u32 uniform u16 << 16.
 *          [10] out i16 result[K,n] -- read ONLY with BRL_F_RESULT_I16: the step result in 2 bytes,
 *              result = 2 * rewards[player 0] + terminated.  Lossless on this path: pgx rewards are
 *              s * [+1,+1,-1,-1] by player id (players 0/1 partners) with s an integer duplicate score,
 *              |s| <= 7600, and random-legal play never takes the illegal-action branch.  Decode:
 *              terminated = result & 1, s = result >> 1 (arithmetic) -- brl_result16_decode. */
int32_t brl_rollout_random(brl_stream_t, void **buffers, const void *opaque, size_t opaque_len);

/* buffers: [0] in f32 a_rewards[n,4]  [1] in f32 b_rewards[n,4]  [2] out f32 imp[n,4] */
int32_t brl_imp_reward(brl_stream_t, void **buffers, const void *opaque, size_t opaque_len);

/* T = k_steps; all arrays time-major [T, n].
 * buffers: [0] in u8 done  [1] in f32 value  [2] in f32 reward  [3] in f32 last_val[n]
 *          [4] out f32 advantages  [5] out f32 targets */
int32_t brl_gae(brl_stream_t, void **buffers, const void *opaque, size_t opaque_len);

/* buffers: [0] in f32 logits[n,38]  [1] in u8 mask[n,38] (NULL -> unmasked)
 *          [2] out i32 action[n]  [3] out f32 log_prob[n] */
int32_t brl_categorical(brl_stream_t, void **buffers, const void *opaque, size_t opaque_len);

/* buffers: [0] in f32 x[n] (per-env return)  [1] inout f64 sums[8] += {n, sum x, sum x^2, #(x>0), 0,0,0,0} */
int32_t brl_match_stats(brl_stream_t, void **buffers, const void *opaque, size_t opaque_len);

/* Unpack private fields (any output may be NULL).
 * buffers: [0] in state  [1] i32 deal  [2] i32 dealer  [3] i8 shuffled_players[n,4]  [4] u8 vul[n,2]
 *          [5] i32 last_bid  [6] i32 last_bidder  [7] u8 call_x  [8] u8 call_xx  [9] i32 pass_num
 *          [10] i32 step_count  [11] u64 rng_key */
int32_t brl_state_fields(brl_stream_t, void **buffers, const void *opaque, size_t opaque_len);

/* Transition bookkeeping of roll_out._env_step (src/roll_out.py:85-94):
 * buffers: [0] in f32 rewards[n,4]  [1] in i8 actor[n]  [2] out f32 reward[n] = rewards[actor] / scale
 *          with BRL_F_COUNT_DONE: [3] in u8 done[n]  [4] inout u64 terminated_count[1] += sum(done)
 * scale passed in BrlParams.gamma. */
int32_t brl_gather_reward(brl_stream_t, void **buffers, const void *opaque, size_t opaque_len);

/* ---- policy / value net on the tensor cores (src/models.py:23-33, "DeepMind" 4x1024 ReLU) ----
 * Weights are packed once (transposed to K-major bf16 hi/lo pairs); the forward is five
 * TMA + tcgen05 GEMM launches with bias/ReLU/requantisation fused into the epilogues.
 * Default arithmetic: 3-term bf16 split accumulated in fp32 (fp32-class results);
 * BRL_F_MLP_BF16 = single bf16 product. */
int64_t brl_mlp_packed_bytes(void);
int64_t brl_mlp_scratch_bytes(int64_t n_envs);
/* buffers: [0..5] in f32 w_l[in,out] (haiku `linear`, `linear_1`..`linear_5`; y = x @ w + b)
 *          [6..11] in f32 b_l[out]   [12] out packed parameters (brl_mlp_packed_bytes()) */
int32_t brl_mlp_pack(brl_stream_t, void **buffers, const void *opaque, size_t opaque_len);
/* buffers: [0] in obs[n,480] f32 (or u8 with BRL_F_OBS_U8)  [1] out bf16 obs[n,480] */
int32_t brl_obs_to_bf16(brl_stream_t, void **buffers, const void *opaque, size_t opaque_len);
/* buffers: [0] in bf16 obs[n,480]  [1] in packed parameters  [2] scratch (brl_mlp_scratch_bytes(n))
 *          [3] out f32 logits[n,38]  [4] out f32 value[n] */
int32_t brl_mlp_forward(brl_stream_t, void **buffers, const void *opaque, size_t opaque_len);
/* forward + masked categorical in one call -- `logits, value = forward.apply(params, obs); pi = Categorical(where(mask,
 * logits, -inf)); action = pi.sample(seed) | pi.mode(); log_prob = pi.log_prob(action)` (src/roll_out.py:73-81,
 * src/utils.py:78-88,150-160, src/evaluation.py:124-133).  From 4096 envs on this is ONE persistent launch: the thread
 * that drains a head-tile row holds its 38 logits and samples in place (same Philox noise as brl_categorical).
 * buffers: [0] in bf16 obs[n,480]  [1] in packed parameters  [2] scratch  [3] in u8 mask[n,38] (NULL -> unmasked)
 *          [4] out i32 action[n]  [5] out f32 log_prob[n] (NULL ok)  [6] out f32 value[n] (NULL ok)
 *          [7] out f32 logits[n,38] (NULL ok)  [8] in u64 seed_salt[1] -- read ONLY with BRL_F_SEED_SALT
 * params: flags BRL_F_SAMPLE / BRL_F_MLP_BF16, seed, env_offset, step as for brl_categorical. */
int32_t brl_policy_act(brl_stream_t, void **buffers, const void *opaque, size_t opaque_len);
/* brl_policy_act on a LIST of envs: n_envs = number of listed rows.  obs / mask / action / log_prob are the full per-env
 * arrays; only the listed envs are read and written.  This is how the evaluation loops (src/evaluation.py:124-151) avoid
 * the reference's vmap waste of running both teams' nets on every env and of forwarding finished envs.
 * buffers: [0] in bf16 obs[total,480]  [1] in packed parameters  [2] scratch (brl_mlp_rows_scratch_bytes(n_envs))
 *          [3] in u8 mask[total,38] (NULL -> unmasked)  [4] out i32 action[total]  [5] in i32 row_index[n_envs]
 *          [6] out f32 log_prob[total] (NULL ok)  [7] out f32 logits[total,38] (NULL ok; rows of the listed envs)
 * The Gumbel noise of BRL_F_SAMPLE is keyed by env_offset + row_index[r], i.e. identical to the unlisted call. */
int64_t brl_mlp_rows_scratch_bytes(int64_t n_rows);
int32_t brl_policy_act_rows(brl_stream_t, void **buffers, const void *opaque, size_t opaque_len);
/* Rows of the envs still playing, split by the team of the player to act (players 0/1 = team 1, 2/3 = team 2).
 * buffers: [0] in i8 current_player[n]  [1] in u8 done[n] (NULL -> all live)  [2] out i32 rows_team1[n]
 *          [3] out i32 rows_team2[n]  [4] out i32 counts[2] */
int32_t brl_team_rows(brl_stream_t, void **buffers, const void *opaque, size_t opaque_len);

/* ---- PPO update, non-GEMM part (src/update.py:74-242, ppo.py:195-211) ----------------------- */
#define BRL_PPO_VALUE_CLIPPING  0x1 /* config["value_clipping"]  (src/update.py:47-62) */
#define BRL_PPO_REWARD_SCALING  0x2 /* config["reward_scaling"]: (gae - mean) / (std + 1e-8) per minibatch (src/update.py:31-45) */
#define BRL_PPO_UNMASKED_POLICY 0x4 /* actor_illegal_action_penalty mode: log-prob from the unmasked softmax (src/update.py:18-24) */
#define BRL_PPO_ILLEGAL_STAT    0x8 /* brl_ppo_grad with illegal_l2_coef == 0: also form the logged illegal_action_loss (a 38 x 38
                                       eigen-problem per minibatch, ~19 us); clear: stats[6] = NaN.  With a non-zero coefficient
                                       and in brl_ppo_loss the norm is always formed (the gradient needs it). */

typedef struct BrlPpoParams {
    int64_t batch;         /* samples in this minibatch                                  */
    int64_t total;         /* rows of the flat [T * n_envs] trajectory the index addresses */
    float clip_eps;        /* config["clip_eps"]                                         */
    float ent_coef;        /* config["ent_coef"]                                         */
    float vf_coef;         /* config["vf_coef"]                                          */
    float illegal_l2_coef; /* config["illegal_action_l2norm_coef"]                       */
    int32_t flags;         /* BRL_PPO_*                                                  */
    int32_t reserved;      /* 0.  brl_ppo_grad reads experiment bits here (0 = measured defaults): 4 / 16 = the other tile
                              width (128x128 / 128x64) in the fused forward / backward launch, 8 = per-tile time stamps
                              into the scratch (brl_mlp_train_trace_offset) */
} BrlPpoParams;

#define BRL_PPO_SCRATCH_BYTES 188416 /* scratch of brl_ppo_loss / brl_ppo_grad (buffers[12]) */

typedef struct BrlAdamParams {
    int64_t n;             /* elements of the flat parameter buffer                      */
    int32_t step;          /* 1-based optimizer step count (bias correction)             */
    float lr;              /* learning rate for THIS step (schedules are evaluated by the caller, ppo.py:186-192) */
    float beta1, beta2, eps; /* optax.adam defaults 0.9 / 0.999, eps = 1e-5 (ppo.py:198) */
    float max_grad_norm;   /* optax.clip_by_global_norm; <= 0 disables                   */
} BrlAdamParams;

/* _loss_fn of src/update.py:91-162 from the logits / value the net produced for one minibatch, fused
 * with its backward: d total_loss / d logits and d total_loss / d value come out of the same pass.
 * opaque = BrlPpoParams.
 * buffers: [0] in f32 logits[B,38]  [1] in f32 value[B]  [2] in i32 index[B] (row of each sample in the
 *          flat trajectory; NULL = identity)  [3] in u8 mask[total,38]  [4] in i32 action[total]
 *          [5] in f32 old_log_prob[total]  [6] in f32 old_value[total]  [7] in f32 advantages[total]
 *          [8] in f32 targets[total]  [9] out f32 dlogits[B,38]  [10] out f32 dvalue[B]
 *          [11] out f32 stats[8] = {total_loss, value_loss, loss_actor, entropy, approx_kl, clipfracs,
 *               illegal_action_loss, 0}  [12] scratch, BRL_PPO_SCRATCH_BYTES, 16-byte aligned (f64[16] accumulators, then the
 *               38 x 38 Gram partials of the illegal-probability matrix: illegal_action_loss is its SPECTRAL norm / 2,
 *               jnp.linalg.norm(X, ord=2) of a 2-D array, src/update.py:141-142) */
int32_t brl_ppo_loss(brl_stream_t, void **buffers, const void *opaque, size_t opaque_len);
/* optax.chain(clip_by_global_norm(max_grad_norm), adam(lr, eps)) over one flat fp32 buffer.  opaque = BrlAdamParams.
 * buffers: [0] inout f32 params[n]  [1] in f32 grads[n]  [2] inout f32 m[n]  [3] inout f32 v[n]  [4] scratch f64[1] */
int32_t brl_adam_clip(brl_stream_t, void **buffers, const void *opaque, size_t opaque_len);
/* brl_adam_clip with the gradient's sum of squares supplied by the caller: [4] in f64[1] sum g^2 (brl_ppo_grad leaves it in
 * element 14 of its f64[16] scratch, accumulated by the weight- and bias-gradient epilogues), no extra pass over the gradient. */
int32_t brl_adam_apply(brl_stream_t, void **buffers, const void *opaque, size_t opaque_len);
/* minibatch gather (jnp.take(x, permutation, axis=0), src/update.py:194-199): n_envs = B rows, k_steps = bytes per row.
 * buffers: [0] in src[total, row]  [1] in i32 index[B]  [2] out dst[B, row] */
int32_t brl_gather_rows(brl_stream_t, void **buffers, const void *opaque, size_t opaque_len);

/* ---- PPO update, the trip through the net (src/update.py:91-167): forward + backward on tcgen05 ---- */
#define BRL_PPO_OBS_U8   0x10 /* brl_ppo_grad: trajectory observations are u8 / bool 0-1 (pgx dtype) */
#define BRL_PPO_OBS_BF16 0x20 /* brl_ppo_grad: trajectory observations are bf16 (what the tensor-core rollout records) */
/* Flat fp32 parameter / gradient buffer: haiku modules actor_critic/linear, linear_1 .. linear_5 in that order, each
 * w[in,out] then b[out] (480x1024, 3 x 1024x1024, 1024x38, 1024x1): brl_mlp_num_params() = 3,681,319 floats. */
int64_t brl_mlp_num_params(void);
int64_t brl_mlp_train_blob_bytes(void);
int64_t brl_mlp_train_scratch_bytes(int64_t batch);
int64_t brl_mlp_train_trace_offset(int64_t batch); /* debug: offset in the scratch of the u64[2][4096][8] tile time stamps written when
                                                        BrlPpoParams.reserved bit 3 is set (scripts/exp_train_trace.py) */
/* flat params -> training blob: per layer (4 hidden + the 64-wide policy/value head tile) W[in, out_pad] as bf16 hi / lo in
 * haiku's own orientation -- the forward reads it as an MN-major tensor-core operand, the input-gradient GEMM as a K-major
 * one, so training keeps ONE copy of the weights and never transposes -- and the bias in fp32.  Needed once per parameter
 * set; inside the update loop brl_mlp_adam_step keeps the blob current.  opaque = BrlParams (fields unused).
 * buffers: [0] in f32 params[brl_mlp_num_params()]  [1] out blob[brl_mlp_train_blob_bytes()] */
int32_t brl_mlp_pack_train(brl_stream_t, void **buffers, const void *opaque, size_t opaque_len);
/* optimizer.update + optax.apply_updates (src/update.py:168-169, ppo.py:195-211) for the flat MLP parameters, with the
 * refreshed parameters written into the training blob in the same pass: brl_adam_apply + brl_mlp_pack_train as one kernel.
 * opaque = BrlAdamParams (n = brl_mlp_num_params()).
 * buffers: [0] inout f32 params[n]  [1] in f32 grads[n]  [2] inout f32 m[n]  [3] inout f32 v[n]
 *          [4] in f64[1] sum of squares of grads (element 14 of brl_ppo_grad's scratch)  [5] inout blob */
int32_t brl_mlp_adam_step(brl_stream_t, void **buffers, const void *opaque, size_t opaque_len);
/* One minibatch of `jax.value_and_grad(_loss_fn, has_aux=True)(params, traj_batch, gae, targets)` (src/update.py:164-167):
 * take(index) of the observations, forward with the activations kept, brl_ppo_loss, backward GEMMs (input gradients with
 * the ReLU mask fused, weight gradients contracted over the batch, bias gradients), all products as three-term bf16
 * splits accumulated in fp32 on the tensor cores.  opaque = BrlPpoParams (flags: BRL_PPO_* incl. BRL_PPO_OBS_*).
 * buffers: [0] in obs[total,480] (f32 unless BRL_PPO_OBS_*)  [1] in blob (brl_mlp_pack_train)
 *          [2] scratch[brl_mlp_train_scratch_bytes(B)]  [3] in i32 index[B] (NULL = identity)  [4] in u8 mask[total,38]
 *          [5] in i32 action[total]  [6] in f32 old_log_prob[total]  [7] in f32 old_value[total]
 *          [8] in f32 advantages[total]  [9] in f32 targets[total]  [10] out f32 grads[brl_mlp_num_params()]
 *          [11] out f32 stats[8] (as brl_ppo_loss)  [12] scratch, BRL_PPO_SCRATCH_BYTES (as brl_ppo_loss); on return its f64
 *               element 14 = sum of squares of grads (the input of brl_adam_apply) */
int32_t brl_ppo_grad(brl_stream_t, void **buffers, const void *opaque, size_t opaque_len);

/* ---- full evaluation statistics (src/evaluation.py:207-1032) ---------------------------------- */
#define BRL_F_EVAL_INDICATOR_BIDS 0x0400 /* non-duplicate `evaluate`: bid histogram is .set(1), not += 1 (src/evaluation.py:345-354) */
#define BRL_EVAL_ACC_COLS 76    /* per-env log row: ill[2] steps[2] passes[2] actor_bid[35] opp_bid[35] (team 1 first) */
int32_t brl_eval_num_sums(void); /* length of the partial-sum vector of brl_eval_summary (239) */
/* One evaluation step: action = masked argmax of the ACTING team's logits (team 1 = players 0/1) and the
 * per-env log update of update_log_info, skipped for finished envs.
 * buffers: [0] in f32 logits_team1[n,38]  [1] in f32 logits_team2[n,38] (NULL = free-run opponent: always Pass)
 *          [2] in u8 mask[n,38]  [3] in i8 current_player[n]  [4] in u8 terminated[n]
 *          [5] out i32 action[n]  [6] inout f32 acc[n,76] */
int32_t brl_eval_act_log(brl_stream_t, void **buffers, const void *opaque, size_t opaque_len);
/* End of match: every mean of the log_info tuple as f64 partial sums (sums += ...; all-reduce, then divide).
 * sums layout: [0] n [1] sum cum [2] sum cum^2 [3] #(cum>0) [4] sum ill1/steps1 [5] sum ill2/steps2 [6] sum step_count
 *   [7] sum pass1/steps1 [8] sum pass2/steps2 [9] sum tableA rewards[0] [10] sum tableB rewards[0]
 *   [11+9t ..] per table t: pass_out, x1, xx1, x2, xx2, make1, make2, down1, down2
 *   [29..98] bid histograms team 1 then team 2   [99+70t ..] contract histograms of table t, team 1 then team 2
 * buffers: [0] in f32 acc[n,76]  [1] in f32 cum_return[n]  [2] in i32 step_count[n]
 *          [3..8] table A: i32 last_bid, i32 last_bidder, u8 call_x, u8 call_xx, f32 rewards[n,4] (NULL: sign from
 *                 cum_return), i32 pass_num (NULL: not required for a pass-out)
 *          [9..14] table B likewise ([9] NULL = single-table evaluate)   [15] inout f64 sums[239] */
int32_t brl_eval_summary(brl_stream_t, void **buffers, const void *opaque, size_t opaque_len);

/* -------------------------------------------------------------------------
 * Legacy XLA GPU custom-call targets (API_VERSION_STATUS_RETURNING -- the convention
 * of jax/jaxlib 0.4.23, the version brl pins in requirements.txt:25-26).  `buffers` arrive in
 * XLA's order -- the op's inputs (operands) first, then its outputs (results), each in the
 * order the op documents; a buffer the op updates in place is an operand AND an aliased result
 * (input_output_aliases) -- and are re-ordered into the op's list by csrc/xla_ffi_shim.cc from
 * the per-op layout string `brl_xla_layout(name)` returns (one char per op buffer: i input,
 * o output, s scratch output, x in place).  `opaque` = the bytes of the op's params struct.  A
 * failing op sets the XLA status through XlaCustomCallStatusSetFailure, looked up in the process
 * at call time; without that symbol the failure is counted in brl_xla_unreported_failures() and
 * stays in brl_last_error().  Register with xla_client.register_custom_call_target(name,
 * capsule, "CUDA").  Typed-FFI handlers `<op>_ffi` for the same table are built
 * only when the XLA FFI headers are present.
 * ------------------------------------------------------------------------- */
const char *brl_xla_layout(const char *op_name);
long long brl_xla_unreported_failures(void);
struct XlaCustomCallStatus_;
void brl_make_keys_xla(brl_stream_t, void **buffers, const char *opaque, size_t opaque_len, struct XlaCustomCallStatus_ *status);
void brl_init_xla(brl_stream_t, void **buffers, const char *opaque, size_t opaque_len, struct XlaCustomCallStatus_ *status);
void brl_reset_fields_xla(brl_stream_t, void **buffers, const char *opaque, size_t opaque_len, struct XlaCustomCallStatus_ *status);
void brl_step_xla(brl_stream_t, void **buffers, const char *opaque, size_t opaque_len, struct XlaCustomCallStatus_ *status);
void brl_duplicate_step_xla(brl_stream_t, void **buffers, const char *opaque, size_t opaque_len, struct XlaCustomCallStatus_ *status);
void brl_duplicate_init_xla(brl_stream_t, void **buffers, const char *opaque, size_t opaque_len, struct XlaCustomCallStatus_ *status);
void brl_observe_xla(brl_stream_t, void **buffers, const char *opaque, size_t opaque_len, struct XlaCustomCallStatus_ *status);
void brl_legal_mask_xla(brl_stream_t, void **buffers, const char *opaque, size_t opaque_len, struct XlaCustomCallStatus_ *status);
void brl_rollout_random_xla(brl_stream_t, void **buffers, const char *opaque, size_t opaque_len, struct XlaCustomCallStatus_ *status);
void brl_imp_reward_xla(brl_stream_t, void **buffers, const char *opaque, size_t opaque_len, struct XlaCustomCallStatus_ *status);
void brl_gae_xla(brl_stream_t, void **buffers, const char *opaque, size_t opaque_len, struct XlaCustomCallStatus_ *status);
void brl_categorical_xla(brl_stream_t, void **buffers, const char *opaque, size_t opaque_len, struct XlaCustomCallStatus_ *status);
void brl_match_stats_xla(brl_stream_t, void **buffers, const char *opaque, size_t opaque_len, struct XlaCustomCallStatus_ *status);
void brl_state_fields_xla(brl_stream_t, void **buffers, const char *opaque, size_t opaque_len, struct XlaCustomCallStatus_ *status);
void brl_gather_reward_xla(brl_stream_t, void **buffers, const char *opaque, size_t opaque_len, struct XlaCustomCallStatus_ *status);
void brl_mlp_pack_xla(brl_stream_t, void **buffers, const char *opaque, size_t opaque_len, struct XlaCustomCallStatus_ *status);
void brl_obs_to_bf16_xla(brl_stream_t, void **buffers, const char *opaque, size_t opaque_len, struct XlaCustomCallStatus_ *status);
void brl_mlp_forward_xla(brl_stream_t, void **buffers, const char *opaque, size_t opaque_len, struct XlaCustomCallStatus_ *status);
void brl_policy_act_xla(brl_stream_t, void **buffers, const char *opaque, size_t opaque_len, struct XlaCustomCallStatus_ *status);
void brl_policy_act_rows_xla(brl_stream_t, void **buffers, const char *opaque, size_t opaque_len, struct XlaCustomCallStatus_ *status);
void brl_team_rows_xla(brl_stream_t, void **buffers, const char *opaque, size_t opaque_len, struct XlaCustomCallStatus_ *status);
void brl_ppo_loss_xla(brl_stream_t, void **buffers, const char *opaque, size_t opaque_len, struct XlaCustomCallStatus_ *status);
void brl_adam_clip_xla(brl_stream_t, void **buffers, const char *opaque, size_t opaque_len, struct XlaCustomCallStatus_ *status);
void brl_adam_apply_xla(brl_stream_t, void **buffers, const char *opaque, size_t opaque_len, struct XlaCustomCallStatus_ *status);
void brl_gather_rows_xla(brl_stream_t, void **buffers, const char *opaque, size_t opaque_len, struct XlaCustomCallStatus_ *status);
void brl_mlp_pack_train_xla(brl_stream_t, void **buffers, const char *opaque, size_t opaque_len, struct XlaCustomCallStatus_ *status);
void brl_mlp_adam_step_xla(brl_stream_t, void **buffers, const char *opaque, size_t opaque_len, struct XlaCustomCallStatus_ *status);
void brl_ppo_grad_xla(brl_stream_t, void **buffers, const char *opaque, size_t opaque_len, struct XlaCustomCallStatus_ *status);
void brl_eval_act_log_xla(brl_stream_t, void **buffers, const char *opaque, size_t opaque_len, struct XlaCustomCallStatus_ *status);
void brl_eval_summary_xla(brl_stream_t, void **buffers, const char *opaque, size_t opaque_len, struct XlaCustomCallStatus_ *status);

/* -------------------------------------------------------------------------
 * Host-buffer convenience layer (the call a non-JAX host makes): the library
 * owns device state + pinned staging; inputs/outputs are HOST pointers and the
 * copies are part of the call.  Synchronous.  Used for bench.py's `e2e`.
 * ------------------------------------------------------------------------- */
typedef struct BrlEnv BrlEnv;

BrlEnv *brl_env_create(int64_t n_envs, int64_t env_offset, const uint8_t *deal_table_host,
                       int32_t n_deals, uint64_t seed, int32_t flags);
void brl_env_destroy(BrlEnv *env);
/* init from philox keys of (seed, env_offset+i); outputs as brl_env_step_host */
int32_t brl_env_init_host(BrlEnv *env, void *obs, uint8_t *mask, float *rewards, uint8_t *terminated,
                          int8_t *current_player);
/* action: host i32[n] (NULL with BRL_F_RANDOM_ACTION); outputs: host buffers (any may be NULL).
 * obs dtype follows the env's flags (f32 default, BRL_F_OBS_U8 = pgx bool). */
int32_t brl_env_step_host(BrlEnv *env, const int32_t *action, void *obs, uint8_t *mask, float *rewards,
                          uint8_t *terminated, int8_t *current_player);
/* One rollout of k_steps auto-reset random-legal steps (the shape of roll_out's scan,
 * src/roll_out.py:105).  HOST in: uniforms u32[k_steps, n] (the randomness of the action
 * choice; NULL -> in-kernel Philox).  HOST out (any may be NULL): rewards f32[k_steps,n,4],
 * terminated u8[k_steps,n], stats u64[4] (as brl_rollout_random).  The observation / mask
 * trajectories stay in HBM for a device-resident consumer, exactly as `traj_batch` does in
 * the reference; `brl_env_trajectory` exposes their device pointers.
 * If `rewards` / `terminated` are pinned host memory the kernel writes them in place over PCIe
 * (no copy phase); pageable buffers are staged in HBM and copied under the next chunk of steps. */
int32_t brl_env_rollout_host(BrlEnv *env, int32_t k_steps, const uint32_t *uniforms, float *rewards,
                             uint8_t *terminated, uint64_t *stats);
/* Pipelined form: enqueue and return a ticket (> 0; negative = BRL_E_*).  Call c's input copy runs under call
 * c-1's kernel and its result copy under call c+1's kernel.  The host buffers of a call are valid after
 * brl_env_wait(env, ticket); keep at most BRL_ENV_PIPELINE_DEPTH calls in flight, each with its own set of host
 * buffers.  Three in flight keep the GPU fed (with two, the host learns that D2H(c-1) finished only as kernel c
 * ends -- too late to enqueue c+1 without a bubble). */
#define BRL_ENV_PIPELINE_DEPTH 4
int64_t brl_env_rollout_host_async(BrlEnv *env, int32_t k_steps, const uint32_t *uniforms, float *rewards,
                                   uint8_t *terminated, uint64_t *stats);
int32_t brl_env_wait(BrlEnv *env, int64_t ticket);
/* Compact-payload forms (same rollout, same device-resident trajectory incl. f32 rewards / u8 terminated, see
 * brl_env_trajectory): what crosses PCIe is 2 bytes per env-step each way instead of 4 in + 17 out.
 * HOST in: uniforms u32[k_steps,n], or u16[k_steps,n] when the env was created with BRL_F_UNIFORM_U16 (NULL -> in-kernel
 * Philox).  HOST out: result i16[k_steps,n] = 2 * rewards[player 0] + terminated (lossless, see brl_rollout_random
 * buffers[10]; this is what roll_out's consumer reads: rewards[actor] and done, src/roll_out.py:86-94), stats u64[4].
 * Tickets share brl_env_wait and the BRL_ENV_PIPELINE_DEPTH staging slots with brl_env_rollout_host_async. */
int64_t brl_env_rollout_host_compact_async(BrlEnv *env, int32_t k_steps, const void *uniforms, int16_t *result,
                                           uint64_t *stats);
int32_t brl_env_rollout_host_compact(BrlEnv *env, int32_t k_steps, const void *uniforms, int16_t *result, uint64_t *stats);
/* host-side decode of `rows` compact results into rewards f32[rows,4] / terminated u8[rows] (either may be NULL) */
void brl_result16_decode(const int16_t *result, int64_t rows, float *rewards, uint8_t *terminated);
/* device pointers of the last rollout's trajectory: obs, mask, rewards, terminated, current_player, action */
int32_t brl_env_trajectory(BrlEnv *env, void **out_ptrs6);

#ifdef __cplusplus
}
#endif
#endif /* BRL_B200_H */
