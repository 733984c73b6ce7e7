/*
 * brl_oracle.h -- CPU restatement of the bridge-bidding hot path of harukaki/brl.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it, and only as the checker / the timed CPU baseline.
 *
 * The reference env (pgx==1.4.0 `bridge_bidding`, requirements.txt:40) is an
 * un-vendored pip dependency that is absent from /root/reference and from this
 * image, so this file restates the algorithm from the files that ARE in the
 * reference tree (paths relative to /root/reference):
 *   auction update / legality / termination / declarer
 *       submodule/bridge_env/bridge_env/bidding_phase.py:119-206
 *   observation layout      wb5/utils.py:15-52
 *   duplicate score         submodule/bridge_env/bridge_env/score.py:5-106
 *   declarer vulnerability  submodule/bridge_env/bridge_env/contract.py:94-106
 *   IMP reward vector       src/duplicate.py:15-70
 *   duplicate table swap    src/duplicate.py:73-192
 *   auto reset              src/utils.py:9-58
 *   GAE                     src/gae.py:20-39
 *   match statistics        src/evaluation.py:199-201
 *
 * PARITY PINNING: the in-tree parts are pinned by tests/golden/ (vectors produced
 * by running the reference's own Python -- bridge_env BiddingPhase, calc_score,
 * score_to_imp and wb5/utils.convert_obs -- see tests/golden/make_golden.py).
 * pgx-only conventions that nothing in the tree fixes (PRNG -> deal/dealer/vul/
 * seating mapping, state contents AT a terminal, illegal-action penalty vector)
 * are "parity unpinned"; they are isolated in the ORC_CONV_* switches below.
 */
#ifndef BRL_ORACLE_H
#define BRL_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NUM_ACTIONS 38
#define ORC_OBS_DIM 480
#define ORC_MAX_CALLS 320
#define ORC_DEAL_ROW_BYTES 48 /* 4 x u64 hand masks + 10 B DD nibbles + pad */

/* pgx-only conventions -- unpinned, kept switchable. */
#define ORC_CONV_TERMINAL_MASK_ALL_TRUE 1   /* pgx core: mask all-True at terminal */
#define ORC_CONV_TERMINAL_ADVANCES_PLAYER 0 /* current_player frozen at terminal */

/* "Fat" per-env state in the spirit of pgx's State pytree (field names follow
 * the ones brl reads: src/duplicate.py:120-128, src/evaluation.py:97-112). */
typedef struct orc_state {
    int32_t deal;                 /* row of the deal / double-dummy table      */
    uint64_t rng_key;             /* `_rng_key` (src/utils.py:49)              */
    int8_t shuffled_players[4];   /* `_shuffled_players[seat]` = player id      */
    int32_t dealer;               /* `_dealer` (seat 0..3 = N,E,S,W)           */
    uint8_t vul_ns, vul_ew;       /* `_vul_NS`, `_vul_EW`                      */
    int32_t turn;                 /* number of calls made so far               */
    int8_t bid_history[ORC_MAX_CALLS]; /* actions in call order                */
    int32_t last_bid;             /* `_last_bid` = action-3, -1 if none        */
    int32_t last_bidder;          /* `_last_bidder` player id, -1 if none      */
    uint8_t call_x, call_xx;      /* `_call_x`, `_call_xx`                     */
    int32_t pass_num;             /* `_pass_num` consecutive passes            */
    int8_t declarer_check[2][5];  /* first SEAT of pair (0=NS,1=EW) per strain */
    uint8_t available[ORC_NUM_ACTIONS]; /* running legality (bidding_phase.py:51) */
    int8_t current_player;        /* player id to act                          */
    uint8_t terminated, truncated;
    int32_t step_count;           /* `_step_count`                             */
    float rewards[4];             /* by player id                              */
    uint8_t legal_action_mask[ORC_NUM_ACTIONS];
    uint8_t observation[ORC_OBS_DIM];
} orc_state;

/* src/duplicate.py:138-144 */
typedef struct orc_table_info {
    uint8_t terminated;
    float rewards[4];
    int32_t last_bid;
    int32_t last_bidder;
    uint8_t call_x, call_xx;
} orc_table_info;

typedef struct orc_env_params {
    const uint8_t *deal_table; /* n_deals rows of ORC_DEAL_ROW_BYTES */
    int32_t n_deals;
    float illegal_penalty;     /* reward of the offender (pgx default -1) */
    float illegal_bonus;       /* reward of every other player            */
} orc_env_params;

/* ---- scalar building blocks -------------------------------------------- */
void orc_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
int32_t orc_score(int32_t bid /*0..34*/, int x, int xx, int vul, int tricks);
int32_t orc_imp(int32_t point_difference);
void orc_imp_reward(const float a[4], const float b[4], float out[4]);
size_t orc_state_size(void);

/* episode draw shared with the CUDA path (conventions, not pgx parity) */
void orc_draw_episode(uint64_t key, int32_t n_deals, uint64_t *new_key, int32_t *deal,
                      int32_t *dealer, int32_t *vul_ns, int32_t *vul_ew, int32_t *seating);
void orc_seating_to_players(int32_t seating, int8_t out[4]);
uint64_t orc_make_key(uint64_t seed, uint64_t global_env_index);
int32_t orc_random_legal_action(const uint8_t mask[ORC_NUM_ACTIONS], uint64_t seed,
                                uint64_t global_env_index, uint32_t step);

/* ---- per-env env surface ------------------------------------------------ */
void orc_reset_fields(orc_state *s, const orc_env_params *p, int32_t deal, int32_t dealer,
                      int vul_ns, int vul_ew, const int8_t players[4], uint64_t rng_key);
void orc_init(orc_state *s, const orc_env_params *p, uint64_t key);
void orc_observe(const orc_state *s, const orc_env_params *p, int player_id, uint8_t out[ORC_OBS_DIM]);
void orc_step(orc_state *s, const orc_env_params *p, int32_t action);
void orc_step_autoreset(orc_state *s, const orc_env_params *p, int32_t action);
void orc_duplicate_init(orc_state *s, const orc_env_params *p);
void orc_duplicate_step(orc_state *s, const orc_env_params *p, int32_t action,
                        orc_table_info *a, orc_table_info *b);
void orc_table_info_from_state(const orc_state *s, orc_table_info *t);

/* ---- batched wrappers (optionally OpenMP-threaded; n_threads<=0 -> all) -- */
void orc_init_batch(orc_state *s, const orc_env_params *p, const uint64_t *keys, int64_t n, int n_threads);
void orc_step_batch(orc_state *s, const orc_env_params *p, const int32_t *actions, int64_t n,
                    int autoreset, int n_threads);
void orc_duplicate_step_batch(orc_state *s, const orc_env_params *p, const int32_t *actions,
                              orc_table_info *a, orc_table_info *b, int64_t n, int n_threads);
/* copy the Env surface out into SoA arrays (any pointer may be NULL) */
void orc_export(const orc_state *s, int64_t n, float *obs_f32, uint8_t *obs_u8, uint8_t *mask,
                float *rewards, uint8_t *terminated, int8_t *current_player);
void orc_export_private(const orc_state *s, int64_t n, int32_t *deal, int32_t *dealer,
                        int8_t *shuffled, uint8_t *vul, int32_t *last_bid, int32_t *last_bidder,
                        uint8_t *call_x, uint8_t *call_xx, int32_t *pass_num, int32_t *step_count,
                        uint64_t *rng_key);
/* random-legal rollout used as the timed CPU baseline: K autoreset steps over n
 * envs, writing the same outputs the CUDA rollout writes (pointers may be NULL);
 * returns the number of terminal steps seen. */
int64_t orc_rollout_random(orc_state *s, const orc_env_params *p, int64_t n, int64_t env_offset,
                           uint64_t seed, uint32_t step0, int32_t k_steps, float *obs_f32,
                           uint8_t *mask, float *rewards, uint8_t *terminated,
                           int8_t *current_player, int32_t *actions, int n_threads);

/* ---- algorithms ---------------------------------------------------------- */
void orc_gae(const uint8_t *done, const float *value, const float *reward, const float *last_val,
             int32_t t_steps, int64_t n, float gamma, float lam, float *adv, float *targets);
/* masked categorical (src/roll_out.py:27-30,79-81): mode / Gumbel-argmax sample / log-prob */
void orc_categorical(const float *logits, const uint8_t *mask, int64_t n, int sample, uint64_t seed,
                     uint64_t env_offset, uint32_t step, int32_t *action, float *log_prob);
void orc_match_stats(const double *cum_return, int64_t n, double out[3]);

#ifdef __cplusplus
}
#endif
#endif
