/*
 * brl_oracle.c -- CPU restatement (plain C) of brl's bridge-bidding hot path.
 * TEST INFRASTRUCTURE ONLY -- see brl_oracle.h for scope, provenance and pinning.
 *
 * Deliberately written the way the reference computes things (call list,
 * running availability vector, observation rebuilt from the call history on
 * every step) and NOT the way the CUDA path does (packed nibble history), so the
 * two implementations are independent.
 */
#include "brl_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

/* ------------------------------------------------------------------------- */
/* Philox4x32-10 (Salmon et al., SC'11), the counter-based generator both the  */
/* oracle and the CUDA kernels use for episode draws and random-legal actions. */
/* ------------------------------------------------------------------------- */
void orc_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

#define ORC_TAG_KEY  0x4B455930u /* "KEY0" */
#define ORC_TAG_INIT 0x494E4954u /* "INIT" */
#define ORC_TAG_ACT  0x41435430u /* "ACT0" */
#define ORC_TAG_GUM  0x47554D30u /* "GUM0" */

static uint32_t mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }

uint64_t orc_make_key(uint64_t seed, uint64_t g) {
    uint32_t ctr[4] = {(uint32_t)g, (uint32_t)(g >> 32), ORC_TAG_KEY, 0};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t r[4];
    orc_philox4x32(ctr, key, r);
    return (uint64_t)r[0] | ((uint64_t)r[1] << 32);
}

/* init(key): "split" the key (keep one half in _rng_key, src/utils.py:49) and
 * draw deal row, dealer, vulnerabilities and seating from the other half. */
void orc_draw_episode(uint64_t key, int32_t n_deals, uint64_t *new_key, int32_t *deal,
                      int32_t *dealer, int32_t *vul_ns, int32_t *vul_ew, int32_t *seating) {
    uint32_t ctr[4] = {(uint32_t)key, (uint32_t)(key >> 32), ORC_TAG_INIT, 0};
    uint32_t k[2] = {0x62726C5Fu, 0x62323030u}; /* "brl_" "b200" */
    uint32_t r[4];
    orc_philox4x32(ctr, k, r);
    *new_key = (uint64_t)r[0] | ((uint64_t)r[1] << 32);
    *deal = (int32_t)mulhi32(r[2], (uint32_t)n_deals);
    *dealer = (int32_t)(r[3] & 3u);
    *vul_ns = (int32_t)((r[3] >> 2) & 1u);
    *vul_ew = (int32_t)((r[3] >> 3) & 1u);
    *seating = (int32_t)((r[3] >> 4) & 7u);
}

/* A valid seating puts one team (ids {0,1} or {2,3}) on N/S and the other on
 * E/W (examples src/duplicate.py:81-82, wb5/utils.py:72). */
void orc_seating_to_players(int32_t seating, int8_t out[4]) {
    int t = seating & 1, b1 = (seating >> 1) & 1, b2 = (seating >> 2) & 1;
    out[0] = (int8_t)(2 * t + b1);
    out[2] = (int8_t)(2 * t + (1 - b1));
    out[1] = (int8_t)(2 * (1 - t) + b2);
    out[3] = (int8_t)(2 * (1 - t) + (1 - b2));
}

int32_t orc_random_legal_action(const uint8_t mask[ORC_NUM_ACTIONS], uint64_t seed,
                                uint64_t g, uint32_t step) {
    uint32_t ctr[4] = {(uint32_t)g, (uint32_t)(g >> 32), ORC_TAG_ACT, step};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t r[4];
    orc_philox4x32(ctr, key, r);
    int n_legal = 0;
    for (int a = 0; a < ORC_NUM_ACTIONS; ++a) n_legal += mask[a] != 0;
    int k = (int)mulhi32(r[0], (uint32_t)n_legal);
    for (int a = 0; a < ORC_NUM_ACTIONS; ++a) {
        if (mask[a]) {
            if (k == 0) return a;
            --k;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* Scoring -- submodule/bridge_env/bridge_env/score.py:5-106                  */
/* ------------------------------------------------------------------------- */
static const int32_t DOWN[13] = {-50, -100, -150, -200, -250, -300, -350, -400, -450, -500, -550, -600, -650};
static const int32_t DOWN_VUL[13] = {-100, -200, -300, -400, -500, -600, -700, -800, -900, -1000, -1100, -1200, -1300};
static const int32_t DOWN_X[13] = {-100, -300, -500, -800, -1100, -1400, -1700, -2000, -2300, -2600, -2900, -3200, -3500};
static const int32_t DOWN_X_VUL[13] = {-200, -500, -800, -1100, -1400, -1700, -2000, -2300, -2600, -2900, -3200, -3500, -3800};
static const int32_t DOWN_XX[13] = {-200, -600, -1000, -1600, -2200, -2800, -3400, -4000, -4600, -5200, -5800, -6400, -7000};
static const int32_t DOWN_XX_VUL[13] = {-400, -1000, -1600, -2200, -2800, -3400, -4000, -4600, -5200, -5800, -6400, -7000, -7600};

/* score.py:50-106 calc_bid_score; bid = 0..34 (level = bid/5+1, strain = bid%5:
 * C,D,H,S,NT -- bid.py:142,155) */
int32_t orc_score(int32_t bid, int x, int xx, int vul, int tricks) {
    int level = bid / 5 + 1;
    int strain = bid % 5;
    if (level + 6 > tricks) { /* down */
        int down_n = level + 6 - tricks;
        if (xx) return vul ? DOWN_XX_VUL[down_n - 1] : DOWN_XX[down_n - 1];
        if (x) return vul ? DOWN_X_VUL[down_n - 1] : DOWN_X[down_n - 1];
        return vul ? DOWN_VUL[down_n - 1] : DOWN[down_n - 1];
    }
    int over = tricks - level - 6;
    int score, per_over;
    if (strain <= 1) { score = 20 * level; per_over = 20; }        /* minor */
    else if (strain <= 3) { score = 30 * level; per_over = 30; }   /* major */
    else { score = 30 * level + 10; per_over = 30; }               /* NT    */
    if (xx) score *= 4; else if (x) score *= 2;
    if (score >= 100) {                 /* game bonus */
        score += vul ? 450 : 250;
        if (level >= 6) {               /* small slam */
            score += vul ? 750 : 500;
            if (level == 7) score += vul ? 750 : 500; /* grand slam */
        }
    }
    score += 50;                        /* make bonus */
    if (x || xx) {
        score += 50;
        if (xx) { score += 50; per_over = vul ? 400 : 200; }
        else per_over = vul ? 200 : 100;
    }
    return score + per_over * over;
}

/* score.py:43-47,128-137 / src/duplicate.py:46-69 */
static const int32_t IMP_LIST[24] = {20, 50, 90, 130, 170, 220, 270, 320, 370, 430, 500, 600,
                                     750, 900, 1100, 1300, 1500, 1750, 2000, 2250, 2500, 3000, 3500, 4000};

int32_t orc_imp(int32_t d) {
    int win = d >= 0;
    int32_t ad = d < 0 ? -d : d;
    int imp = 0;
    while (imp < 24 && ad >= IMP_LIST[imp]) ++imp;
    return win ? imp : -imp;
}

/* src/duplicate.py:15-70 _imp_reward */
void orc_imp_reward(const float a[4], const float b[4], float out[4]) {
    float d = a[0] + b[0];
    float win = d >= 0.0f ? 1.0f : -1.0f;
    float ad = fabsf(d);
    int imp = 0;
    while (imp < 24 && ad >= (float)IMP_LIST[imp]) ++imp;
    out[0] = (float)imp * win;
    out[1] = (float)imp * win;
    out[2] = -(float)imp * win;
    out[3] = -(float)imp * win;
}

size_t orc_state_size(void) { return sizeof(orc_state); }

/* ------------------------------------------------------------------------- */
/* deal table accessors (row layout: 4 x u64 hand masks in OpenSpiel card      */
/* order rank*4+suit, then 20 DD nibbles seat*5+strain)                        */
/* ------------------------------------------------------------------------- */
static uint64_t deal_hand_mask(const orc_env_params *p, int32_t deal, int seat) {
    uint64_t m;
    memcpy(&m, p->deal_table + (size_t)deal * ORC_DEAL_ROW_BYTES + 8 * seat, 8);
    return m;
}
static int deal_dd_tricks(const orc_env_params *p, int32_t deal, int seat, int strain) {
    int i = seat * 5 + strain;
    uint8_t byte = p->deal_table[(size_t)deal * ORC_DEAL_ROW_BYTES + 32 + i / 2];
    return (i & 1) ? (byte >> 4) : (byte & 15);
}

static int seat_of_player(const orc_state *s, int player_id) {
    for (int seat = 0; seat < 4; ++seat)
        if (s->shuffled_players[seat] == player_id) return seat;
    return 0;
}

/* ------------------------------------------------------------------------- */
/* Observation -- wb5/utils.py:15-52 (convert_vul, convert_history,            */
/* convert_hand, convert_obs), rebuilt from the call list on every call.       */
/* ------------------------------------------------------------------------- */
void orc_observe(const orc_state *s, const orc_env_params *p, int player_id, uint8_t out[ORC_OBS_DIM]) {
    memset(out, 0, ORC_OBS_DIM);
    int obs_seat = seat_of_player(s, player_id);
    /* wb5/utils.py:15-16: [not us_vul, us_vul, not them_vul, them_vul]; seats of
     * equal parity are partners (player.py:104-111), N/S = even seats. */
    int us_vul = (obs_seat % 2 == 0) ? s->vul_ns : s->vul_ew;
    int them_vul = (obs_seat % 2 == 0) ? s->vul_ew : s->vul_ns;
    out[0] = !us_vul; out[1] = us_vul; out[2] = !them_vul; out[3] = them_vul;
    /* wb5/utils.py:28-46 */
    uint8_t *hist = out + 4;
    int last_bid = 0; /* 1-based like Bid.value; 0 = none yet */
    for (int i = 0; i < s->turn; ++i) {
        int rel = ((i + s->dealer) % 4 + (4 - obs_seat)) % 4;
        int a = s->bid_history[i];
        if (a >= 3) {
            last_bid = a - 2;
            hist[4 + (last_bid - 1) * 12 + rel] = 1;
        } else if (a == 0) {
            if (last_bid == 0) hist[rel] = 1;
        } else if (a == 1) {
            hist[4 + (last_bid - 1) * 12 + 4 + rel] = 1;
        } else {
            hist[4 + (last_bid - 1) * 12 + 8 + rel] = 1;
        }
    }
    /* wb5/utils.py:18-26 -- the deal table already stores OpenSpiel card order */
    uint64_t hand = deal_hand_mask(p, s->deal, obs_seat);
    for (int c = 0; c < 52; ++c) out[428 + c] = (uint8_t)((hand >> c) & 1u);
}

/* ------------------------------------------------------------------------- */
/* init / reset                                                                */
/* ------------------------------------------------------------------------- */
void orc_reset_fields(orc_state *s, const orc_env_params *p, int32_t deal, int32_t dealer,
                      int vul_ns, int vul_ew, const int8_t players[4], uint64_t rng_key) {
    memset(s, 0, sizeof(*s));
    s->deal = deal;
    s->rng_key = rng_key;
    memcpy(s->shuffled_players, players, 4);
    s->dealer = dealer;
    s->vul_ns = (uint8_t)vul_ns;
    s->vul_ew = (uint8_t)vul_ew;
    s->last_bid = -1;
    s->last_bidder = -1;
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 5; ++j) s->declarer_check[i][j] = -1;
    /* bidding_phase.py:51-52: everything available except X and XX */
    for (int a = 0; a < ORC_NUM_ACTIONS; ++a) s->available[a] = 1;
    s->available[1] = 0;
    s->available[2] = 0;
    memcpy(s->legal_action_mask, s->available, ORC_NUM_ACTIONS);
    s->current_player = players[dealer];
    orc_observe(s, p, s->current_player, s->observation);
}

void orc_init(orc_state *s, const orc_env_params *p, uint64_t key) {
    uint64_t new_key;
    int32_t deal, dealer, vns, vew, seating;
    int8_t players[4];
    orc_draw_episode(key, p->n_deals, &new_key, &deal, &dealer, &vns, &vew, &seating);
    orc_seating_to_players(seating, players);
    orc_reset_fields(s, p, deal, dealer, vns, vew, players, new_key);
}

/* ------------------------------------------------------------------------- */
/* step -- pgx core.Env.step wrapped around bidding_phase.py:119-180           */
/* ------------------------------------------------------------------------- */
static void orc_terminal_rewards(orc_state *s, const orc_env_params *p) {
    /* bidding_phase.py:182-206 contract(); score.py:109-125 calc_score */
    if (s->last_bid < 0) { /* passed out */
        for (int i = 0; i < 4; ++i) s->rewards[i] = 0.0f;
        return;
    }
    int strain = s->last_bid % 5;
    int bidder_seat = seat_of_player(s, s->last_bidder);
    int pair = bidder_seat % 2; /* 0 = N/S, 1 = E/W */
    int declarer_seat = s->declarer_check[pair][strain];
    int vul = pair == 0 ? s->vul_ns : s->vul_ew; /* contract.py:94-106 */
    int tricks = deal_dd_tricks(p, s->deal, declarer_seat, strain);
    int32_t score = orc_score(s->last_bid, s->call_x, s->call_xx, vul, tricks);
    int declarer_team = s->shuffled_players[declarer_seat] / 2;
    for (int id = 0; id < 4; ++id) s->rewards[id] = (id / 2 == declarer_team) ? (float)score : -(float)score;
}

void orc_step(orc_state *s, const orc_env_params *p, int32_t action) {
    /* pgx core: stepping a finished env is a zero-reward no-op (relied upon by
     * src/evaluation.py:120-122). */
    if (s->terminated || s->truncated) {
        for (int i = 0; i < 4; ++i) s->rewards[i] = 0.0f;
        return;
    }
    int illegal = action < 0 || action >= ORC_NUM_ACTIONS || !s->legal_action_mask[action];
    int actor = s->current_player;
    s->step_count += 1;
    for (int i = 0; i < 4; ++i) s->rewards[i] = 0.0f;

    if (illegal) {
        /* A.6.2 (unpinned): illegal => immediate termination with a penalty vector */
        for (int i = 0; i < 4; ++i) s->rewards[i] = p->illegal_bonus;
        s->rewards[actor] = p->illegal_penalty;
        s->terminated = 1;
    } else {
        int seat = (s->dealer + s->turn) % 4; /* seat of the caller */
        int finished = 0;
        if (action == 0) { /* Pass: bidding_phase.py:139-146 */
            if (s->turn >= 3 && s->bid_history[s->turn - 1] == 0 && s->bid_history[s->turn - 2] == 0)
                finished = 1;
            s->pass_num += 1;
        } else if (action == 1) { /* X */
            s->call_x = 1;
            s->pass_num = 0;
        } else if (action == 2) { /* XX */
            s->call_xx = 1;
            s->pass_num = 0;
        } else { /* regular bid: bidding_phase.py:151-163 */
            int bid = action - 3;
            s->last_bidder = actor;
            s->last_bid = bid;
            if (s->declarer_check[seat % 2][bid % 5] < 0) s->declarer_check[seat % 2][bid % 5] = (int8_t)seat;
            s->call_x = 0;
            s->call_xx = 0;
            for (int a = 3; a <= action; ++a) s->available[a] = 0;
            s->pass_num = 0;
        }
        s->bid_history[s->turn] = (int8_t)action;
        s->turn += 1;
        if (finished) {
            s->terminated = 1;
            orc_terminal_rewards(s, p);
#if ORC_CONV_TERMINAL_ADVANCES_PLAYER
            s->current_player = s->shuffled_players[(s->dealer + s->turn) % 4];
#endif
        } else {
            int next_seat = (s->dealer + s->turn) % 4;
            s->current_player = s->shuffled_players[next_seat];
            if (s->last_bidder >= 0) { /* bidding_phase.py:166-178 */
                int bidder_seat = seat_of_player(s, s->last_bidder);
                int partner = (next_seat % 2) == (bidder_seat % 2);
                s->available[1] = (!s->call_x && !s->call_xx && !partner) ? 1 : 0;
                s->available[2] = (s->call_x && !s->call_xx && partner) ? 1 : 0;
            }
        }
    }
    if (s->terminated && ORC_CONV_TERMINAL_MASK_ALL_TRUE)
        memset(s->legal_action_mask, 1, ORC_NUM_ACTIONS);
    else
        memcpy(s->legal_action_mask, s->available, ORC_NUM_ACTIONS);
    orc_observe(s, p, s->current_player, s->observation);
}

/* src/utils.py:33-56 auto_reset */
void orc_step_autoreset(orc_state *s, const orc_env_params *p, int32_t action) {
    if (s->terminated || s->truncated) {
        s->step_count = 0;
        s->terminated = 0;
        s->truncated = 0;
        for (int i = 0; i < 4; ++i) s->rewards[i] = 0.0f;
    }
    orc_step(s, p, action);
    if (s->terminated || s->truncated) {
        uint8_t term = s->terminated, trunc = s->truncated;
        float rew[4];
        memcpy(rew, s->rewards, sizeof(rew));
        orc_init(s, p, s->rng_key);
        s->terminated = term;
        s->truncated = trunc;
        memcpy(s->rewards, rew, sizeof(rew));
    }
}

/* ------------------------------------------------------------------------- */
/* duplicate -- src/duplicate.py:73-192                                        */
/* ------------------------------------------------------------------------- */
void orc_duplicate_init(orc_state *s, const orc_env_params *p) {
    /* src/duplicate.py:113-135: seats handed to the other team, same deal/
     * dealer/vul, everything else back to the State defaults (rng key too). */
    static const int ix[4] = {1, 0, 3, 2};
    int8_t players[4];
    for (int i = 0; i < 4; ++i) players[i] = s->shuffled_players[ix[i]];
    orc_reset_fields(s, p, s->deal, s->dealer, s->vul_ns, s->vul_ew, players, 0);
}

void orc_table_info_from_state(const orc_state *s, orc_table_info *t) {
    t->terminated = s->terminated;
    memcpy(t->rewards, s->rewards, sizeof(t->rewards));
    t->last_bid = s->last_bid;
    t->last_bidder = s->last_bidder;
    t->call_x = s->call_x;
    t->call_xx = s->call_xx;
}

void orc_duplicate_step(orc_state *s, const orc_env_params *p, int32_t action,
                        orc_table_info *a, orc_table_info *b) {
    orc_step(s, p, action);                                         /* :149 */
    int a_just = !a->terminated && s->terminated;                   /* :152 */
    int b_just = a->terminated && s->terminated && !b->terminated;  /* :158 */
    orc_table_info snap;
    orc_table_info_from_state(s, &snap);
    if (b_just) {
        float imp[4];
        orc_imp_reward(a->rewards, s->rewards, imp);                /* :159-161 */
        memcpy(s->rewards, imp, sizeof(imp));
    } else {
        if (a_just) orc_duplicate_init(s, p);                       /* :151-155 */
        for (int i = 0; i < 4; ++i) s->rewards[i] = 0.0f;           /* :162 */
    }
    if (b_just) *b = snap;                                          /* :165-176 */
    if (a_just) *a = snap;                                          /* :177-188 */
}

/* ------------------------------------------------------------------------- */
/* batched wrappers                                                            */
/* ------------------------------------------------------------------------- */
/* Minimal pthread parallel-for: [0,n) split into contiguous chunks. */
typedef void (*orc_range_fn)(void *ctx, int64_t lo, int64_t hi, int tid);
typedef struct { orc_range_fn fn; void *ctx; int64_t lo, hi; int tid; } orc_job;
static void *orc_job_main(void *arg) {
    orc_job *j = (orc_job *)arg;
    j->fn(j->ctx, j->lo, j->hi, j->tid);
    return NULL;
}
static int resolve_threads(int n_threads) {
    if (n_threads > 0) return n_threads > 256 ? 256 : n_threads;
    long c = sysconf(_SC_NPROCESSORS_ONLN);
    return c < 1 ? 1 : (c > 256 ? 256 : (int)c);
}
static void orc_parallel_for(int64_t n, int n_threads, orc_range_fn fn, void *ctx) {
    int nt = resolve_threads(n_threads);
    if ((int64_t)nt > n) nt = n > 0 ? (int)n : 1;
    if (nt <= 1) { fn(ctx, 0, n, 0); return; }
    pthread_t th[256];
    orc_job jobs[256];
    int started[256];
    for (int t = 0; t < nt; ++t) {
        jobs[t].fn = fn; jobs[t].ctx = ctx; jobs[t].tid = t;
        jobs[t].lo = n * t / nt; jobs[t].hi = n * (t + 1) / nt;
        started[t] = pthread_create(&th[t], NULL, orc_job_main, &jobs[t]) == 0;
        if (!started[t]) orc_job_main(&jobs[t]);
    }
    for (int t = 0; t < nt; ++t) if (started[t]) pthread_join(th[t], NULL);
}

typedef struct {
    orc_state *s; const orc_env_params *p; const uint64_t *keys; const int32_t *actions;
    orc_table_info *a, *b; int autoreset;
} orc_batch_ctx;

static void init_range(void *vc, int64_t lo, int64_t hi, int tid) {
    orc_batch_ctx *c = (orc_batch_ctx *)vc; (void)tid;
    for (int64_t i = lo; i < hi; ++i) orc_init(&c->s[i], c->p, c->keys[i]);
}
static void step_range(void *vc, int64_t lo, int64_t hi, int tid) {
    orc_batch_ctx *c = (orc_batch_ctx *)vc; (void)tid;
    for (int64_t i = lo; i < hi; ++i) {
        if (c->autoreset) orc_step_autoreset(&c->s[i], c->p, c->actions[i]);
        else orc_step(&c->s[i], c->p, c->actions[i]);
    }
}
static void dup_range(void *vc, int64_t lo, int64_t hi, int tid) {
    orc_batch_ctx *c = (orc_batch_ctx *)vc; (void)tid;
    for (int64_t i = lo; i < hi; ++i) orc_duplicate_step(&c->s[i], c->p, c->actions[i], &c->a[i], &c->b[i]);
}

void orc_init_batch(orc_state *s, const orc_env_params *p, const uint64_t *keys, int64_t n, int n_threads) {
    orc_batch_ctx c = {s, p, keys, NULL, NULL, NULL, 0};
    orc_parallel_for(n, n_threads, init_range, &c);
}

void orc_step_batch(orc_state *s, const orc_env_params *p, const int32_t *actions, int64_t n,
                    int autoreset, int n_threads) {
    orc_batch_ctx c = {s, p, NULL, actions, NULL, NULL, autoreset};
    orc_parallel_for(n, n_threads, step_range, &c);
}

void orc_duplicate_step_batch(orc_state *s, const orc_env_params *p, const int32_t *actions,
                              orc_table_info *a, orc_table_info *b, int64_t n, int n_threads) {
    orc_batch_ctx c = {s, p, NULL, actions, a, b, 0};
    orc_parallel_for(n, n_threads, dup_range, &c);
}

static void export_one(const orc_state *s, int64_t i, float *obs_f32, uint8_t *obs_u8, uint8_t *mask,
                       float *rewards, uint8_t *terminated, int8_t *current_player) {
    if (obs_f32) for (int k = 0; k < ORC_OBS_DIM; ++k) obs_f32[i * ORC_OBS_DIM + k] = (float)s->observation[k];
    if (obs_u8) memcpy(obs_u8 + i * ORC_OBS_DIM, s->observation, ORC_OBS_DIM);
    if (mask) memcpy(mask + i * ORC_NUM_ACTIONS, s->legal_action_mask, ORC_NUM_ACTIONS);
    if (rewards) memcpy(rewards + i * 4, s->rewards, 4 * sizeof(float));
    if (terminated) terminated[i] = s->terminated;
    if (current_player) current_player[i] = s->current_player;
}

void orc_export(const orc_state *s, int64_t n, float *obs_f32, uint8_t *obs_u8, uint8_t *mask,
                float *rewards, uint8_t *terminated, int8_t *current_player) {
    for (int64_t i = 0; i < n; ++i) export_one(&s[i], i, obs_f32, obs_u8, mask, rewards, terminated, current_player);
}

void orc_export_private(const orc_state *s, int64_t n, int32_t *deal, int32_t *dealer,
                        int8_t *shuffled, uint8_t *vul, int32_t *last_bid, int32_t *last_bidder,
                        uint8_t *call_x, uint8_t *call_xx, int32_t *pass_num, int32_t *step_count,
                        uint64_t *rng_key) {
    for (int64_t i = 0; i < n; ++i) {
        if (deal) deal[i] = s[i].deal;
        if (dealer) dealer[i] = s[i].dealer;
        if (shuffled) memcpy(shuffled + 4 * i, s[i].shuffled_players, 4);
        if (vul) { vul[2 * i] = s[i].vul_ns; vul[2 * i + 1] = s[i].vul_ew; }
        if (last_bid) last_bid[i] = s[i].last_bid;
        if (last_bidder) last_bidder[i] = s[i].last_bidder;
        if (call_x) call_x[i] = s[i].call_x;
        if (call_xx) call_xx[i] = s[i].call_xx;
        if (pass_num) pass_num[i] = s[i].pass_num;
        if (step_count) step_count[i] = s[i].step_count;
        if (rng_key) rng_key[i] = s[i].rng_key;
    }
}

typedef struct {
    orc_state *s; const orc_env_params *p; int64_t n, env_offset; uint64_t seed; uint32_t step0;
    int32_t k_steps; float *obs_f32; uint8_t *mask; float *rewards; uint8_t *terminated;
    int8_t *current_player; int32_t *actions; int64_t n_term[256];
} orc_rollout_ctx;

static void rollout_range(void *vc, int64_t lo, int64_t hi, int tid) {
    orc_rollout_ctx *c = (orc_rollout_ctx *)vc;
    int64_t n_term = 0;
    for (int64_t i = lo; i < hi; ++i) {
        orc_state *s = &c->s[i];
        for (int32_t t = 0; t < c->k_steps; ++t) {
            int32_t a = orc_random_legal_action(s->legal_action_mask, c->seed,
                                                (uint64_t)(c->env_offset + i), c->step0 + (uint32_t)t);
            orc_step_autoreset(s, c->p, a);
            n_term += s->terminated;
            int64_t row = (int64_t)t * c->n + i;
            if (c->actions) c->actions[row] = a;
            export_one(s, row, c->obs_f32, NULL, c->mask, c->rewards, c->terminated, c->current_player);
        }
    }
    c->n_term[tid] += n_term;
}

int64_t orc_rollout_random(orc_state *s, const orc_env_params *p, int64_t n, int64_t env_offset,
                           uint64_t seed, uint32_t step0, int32_t k_steps, float *obs_f32,
                           uint8_t *mask, float *rewards, uint8_t *terminated,
                           int8_t *current_player, int32_t *actions, int n_threads) {
    orc_rollout_ctx c;
    memset(&c, 0, sizeof(c));
    c.s = s; c.p = p; c.n = n; c.env_offset = env_offset; c.seed = seed; c.step0 = step0;
    c.k_steps = k_steps; c.obs_f32 = obs_f32; c.mask = mask; c.rewards = rewards;
    c.terminated = terminated; c.current_player = current_player; c.actions = actions;
    orc_parallel_for(n, n_threads, rollout_range, &c);
    int64_t total = 0;
    for (int t = 0; t < 256; ++t) total += c.n_term[t];
    return total;
}

/* ------------------------------------------------------------------------- */
/* GAE -- src/gae.py:20-39 (reverse scan; fp32 like the reference)             */
/* ------------------------------------------------------------------------- */
void orc_gae(const uint8_t *done, const float *value, const float *reward, const float *last_val,
             int32_t t_steps, int64_t n, float gamma, float lam, float *adv, float *targets) {
    for (int64_t i = 0; i < n; ++i) {
        float gae = 0.0f, next_value = last_val[i];
        for (int32_t t = t_steps - 1; t >= 0; --t) {
            int64_t k = (int64_t)t * n + i;
            float nd = 1.0f - (float)done[k];
            float delta = reward[k] + gamma * next_value * nd - value[k];
            gae = delta + gamma * lam * nd * gae;
            next_value = value[k];
            adv[k] = gae;
            targets[k] = gae + value[k];
        }
    }
}

/* ------------------------------------------------------------------------- */
/* masked categorical -- src/roll_out.py:27-30,79-81; src/evaluation.py:128-133 */
/* where(mask, logits, -inf); mode = first argmax; sample = Gumbel-argmax       */
/* ------------------------------------------------------------------------- */
void orc_categorical(const float *logits, const uint8_t *mask, int64_t n, int sample, uint64_t seed,
                     uint64_t env_offset, uint32_t step, int32_t *action, float *log_prob) {
    for (int64_t i = 0; i < n; ++i) {
        const float *l = logits + i * ORC_NUM_ACTIONS;
        const uint8_t *m = mask + i * ORC_NUM_ACTIONS;
        double best = -INFINITY;
        int best_a = 0;
        double mx = -INFINITY;
        for (int a = 0; a < ORC_NUM_ACTIONS; ++a) {
            if (!m[a]) continue;
            if ((double)l[a] > mx) mx = (double)l[a];
            double v = (double)l[a];
            if (sample) {
                uint64_t g = env_offset + (uint64_t)i;
                uint32_t ctr[4] = {(uint32_t)g, (uint32_t)(g >> 32), ORC_TAG_GUM + (uint32_t)(a / 4), step};
                uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
                uint32_t r[4];
                orc_philox4x32(ctr, key, r);
                float u = ((float)(r[a % 4] >> 8) + 0.5f) * (1.0f / 16777216.0f);
                v += -log(-log((double)u));
            }
            if (v > best) { best = v; best_a = a; }
        }
        double se = 0.0;
        for (int a = 0; a < ORC_NUM_ACTIONS; ++a)
            if (m[a]) se += exp((double)l[a] - mx);
        if (action) action[i] = best_a;
        if (log_prob) log_prob[i] = (float)((double)l[best_a] - mx - log(se));
    }
}

/* src/evaluation.py:199-201: mean, std(ddof=1)/sqrt(N), win rate */
void orc_match_stats(const double *x, int64_t n, double out[3]) {
    double sum = 0.0, win = 0.0;
    for (int64_t i = 0; i < n; ++i) { sum += x[i]; win += x[i] > 0.0; }
    double mean = sum / (double)n, ss = 0.0;
    for (int64_t i = 0; i < n; ++i) ss += (x[i] - mean) * (x[i] - mean);
    out[0] = mean;
    out[1] = sqrt(ss / (double)(n - 1)) / sqrt((double)n);
    out[2] = win / (double)n;
}
