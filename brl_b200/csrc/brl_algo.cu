// brl_algo.cu -- the small per-env algorithms around the environment: GAE reverse
// scan, masked categorical (mode / Gumbel-argmax sample / log-prob), IMP reward,
// match-statistic partial sums, transition reward gather.  sm_100a only.
#include <math.h>

#include "common.h"
#include "env_device.cuh"

namespace brl {

static thread_local char g_last_error[512] = {0};
char* last_error_buffer() { return g_last_error; }

// ---- GAE -- src/gae.py:20-39 ---------------------------------------------------------------
// One thread per env walks T..0; every load/store is coalesced over the env axis
// ([T, n] time-major).  17 B per (t, env) of traffic; loads of a chunk are issued
// before the dependent chain so the scan is not latency-serialised per step.
// __fmul_rn/__fadd_rn keep the reference's operation order without FMA contraction.
constexpr int kGaeChunk = 8;

__global__ void __launch_bounds__(128) k_gae(const uint8_t* __restrict__ done, const float* __restrict__ value,
                                             const float* __restrict__ reward, const float* __restrict__ last_val,
                                             float* __restrict__ adv, float* __restrict__ targets, int T, int64_t n,
                                             float gamma, float lam) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float gae = 0.0f, next_value = last_val[i];
    const float gl = __fmul_rn(gamma, lam);
    for (int t0 = T; t0 > 0; t0 -= kGaeChunk) {
        float v[kGaeChunk], r[kGaeChunk];
        uint8_t d[kGaeChunk];
#pragma unroll
        for (int k = 0; k < kGaeChunk; ++k) {
            int t = t0 - 1 - k;
            if (t >= 0) {
                int64_t idx = (int64_t)t * n + i;
                v[k] = value[idx];
                r[k] = reward[idx];
                d[k] = done[idx];
            }
        }
#pragma unroll
        for (int k = 0; k < kGaeChunk; ++k) {
            int t = t0 - 1 - k;
            if (t >= 0) {
                int64_t idx = (int64_t)t * n + i;
                float nd = 1.0f - (float)d[k];
                float delta = __fadd_rn(__fadd_rn(r[k], __fmul_rn(__fmul_rn(gamma, next_value), nd)), -v[k]);
                gae = __fadd_rn(delta, __fmul_rn(__fmul_rn(gl, nd), gae));
                next_value = v[k];
                adv[idx] = gae;
                targets[idx] = __fadd_rn(gae, v[k]);
            }
        }
    }
}

// ---- masked categorical -- src/roll_out.py:27-30,79-81; src/evaluation.py:128-133 -----------
// where(mask, logits, -inf); mode = first argmax; sample = argmax(logits + Gumbel);
// log_prob = log_softmax(masked logits)[action].  One warp per env row.
__global__ void __launch_bounds__(128) k_categorical(const float* __restrict__ logits, const uint8_t* __restrict__ mask,
                                                     int32_t* __restrict__ action, float* __restrict__ log_prob,
                                                     int64_t n, int sample, uint64_t seed, int64_t env_offset,
                                                     uint32_t step, const int32_t* __restrict__ row_index,
                                                     float* __restrict__ logits_by_env) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < n; r += n_warps) {
        // row_index: logits are compact ([n, 38] for the listed rows); mask / action / log_prob / RNG counter are by env
        const int64_t row = row_index ? (int64_t)row_index[r] : r;
        const float* l = logits + r * kNumActions;
        const uint8_t* m = mask ? mask + row * kNumActions : nullptr;
        float best = -INFINITY, mx = -INFINITY;
        int best_a = kNumActions;
        float lv[2];
        bool ok[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            int a = lane + 32 * k;
            ok[k] = a < kNumActions && (m == nullptr || m[a] != 0);
            if (logits_by_env && a < kNumActions) logits_by_env[row * kNumActions + a] = l[a];
            lv[k] = ok[k] ? l[a] : -INFINITY;
            if (ok[k]) {
                float v = lv[k];
                if (sample) {
                    uint64_t g = (uint64_t)(env_offset + row);
                    uint4 r = philox4x32(make_uint4((uint32_t)g, (uint32_t)(g >> 32), kTagGum + (uint32_t)(a >> 2), step),
                                         make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
                    uint32_t w = (a & 3) == 0 ? r.x : ((a & 3) == 1 ? r.y : ((a & 3) == 2 ? r.z : r.w));
                    float u = ((float)(w >> 8) + 0.5f) * (1.0f / 16777216.0f);
                    v += -logf(-logf(u));
                }
                if (v > best) { best = v; best_a = a; }  // k ascending: ties keep the lower index
                mx = fmaxf(mx, lv[k]);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float ob = __shfl_xor_sync(0xffffffffu, best, o);
            int oa = __shfl_xor_sync(0xffffffffu, best_a, o);
            if (ob > best || (ob == best && oa < best_a)) { best = ob; best_a = oa; }
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        float se = 0.0f;
#pragma unroll
        for (int k = 0; k < 2; ++k)
            if (ok[k]) se += expf(lv[k] - mx);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
        if (best_a >= kNumActions) best_a = 0;  // no legal action: cannot happen for a valid mask
        float la = __shfl_sync(0xffffffffu, best_a >= 32 ? lv[1] : lv[0], best_a & 31);
        if (lane == 0) {
            if (action) action[row] = best_a;
            if (log_prob) log_prob[row] = la - mx - logf(se);
        }
    }
}

// ---- live-env compaction by acting team -- src/evaluation.py:124-151 ---------------------------
// The reference's vmapped loop runs BOTH teams' nets on every env every iteration and selects by
// `current_player < 2`, and keeps stepping finished envs until the slowest auction ends.  Here the envs still playing
// are listed per acting team (players 0/1 = team 1) so each net runs on exactly the rows it decides.  One block =
// 1024 consecutive envs: ballot + popc give every env its slot inside the block, one atomicAdd per block and team
// reserves the block's range (rows ascend within a block; block order is the order of arrival -- the consumers are
// row-order independent).
__global__ void __launch_bounds__(1024) k_team_rows(const int8_t* __restrict__ player, const uint8_t* __restrict__ done,
                                                    int32_t* __restrict__ rows1, int32_t* __restrict__ rows2,
                                                    int32_t* __restrict__ counts, int64_t n) {
    __shared__ int warp_cnt[2][32];
    __shared__ int base[2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x;
    const bool live = i < n && (done == nullptr || done[i] == 0);
    const bool t1 = live && player[i] < 2, t2 = live && !t1;
    const unsigned b1 = __ballot_sync(0xffffffffu, t1), b2 = __ballot_sync(0xffffffffu, t2);
    if (lane == 0) { warp_cnt[0][warp] = __popc(b1); warp_cnt[1][warp] = __popc(b2); }
    __syncthreads();
    if (warp < 2) {  // exclusive scan of the 32 warp counts of team `warp`
        const int c = warp_cnt[warp][lane];
        int incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        warp_cnt[warp][lane] = incl - c;
        if (lane == 31) base[warp] = incl > 0 ? atomicAdd(&counts[warp], incl) : 0;
    }
    __syncthreads();
    const unsigned below = (1u << lane) - 1u;
    if (t1) rows1[base[0] + warp_cnt[0][warp] + __popc(b1 & below)] = (int32_t)i;
    if (t2) rows2[base[1] + warp_cnt[1][warp] + __popc(b2 & below)] = (int32_t)i;
}

// ---- _imp_reward -- src/duplicate.py:15-70 ----------------------------------------------------
__global__ void __launch_bounds__(256) k_imp_reward(const float4* __restrict__ a, const float4* __restrict__ b,
                                                    float4* __restrict__ out, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = imp_of_difference(a[i].x + b[i].x);
    out[i] = make_float4(s, s, -s, -s);
}

// ---- match statistics partial sums -- src/evaluation.py:199-201 -------------------------------
// sums[0..3] += {n, sum x, sum x^2, #(x > 0)} in double; the caller all-reduces the 8
// doubles across ranks (one NCCL all-reduce) and forms mean / SE(ddof=1) / win-rate.
__global__ void __launch_bounds__(256) k_match_stats(const float* __restrict__ x, double* __restrict__ sums, int64_t n) {
    double s1 = 0.0, s2 = 0.0, w = 0.0, c = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double v = (double)x[i];
        s1 += v;
        s2 += v * v;
        w += v > 0.0 ? 1.0 : 0.0;
        c += 1.0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        w += __shfl_xor_sync(0xffffffffu, w, o);
        c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    __shared__ double sh[4][8];
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { sh[0][warp] = c; sh[1][warp] = s1; sh[2][warp] = s2; sh[3][warp] = w; }
    __syncthreads();
    if (threadIdx.x < 4) {
        double t = 0.0;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += sh[threadIdx.x][k];
        atomicAdd(&sums[threadIdx.x], t);
    }
}

// ---- roll_out bookkeeping -- src/roll_out.py:85-94: rewards[actor] / reward_scale -------------
// With `done` / `count`: also terminated_count += sum(done) (src/roll_out.py:85), one atomic per block.
__global__ void __launch_bounds__(256) k_gather_reward(const float* __restrict__ rewards, const int8_t* __restrict__ actor,
                                                       float* __restrict__ out, int64_t n, float scale,
                                                       const uint8_t* __restrict__ done, unsigned long long* __restrict__ count) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = rewards[4 * i + (actor[i] & 3)] / scale;
    if (count != nullptr) {
        const int c = __syncthreads_count(i < n && done[i] != 0);
        if (threadIdx.x == 0 && c > 0) atomicAdd(count, (unsigned long long)c);
    }
}


int32_t launch_categorical(cudaStream_t stream, const float* logits, const unsigned char* mask, int32_t* action, float* log_prob,
                           int64_t n, int sample, uint64_t seed, int64_t env_offset, uint32_t step, const int32_t* row_index, float* logits_by_env) {
    unsigned grid = (unsigned)((n + 3) / 4);  // one warp per env row
    if (grid > 148u * 16u) grid = 148u * 16u;
    k_categorical<<<grid, 128, 0, stream>>>(logits, mask, action, log_prob, n, sample, seed, env_offset, step, row_index, logits_by_env);
    return check_launch("brl_categorical");
}

}  // namespace brl

using namespace brl;

extern "C" {

const char* brl_last_error(void) { return last_error_buffer(); }
int32_t brl_abi_version(void) { return BRL_ABI_VERSION; }

int32_t brl_gae(brl_stream_t stream, void** b, const void* opaque, size_t len) {
    int32_t rc;
    const BrlParams* p = get_params(opaque, len, &rc);
    if (!p) return rc;
    if (p->n_envs == 0) return BRL_OK;  // an empty batch is a no-op; its buffers may be NULL (zero-size XLA / torch buffers)
    static const char* names[] = {"done", "value", "reward", "last_val", "advantages", "targets"};
    for (int k = 0; k < 6; ++k)
        if (b[k] == nullptr) return fail(BRL_E_BUFFER, "brl_gae: buffer '%s' is NULL", names[k]);
    if (p->k_steps <= 0) return fail(BRL_E_OPAQUE, "brl_gae: k_steps (T) must be > 0");
    if (p->n_envs == 0) return BRL_OK;
    k_gae<<<(unsigned)((p->n_envs + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        static_cast<const uint8_t*>(b[0]), static_cast<const float*>(b[1]), static_cast<const float*>(b[2]),
        static_cast<const float*>(b[3]), static_cast<float*>(b[4]), static_cast<float*>(b[5]), p->k_steps, p->n_envs,
        p->gamma, p->gae_lambda);
    return check_launch("brl_gae");
}

int32_t brl_categorical(brl_stream_t stream, void** b, const void* opaque, size_t len) {
    int32_t rc;
    const BrlParams* p = get_params(opaque, len, &rc);
    if (!p) return rc;
    if (p->n_envs == 0) return BRL_OK;  // an empty batch is a no-op; its buffers may be NULL (zero-size XLA / torch buffers)
    if (b[0] == nullptr) return fail(BRL_E_BUFFER, "brl_categorical: buffer 'logits' is NULL");
    if (p->n_envs == 0) return BRL_OK;
    return launch_categorical((cudaStream_t)stream, static_cast<const float*>(b[0]), static_cast<const uint8_t*>(b[1]),
                              static_cast<int32_t*>(b[2]), static_cast<float*>(b[3]), p->n_envs,
                              (p->flags & BRL_F_SAMPLE) ? 1 : 0, p->seed, p->env_offset, p->step, nullptr);
}

int32_t brl_team_rows(brl_stream_t stream, void** b, const void* opaque, size_t len) {
    int32_t rc;
    const BrlParams* p = get_params(opaque, len, &rc);
    if (!p) return rc;
    if (p->n_envs == 0) {  // an empty batch: both lists are empty
        if (b[4] != nullptr) cudaMemsetAsync(b[4], 0, 2 * sizeof(int32_t), (cudaStream_t)stream);
        return check_launch("brl_team_rows");
    }
    BRL_REQUIRE(b[0], "current_player");
    BRL_REQUIRE(b[2], "rows_team1");
    BRL_REQUIRE(b[3], "rows_team2");
    BRL_REQUIRE(b[4], "counts");
    if (cudaMemsetAsync(b[4], 0, 2 * sizeof(int32_t), (cudaStream_t)stream) != cudaSuccess) return check_launch("brl_team_rows");
    if (p->n_envs == 0) return BRL_OK;
    unsigned grid = (unsigned)((p->n_envs + 1023) / 1024);
    k_team_rows<<<grid, 1024, 0, (cudaStream_t)stream>>>(static_cast<const int8_t*>(b[0]), static_cast<const uint8_t*>(b[1]),
                                                         static_cast<int32_t*>(b[2]), static_cast<int32_t*>(b[3]),
                                                         static_cast<int32_t*>(b[4]), p->n_envs);
    return check_launch("brl_team_rows");
}

int32_t brl_imp_reward(brl_stream_t stream, void** b, const void* opaque, size_t len) {
    int32_t rc;
    const BrlParams* p = get_params(opaque, len, &rc);
    if (!p) return rc;
    if (p->n_envs == 0) return BRL_OK;  // an empty batch is a no-op; its buffers may be NULL (zero-size XLA / torch buffers)
    BRL_REQUIRE(b[0], "a_rewards");
    BRL_REQUIRE(b[1], "b_rewards");
    BRL_REQUIRE(b[2], "imp");
    if (p->n_envs == 0) return BRL_OK;
    k_imp_reward<<<(unsigned)((p->n_envs + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        static_cast<const float4*>(b[0]), static_cast<const float4*>(b[1]), static_cast<float4*>(b[2]), p->n_envs);
    return check_launch("brl_imp_reward");
}

int32_t brl_match_stats(brl_stream_t stream, void** b, const void* opaque, size_t len) {
    int32_t rc;
    const BrlParams* p = get_params(opaque, len, &rc);
    if (!p) return rc;
    if (p->n_envs == 0) return BRL_OK;  // an empty batch is a no-op; its buffers may be NULL (zero-size XLA / torch buffers)
    if (b[0] == nullptr || b[1] == nullptr) return fail(BRL_E_BUFFER, "brl_match_stats: NULL buffer");
    if (p->n_envs == 0) return BRL_OK;
    unsigned grid = (unsigned)((p->n_envs + 255) / 256);
    if (grid > 148u * 4u) grid = 148u * 4u;
    k_match_stats<<<grid, 256, 0, (cudaStream_t)stream>>>(static_cast<const float*>(b[0]), static_cast<double*>(b[1]),
                                                          p->n_envs);
    return check_launch("brl_match_stats");
}

int32_t brl_gather_reward(brl_stream_t stream, void** b, const void* opaque, size_t len) {
    int32_t rc;
    const BrlParams* p = get_params(opaque, len, &rc);
    if (!p) return rc;
    if (p->n_envs == 0) return BRL_OK;  // an empty batch is a no-op; its buffers may be NULL (zero-size XLA / torch buffers)
    if (b[0] == nullptr || b[1] == nullptr || b[2] == nullptr) return fail(BRL_E_BUFFER, "brl_gather_reward: NULL buffer");
    if (p->n_envs == 0) return BRL_OK;
    const bool counting = (p->flags & BRL_F_COUNT_DONE) != 0;
    if (counting && (b[3] == nullptr || b[4] == nullptr || (reinterpret_cast<uintptr_t>(b[4]) & 7u)))
        return fail(BRL_E_BUFFER, "brl_gather_reward: BRL_F_COUNT_DONE needs buffers [3] done and [4] count (8-byte aligned)");
    k_gather_reward<<<(unsigned)((p->n_envs + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        static_cast<const float*>(b[0]), static_cast<const int8_t*>(b[1]), static_cast<float*>(b[2]), p->n_envs,
        p->gamma, counting ? static_cast<const uint8_t*>(b[3]) : nullptr,
        counting ? static_cast<unsigned long long*>(b[4]) : nullptr);
    return check_launch("brl_gather_reward");
}

}  // extern "C"
