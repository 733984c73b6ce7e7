// brl_ppo.cu -- the PPO update's non-GEMM arithmetic (src/update.py:74-242,
// ppo.py:195-211): the clipped-surrogate / clipped-value / entropy loss head fused with its
// own backward (d loss / d logits, d loss / d value in one pass over the minibatch), and the
// optimizer step (optax.clip_by_global_norm + optax.adam) over a flat parameter buffer.
// The MLP forward/backward GEMMs between them are plain library GEMMs (cuBLAS via torch).
#include <cuda_bf16.h>
#include <math.h>

#include "common.h"
#include "ppo_gram.cuh"

namespace brl {

constexpr int kA = 38;

struct PpoArgs {
    const float* logits;       // [B, 38]
    const float* value;        // [B]
    const int32_t* index;      // [B] rows of the flat trajectory (NULL = identity)
    const uint8_t* mask;       // [total, 38]
    const int32_t* action;     // [total]
    const float* old_log_prob; // [total]
    const float* old_value;    // [total]
    const float* adv;          // [total]
    const float* targets;      // [total]
    float* dlogits;            // [B, 38]
    float* dvalue;             // [B]
    float* stats;              // [8]
    double* acc;               // [16] scratch accumulators; the rest of the BRL_PPO_SCRATCH_BYTES scratch follows:
    double* spec;              // [40] {sigma_max, v[38], 0} of the illegal-probability matrix (k_ppo_illegal_gram)
    float* gram_part;          // [kGramBlocks][38 * 38] per-block partial Gram matrices
    int defer_illegal;         // the statistic is formed later (a spare block of k_bias_grad); only valid with ill_coef == 0
    int skip_illegal;          // ... or not at all (BRL_PPO_ILLEGAL_STAT clear): stats[6] = NaN
    // optional (brl_ppo_grad): d loss / d (logits, value) also as bf16 hi / lo rows [B, 64] (38 logits, value, zeros),
    // the A operand of the head's input-gradient GEMM and the B operand of its weight-gradient GEMM
    unsigned short* dz_hi;
    unsigned short* dz_lo;
    int64_t B;
    float clip_eps, ent_coef, vf_coef, ill_coef;
    int value_clipping, reward_scaling, masked_policy;
};

// acc slots
enum { kAccAdv = 0, kAccAdv2, kAccIll, kAccActor, kAccValue, kAccEnt, kAccKl, kAccClip, kAccGramTicket = 13, kAccTicket = 15 };
static_assert(128 + 40 * 8 <= 512 && 512 + kGramBlocks * kGramN * 4 <= BRL_PPO_SCRATCH_BYTES, "scratch layout");

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// one warp per sample; lane owns actions lane and lane + 32
struct Row {
    float l[2];
    bool in[2], legal[2];
    __device__ __forceinline__ void load(const PpoArgs& a, int64_t b, int64_t src, int lane) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            int j = lane + 32 * k;
            in[k] = j < kA;
            l[k] = in[k] ? a.logits[b * kA + j] : -INFINITY;
            legal[k] = in[k] && a.mask[src * kA + j] != 0;
        }
    }
    // softmax over `sel` entries: returns log-partition; p[k] = probabilities (0 outside sel)
    __device__ __forceinline__ float softmax(const bool sel[2], float p[2], float logp[2]) const {
        float mx = warp_max(fmaxf(sel[0] ? l[0] : -INFINITY, sel[1] ? l[1] : -INFINITY));
        float e0 = sel[0] ? expf(l[0] - mx) : 0.0f, e1 = sel[1] ? expf(l[1] - mx) : 0.0f;
        float lse = mx + logf(warp_sum(e0 + e1));
        logp[0] = sel[0] ? l[0] - lse : -INFINITY;
        logp[1] = sel[1] ? l[1] - lse : -INFINITY;
        p[0] = sel[0] ? expf(logp[0]) : 0.0f;
        p[1] = sel[1] ? expf(logp[1]) : 0.0f;
        return lse;
    }
};

// sum over the minibatch of advantages and advantages^2 (reward scaling, src/update.py:35-36)
__global__ void __launch_bounds__(128) k_ppo_prepass(const PpoArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    double s_adv = 0.0, s_adv2 = 0.0;
    for (int64_t b = warp * 32 + lane; b < a.B; b += nw * 32) {
        const double g = (double)a.adv[a.index ? a.index[b] : b];
        s_adv += g;
        s_adv2 += g * g;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s_adv += __shfl_xor_sync(0xffffffffu, s_adv, o);
        s_adv2 += __shfl_xor_sync(0xffffffffu, s_adv2, o);
    }
    if (lane == 0) {
        atomicAdd(&a.acc[kAccAdv], s_adv);
        atomicAdd(&a.acc[kAccAdv2], s_adv2);
    }
}

// illegal_action_loss as the spectral norm (ppo_gram.cuh): partials + tail before the loss kernel, which then has sigma_1 / v_1
// for the statistic and the gradient
__global__ void __launch_bounds__(1024) k_ppo_illegal_gram(const GramArgs g) {
    __shared__ GramSmem sm;
    pdl_trigger();
    pdl_wait();
    gram_block<1024>(g, (int)blockIdx.x, (int)gridDim.x, sm);
}
// the partial Gram matrices only (deferred statistic: the tail runs in a spare block of k_bias_grad)
__global__ void __launch_bounds__(1024) k_ppo_gram_partial(const GramArgs g) {
    __shared__ float X[64 * kGramPad];
    pdl_trigger();
    pdl_wait();
    gram_partial<1024>(g, (int)blockIdx.x, (int)gridDim.x, X);
}

__device__ __forceinline__ void split_store(const PpoArgs& a, int64_t idx, float v) {
    const __nv_bfloat16 h = __float2bfloat16_rn(v), l = __float2bfloat16_rn(v - __bfloat162float(h));
    a.dz_hi[idx] = *reinterpret_cast<const unsigned short*>(&h);
    a.dz_lo[idx] = *reinterpret_cast<const unsigned short*>(&l);
}

// stats = {total_loss, value_loss, loss_actor, entropy, approx_kl, clipfracs, illegal_action_loss, 0}
// (the aux tuple of _loss_fn, src/update.py:144-162); run by the last block of k_ppo_loss
__device__ __forceinline__ void ppo_finalize(const PpoArgs& a) {
    const volatile double* acc = a.acc;
    const double B = (double)a.B;
    const float value_loss = (float)(acc[kAccValue] / B), loss_actor = (float)(acc[kAccActor] / B);
    const float entropy = (float)(acc[kAccEnt] / B), ill = a.defer_illegal ? 0.0f : 0.5f * (float)((const volatile double*)a.spec)[0];
    a.stats[0] = loss_actor + a.vf_coef * value_loss - a.ent_coef * entropy + a.ill_coef * ill;
    a.stats[1] = value_loss;
    a.stats[2] = loss_actor;
    a.stats[3] = entropy;
    a.stats[4] = (float)(acc[kAccKl] / B);
    a.stats[5] = (float)(acc[kAccClip] / B);
    if (a.skip_illegal) a.stats[6] = __int_as_float(0x7fc00000);  // NaN: not computed, never a stale or wrong number
    else if (!a.defer_illegal) a.stats[6] = ill;
    a.stats[7] = 0.0f;
}

__global__ void __launch_bounds__(128) k_ppo_loss(const PpoArgs a, int have_prepass) {
    pdl_trigger();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const float invB = 1.0f / (float)a.B;
    float g_mean = 0.0f, g_istd = 1.0f;
    if (have_prepass) {
        double m = a.acc[kAccAdv] / (double)a.B;
        double var = a.acc[kAccAdv2] / (double)a.B - m * m;
        g_mean = (float)m;
        g_istd = 1.0f / ((float)sqrt(var > 0.0 ? var : 0.0) + 1e-8f);
    }
    // spectral norm of the illegal-probability matrix and its top right singular vector (k_ppo_illegal_gram)
    const float sigma = a.defer_illegal ? 0.0f : (float)a.spec[0];
    const bool ill_grad = a.ill_coef != 0.0f && sigma > 0.0f;
    float v1[2] = {0.0f, 0.0f};
    if (ill_grad) {
        v1[0] = (float)a.spec[1 + lane];
        v1[1] = lane < kA - 32 ? (float)a.spec[33 + lane] : 0.0f;
    }
    double s_actor = 0.0, s_value = 0.0, s_ent = 0.0, s_kl = 0.0, s_clip = 0.0;
    for (int64_t b = warp; b < a.B; b += nw) {
        const int64_t src = a.index ? a.index[b] : b;
        Row r;
        r.load(a, b, src, lane);
        // masked distribution: where(mask, logits, -inf) (src/update.py:12-16,132-137)
        float p[2], logp[2];
        r.softmax(r.legal, p, logp);
        // unmasked distribution (src/update.py:139; also the policy when masking is off, :20-24)
        float q[2], logq[2];
        r.softmax(r.in, q, logq);
        const int act = a.action[src];
        const float* pol_p = a.masked_policy ? p : q;
        const float* pol_logp = a.masked_policy ? logp : logq;
        float lp_a = __shfl_sync(0xffffffffu, act >= 32 ? pol_logp[1] : pol_logp[0], act & 31);
        // actor loss (src/update.py:115-131)
        float gae = a.adv[src];
        if (a.reward_scaling) gae = (gae - g_mean) * g_istd;
        const float logratio = lp_a - a.old_log_prob[src];
        const float ratio = expf(logratio);
        const float l1 = ratio * gae, l2 = fminf(fmaxf(ratio, 1.0f - a.clip_eps), 1.0f + a.clip_eps) * gae;
        const float loss_actor = -fminf(l1, l2);
        const bool inside = ratio >= 1.0f - a.clip_eps && ratio <= 1.0f + a.clip_eps;
        const float dl_dratio = (l1 < l2 || inside) ? -gae : 0.0f;  // min picks the unclipped term, or both coincide
        const float c_ratio = dl_dratio * ratio * invB;               // d mean(loss_actor) / d log_prob
        // entropy of the masked distribution (src/update.py:136-137)
        float ent = -((r.legal[0] && p[0] > 0.0f ? p[0] * logp[0] : 0.0f) + (r.legal[1] && p[1] > 0.0f ? p[1] * logp[1] : 0.0f));
        ent = warp_sum(ent);
        // illegal-action loss (src/update.py:138-142): x = unmasked probabilities of the illegal actions, t = x . v_1;
        // d (coef * sigma_1 / 2) / d logits_j = coef / 2 * (t / sigma_1) * (x_j v_j - q_j t)
        float qi[2] = {r.in[0] && !r.legal[0] ? q[0] : 0.0f, r.in[1] && !r.legal[1] ? q[1] : 0.0f};
        const float t_i = ill_grad ? warp_sum(qi[0] * v1[0] + qi[1] * v1[1]) : 0.0f;
        // d total / d logits
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int j = lane + 32 * k;
            if (!r.in[k]) continue;
            float d = c_ratio * ((j == act ? 1.0f : 0.0f) - pol_p[k]);
            if (r.legal[k] && p[k] > 0.0f) d += a.ent_coef * invB * p[k] * (logp[k] + ent);  // -ent_coef * dH/dl
            if (ill_grad) d += a.ill_coef * 0.5f * (t_i / sigma) * (qi[k] * v1[k] - q[k] * t_i);
            a.dlogits[b * kA + j] = d;
            if (a.dz_hi) split_store(a, b * 64 + j, d);
        }
        if (a.dz_hi && lane >= 7) split_store(a, b * 64 + 32 + lane, 0.0f);  // columns 39..63 (38 is the value, below)
        if (lane == 0) {
            // value loss (src/update.py:47-72)
            const float v = a.value[b], t = a.targets[src];
            float vl, dv;
            if (a.value_clipping) {
                const float ov = a.old_value[src];
                const float diff = v - ov;
                const float vc = ov + fminf(fmaxf(diff, -a.clip_eps), a.clip_eps);
                const float e1 = (v - t) * (v - t), e2 = (vc - t) * (vc - t);
                vl = 0.5f * fmaxf(e1, e2);
                const bool unclipped = diff >= -a.clip_eps && diff <= a.clip_eps;
                dv = (e1 > e2 || unclipped) ? (v - t) : 0.0f;  // the clipped branch is constant in v when it binds
            } else {
                vl = 0.5f * (v - t) * (v - t);
                dv = v - t;
            }
            a.dvalue[b] = a.vf_coef * dv * invB;
            if (a.dz_hi) split_store(a, b * 64 + 38, a.vf_coef * dv * invB);
            s_actor += (double)loss_actor;
            s_value += (double)vl;
            s_ent += (double)ent;
            s_kl += (double)((ratio - 1.0f) - logratio);
            s_clip += fabsf(ratio - 1.0f) > a.clip_eps ? 1.0 : 0.0;
        }
    }
    // block-level sums first (one set of same-address f64 atomics per block, not per warp), then the last block to
    // arrive (ticket in acc[kAccTicket]) forms the statistics: no separate finalize launch
    __shared__ double part[4][5];
    __shared__ bool last_block;
    const int w = threadIdx.x >> 5;
    if (lane == 0) {
        part[w][0] = s_actor; part[w][1] = s_value; part[w][2] = s_ent; part[w][3] = s_kl; part[w][4] = s_clip;
    }
    __syncthreads();
    if (threadIdx.x < 5) {
        const int k = threadIdx.x;
        const double s = part[0][k] + part[1][k] + part[2][k] + part[3][k];
        const int slot = k == 0 ? kAccActor : k == 1 ? kAccValue : k == 2 ? kAccEnt : k == 3 ? kAccKl : kAccClip;
        atomicAdd(&a.acc[slot], s);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned long long t = atomicAdd(reinterpret_cast<unsigned long long*>(&a.acc[kAccTicket]), 1ull);
        last_block = t == (unsigned long long)gridDim.x - 1ull;
    }
    __syncthreads();
    if (last_block && threadIdx.x == 0) {
        __threadfence();
        ppo_finalize(a);
    }
}


// ---- optimizer: optax.chain(clip_by_global_norm(c), adam(lr, eps=1e-5)) (ppo.py:195-211) -------
// VEC = 4: 128-bit accesses over the first n / 4 * 4 elements (16-byte-aligned buffers), the <= 3 tail elements by block 0
template <int VEC>
__global__ void __launch_bounds__(256) k_sumsq(const float* __restrict__ g, int64_t n, double* __restrict__ out) {
    double s = 0.0;
    const int64_t n_items = n / VEC;
    if (VEC == 4 && blockIdx.x == 0 && n_items * VEC + threadIdx.x < n) {
        const double v = (double)g[n_items * VEC + threadIdx.x];
        s += v * v;
    }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_items; i += (int64_t)gridDim.x * blockDim.x) {
        if (VEC == 4) {
            const float4 v = reinterpret_cast<const float4*>(g)[i];
            s += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
        } else {
            const double v = (double)g[i];
            s += v * v;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __shared__ double sh[8];
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) t += sh[k];
        atomicAdd(out, t);
    }
}

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float scale, float lr, float b1, float b2, float eps,
                                         float bc1, float bc2) {
    const float gi = g * scale;
    const float mi = b1 * m + (1.0f - b1) * gi;
    const float vi = b2 * v + (1.0f - b2) * gi * gi;
    m = mi;
    v = vi;
    p = p - lr * (mi / bc1) / (sqrtf(vi / bc2) + eps);
}

template <int VEC>
__global__ void __launch_bounds__(256) k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                              float* __restrict__ v, int64_t n, const double* __restrict__ sumsq,
                                              float max_norm, float lr, float b1, float b2, float eps, float bc1, float bc2) {
    // optax.clip_by_global_norm: g * (max_norm / norm) only when norm >= max_norm
    float scale = 1.0f;
    if (max_norm > 0.0f) {
        const float norm = (float)sqrt(*sumsq);
        if (!(norm < max_norm)) scale = max_norm / norm;
    }
    const int64_t n_items = n / VEC;
    if (VEC == 4 && blockIdx.x == 0 && n_items * VEC + threadIdx.x < n) {
        const int64_t i = n_items * VEC + threadIdx.x;
        adam_one(p[i], g[i], m[i], v[i], scale, lr, b1, b2, eps, bc1, bc2);
    }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_items; i += (int64_t)gridDim.x * blockDim.x) {
        if (VEC == 4) {
            float4 pi = reinterpret_cast<float4*>(p)[i], mi = reinterpret_cast<float4*>(m)[i], vi = reinterpret_cast<float4*>(v)[i];
            const float4 gi = reinterpret_cast<const float4*>(g)[i];
            adam_one(pi.x, gi.x, mi.x, vi.x, scale, lr, b1, b2, eps, bc1, bc2);
            adam_one(pi.y, gi.y, mi.y, vi.y, scale, lr, b1, b2, eps, bc1, bc2);
            adam_one(pi.z, gi.z, mi.z, vi.z, scale, lr, b1, b2, eps, bc1, bc2);
            adam_one(pi.w, gi.w, mi.w, vi.w, scale, lr, b1, b2, eps, bc1, bc2);
            reinterpret_cast<float4*>(p)[i] = pi;
            reinterpret_cast<float4*>(m)[i] = mi;
            reinterpret_cast<float4*>(v)[i] = vi;
        } else {
            adam_one(p[i], g[i], m[i], v[i], scale, lr, b1, b2, eps, bc1, bc2);
        }
    }
}

// rows of a [total, width] matrix selected by index -> [B, width] (minibatch of src/update.py:188-206)
template <class T>
__global__ void __launch_bounds__(256) k_gather_rows(const T* __restrict__ src, const int32_t* __restrict__ index,
                                                     T* __restrict__ dst, int64_t B, int width_vec) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * width_vec) return;
    const int64_t b = i / width_vec;
    const int c = (int)(i - b * width_vec);
    dst[b * width_vec + c] = src[(int64_t)index[b] * width_vec + c];
}

}  // namespace brl

using namespace brl;

extern "C" {

int32_t brl_ppo_loss(brl_stream_t stream, void** b, const void* opaque, size_t len) {
    if (opaque == nullptr || len != sizeof(BrlPpoParams))
        return fail(BRL_E_OPAQUE, "brl_ppo_loss: opaque must be one BrlPpoParams (%zu bytes), got %zu", sizeof(BrlPpoParams), len);
    return brl::launch_ppo_loss((cudaStream_t)stream, b, static_cast<const BrlPpoParams*>(opaque), nullptr, nullptr, false, false);
}

}  // extern "C"

// blocks of `threads` threads that form the partial Gram matrices (2 samples per warp and chunk)
int brl::gram_blocks(int64_t B, int threads) {
    const int64_t chunk = 2 * (threads / 32), want = (B + chunk - 1) / chunk;
    return (int)(want < kGramBlocks ? want : kGramBlocks);
}

brl::GramArgs brl::gram_args(void** b, const BrlPpoParams* p, float* stat) {
    GramArgs g{};
    g.logits = static_cast<const float*>(b[0]);
    g.index = static_cast<const int32_t*>(b[2]);
    g.mask = static_cast<const uint8_t*>(b[3]);
    g.B = p->batch;
    g.ticket = reinterpret_cast<unsigned long long*>(static_cast<double*>(b[12]) + kAccGramTicket);
    g.spec = reinterpret_cast<double*>(static_cast<unsigned char*>(b[12]) + 128);
    g.part = reinterpret_cast<float*>(static_cast<unsigned char*>(b[12]) + 512);
    g.stat = stat;
    g.tol = stat ? 3e-4f : 1e-6f;  // statistic only (deferred) vs v_1 feeding the loss gradient
    return g;
}

int32_t brl::launch_ppo_loss(cudaStream_t stream, void** b, const BrlPpoParams* p, void* dz_hi, void* dz_lo, bool acc_zeroed,
                             bool defer_illegal) {
    if (p->batch <= 0) return fail(BRL_E_OPAQUE, "brl_ppo_loss: batch must be > 0");
    static const char* names[] = {"logits", "value", "index", "mask", "action", "old_log_prob", "old_value", "advantages",
                                  "targets", "dlogits", "dvalue", "stats", "scratch"};
    for (int k = 0; k < 13; ++k)
        if (b[k] == nullptr && k != 2) return fail(BRL_E_BUFFER, "brl_ppo_loss: buffer '%s' is NULL", names[k]);
    PpoArgs a{};
    a.logits = static_cast<const float*>(b[0]);
    a.value = static_cast<const float*>(b[1]);
    a.index = static_cast<const int32_t*>(b[2]);
    a.mask = static_cast<const uint8_t*>(b[3]);
    a.action = static_cast<const int32_t*>(b[4]);
    a.old_log_prob = static_cast<const float*>(b[5]);
    a.old_value = static_cast<const float*>(b[6]);
    a.adv = static_cast<const float*>(b[7]);
    a.targets = static_cast<const float*>(b[8]);
    a.dlogits = static_cast<float*>(b[9]);
    a.dvalue = static_cast<float*>(b[10]);
    a.stats = static_cast<float*>(b[11]);
    a.acc = static_cast<double*>(b[12]);
    if (reinterpret_cast<uintptr_t>(b[12]) & 15u) return fail(BRL_E_BUFFER, "brl_ppo_loss: scratch is not 16-byte aligned");
    a.spec = reinterpret_cast<double*>(static_cast<unsigned char*>(b[12]) + 128);
    a.gram_part = reinterpret_cast<float*>(static_cast<unsigned char*>(b[12]) + 512);
    a.dz_hi = static_cast<unsigned short*>(dz_hi);
    a.dz_lo = static_cast<unsigned short*>(dz_lo);
    a.B = p->batch;
    a.clip_eps = p->clip_eps;
    a.ent_coef = p->ent_coef;
    a.vf_coef = p->vf_coef;
    a.ill_coef = p->illegal_l2_coef;
    a.value_clipping = (p->flags & BRL_PPO_VALUE_CLIPPING) != 0;
    a.reward_scaling = (p->flags & BRL_PPO_REWARD_SCALING) != 0;
    a.masked_policy = (p->flags & BRL_PPO_UNMASKED_POLICY) == 0;
    cudaStream_t s = (cudaStream_t)stream;
    if (!acc_zeroed && cudaMemsetAsync(a.acc, 0, 16 * sizeof(double), s) != cudaSuccess) return check_launch("brl_ppo_loss");
    unsigned grid = (unsigned)((a.B + 3) / 4);
    if (grid > 148u * 8u) grid = 148u * 8u;
    const int prepass = a.reward_scaling;
    if (prepass) k_ppo_prepass<<<grid > 32u ? 32u : grid, 128, 0, s>>>(a);
    a.defer_illegal = defer_illegal && a.ill_coef == 0.0f;
    a.skip_illegal = a.defer_illegal && !(p->flags & BRL_PPO_ILLEGAL_STAT);
    if (!a.skip_illegal) {
        const GramArgs g = brl::gram_args(b, p, nullptr);
        const unsigned gram_blocks = (unsigned)brl::gram_blocks(a.B, 1024);
        auto kern = a.defer_illegal ? k_ppo_gram_partial : k_ppo_illegal_gram;
        if (prepass) kern<<<gram_blocks, 1024, 0, s>>>(g);  // a plain launch may not be followed by a PDL one that skips it
        else launch_pdl(kern, dim3(gram_blocks), dim3(1024), 0, s, g);
    }
    launch_pdl(k_ppo_loss, dim3(grid), dim3(128), 0, s, a, prepass);
    return check_launch("brl_ppo_loss");
}

extern "C" {

static int32_t adam_impl(brl_stream_t stream, void** b, const void* opaque, size_t len, bool presummed);
int32_t brl_adam_clip(brl_stream_t stream, void** b, const void* opaque, size_t len) { return adam_impl(stream, b, opaque, len, false); }
int32_t brl_adam_apply(brl_stream_t stream, void** b, const void* opaque, size_t len) { return adam_impl(stream, b, opaque, len, true); }

static int32_t adam_impl(brl_stream_t stream, void** b, const void* opaque, size_t len, bool presummed) {
    if (opaque == nullptr || len != sizeof(BrlAdamParams))
        return fail(BRL_E_OPAQUE, "brl_adam_clip: opaque must be one BrlAdamParams (%zu bytes), got %zu", sizeof(BrlAdamParams), len);
    const BrlAdamParams* p = static_cast<const BrlAdamParams*>(opaque);
    static const char* names[] = {"params", "grads", "m", "v", "scratch"};
    for (int k = 0; k < 5; ++k)
        if (b[k] == nullptr) return fail(BRL_E_BUFFER, "brl_adam_clip: buffer '%s' is NULL", names[k]);
    if (p->n <= 0 || p->step <= 0) return fail(BRL_E_OPAQUE, "brl_adam_clip: n and step (1-based) must be > 0");
    cudaStream_t s = (cudaStream_t)stream;
    double* sumsq = static_cast<double*>(b[4]);
    float* pp = static_cast<float*>(b[0]);
    const float* gg = static_cast<const float*>(b[1]);
    float* mm = static_cast<float*>(b[2]);
    float* vv = static_cast<float*>(b[3]);
    const bool aligned = ((reinterpret_cast<uintptr_t>(pp) | reinterpret_cast<uintptr_t>(gg) | reinterpret_cast<uintptr_t>(mm) |
                           reinterpret_cast<uintptr_t>(vv)) & 15u) == 0;
    const int64_t items = aligned ? (p->n + 3) / 4 : p->n;
    unsigned grid = (unsigned)((items + 255) / 256);
    if (grid > 148u * 8u) grid = 148u * 8u;
    if (!presummed && cudaMemsetAsync(sumsq, 0, sizeof(double), s) != cudaSuccess) return check_launch("brl_adam_clip");
    const float bc1 = 1.0f - powf(p->beta1, (float)p->step), bc2 = 1.0f - powf(p->beta2, (float)p->step);
    if (aligned) {
        if (!presummed) k_sumsq<4><<<grid, 256, 0, s>>>(gg, p->n, sumsq);
        k_adam<4><<<grid, 256, 0, s>>>(pp, gg, mm, vv, p->n, sumsq, p->max_grad_norm, p->lr, p->beta1, p->beta2, p->eps, bc1, bc2);
    } else {
        if (!presummed) k_sumsq<1><<<grid, 256, 0, s>>>(gg, p->n, sumsq);
        k_adam<1><<<grid, 256, 0, s>>>(pp, gg, mm, vv, p->n, sumsq, p->max_grad_norm, p->lr, p->beta1, p->beta2, p->eps, bc1, bc2);
    }
    return check_launch("brl_adam_clip");
}

int32_t brl_gather_rows(brl_stream_t stream, void** b, const void* opaque, size_t len) {
    int32_t rc;
    const BrlParams* p = get_params(opaque, len, &rc);
    if (!p) return rc;
    if (p->n_envs == 0) return BRL_OK;  // an empty batch is a no-op; its buffers may be NULL (zero-size XLA / torch buffers)
    BRL_REQUIRE(b[0], "src");
    if (b[1] == nullptr) return fail(BRL_E_BUFFER, "brl_gather_rows: buffer 'index' is NULL");
    BRL_REQUIRE(b[2], "dst");
    const int64_t B = p->n_envs;
    const int row_bytes = p->k_steps;  // bytes per row
    if (row_bytes <= 0) return fail(BRL_E_OPAQUE, "brl_gather_rows: k_steps (row bytes) must be > 0");
    if (B == 0) return BRL_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if (row_bytes % 16 == 0) {
        const int w = row_bytes / 16;
        k_gather_rows<uint4><<<(unsigned)((B * w + 255) / 256), 256, 0, s>>>(static_cast<const uint4*>(b[0]), static_cast<const int32_t*>(b[1]),
                                                                              static_cast<uint4*>(b[2]), B, w);
    } else if (row_bytes % 2 == 0) {
        const int w = row_bytes / 2;
        k_gather_rows<uint16_t><<<(unsigned)((B * w + 255) / 256), 256, 0, s>>>(static_cast<const uint16_t*>(b[0]), static_cast<const int32_t*>(b[1]),
                                                                                 static_cast<uint16_t*>(b[2]), B, w);
    } else {
        k_gather_rows<uint8_t><<<(unsigned)((B * row_bytes + 255) / 256), 256, 0, s>>>(static_cast<const uint8_t*>(b[0]), static_cast<const int32_t*>(b[1]),
                                                                                       static_cast<uint8_t*>(b[2]), B, row_bytes);
    }
    return check_launch("brl_gather_rows");
}

}  // extern "C"
