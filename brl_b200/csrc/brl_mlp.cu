// brl_mlp.cu -- the policy/value net of the rollout and evaluation loops on the 5th-gen
// tensor cores: the "DeepMind" 480 -> 4 x 1024 ReLU -> {38 logits, 1 value} MLP
// (src/models.py:23-33) as five TMA + tcgen05 GEMM launches.  sm_100a only.
//
// The reference runs this net in fp32 (haiku default).  A single bf16 product would
// change near-tie argmax decisions, so the default mode here is a THREE-TERM bf16 split
//     x . w  ~=  x_hi . w_hi  +  x_lo . w_hi  +  x_hi . w_lo        (hi = bf16(v), lo = bf16(v - hi))
// accumulated in fp32 in tensor memory: ~2^-16 relative error per product, i.e. fp32-class
// results at 1/3 of the bf16 tensor rate (still ~6x the fp32 SIMT GEMM rate).  The first
// layer's input is the 0/1 observation, exact in bf16, so it needs only two terms.
// BRL_F_MLP_BF16 selects the plain single-product mode.
//
// One layer = one launch of k_mlp_layer: C[M, N] = act(A[M, K] . Wt[N, K]^T + b).
//   Persistent, one CTA per SM; CTA tile 128 (envs) x BN (features), K in blocks of 64 bf16 = one 128-byte
//   swizzle row; two TMEM accumulators so the epilogue of one tile overlaps the main loop of the next.
//   warp 0    TMA producer: per K block, bulk-tensor loads of the A / Wt (hi, lo) boxes into a
//             ring of shared-memory stages (128B swizzle), completion on an mbarrier;
//   warp 1    MMA issuer: one lane issues tcgen05.mma (UMMA 128 x BN x 16, kind::f16, fp32
//             accumulate in TMEM); tcgen05.commit hands the stage back / signals the epilogue;
//   warps 2-5 epilogue: tcgen05.ld the accumulator (thread = one env row), + bias, ReLU,
//             split into bf16 hi / lo for the next layer (or fp32 logits + value for the head).
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.h"
#include "env_device.cuh"
#include "mlp_device.cuh"

namespace brl {

// w[in, out] fp32 (haiku layout, y = x @ w + b) -> Wt_hi / Wt_lo [out_pad, in] bf16 (K-major rows)
__global__ void __launch_bounds__(256) k_mlp_pack(const float* __restrict__ w, const float* __restrict__ b,
                                                  const float* __restrict__ w2, const float* __restrict__ b2,
                                                  __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                                  float* __restrict__ bias, int k_in, int n_out, int n_src, int n_src2) {
    // rows [0, n_src) come from w, rows [n_src, n_src + n_src2) from w2 (the value head), the rest are zero
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)n_out * k_in) return;
    int n = (int)(idx / k_in), k = (int)(idx % k_in);
    float v = 0.0f;
    if (n < n_src) v = w[(size_t)k * n_src + n];
    else if (n < n_src + n_src2) v = w2[(size_t)k * n_src2 + (n - n_src)];
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[idx] = h;
    lo[idx] = __float2bfloat16_rn(v - __bfloat162float(h));
    if (k == 0) bias[n] = n < n_src ? b[n] : (n < n_src + n_src2 ? b2[n - n_src] : 0.0f);
}

// observation f32 / u8 (0/1) -> bf16, for callers whose env emits the reference dtypes
template <class T>
__global__ void __launch_bounds__(256) k_obs_to_bf16(const T* __restrict__ in, __nv_bfloat16* __restrict__ out, int64_t n) {
    int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (i >= n) return;
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        uint32_t a = (float)in[i + 2 * k] != 0.0f ? 0x3F80u : 0u, c = (float)in[i + 2 * k + 1] != 0.0f ? 0x3F80u : 0u;
        w[k] = a | (c << 16);
    }
    *reinterpret_cast<uint4*>(out + i) = make_uint4(w[0], w[1], w[2], w[3]);
}


// ---- epilogues: thread = one env row of the accumulator (TMEM lane), shared by all layer kernels ----------
// hidden layer: + bias, ReLU, split into bf16 hi / lo for the next layer (128-bit stores)
#ifdef BRL_MLP_TRACE
__device__ unsigned long long g_mlp_epi[4];  // debug: SM clocks warp 2 spent in {tcgen05.ld + wait, arithmetic, stores, chunks}
#define BRL_EPI_T0() long long _e0 = clock64()
#define BRL_EPI_ACC(k) do { long long _e1 = clock64(); if ((threadIdx.x & 31) == 0 && (threadIdx.x >> 5) == 2 && blockIdx.x == 0) atomicAdd(&g_mlp_epi[k], (unsigned long long)(_e1 - _e0)); _e0 = _e1; } while (0)
#else
#define BRL_EPI_T0()
#define BRL_EPI_ACC(k)
#endif

// SMEM_BIAS: `bias` points into shared memory (the fused forward stages every layer's bias once per CTA: with ~194 KB of the
// SM given to the operand ring, what is left of L1 does not keep the 4 KB bias vector against the tile's own 131 KB of
// stores, and 16 global loads per 32-column chunk each waited a full L2 round trip -- scripts/exp_mlp_trace.py)
template <bool SMEM_BIAS = false>
__device__ __forceinline__ void epilogue_hidden_row(uint32_t t_row, int n_cols, const float* __restrict__ bias, bool row_ok,
                                                    __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo) {
    BRL_EPI_T0();
#pragma unroll 1
    for (int c0 = 0; c0 < n_cols; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(t_row + (uint32_t)c0, r);
        BRL_EPI_ACC(0);
        if (row_ok) {
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
                const float2 bb = SMEM_BIAS ? reinterpret_cast<const float2*>(bias + c0)[jj]
                                            : __ldg(reinterpret_cast<const float2*>(bias + c0) + jj);  // warp-uniform
                float x0 = fmaxf(__uint_as_float(r[2 * jj]) + bb.x, 0.0f);
                float x1 = fmaxf(__uint_as_float(r[2 * jj + 1]) + bb.y, 0.0f);
                __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
                hi[jj] = *reinterpret_cast<uint32_t*>(&h);
                lo[jj] = pack_bf16x2(x0 - __low2float(h), x1 - __high2float(h));
            }
            BRL_EPI_ACC(1);
            uint4* ph = reinterpret_cast<uint4*>(out_hi + c0);
#pragma unroll
            for (int v = 0; v < 4; ++v) ph[v] = make_uint4(hi[4 * v], hi[4 * v + 1], hi[4 * v + 2], hi[4 * v + 3]);
            if (out_lo) {
                uint4* pl = reinterpret_cast<uint4*>(out_lo + c0);
#pragma unroll
                for (int v = 0; v < 4; ++v) pl[v] = make_uint4(lo[4 * v], lo[4 * v + 1], lo[4 * v + 2], lo[4 * v + 3]);
            }
            BRL_EPI_ACC(2);
        }
    }
}

// Hidden-layer epilogue of the fused forward with TMA stores.  A thread owns one accumulator row; written straight to global
// memory its 16-byte pieces land in 32 different rows per warp instruction (32 LSU wavefronts each, 8192 per 128 x 256
// tile: ~4 us of the SM's load/store pipe -- the epilogue took 10-12 us per tile, longer than layer 0's 6 us main loop and on
// every layer's dependency path, scripts/exp_mlp_trace.py).  Here each 32-column chunk is staged in shared memory as the
// 32-row x 64-byte box of a 64B-swizzled tensor map (lane = row, conflict-free 128-bit stores) and one lane issues a bulk
// tensor store for the hi and the lo block; rows past M are clipped by the map.  `stage` = this warp's 2 x 2 KB.
__device__ __forceinline__ void epilogue_hidden_tma(uint32_t t_row, int n_cols, const float* __restrict__ bias, int lane,
                                                    uint32_t stage, const CUtensorMap* st_hi, const CUtensorMap* st_lo,
                                                    int col0, int row0) {
    BRL_EPI_T0();
    const uint32_t sw = (uint32_t)((lane >> 1) & 3);  // 64B swizzle: 16-byte chunk index ^= bits 7-8 of the byte offset
#pragma unroll 1
    for (int c0 = 0; c0 < n_cols; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(t_row + (uint32_t)c0, r);
        BRL_EPI_ACC(0);
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) {
            const float2 bb = __ldg(reinterpret_cast<const float2*>(bias + c0) + jj);  // warp-uniform
            float x0 = fmaxf(__uint_as_float(r[2 * jj]) + bb.x, 0.0f);
            float x1 = fmaxf(__uint_as_float(r[2 * jj + 1]) + bb.y, 0.0f);
            __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
            hi[jj] = *reinterpret_cast<uint32_t*>(&h);
            lo[jj] = pack_bf16x2(x0 - __low2float(h), x1 - __high2float(h));
        }
        BRL_EPI_ACC(1);
        if (c0 > 0) {  // the previous chunk's stores have read the staging blocks
            if (lane == 0) bulk_wait_read0();
            __syncwarp();
        }
        const uint32_t row_hi = stage + (uint32_t)lane * 64u, row_lo = row_hi + 2048u;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const uint32_t off = ((uint32_t)v ^ sw) * 16u;
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_hi + off), "r"(hi[4 * v]), "r"(hi[4 * v + 1]),
                         "r"(hi[4 * v + 2]), "r"(hi[4 * v + 3]) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_lo + off), "r"(lo[4 * v]), "r"(lo[4 * v + 1]),
                         "r"(lo[4 * v + 2]), "r"(lo[4 * v + 3]) : "memory");
        }
        fence_async_smem();  // generic-proxy writes -> async-proxy (TMA) reads
        __syncwarp();
        if (lane == 0) {
            tma_store_2d(st_hi, stage, col0 + c0, row0);
            tma_store_2d(st_lo, stage + 2048u, col0 + c0, row0);
            bulk_commit();
        }
        BRL_EPI_ACC(2);
    }
}

// head tile with the masked categorical of src/roll_out.py:27-30,79-81 fused in: the thread that owns an env row
// holds its 38 logits in registers, so where(mask, logits, -inf), the Gumbel-argmax sample (same Philox stream as
// brl_categorical: counter (env, GUM0 + a / 4, step), key = seed) or the first-argmax mode, and
// log_softmax(masked)[action] cost no extra launch and no logits round trip.
struct ActArgs {
    const uint8_t* mask;   // [n, 38] or NULL (unmasked)
    int32_t* action;       // [n]; NULL = no categorical fused (plain forward)
    float* log_prob;       // [n] or NULL
    uint64_t seed;
    int64_t env_offset;
    uint32_t step;
    int sample;
    const int32_t* row_index;  // [n] or NULL: row r of the batch is env row_index[r] (mask / action / log_prob / value / RNG by env)
    const uint64_t* seed_salt; // device u64 or NULL: XORed into `seed` (a captured CUDA graph replays with fresh noise)
};

__device__ __forceinline__ void epilogue_head_act_row(uint32_t t_row, const float* __restrict__ bias, bool row_ok, int64_t row,
                                                      float* __restrict__ logits, float* __restrict__ value, const ActArgs& act) {
    uint32_t r0[32], r1[32];
    tmem_ld32(t_row, r0);
    tmem_ld32(t_row + 32u, r1);
    if (!row_ok) return;
    float l[40];
#pragma unroll
    for (int k = 0; k < 32; ++k) l[k] = __uint_as_float(r0[k]) + __ldg(bias + k);
#pragma unroll
    for (int k = 0; k < 7; ++k) l[32 + k] = __uint_as_float(r1[k]) + __ldg(bias + 32 + k);
    l[39] = 0.0f;
    if (act.row_index) row = act.row_index[row];  // a listed batch: every per-env array below is addressed by env
    if (logits) {
        float2* pl = reinterpret_cast<float2*>(logits + row * 38);
#pragma unroll
        for (int jj = 0; jj < 19; ++jj) pl[jj] = make_float2(l[2 * jj], l[2 * jj + 1]);
    }
    if (value) value[row] = l[38];
    if (act.action == nullptr) return;
    uint64_t legal = ~0ull;
    if (act.mask) {
        const uint16_t* pm = reinterpret_cast<const uint16_t*>(act.mask + row * 38);  // 38 * row is even
        legal = 0;
#pragma unroll
        for (int jj = 0; jj < 19; ++jj) {
            const uint32_t m2 = pm[jj];
            legal |= (uint64_t)((m2 & 0xFFu) != 0) << (2 * jj) | (uint64_t)((m2 >> 8) != 0) << (2 * jj + 1);
        }
    }
    const uint64_t g = (uint64_t)(act.env_offset + row);
    const uint64_t seed = act.seed_salt ? (act.seed ^ __ldg(act.seed_salt)) : act.seed;
    float best = -INFINITY, mx = -INFINITY, la = 0.0f;
    int best_a = kNumActions;
#pragma unroll
    for (int grp = 0; grp < 10; ++grp) {
        uint32_t w[4] = {0u, 0u, 0u, 0u};
        if (act.sample) {
            const uint4 rr = philox4x32(make_uint4((uint32_t)g, (uint32_t)(g >> 32), kTagGum + (uint32_t)grp, act.step),
                                        make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
            w[0] = rr.x; w[1] = rr.y; w[2] = rr.z; w[3] = rr.w;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int aidx = 4 * grp + e;
            if (aidx < kNumActions && ((legal >> aidx) & 1ull)) {
                float v = l[aidx];
                if (act.sample) {
                    const float u = ((float)(w[e] >> 8) + 0.5f) * (1.0f / 16777216.0f);
                    v += -logf(-logf(u));
                }
                if (v > best) { best = v; best_a = aidx; la = l[aidx]; }  // ascending: ties keep the lower index
                mx = fmaxf(mx, l[aidx]);
            }
        }
    }
    float se = 0.0f;
#pragma unroll
    for (int aidx = 0; aidx < kNumActions; ++aidx)
        if ((legal >> aidx) & 1ull) se += expf(l[aidx] - mx);
    if (best_a >= kNumActions) { best_a = 0; la = l[0]; }  // no legal action: cannot happen for a valid mask
    act.action[row] = best_a;
    if (act.log_prob) act.log_prob[row] = la - mx - logf(se);
}

// ---- one layer ------------------------------------------------------------------------------
struct LayerArgs {
    const float* bias;           // [n_pad]
    __nv_bfloat16* out_hi;       // hidden: [M, n_total] bf16
    __nv_bfloat16* out_lo;       // hidden, split mode: residual (may be NULL)
    float* logits;               // head: [M, 38]
    float* value;                // head: [M]
    int M, n_total, k_blocks;
    int n_tiles_n, n_tiles;      // tiles along N, total tiles of this layer
};

template <int BN, bool SPLIT_A, bool SPLIT_W>
struct LayerCfg {
    static constexpr uint32_t kABytes = kBM * kBK * 2, kWBytes = BN * kBK * 2;
    static constexpr uint32_t kStageBytes = kABytes * (SPLIT_A ? 2 : 1) + kWBytes * (SPLIT_W ? 2 : 1);
    static constexpr int kStages = (kSmemBudget / kStageBytes) > 8 ? 8 : (int)(kSmemBudget / kStageBytes);
    static constexpr uint32_t kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
};

template <int BN, bool SPLIT_A, bool SPLIT_W, bool HEAD>
__global__ void __launch_bounds__(kMlpThreads, 1)
k_mlp_layer(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
            const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo, const LayerArgs a) {
    // PERSISTENT: one CTA per SM walks tiles blockIdx.x, blockIdx.x + gridDim.x, ... (n-tile fastest, so the
    // CTAs running together share A row-blocks in L2).  The smem stage ring runs straight across tile
    // boundaries and the accumulator is double-buffered in TMEM, so the epilogue of tile j overlaps the
    // TMA + MMA main loop of tile j + 1.
    using Cfg = LayerCfg<BN, SPLIT_A, SPLIT_W>;
    constexpr int S = Cfg::kStages;
    constexpr uint32_t kTmemCols = 2 * BN;  // two accumulators of BN fp32 columns (power of two >= 32)
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t base = (smem_addr(smem_dyn) + 1023u) & ~1023u;  // 128B swizzle atoms need 1024-byte alignment
    const uint32_t bar_base = base + S * Cfg::kStageBytes;         // full[S], empty[S], tmem_full[2], tmem_empty[2]
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (S + s); };
    auto tmem_full_bar = [&](int b) { return bar_base + 8u * (2 * S + b); };
    auto tmem_empty_bar = [&](int b) { return bar_base + 8u * (2 * S + 2 + b); };
    __shared__ uint32_t tmem_base_s;  // written by tcgen05.alloc

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles_n = a.n_tiles_n, n_tiles = a.n_tiles;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_a_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_w_hi) : "memory");
        if (SPLIT_A) asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_a_lo) : "memory");
        if (SPLIT_W) asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_w_lo) : "memory");
        for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tmem_full_bar(b), 1); mbar_init(tmem_empty_bar(b), 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(smem_addr(&tmem_base_s), kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_acc = *reinterpret_cast<volatile uint32_t*>(&tmem_base_s);

    if (warp == 0) {
        if (lane == 0) {  // ===== TMA producer =====
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int m0 = (tile / n_tiles_n) * kBM, n0 = (tile % n_tiles_n) * BN;
                for (int kb = 0; kb < a.k_blocks; ++kb, ++it) {
                    const int s = (int)(it % S);
                    mbar_wait(empty_bar(s), ((it / S) & 1u) ^ 1u);
                    mbar_expect_tx(full_bar(s), Cfg::kStageBytes);
                    uint32_t dst = base + s * Cfg::kStageBytes;
                    tma_load_2d(dst, &tm_a_hi, full_bar(s), kb * kBK, m0);
                    dst += Cfg::kABytes;
                    if (SPLIT_A) { tma_load_2d(dst, &tm_a_lo, full_bar(s), kb * kBK, m0); dst += Cfg::kABytes; }
                    tma_load_2d(dst, &tm_w_hi, full_bar(s), kb * kBK, n0);
                    dst += Cfg::kWBytes;
                    if (SPLIT_W) tma_load_2d(dst, &tm_w_lo, full_bar(s), kb * kBK, n0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ===== MMA issuer =====
            constexpr uint32_t idesc = umma_idesc_bf16(kBM, BN);
            uint32_t it = 0, j = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++j) {
                const uint32_t buf = j & 1u;
                mbar_wait(tmem_empty_bar(buf), ((j >> 1) & 1u) ^ 1u);  // the epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t acc = tmem_acc + buf * BN;
                for (int kb = 0; kb < a.k_blocks; ++kb, ++it) {
                    const int s = (int)(it % S);
                    mbar_wait(full_bar(s), (it / S) & 1u);
                    tc_fence_after();
                    const uint32_t sa_hi = base + s * Cfg::kStageBytes;
                    const uint32_t sa_lo = sa_hi + Cfg::kABytes;
                    const uint32_t sw_hi = sa_hi + Cfg::kABytes * (SPLIT_A ? 2 : 1);
                    const uint32_t sw_lo = sw_hi + Cfg::kWBytes;
#pragma unroll
                    for (int k = 0; k < kBK / kUmmaK; ++k) {
                        const uint32_t koff = (uint32_t)k * kUmmaK * 2;  // bytes along K inside the 128-byte row
                        const uint64_t da_hi = umma_desc_sw128(sa_hi + koff), dw_hi = umma_desc_sw128(sw_hi + koff);
                        umma_bf16(acc, da_hi, dw_hi, idesc, (kb | k) != 0);
                        if (SPLIT_A) umma_bf16(acc, umma_desc_sw128(sa_lo + koff), dw_hi, idesc, 1u);
                        if (SPLIT_W) umma_bf16(acc, da_hi, umma_desc_sw128(sw_lo + koff), idesc, 1u);
                    }
                    umma_commit(empty_bar(s));  // the stage is free once these MMAs have read it
                }
                umma_commit(tmem_full_bar(buf));
            }
        }
    } else {  // ===== epilogue: warps 2..5, TMEM lane quarter = warp % 4 =====
        const int q = warp & 3;
        uint32_t j = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++j) {
            const int m0 = (tile / n_tiles_n) * kBM, n0 = (tile % n_tiles_n) * BN;
            const uint32_t buf = j & 1u;
            mbar_wait(tmem_full_bar(buf), (j >> 1) & 1u);
            tc_fence_after();
            const int row = m0 + q * 32 + lane;
            const uint32_t t_row = tmem_acc + buf * BN + ((uint32_t)(q * 32) << 16);
            const float* bias = a.bias + n0;
            if (!HEAD)
                epilogue_hidden_row(t_row, BN, bias, row < a.M, a.out_hi + (size_t)row * a.n_total + n0,
                                    a.out_lo ? a.out_lo + (size_t)row * a.n_total + n0 : nullptr);
            else
                epilogue_head_row(t_row, bias, row < a.M, a.logits + (size_t)row * 38, a.value + row);
            // all of this warp's tcgen05.ld have completed (wait::ld): hand the accumulator back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tmem_empty_bar(buf)) : "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_acc, kTmemCols);
}

// ---- one hidden layer on CTA PAIRS (cta_group::2) ------------------------------------------------
// The 1-CTA tiles above are bound by the L2 -> SM operand fill (~64 B/clk/SM against the split mode's
// 83 B/clk demand at 128 x 128).  A CTA pair (cluster of 2 = the two SMs of a TPC) computes a 256 x BN tile
// with ONE tcgen05.mma.cta_group::2 stream issued by the leader: each CTA stages its own 128 env rows of A and
// only HALF of the Wt rows (the MMA reads the other half out of the peer's shared memory), and accumulates its
// own 128 x BN block in its own TMEM.  Operand bytes per MMA clock: 42 B/clk/SM at BN = 256.
//   full[s]        lives in the leader; both CTAs' TMA loads complete_tx on it (cta_group::2, peer bit masked);
//   empty[s]       one per CTA, released by the leader's multicast tcgen05.commit;
//   tmem_full[b]   one per CTA, multicast commit;   tmem_empty[b]  in the leader, 8 arrivals (4 epilogue warps x 2).
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // shared::cluster address of the same offset in the pair's even CTA

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {  // arrives on `bar`'s offset in BOTH CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}

template <int BN, bool SPLIT_A, bool SPLIT_W>
struct PairCfg {
    static constexpr uint32_t kABytes = kBM * kBK * 2, kWBytes = (BN / 2) * kBK * 2;  // per CTA: own A rows, half of Wt
    static constexpr uint32_t kStageBytes = kABytes * (SPLIT_A ? 2 : 1) + kWBytes * (SPLIT_W ? 2 : 1);
    static constexpr int kStages = (kSmemBudget / kStageBytes) > 8 ? 8 : (int)(kSmemBudget / kStageBytes);
    static constexpr uint32_t kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
};

template <int BN, bool SPLIT_A, bool SPLIT_W>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kMlpThreads, 1)
k_mlp_layer_pair(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                 const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo, const LayerArgs a) {
    using Cfg = PairCfg<BN, SPLIT_A, SPLIT_W>;
    constexpr int S = Cfg::kStages;
    constexpr uint32_t kTmemCols = 2 * BN;
    static_assert(kTmemCols <= 512, "two accumulators must fit tensor memory");
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t base = (smem_addr(smem_dyn) + 1023u) & ~1023u;  // same offset in both CTAs of the pair
    const uint32_t bar_base = base + S * Cfg::kStageBytes;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (S + s); };
    auto tmem_full_bar = [&](int b) { return bar_base + 8u * (2 * S + b); };
    auto tmem_empty_bar = [&](int b) { return bar_base + 8u * (2 * S + 2 + b); };
    const uint32_t store_base = (bar_base + 256u + 1023u) & ~1023u;  // 8 epilogue warps x (hi, lo) x 2 KB store staging
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int n_tiles_n = a.n_tiles_n, n_tiles = a.n_tiles;       // pair tiles: 256 rows x BN columns
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_a_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_w_hi) : "memory");
        if (SPLIT_A) asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_a_lo) : "memory");
        if (SPLIT_W) asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_w_lo) : "memory");
        for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tmem_full_bar(b), 1); mbar_init(tmem_empty_bar(b), 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {  // the same warp of both CTAs allocates the pair's tensor memory
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&tmem_base_s)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();  // barriers of both CTAs initialised before any remote arrive / multicast commit
    tc_fence_after();
    const uint32_t tmem_acc = *reinterpret_cast<volatile uint32_t*>(&tmem_base_s);

    if (warp == 0) {
        if (lane == 0) {  // ===== TMA producer (both CTAs: own A rows, own half of the Wt rows) =====
            const uint32_t lead_full0 = full_bar(0) & kPeerBitMask;
            uint32_t it = 0;
            for (int tile = pair; tile < n_tiles; tile += n_pairs) {
                const int m0 = (tile / n_tiles_n) * (2 * kBM) + (int)rank * kBM;
                const int n0 = (tile % n_tiles_n) * BN + (int)rank * (BN / 2);
                for (int kb = 0; kb < a.k_blocks; ++kb, ++it) {
                    const int s = (int)(it % S);
                    mbar_wait(empty_bar(s), ((it / S) & 1u) ^ 1u);
                    if (leader) mbar_expect_tx(full_bar(s), 2u * Cfg::kStageBytes);  // both CTAs' bytes land on the leader's barrier
                    const uint32_t fb = lead_full0 + 8u * s;
                    uint32_t dst = base + s * Cfg::kStageBytes;
                    tma_load_2d_pair(dst, &tm_a_hi, fb, kb * kBK, m0);
                    dst += Cfg::kABytes;
                    if (SPLIT_A) { tma_load_2d_pair(dst, &tm_a_lo, fb, kb * kBK, m0); dst += Cfg::kABytes; }
                    tma_load_2d_pair(dst, &tm_w_hi, fb, kb * kBK, n0);
                    dst += Cfg::kWBytes;
                    if (SPLIT_W) tma_load_2d_pair(dst, &tm_w_lo, fb, kb * kBK, n0);
                }
            }
        }
    } else if (warp == 1) {
        if (leader && lane == 0) {  // ===== MMA issuer: the leader drives both SMs' tensor cores =====
            constexpr uint32_t idesc = umma_idesc_bf16(2 * kBM, BN);
            uint32_t it = 0, j = 0;
            for (int tile = pair; tile < n_tiles; tile += n_pairs, ++j) {
                const uint32_t buf = j & 1u;
                mbar_wait(tmem_empty_bar(buf), ((j >> 1) & 1u) ^ 1u);  // both CTAs' epilogues have drained this accumulator
                tc_fence_after();
                const uint32_t acc = tmem_acc + buf * BN;
                for (int kb = 0; kb < a.k_blocks; ++kb, ++it) {
                    const int s = (int)(it % S);
                    mbar_wait(full_bar(s), (it / S) & 1u);
                    tc_fence_after();
                    const uint32_t sa_hi = base + s * Cfg::kStageBytes;
                    const uint32_t sa_lo = sa_hi + Cfg::kABytes;
                    const uint32_t sw_hi = sa_hi + Cfg::kABytes * (SPLIT_A ? 2 : 1);
                    const uint32_t sw_lo = sw_hi + Cfg::kWBytes;
#pragma unroll
                    for (int k = 0; k < kBK / kUmmaK; ++k) {
                        const uint32_t koff = (uint32_t)k * kUmmaK * 2;
                        const uint64_t da_hi = umma_desc_sw128(sa_hi + koff), dw_hi = umma_desc_sw128(sw_hi + koff);
                        umma_bf16_pair(acc, da_hi, dw_hi, idesc, (kb | k) != 0);
                        if (SPLIT_A) umma_bf16_pair(acc, umma_desc_sw128(sa_lo + koff), dw_hi, idesc, 1u);
                        if (SPLIT_W) umma_bf16_pair(acc, da_hi, umma_desc_sw128(sw_lo + koff), idesc, 1u);
                    }
                    umma_commit_pair(empty_bar(s));
                }
                umma_commit_pair(tmem_full_bar(buf));
            }
        }
    } else {  // ===== epilogue: warps 2..5 of each CTA drain that CTA's own 128 rows =====
        const int q = warp & 3;
        const uint32_t lead_tmem_empty0 = tmem_empty_bar(0) & kPeerBitMask;
        uint32_t j = 0;
        for (int tile = pair; tile < n_tiles; tile += n_pairs, ++j) {
            const int m0 = (tile / n_tiles_n) * (2 * kBM) + (int)rank * kBM, n0 = (tile % n_tiles_n) * BN;
            const uint32_t buf = j & 1u;
            mbar_wait(tmem_full_bar(buf), (j >> 1) & 1u);
            tc_fence_after();
            const int row = m0 + q * 32 + lane;
            const uint32_t t_row = tmem_acc + buf * BN + ((uint32_t)(q * 32) << 16);
            const float* bias = a.bias + n0;
            epilogue_hidden_row(t_row, BN, bias, row < a.M, a.out_hi + (size_t)row * a.n_total + n0,
                                a.out_lo ? a.out_lo + (size_t)row * a.n_total + n0 : nullptr);
            tc_fence_before();
            __syncwarp();
            if (lane == 0)  // remote arrive on the leader's barrier (local for the leader itself)
                asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(lead_tmem_empty0 + 8u * buf) : "memory");
        }
    }
    tc_fence_before();
    cluster_sync_all();  // no CTA of the pair may exit (or free tensor memory) while the other still uses its smem / barriers
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(kTmemCols) : "memory");
}

// ---- the whole forward as ONE persistent launch on CTA pairs ----------------------------------------
// Five separate launches lose (i) a wave-quantisation tail per layer (8192 envs = 128 pair tiles on 74 pairs =
// 1.73 waves) and (ii) the drain / launch / pipeline-fill gap between layers.  Here all 4 x 128 + 32 tiles of the
// five layers form one list, walked round-robin by the persistent pairs; a tile of layer l + 1 needs only the
// 128-row block of layer l's output it reads, so the next layer's first tiles fill the previous layer's tail.
// Dependency tracking: every epilogue warp publishes (st.global ... fence, red.add) into ready[l][128-row block];
// the TMA producer acquires the count (16 = 4 n-tiles x 4 warps) and crosses into the async proxy before its
// first A load of the tile.  Tiles are claimed in list order and depend only on earlier tiles, and the grid never
// exceeds the co-resident cluster count, so the waits cannot deadlock (spins are bounded: a bug traps).
struct FusedArgs {
    CUtensorMap a_hi[5], a_lo[5], w_hi[5], w_lo[5];
    CUtensorMap st_hi[4], st_lo[4];  // the hidden layers' outputs as TMA-store destinations (32 x 32 boxes)
    const float* bias[5];
    __nv_bfloat16* out_hi[4];
    __nv_bfloat16* out_lo[4];
    float* logits;
    float* value;
    uint32_t* ready;  // [4][2 * nmb], zeroed before the launch
    int M, nmb, tiles_per_layer, n_tiles;
    ActArgs act;
};

constexpr int kFusedBN = 256, kFusedTilesN = kHidden / kFusedBN;
// Epilogue warps of the fused forward: TWO per TMEM lane quadrant, each draining half of the tile's 256 columns.  One
// warp per quadrant sits alone on its SM sub-partition and runs the bias / ReLU / bf16-split chain at ~7 clocks per
// instruction (scripts/exp_mlp_trace.py: 10-12 us per tile, longer than layer 0's 6 us main loop and on every layer's
// dependency path); two warps of a quadrant share a sub-partition and interleave.
constexpr int kFusedEpiWarps = 8, kFusedThreads = 32 * (2 + kFusedEpiWarps);
constexpr uint32_t kReadyPerBlock = kFusedTilesN * kFusedEpiWarps;  // n-tiles x epilogue warps

template <bool SPLIT>
struct FusedCfg {
    static constexpr uint32_t kABytes = kBM * kBK * 2, kWBytes = (kFusedBN / 2) * kBK * 2, kWHeadBytes = (kHeadPad / 2) * kBK * 2;
    static constexpr uint32_t kStageBytes = (kABytes + kWBytes) * (SPLIT ? 2 : 1);
    static constexpr int kStages = (kSmemBudget / kStageBytes) > 8 ? 8 : (int)(kSmemBudget / kStageBytes);
    static constexpr uint32_t kStoreStageBytes = 8 /*epilogue warps*/ * 2 /*hi, lo*/ * 2048;  // 32 rows x 64 B per warp and array
    static constexpr uint32_t kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/ + 1024 /*align*/ + kStoreStageBytes;
    static_assert(kSmemBytes <= 227 * 1024, "operand ring + store staging exceed the SM's shared memory");
};

struct FusedTile { int layer, mb, nt; };
__device__ __forceinline__ FusedTile fused_tile(int t, int tiles_per_layer) {
    FusedTile f;
    f.layer = t / tiles_per_layer;
    if (f.layer < 4) {
        const int r = t - f.layer * tiles_per_layer;
        f.mb = r / kFusedTilesN;
        f.nt = r % kFusedTilesN;
    } else {
        f.layer = 4;
        f.mb = t - 4 * tiles_per_layer;
        f.nt = 0;
    }
    return f;
}

#ifdef BRL_MLP_TRACE
// debug build (scripts/exp_mlp_trace.py): per-tile %globaltimer stamps of the fused forward, written by the leader CTA's role lanes
// [0] dependency wait begins [1] dependency ready [2] last TMA issued [3] accumulator free, first MMA [4] last MMA committed
// [5] accumulator seen by the epilogue [6] tile published [7] pair
__device__ unsigned long long g_mlp_trace[1024][8];
__device__ __forceinline__ unsigned long long mlp_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define BRL_MLP_STAMP(tile, k) do { if (leader && (tile) < 1024) g_mlp_trace[tile][k] = mlp_now(); } while (0)
#else
#define BRL_MLP_STAMP(tile, k)
#endif

template <bool SPLIT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kFusedThreads, 1) k_mlp_fused(const __grid_constant__ FusedArgs a) {
    using Cfg = FusedCfg<SPLIT>;
    constexpr int S = Cfg::kStages;
    constexpr uint32_t kTmemCols = 2 * kFusedBN;
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t base = (smem_addr(smem_dyn) + 1023u) & ~1023u;
    const uint32_t bar_base = base + S * Cfg::kStageBytes;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (S + s); };
    auto tmem_full_bar = [&](int b) { return bar_base + 8u * (2 * S + b); };
    auto tmem_empty_bar = [&](int b) { return bar_base + 8u * (2 * S + 2 + b); };
    const uint32_t store_base = (bar_base + 256u + 1023u) & ~1023u;  // 8 epilogue warps x (hi, lo) x 2 KB store staging
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int n_tiles = a.n_tiles, tpl = a.tiles_per_layer;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tmem_full_bar(b), 1); mbar_init(tmem_empty_bar(b), 2 * kFusedEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&tmem_base_s)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_acc = *reinterpret_cast<volatile uint32_t*>(&tmem_base_s);

    if (warp == 0) {
        if (lane == 0) {  // ===== TMA producer (both CTAs) =====
            const uint32_t lead_full0 = full_bar(0) & kPeerBitMask;
            uint32_t it = 0;
            for (int tile = pair; tile < n_tiles; tile += n_pairs) {
                const FusedTile f = fused_tile(tile, tpl);
                const bool split_a = SPLIT && f.layer > 0;  // the 0/1 observation is exact in bf16
                const int k_blocks = f.layer == 0 ? (kObsDimM + kBK - 1) / kBK : kHidden / kBK;
                const uint32_t w_bytes = f.layer == 4 ? Cfg::kWHeadBytes : Cfg::kWBytes;
                const uint32_t tx = Cfg::kABytes * (split_a ? 2u : 1u) + w_bytes * (SPLIT ? 2u : 1u);
                const int m0 = f.mb * (2 * kBM) + (int)rank * kBM;
                const int n0 = f.layer == 4 ? (int)rank * (kHeadPad / 2) : f.nt * kFusedBN + (int)rank * (kFusedBN / 2);
                const CUtensorMap* ma_hi = &a.a_hi[f.layer];
                const CUtensorMap* ma_lo = &a.a_lo[f.layer];
                const CUtensorMap* mw_hi = &a.w_hi[f.layer];
                const CUtensorMap* mw_lo = &a.w_lo[f.layer];
                BRL_MLP_STAMP(tile, 0);
                if (f.layer > 0) {  // the previous layer's rows [m0, m0 + 128) must be complete (all 4 n-tiles)
                    const uint32_t* flag = a.ready + (size_t)(f.layer - 1) * 2 * a.nmb + 2 * f.mb + rank;
                    uint32_t v = 0;
                    for (uint32_t spin = 0;; ++spin) {
                        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
                        if (v >= kReadyPerBlock) break;
                        if (spin > (1u << 24)) __trap();
                        __nanosleep(64);
                    }
                    asm volatile("fence.proxy.async;" ::: "memory");  // generic-proxy writes -> async-proxy (TMA) reads
                }
                BRL_MLP_STAMP(tile, 1);
                for (int kb = 0; kb < k_blocks; ++kb, ++it) {
                    const int s = (int)(it % S);
                    mbar_wait(empty_bar(s), ((it / S) & 1u) ^ 1u);
                    if (leader) mbar_expect_tx(full_bar(s), 2u * tx);
                    const uint32_t fb = lead_full0 + 8u * s;
                    const uint32_t sa = base + s * Cfg::kStageBytes;
                    const uint32_t sw = sa + Cfg::kABytes * (SPLIT ? 2 : 1);
                    tma_load_2d_pair(sa, ma_hi, fb, kb * kBK, m0);
                    if (split_a) tma_load_2d_pair(sa + Cfg::kABytes, ma_lo, fb, kb * kBK, m0);
                    tma_load_2d_pair(sw, mw_hi, fb, kb * kBK, n0);
                    if (SPLIT) tma_load_2d_pair(sw + Cfg::kWBytes, mw_lo, fb, kb * kBK, n0);
                }
                BRL_MLP_STAMP(tile, 2);
            }
        }
    } else if (warp == 1) {
        if (leader && lane == 0) {  // ===== MMA issuer =====
            uint32_t it = 0, j = 0;
            for (int tile = pair; tile < n_tiles; tile += n_pairs, ++j) {
                const FusedTile f = fused_tile(tile, tpl);
                const bool split_a = SPLIT && f.layer > 0;
                const int k_blocks = f.layer == 0 ? (kObsDimM + kBK - 1) / kBK : kHidden / kBK;
                const uint32_t idesc = f.layer == 4 ? umma_idesc_bf16(2 * kBM, kHeadPad) : umma_idesc_bf16(2 * kBM, kFusedBN);
                const uint32_t buf = j & 1u;
                mbar_wait(tmem_empty_bar(buf), ((j >> 1) & 1u) ^ 1u);
                tc_fence_after();
                BRL_MLP_STAMP(tile, 3);
                const uint32_t acc = tmem_acc + buf * kFusedBN;
                for (int kb = 0; kb < k_blocks; ++kb, ++it) {
                    const int s = (int)(it % S);
                    mbar_wait(full_bar(s), (it / S) & 1u);
                    tc_fence_after();
                    const uint32_t sa_hi = base + s * Cfg::kStageBytes;
                    const uint32_t sa_lo = sa_hi + Cfg::kABytes;
                    const uint32_t sw_hi = sa_hi + Cfg::kABytes * (SPLIT ? 2 : 1);
                    const uint32_t sw_lo = sw_hi + Cfg::kWBytes;
#pragma unroll
                    for (int k = 0; k < kBK / kUmmaK; ++k) {
                        const uint32_t koff = (uint32_t)k * kUmmaK * 2;
                        const uint64_t da_hi = umma_desc_sw128(sa_hi + koff), dw_hi = umma_desc_sw128(sw_hi + koff);
                        umma_bf16_pair(acc, da_hi, dw_hi, idesc, (kb | k) != 0);
                        if (split_a) umma_bf16_pair(acc, umma_desc_sw128(sa_lo + koff), dw_hi, idesc, 1u);
                        if (SPLIT) umma_bf16_pair(acc, da_hi, umma_desc_sw128(sw_lo + koff), idesc, 1u);
                    }
                    umma_commit_pair(empty_bar(s));
                }
                umma_commit_pair(tmem_full_bar(buf));
                BRL_MLP_STAMP(tile, 4);
            }
        }
    } else {  // ===== epilogue: warp w drains TMEM lane quadrant w % 4, column half (w - 2) / 4 =====
        const int q = warp & 3, half = (warp - 2) >> 2;
        constexpr int kHalfCols = kFusedBN / 2;
        const uint32_t lead_tmem_empty0 = tmem_empty_bar(0) & kPeerBitMask;
        uint32_t j = 0;
        for (int tile = pair; tile < n_tiles; tile += n_pairs, ++j) {
            const FusedTile f = fused_tile(tile, tpl);
            const int m0 = f.mb * (2 * kBM) + (int)rank * kBM;
            const uint32_t buf = j & 1u;
            mbar_wait(tmem_full_bar(buf), (j >> 1) & 1u);
            tc_fence_after();
            if (warp == 2 && lane == 0) { BRL_MLP_STAMP(tile, 5); }
            const int row = m0 + q * 32 + lane;
            const uint32_t t_row = tmem_acc + buf * kFusedBN + ((uint32_t)(q * 32) << 16);
            if (f.layer < 4) {
                const int n0 = f.nt * kFusedBN + half * kHalfCols;
                if (SPLIT) {
                    epilogue_hidden_tma(t_row + (uint32_t)(half * kHalfCols), kHalfCols, a.bias[f.layer] + n0, lane,
                                        store_base + (uint32_t)(warp - 2) * 4096u, &a.st_hi[f.layer], &a.st_lo[f.layer], n0,
                                        m0 + q * 32);
                } else {
                    __nv_bfloat16* oh = a.out_hi[f.layer];
                    epilogue_hidden_row(t_row + (uint32_t)(half * kHalfCols), kHalfCols, a.bias[f.layer] + n0, row < a.M,
                                        oh + (size_t)row * kHidden + n0, nullptr);
                }
            } else if (half == 0) {  // the 64-column head tile: one warp per quadrant
                epilogue_head_act_row(t_row, a.bias[4], row < a.M, row, a.logits, a.value, a.act);
            }
            tc_fence_before();
            __syncwarp();  // orders the other lanes' activation stores before lane 0's fence (sync, one fence, one atomic)
            if (lane == 0) {
                asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(lead_tmem_empty0 + 8u * buf) : "memory");
                if (f.layer < 4) {
                    if (SPLIT) {  // this warp's bulk tensor stores are performed (not just read) before the rows are published
                        bulk_wait0();
                        asm volatile("fence.proxy.async;" ::: "memory");
                    }
                    __threadfence();  // cumulative: the warp's rows are visible GPU-wide before the count below
                    atomicAdd(a.ready + (size_t)f.layer * 2 * a.nmb + 2 * f.mb + rank, 1u);
                }
#ifdef BRL_MLP_TRACE
                if (warp == 2 && leader && tile < 1024) { g_mlp_trace[tile][6] = mlp_now(); g_mlp_trace[tile][7] = (unsigned long long)pair; }
#endif
            }
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(kTmemCols) : "memory");
}


template <int BN, bool SPLIT_A, bool SPLIT_W, bool HEAD>
static int32_t launch_layer(cudaStream_t s, const void* a_hi, const void* a_lo, int k_in, const void* w_hi, const void* w_lo,
                            int n_valid_rows, const LayerArgs& args) {
    using Cfg = LayerCfg<BN, SPLIT_A, SPLIT_W>;
    CUtensorMap ta_hi, ta_lo, tw_hi, tw_lo;
    bool ok = make_map(&ta_hi, a_hi, (uint64_t)args.M, (uint64_t)k_in, (uint64_t)k_in, kBM) &&
              make_map(&tw_hi, w_hi, (uint64_t)n_valid_rows, (uint64_t)k_in, (uint64_t)k_in, BN);
    ta_lo = ta_hi;
    tw_lo = tw_hi;
    if (ok && SPLIT_A) ok = make_map(&ta_lo, a_lo, (uint64_t)args.M, (uint64_t)k_in, (uint64_t)k_in, kBM);
    if (ok && SPLIT_W) ok = make_map(&tw_lo, w_lo, (uint64_t)n_valid_rows, (uint64_t)k_in, (uint64_t)k_in, BN);
    if (!ok) return fail(BRL_E_LAUNCH, "brl_mlp_forward: cuTensorMapEncodeTiled failed");
    auto kern = k_mlp_layer<BN, SPLIT_A, SPLIT_W, HEAD>;
    static bool attr_set_dev[kMaxDevices] = {};  // per device; idempotent, a race only repeats the call
    bool& attr_set = attr_set_dev[device_slot()];
    if (!attr_set) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmemBytes) != cudaSuccess)
            return fail(BRL_E_LAUNCH, "brl_mlp_forward: cannot reserve %u bytes of shared memory", Cfg::kSmemBytes);
        attr_set = true;
    }
    LayerArgs la = args;
    la.n_tiles_n = (n_valid_rows + BN - 1) / BN;
    la.n_tiles = la.n_tiles_n * ((args.M + kBM - 1) / kBM);
    static int n_sm_dev[kMaxDevices] = {};
    int& n_sm = n_sm_dev[device_slot()];
    if (n_sm == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sm <= 0) n_sm = 148;
    }
    const unsigned grid = (unsigned)(la.n_tiles < n_sm ? la.n_tiles : n_sm);  // persistent: one CTA per SM
    kern<<<grid, kMlpThreads, Cfg::kSmemBytes, s>>>(ta_hi, ta_lo, tw_hi, tw_lo, la);
    return BRL_OK;
}

template <int BN, bool SPLIT_A, bool SPLIT_W>
static int32_t launch_layer_pair(cudaStream_t s, const void* a_hi, const void* a_lo, int k_in, const void* w_hi, const void* w_lo,
                                 int n_valid_rows, const LayerArgs& args) {
    using Cfg = PairCfg<BN, SPLIT_A, SPLIT_W>;
    CUtensorMap ta_hi, ta_lo, tw_hi, tw_lo;
    bool ok = make_map(&ta_hi, a_hi, (uint64_t)args.M, (uint64_t)k_in, (uint64_t)k_in, kBM) &&
              make_map(&tw_hi, w_hi, (uint64_t)n_valid_rows, (uint64_t)k_in, (uint64_t)k_in, BN / 2);
    ta_lo = ta_hi;
    tw_lo = tw_hi;
    if (ok && SPLIT_A) ok = make_map(&ta_lo, a_lo, (uint64_t)args.M, (uint64_t)k_in, (uint64_t)k_in, kBM);
    if (ok && SPLIT_W) ok = make_map(&tw_lo, w_lo, (uint64_t)n_valid_rows, (uint64_t)k_in, (uint64_t)k_in, BN / 2);
    if (!ok) return fail(BRL_E_LAUNCH, "brl_mlp_forward: cuTensorMapEncodeTiled failed");
    auto kern = k_mlp_layer_pair<BN, SPLIT_A, SPLIT_W>;
    static bool attr_set_dev[kMaxDevices] = {};
    bool& attr_set = attr_set_dev[device_slot()];
    if (!attr_set) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmemBytes) != cudaSuccess)
            return fail(BRL_E_LAUNCH, "brl_mlp_forward: cannot reserve %u bytes of shared memory", Cfg::kSmemBytes);
        attr_set = true;
    }
    LayerArgs la = args;
    la.n_tiles_n = n_valid_rows / BN;
    la.n_tiles = la.n_tiles_n * ((args.M + 2 * kBM - 1) / (2 * kBM));
    static int n_sm_dev[kMaxDevices] = {};
    int& n_sm = n_sm_dev[device_slot()];
    if (n_sm == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sm <= 0) n_sm = 148;
    }
    const int n_pairs = la.n_tiles < n_sm / 2 ? la.n_tiles : n_sm / 2;  // persistent: one CTA pair per TPC
    kern<<<(unsigned)(2 * n_pairs), kMlpThreads, Cfg::kSmemBytes, s>>>(ta_hi, ta_lo, tw_hi, tw_lo, la);
    return BRL_OK;
}

// counters of the fused forward live behind the four activation buffers of the scratch
static inline size_t fused_ready_bytes(int64_t n_envs) { return (size_t)(4 * 2 * ((n_envs + 2 * kBM - 1) / (2 * kBM))) * sizeof(uint32_t); }

template <bool SPLIT>
static int32_t launch_fused(cudaStream_t s, const void* obs, const unsigned char* blob, const MlpLayout& L, __nv_bfloat16* const buf_hi[2],
                            __nv_bfloat16* const buf_lo[2], uint32_t* ready, float* logits, float* value, int M, const ActArgs& act) {
    using Cfg = FusedCfg<SPLIT>;
    FusedArgs fa{};
    bool ok = true;
    for (int l = 0; l < 5 && ok; ++l) {
        const void* a_hi = l == 0 ? obs : (const void*)buf_hi[(l - 1) & 1];
        const void* a_lo = l == 0 ? obs : (const void*)buf_lo[(l - 1) & 1];
        const uint32_t w_box = l == 4 ? kHeadPad / 2 : kFusedBN / 2;
        ok = make_map(&fa.a_hi[l], a_hi, (uint64_t)M, (uint64_t)L.k_in[l], (uint64_t)L.k_in[l], kBM) &&
             make_map(&fa.w_hi[l], blob + L.w_hi[l], (uint64_t)L.n_out[l], (uint64_t)L.k_in[l], (uint64_t)L.k_in[l], w_box);
        fa.a_lo[l] = fa.a_hi[l];
        fa.w_lo[l] = fa.w_hi[l];
        if (ok && SPLIT && l > 0) ok = make_map(&fa.a_lo[l], a_lo, (uint64_t)M, (uint64_t)L.k_in[l], (uint64_t)L.k_in[l], kBM);
        if (ok && SPLIT) ok = make_map(&fa.w_lo[l], blob + L.w_lo[l], (uint64_t)L.n_out[l], (uint64_t)L.k_in[l], (uint64_t)L.k_in[l], w_box);
        fa.bias[l] = reinterpret_cast<const float*>(blob + L.bias[l]);
        if (l < 4) {
            fa.out_hi[l] = buf_hi[l & 1];
            fa.out_lo[l] = SPLIT ? buf_lo[l & 1] : nullptr;
            if (ok && SPLIT)
                ok = make_map_store(&fa.st_hi[l], buf_hi[l & 1], (uint64_t)M, (uint64_t)kHidden, (uint64_t)kHidden) &&
                     make_map_store(&fa.st_lo[l], buf_lo[l & 1], (uint64_t)M, (uint64_t)kHidden, (uint64_t)kHidden);
        }
    }
    if (!ok) return fail(BRL_E_LAUNCH, "brl_mlp_forward: cuTensorMapEncodeTiled failed");
    static_assert(sizeof(FusedArgs) <= 4096, "kernel parameters");
    fa.logits = logits;
    fa.value = value;
    fa.act = act;
    fa.ready = ready;
    fa.M = M;
    fa.nmb = (M + 2 * kBM - 1) / (2 * kBM);
    fa.tiles_per_layer = fa.nmb * kFusedTilesN;
    fa.n_tiles = 4 * fa.tiles_per_layer + fa.nmb;
    auto kern = k_mlp_fused<SPLIT>;
    static int max_pairs_dev[kMaxDevices] = {};  // co-resident CTA pairs: the dependency waits need every launched pair to be running
    int& max_pairs = max_pairs_dev[device_slot()];
    if (max_pairs == 0) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmemBytes) != cudaSuccess)
            return fail(BRL_E_LAUNCH, "brl_mlp_forward: cannot reserve %u bytes of shared memory", Cfg::kSmemBytes);
        int dev = 0, n_sm = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sm <= 0) n_sm = 148;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)(n_sm & ~1));
        cfg.blockDim = dim3(kFusedThreads);
        cfg.dynamicSmemBytes = Cfg::kSmemBytes;
        cudaLaunchAttribute at{};
        at.id = cudaLaunchAttributeClusterDimension;
        at.val.clusterDim.x = 2;
        at.val.clusterDim.y = 1;
        at.val.clusterDim.z = 1;
        cfg.attrs = &at;
        cfg.numAttrs = 1;
        int n_clusters = 0;
        if (cudaOccupancyMaxActiveClusters(&n_clusters, kern, &cfg) != cudaSuccess || n_clusters <= 0) {
            cudaGetLastError();
            n_clusters = n_sm / 2;
        }
        max_pairs = n_clusters < n_sm / 2 ? n_clusters : n_sm / 2;
    }
    if (cudaMemsetAsync(ready, 0, fused_ready_bytes(M), s) != cudaSuccess) return check_launch("brl_mlp_forward (memset)");
    const int n_pairs = fa.n_tiles < max_pairs ? fa.n_tiles : max_pairs;
    kern<<<(unsigned)(2 * n_pairs), kFusedThreads, Cfg::kSmemBytes, s>>>(fa);
    return BRL_OK;
}

__global__ void __launch_bounds__(256) k_gather_obs_rows(const uint4* __restrict__ src, const int32_t* __restrict__ index,
                                                         uint4* __restrict__ dst, int64_t n, int w) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * w) return;
    const int64_t r = i / w;
    const int c = (int)(i - r * w);
    dst[i] = src[(int64_t)index[r] * w + c];
}

}  // namespace brl

using namespace brl;

extern "C" {

int64_t brl_mlp_packed_bytes(void) { return (int64_t)mlp_layout().total; }
int64_t brl_mlp_scratch_bytes(int64_t n_envs) {  // 4 activation buffers bf16[n, 1024] + the fused forward's tile counters
    return n_envs * (int64_t)kHidden * 2 * 4 + (int64_t)((fused_ready_bytes(n_envs) + 255) & ~(size_t)255);
}

int32_t brl_mlp_pack(brl_stream_t stream, void** b, const void* opaque, size_t len) {
    int32_t rc;
    const BrlParams* p = get_params(opaque, len, &rc);
    if (!p) return rc;
    for (int k = 0; k < 13; ++k)
        if (b[k] == nullptr) return fail(BRL_E_BUFFER, "brl_mlp_pack: buffer %d is NULL", k);
    BRL_REQUIRE(b[12], "packed");
    const MlpLayout L = mlp_layout();
    unsigned char* blob = static_cast<unsigned char*>(b[12]);
    for (int l = 0; l < 5; ++l) {
        const int64_t total = (int64_t)L.n_out[l] * L.k_in[l];
        const bool head = l == 4;
        k_mlp_pack<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
            static_cast<const float*>(b[l]), static_cast<const float*>(b[6 + l]),
            head ? static_cast<const float*>(b[5]) : nullptr, head ? static_cast<const float*>(b[11]) : nullptr,
            reinterpret_cast<__nv_bfloat16*>(blob + L.w_hi[l]), reinterpret_cast<__nv_bfloat16*>(blob + L.w_lo[l]),
            reinterpret_cast<float*>(blob + L.bias[l]), L.k_in[l], L.n_out[l], head ? 38 : kHidden, head ? 1 : 0);
    }
    return check_launch("brl_mlp_pack");
}

int32_t brl_obs_to_bf16(brl_stream_t stream, void** b, const void* opaque, size_t len) {
    int32_t rc;
    const BrlParams* p = get_params(opaque, len, &rc);
    if (!p) return rc;
    if (p->n_envs == 0) return BRL_OK;  // an empty batch is a no-op; its buffers may be NULL (zero-size XLA / torch buffers)
    BRL_REQUIRE(b[0], "obs");
    BRL_REQUIRE(b[1], "obs_bf16");
    if (p->n_envs == 0) return BRL_OK;
    const int64_t n = p->n_envs * kObsDimM;
    const unsigned grid = (unsigned)((n / 8 + 255) / 256);
    if (p->flags & BRL_F_OBS_U8)
        k_obs_to_bf16<uint8_t><<<grid, 256, 0, (cudaStream_t)stream>>>(static_cast<const uint8_t*>(b[0]), static_cast<__nv_bfloat16*>(b[1]), n);
    else
        k_obs_to_bf16<float><<<grid, 256, 0, (cudaStream_t)stream>>>(static_cast<const float*>(b[0]), static_cast<__nv_bfloat16*>(b[1]), n);
    return check_launch("brl_obs_to_bf16");
}

// forward (+ optional fused categorical) shared by brl_mlp_forward and brl_policy_act
static int32_t mlp_forward_impl(const char* who, cudaStream_t s, const BrlParams* p, const void* obs, const unsigned char* blob,
                                __nv_bfloat16* scratch, float* logits, float* value, const ActArgs& act) {
    int32_t rc = BRL_OK;
    if (p->n_envs > (int64_t)1 << 30) return fail(BRL_E_OPAQUE, "%s: n_envs too large", who);
    if (encode_fn() == nullptr) return fail(BRL_E_LAUNCH, "%s: cuTensorMapEncodeTiled not available from the driver", who);
    const bool split = !(p->flags & BRL_F_MLP_BF16);
    const int M = (int)p->n_envs;
    const MlpLayout L = mlp_layout();
    const size_t act_elems = (size_t)M * kHidden;
    __nv_bfloat16* buf_hi[2] = {scratch, scratch + act_elems};
    __nv_bfloat16* buf_lo[2] = {scratch + 2 * act_elems, scratch + 3 * act_elems};
    // one persistent launch for the whole net once the batch fills the machine; flags bit 30 forces it,
    // bit 28 forces per-layer CTA-pair launches, bit 29 forbids both (1-CTA tiles)
    const bool fused = (p->flags & ((1 << 29) | (1 << 28))) ? false : ((p->flags & (1 << 30)) != 0 || M >= kPairMinM);
    if (fused) {
        uint32_t* ready = reinterpret_cast<uint32_t*>(scratch + 4 * act_elems);
        rc = split ? launch_fused<true>(s, obs, blob, L, buf_hi, buf_lo, ready, logits, value, M, act)
                   : launch_fused<false>(s, obs, blob, L, buf_hi, buf_lo, ready, logits, value, M, act);
        if (rc != BRL_OK) return rc;
        return check_launch(who);
    }
    // per-layer launches.  The head reads buffer 1, so buffer 0 is free to hold logits / value the caller did not ask for.
    float* lg = (logits && !act.row_index) ? logits : reinterpret_cast<float*>(buf_hi[0]);  // listed batch: compact here, scattered below
    float* vl = value ? value : reinterpret_cast<float*>(buf_lo[0]);
    const void* in_hi = obs;
    const void* in_lo = nullptr;
    for (int l = 0; l < 5; ++l) {
        LayerArgs a{};
        a.bias = reinterpret_cast<const float*>(blob + L.bias[l]);
        a.M = M;
        a.n_total = kHidden;
        a.k_blocks = (L.k_in[l] + kBK - 1) / kBK;
        const void* w_hi = blob + L.w_hi[l];
        const void* w_lo = blob + L.w_lo[l];
        if (l < 4) {
            a.out_hi = buf_hi[l & 1];
            a.out_lo = split ? buf_lo[l & 1] : nullptr;
            // 128 x 256 tiles move fewer operand bytes per MMA (the L2 -> SM fill rate, ~64 B/clk, is what bounds
            // these 1-CTA tiles): measured faster for the single-product mode and for large batches
            // (scripts/exp_mlp_ab.py); flags bit 26 / 27 force wide / narrow
            const bool wide = (p->flags & (1 << 27)) ? false : ((p->flags & (1 << 26)) != 0 || !split || M >= 32768);
            const bool pair = (p->flags & (1 << 28)) != 0;
            if (pair) {
                if (l == 0) rc = split ? launch_layer_pair<256, false, true>(s, in_hi, in_lo, L.k_in[l], w_hi, w_lo, kHidden, a)
                                       : launch_layer_pair<256, false, false>(s, in_hi, in_lo, L.k_in[l], w_hi, w_lo, kHidden, a);
                else rc = split ? launch_layer_pair<256, true, true>(s, in_hi, in_lo, L.k_in[l], w_hi, w_lo, kHidden, a)
                                : launch_layer_pair<256, false, false>(s, in_hi, in_lo, L.k_in[l], w_hi, w_lo, kHidden, a);
            } else if (wide) {
                if (l == 0) rc = split ? launch_layer<256, false, true, false>(s, in_hi, in_lo, L.k_in[l], w_hi, w_lo, kHidden, a)
                                       : launch_layer<256, false, false, false>(s, in_hi, in_lo, L.k_in[l], w_hi, w_lo, kHidden, a);
                else rc = split ? launch_layer<256, true, true, false>(s, in_hi, in_lo, L.k_in[l], w_hi, w_lo, kHidden, a)
                                : launch_layer<256, false, false, false>(s, in_hi, in_lo, L.k_in[l], w_hi, w_lo, kHidden, a);
            } else if (l == 0) rc = split ? launch_layer<128, false, true, false>(s, in_hi, in_lo, L.k_in[l], w_hi, w_lo, kHidden, a)
                                   : launch_layer<128, false, false, false>(s, in_hi, in_lo, L.k_in[l], w_hi, w_lo, kHidden, a);
            else rc = split ? launch_layer<128, true, true, false>(s, in_hi, in_lo, L.k_in[l], w_hi, w_lo, kHidden, a)
                            : launch_layer<128, false, false, false>(s, in_hi, in_lo, L.k_in[l], w_hi, w_lo, kHidden, a);
            in_hi = a.out_hi;
            in_lo = a.out_lo;
        } else {
            a.logits = lg;
            a.value = vl;
            rc = split ? launch_layer<kHeadPad, true, true, true>(s, in_hi, in_lo, L.k_in[l], w_hi, w_lo, kHeadPad, a)
                       : launch_layer<kHeadPad, false, false, true>(s, in_hi, in_lo, L.k_in[l], w_hi, w_lo, kHeadPad, a);
        }
        if (rc != BRL_OK) return rc;
    }
    rc = check_launch(who);
    if (rc != BRL_OK || act.action == nullptr) return rc;
    if (act.row_index && value != nullptr) return fail(BRL_E_OPAQUE, "%s: a row-indexed batch below the fused-launch size has no value output", who);
    return launch_categorical(s, lg, act.mask, act.action, act.log_prob, M, act.sample, act.seed, act.env_offset, act.step, act.row_index,
                              act.row_index ? logits : nullptr);
}

int32_t brl_mlp_forward(brl_stream_t stream, void** b, const void* opaque, size_t len) {
    int32_t rc;
    const BrlParams* p = get_params(opaque, len, &rc);
    if (!p) return rc;
    if (p->n_envs == 0) return BRL_OK;  // an empty batch is a no-op; its buffers may be NULL (zero-size XLA / torch buffers)
    BRL_REQUIRE(b[0], "obs_bf16");
    BRL_REQUIRE(b[1], "packed");
    BRL_REQUIRE(b[2], "scratch");
    BRL_REQUIRE(b[3], "logits");
    BRL_REQUIRE(b[4], "value");
    if (p->n_envs == 0) return BRL_OK;
    return mlp_forward_impl("brl_mlp_forward", (cudaStream_t)stream, p, b[0], static_cast<const unsigned char*>(b[1]),
                            static_cast<__nv_bfloat16*>(b[2]), static_cast<float*>(b[3]), static_cast<float*>(b[4]), ActArgs{});
}

#ifdef BRL_MLP_TRACE
int32_t brl_debug_mlp_trace(unsigned long long* out_1024x8) {
    return cudaMemcpyFromSymbol(out_1024x8, g_mlp_trace, sizeof(g_mlp_trace)) == cudaSuccess ? 0 : -1;
}
int32_t brl_debug_mlp_epi(unsigned long long* out4, int reset) {
    if (cudaMemcpyFromSymbol(out4, g_mlp_epi, sizeof(g_mlp_epi)) != cudaSuccess) return -1;
    if (reset) { unsigned long long z[4] = {0, 0, 0, 0}; cudaMemcpyToSymbol(g_mlp_epi, z, sizeof(z)); }
    return 0;
}
#endif

int64_t brl_mlp_rows_scratch_bytes(int64_t n_rows) {  // brl_mlp_scratch_bytes + the gathered bf16 observation rows
    return brl_mlp_scratch_bytes(n_rows) + n_rows * (int64_t)kObsDimM * 2;
}

// `team1_forward_pass.apply` + masked Categorical on the LISTED envs only (src/evaluation.py:124-151 without running
// either net on envs it does not decide): gather the rows' observations (one 960-byte bf16 row per thread group), run
// the forward on the compact batch, and let the head epilogue read mask[env] and write action[env] / log_prob[env].
int32_t brl_policy_act_rows(brl_stream_t stream, void** b, const void* opaque, size_t len) {
    int32_t rc;
    const BrlParams* p = get_params(opaque, len, &rc);
    if (!p) return rc;
    if (p->n_envs == 0) return BRL_OK;  // an empty batch is a no-op; its buffers may be NULL (zero-size XLA / torch buffers)
    BRL_REQUIRE(b[0], "obs_bf16");
    BRL_REQUIRE(b[1], "packed");
    BRL_REQUIRE(b[2], "scratch");
    BRL_REQUIRE(b[4], "action");
    BRL_REQUIRE(b[5], "row_index");
    if ((reinterpret_cast<uintptr_t>(b[0]) | reinterpret_cast<uintptr_t>(b[2])) & 15u)
        return fail(BRL_E_BUFFER, "brl_policy_act_rows: obs / scratch must be 16-byte aligned");
    if (b[3] != nullptr && (reinterpret_cast<uintptr_t>(b[3]) & 1u)) return fail(BRL_E_BUFFER, "brl_policy_act_rows: mask is misaligned");
    if (b[7] != nullptr && (reinterpret_cast<uintptr_t>(b[7]) & 7u)) return fail(BRL_E_BUFFER, "brl_policy_act_rows: logits is misaligned");
    if (p->n_envs == 0) return BRL_OK;
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t n = p->n_envs;
    unsigned char* scratch = static_cast<unsigned char*>(b[2]);
    uint4* gathered = reinterpret_cast<uint4*>(scratch + brl_mlp_scratch_bytes(n));
    const int w = kObsDimM * 2 / 16;  // uint4 per bf16 observation row
    k_gather_obs_rows<<<(unsigned)((n * w + 255) / 256), 256, 0, s>>>(static_cast<const uint4*>(b[0]), static_cast<const int32_t*>(b[5]),
                                                                     gathered, n, w);
    ActArgs act{};
    act.mask = static_cast<const uint8_t*>(b[3]);
    act.action = static_cast<int32_t*>(b[4]);
    act.log_prob = static_cast<float*>(b[6]);
    act.seed = p->seed;
    act.env_offset = p->env_offset;
    act.step = p->step;
    act.sample = (p->flags & BRL_F_SAMPLE) ? 1 : 0;
    act.row_index = static_cast<const int32_t*>(b[5]);
    return mlp_forward_impl("brl_policy_act_rows", s, p, gathered, static_cast<const unsigned char*>(b[1]),
                            reinterpret_cast<__nv_bfloat16*>(scratch), static_cast<float*>(b[7]), nullptr, act);
}

int32_t brl_policy_act(brl_stream_t stream, void** b, const void* opaque, size_t len) {
    int32_t rc;
    const BrlParams* p = get_params(opaque, len, &rc);
    if (!p) return rc;
    if (p->n_envs == 0) return BRL_OK;  // an empty batch is a no-op; its buffers may be NULL (zero-size XLA / torch buffers)
    BRL_REQUIRE(b[0], "obs_bf16");
    BRL_REQUIRE(b[1], "packed");
    BRL_REQUIRE(b[2], "scratch");
    BRL_REQUIRE(b[4], "action");
    for (int k : {3, 5, 6, 7})
        if (b[k] != nullptr && (reinterpret_cast<uintptr_t>(b[k]) & (k == 3 ? 1u : 7u)) != 0)
            return fail(BRL_E_BUFFER, "brl_policy_act: buffer %d is misaligned", k);
    if (p->n_envs == 0) return BRL_OK;
    ActArgs act{};
    act.mask = static_cast<const uint8_t*>(b[3]);
    act.action = static_cast<int32_t*>(b[4]);
    act.log_prob = static_cast<float*>(b[5]);
    act.seed = p->seed;
    act.env_offset = p->env_offset;
    act.step = p->step;
    act.sample = (p->flags & BRL_F_SAMPLE) ? 1 : 0;
    if (p->flags & BRL_F_SEED_SALT) {
        act.seed_salt = static_cast<const uint64_t*>(b[8]);
        if (act.seed_salt == nullptr || (reinterpret_cast<uintptr_t>(b[8]) & 7u)) return fail(BRL_E_BUFFER, "brl_policy_act: BRL_F_SEED_SALT needs an 8-byte aligned buffer [8]");
        if (p->n_envs < kPairMinM) return fail(BRL_E_OPAQUE, "brl_policy_act: BRL_F_SEED_SALT is supported by the fused launch only (n_envs >= %d)", kPairMinM);
    }
    return mlp_forward_impl("brl_policy_act", (cudaStream_t)stream, p, b[0], static_cast<const unsigned char*>(b[1]),
                            static_cast<__nv_bfloat16*>(b[2]), static_cast<float*>(b[7]), static_cast<float*>(b[6]), act);
}

}  // extern "C"
