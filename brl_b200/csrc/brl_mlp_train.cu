// brl_mlp_train.cu -- the PPO update's trip through the policy/value net on the 5th-gen tensor cores
// (SURVEY 8f-1): `jax.value_and_grad(_loss_fn)(params, ...)` of src/update.py:91-167 for one minibatch,
// i.e. forward of the "DeepMind" MLP (src/models.py:23-33) with the activations kept, the PPO loss head
// (brl_ppo_loss, csrc/brl_ppo.cu) and the backward GEMMs, producing d total_loss / d params as ONE flat
// fp32 buffer in the optimizer's order.  No library GEMM and no autograd tape: every product runs on
// tcgen05 with the same three-term bf16 split as the rollout forward (brl_mlp.cu), so gradients are
// fp32-class (the reference differentiates in fp32).
//
// All three GEMM shapes are  C[M, N] = A[M, K] . Bt[N, K]^T  with both operands K-major, so one kernel
// (k_gemm_tc: TMA producer warp, single-lane tcgen05.mma issuer, four epilogue warps, persistent tiles,
// double-buffered TMEM accumulator) serves all of them; they differ in the epilogue:
//   forward  h_l  = relu(h_{l-1} . W_l + b_l)      A = h_{l-1} [B, in]      Bt = W_l^T [out, in]   (packed "Wt")
//   dgrad    dz_{l-1} = (dz_l . W_l^T) * (h_{l-1} > 0)   A = dz_l [B, out]  Bt = W_l [in, out]     (packed "Wn")
//   wgrad    dW_l = h_{l-1}^T . dz_l               A = h_{l-1}^T [in, B]    Bt = dz_l^T [out, B]
// The wgrad contracts over the batch, so it needs batch-major ("transposed") copies of the activations
// and of the dz's: the forward / dgrad epilogues write those next to the row-major copies (a warp's 32
// rows of one column are 64 contiguous bytes), which costs no extra launch and no extra read.
// Bias gradients are the row sums of the transposed dz's (one warp per output feature).
#include "common.h"
#include "mlp_device.cuh"

namespace brl {

constexpr int kNumLayers = 5;  // four hidden layers + the fused policy/value head (38 + 1 columns, padded to 64)

// ---- flat fp32 parameter / gradient buffer: LAYERS order, w then b (brl_b200/optim.py flatten_params) ----
struct FlatLayout {
    size_t w[6], b[6], total;
};
__host__ __device__ inline FlatLayout flat_layout() {
    FlatLayout F{};
    size_t off = 0;
    for (int l = 0; l < 6; ++l) {
        const size_t in = l == 0 ? kObsDimM : kHidden, out = l < 4 ? kHidden : (l == 4 ? 38 : 1);
        F.w[l] = off; off += in * out;
        F.b[l] = off; off += out;
    }
    F.total = off;
    return F;
}

// ---- training blob: the forward blob (MlpLayout, usable by brl_mlp_forward as is) + W in its own
// orientation [in, out_pad] as bf16 hi / lo for the dgrad GEMMs of layers 1..4 ------------------------
struct TrainBlob {
    size_t wn_hi[kNumLayers], wn_lo[kNumLayers], total;
};
__host__ __device__ inline TrainBlob train_blob() {
    const MlpLayout L = mlp_layout();
    TrainBlob T{};
    size_t off = L.total;
    for (int l = 1; l < kNumLayers; ++l) {
        const size_t bytes = (size_t)L.k_in[l] * L.n_out[l] * 2;
        T.wn_hi[l] = off; off += bytes;
        T.wn_lo[l] = off; off += bytes;
        off = (off + 255) & ~(size_t)255;
    }
    T.total = off;
    return T;
}

struct PackArgs {  // layouts travel as kernel parameters (constant bank): indexing them by layer costs no local memory
    const float* flat;
    unsigned char* blob;
    MlpLayout L;
    TrainBlob T;
    FlatLayout F;
};

// 32 x 32 tile of layer blockIdx.z: fp32 W[in, out] -> Wn hi/lo [in, out_pad] (same orientation) and, through a
// shared-memory transpose, Wt hi/lo [out_pad, in]; the head layer concatenates the policy (38) and value (1) columns.
__global__ void __launch_bounds__(256) k_pack_train(const __grid_constant__ PackArgs a) {
    const int l = blockIdx.z;
    const MlpLayout& L = a.L;
    const TrainBlob& T = a.T;
    const FlatLayout& F = a.F;
    const int k_in = L.k_in[l], n_pad = L.n_out[l];
    const int k0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
    if (k0 >= k_in || n0 >= n_pad) return;
    const bool head = l == 4;
    const int n_src = head ? 38 : kHidden;
    const float* w = a.flat + F.w[l];
    const float* w2 = head ? a.flat + F.w[5] : nullptr;
    __shared__ float t[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    __nv_bfloat16* wn_hi = l > 0 ? reinterpret_cast<__nv_bfloat16*>(a.blob + T.wn_hi[l]) : nullptr;
    __nv_bfloat16* wn_lo = l > 0 ? reinterpret_cast<__nv_bfloat16*>(a.blob + T.wn_lo[l]) : nullptr;
    for (int j = ty; j < 32; j += 8) {
        const int k = k0 + j, n = n0 + tx;
        float v = 0.0f;
        if (n < n_src) v = w[(size_t)k * n_src + n];
        else if (head && n == 38) v = w2[k];
        t[j][tx] = v;
        if (wn_hi) {
            const __nv_bfloat16 h = __float2bfloat16_rn(v);
            wn_hi[(size_t)k * n_pad + n] = h;
            wn_lo[(size_t)k * n_pad + n] = __float2bfloat16_rn(v - __bfloat162float(h));
        }
    }
    __syncthreads();
    __nv_bfloat16* wt_hi = reinterpret_cast<__nv_bfloat16*>(a.blob + L.w_hi[l]);
    __nv_bfloat16* wt_lo = reinterpret_cast<__nv_bfloat16*>(a.blob + L.w_lo[l]);
    for (int j = ty; j < 32; j += 8) {
        const int n = n0 + j, k = k0 + tx;
        const float v = t[tx][j];
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        wt_hi[(size_t)n * k_in + k] = h;
        wt_lo[(size_t)n * k_in + k] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
    if (blockIdx.y == 0 && ty == 0) {
        const int n = n0 + tx;
        float bv = 0.0f;
        if (n < n_src) bv = a.flat[F.b[l] + n];
        else if (head && n == 38) bv = a.flat[F.b[5]];
        reinterpret_cast<float*>(a.blob + L.bias[l])[n] = bv;
    }
}

// ---- per-minibatch scratch -------------------------------------------------------------------------------
struct TrainScratch {
    size_t obs_r, obs_t;                          // bf16 [B, 480], [480, ldt]
    size_t h_r_hi[4], h_r_lo[4], h_t_hi[4], h_t_lo[4];   // activations of layers 1..4: [B, 1024], [1024, ldt]
    size_t dz_r_hi[2], dz_r_lo[2];                // d loss / d pre-activation, row-major ping-pong [B, 1024]
    size_t dz_t_hi[4], dz_t_lo[4];                // ... batch-major [1024, ldt], kept per layer for the bias sums
    size_t dz5_r_hi, dz5_r_lo, dz5_t_hi, dz5_t_lo;  // head: [B, 64], [64, ldt]
    size_t logits, value, dlogits, dvalue;        // f32 [B, 38], [B]
    size_t total;
    int ldt;                                      // row pitch (elements) of the batch-major arrays
};
__host__ inline TrainScratch train_scratch(int64_t B) {
    TrainScratch S{};
    const size_t ldt = (size_t)((B + 63) / 64 * 64);
    S.ldt = (int)ldt;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) & ~(size_t)255; return o; };
    S.obs_r = take((size_t)B * kObsDimM * 2);
    S.obs_t = take((size_t)kObsDimM * ldt * 2);
    for (int l = 0; l < 4; ++l) {
        S.h_r_hi[l] = take((size_t)B * kHidden * 2);
        S.h_r_lo[l] = take((size_t)B * kHidden * 2);
        S.h_t_hi[l] = take((size_t)kHidden * ldt * 2);
        S.h_t_lo[l] = take((size_t)kHidden * ldt * 2);
        S.dz_t_hi[l] = take((size_t)kHidden * ldt * 2);
        S.dz_t_lo[l] = take((size_t)kHidden * ldt * 2);
    }
    for (int k = 0; k < 2; ++k) {
        S.dz_r_hi[k] = take((size_t)B * kHidden * 2);
        S.dz_r_lo[k] = take((size_t)B * kHidden * 2);
    }
    S.dz5_r_hi = take((size_t)B * kHeadPad * 2);
    S.dz5_r_lo = take((size_t)B * kHeadPad * 2);
    S.dz5_t_hi = take((size_t)kHeadPad * ldt * 2);
    S.dz5_t_lo = take((size_t)kHeadPad * ldt * 2);
    S.logits = take((size_t)B * 38 * 4);
    S.value = take((size_t)B * 4);
    S.dlogits = take((size_t)B * 38 * 4);
    S.dvalue = take((size_t)B * 4);
    S.total = off;
    return S;
}

// ---- minibatch gather of the observation: obs[index[b]] (f32 / u8 0-1, or bf16) -> bf16 [B, 480] and [480, ldt] ----
template <class T>
__device__ __forceinline__ uint16_t obs_bits(T v) { return (float)v != 0.0f ? (uint16_t)0x3F80u : (uint16_t)0u; }
template <>
__device__ __forceinline__ uint16_t obs_bits<__nv_bfloat16>(__nv_bfloat16 v) { return *reinterpret_cast<uint16_t*>(&v); }

constexpr int kGatherCols = 120;  // observation columns per block: grid = (B / 32) x 4 blocks
template <class T>
__global__ void __launch_bounds__(256) k_gather_obs(const T* __restrict__ obs, const int32_t* __restrict__ index, int64_t B, int ldt,
                                                    uint16_t* __restrict__ out_r, uint16_t* __restrict__ out_t) {
    __shared__ uint16_t tile[32][kGatherCols + 2];  // row stride 61 words: column reads are conflict-free
    const int64_t r0 = (int64_t)blockIdx.x * 32;
    const int c0 = blockIdx.y * kGatherCols;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int r = warp; r < 32; r += 8) {
        const int64_t b = r0 + r;
        if (b >= B) break;
        const int64_t src = index ? (int64_t)index[b] : b;
        for (int c = lane; c < kGatherCols; c += 32) {
            const uint16_t v = obs_bits<T>(obs[src * kObsDimM + c0 + c]);
            tile[r][c] = v;
            out_r[b * kObsDimM + c0 + c] = v;
        }
    }
    __syncthreads();
    const int64_t b = r0 + lane;
    if (b < B)
        for (int c = warp; c < kGatherCols; c += 8) out_t[(size_t)(c0 + c) * ldt + b] = tile[lane][c];
}

// d loss / d (logits, value) f32 -> dz5 hi / lo, row-major [B, 64] and batch-major [64, ldt]; columns 39..63 are zero
__global__ void __launch_bounds__(256) k_head_grad_pack(const float* __restrict__ dlogits, const float* __restrict__ dvalue, int64_t B, int ldt,
                                                        __nv_bfloat16* __restrict__ r_hi, __nv_bfloat16* __restrict__ r_lo,
                                                        __nv_bfloat16* __restrict__ t_hi, __nv_bfloat16* __restrict__ t_lo) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * kHeadPad) return;
    const int64_t row = idx / kHeadPad;
    const int c = (int)(idx % kHeadPad);
    const float v = c < 38 ? dlogits[row * 38 + c] : (c == 38 ? dvalue[row] : 0.0f);
    const __nv_bfloat16 h = __float2bfloat16_rn(v), lo = __float2bfloat16_rn(v - __bfloat162float(h));
    r_hi[idx] = h;
    r_lo[idx] = lo;
    t_hi[(size_t)c * ldt + row] = h;
    t_lo[(size_t)c * ldt + row] = lo;
}

// bias gradients: db[c] = sum_b dz[b, c]; one warp per output feature.  Hidden layers: row sums of the batch-major dz
// (hi + lo); head: column sums of the fp32 d loss / d (logits, value) themselves.
struct BiasArgs {
    const __nv_bfloat16* t_hi[4];
    const __nv_bfloat16* t_lo[4];
    const float* dlogits;  // [B, 38]
    const float* dvalue;   // [B]
    float* grads;
    int64_t B;
    int ldt;
    FlatLayout F;
};
__global__ void __launch_bounds__(256) k_bias_grad(const __grid_constant__ BiasArgs a) {
    const int col = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (col >= 4 * kHidden + kHeadValid) return;
    const FlatLayout& F = a.F;
    const int l = col < 4 * kHidden ? col / kHidden : 4, c = col < 4 * kHidden ? col % kHidden : col - 4 * kHidden;
    float s = 0.0f;
    if (l < 4) {
        const __nv_bfloat162* hi = reinterpret_cast<const __nv_bfloat162*>(a.t_hi[l] + (size_t)c * a.ldt);
        const __nv_bfloat162* lo = reinterpret_cast<const __nv_bfloat162*>(a.t_lo[l] + (size_t)c * a.ldt);
        for (int64_t i = lane; 2 * i < a.B; i += 32) {
            const float2 h = __bfloat1622float2(hi[i]), r = __bfloat1622float2(lo[i]);
            s += h.x + r.x;
            if (2 * i + 1 < a.B) s += h.y + r.y;
        }
    } else {
        for (int64_t i = lane; i < a.B; i += 32) s += c < 38 ? a.dlogits[i * 38 + c] : a.dvalue[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
        float* out = l < 4 ? a.grads + F.b[l] + c : (c < 38 ? a.grads + F.b[4] + c : a.grads + F.b[5]);
        *out = s;
    }
}

// ---- the GEMM ---------------------------------------------------------------------------------------------
enum { kEpiFwd = 0, kEpiHead = 1, kEpiDgrad = 2, kEpiWgrad = 3 };

struct GemmArgs {
    int M, k_blocks, n_tiles_n, n_tiles;
    // kEpiFwd / kEpiDgrad: bf16 hi / lo outputs, row-major [M, ld_out] and batch-major [N, ld_t]
    const float* bias;                 // kEpiFwd, kEpiHead
    __nv_bfloat16 *out_hi, *out_lo;
    __nv_bfloat16 *out_t_hi, *out_t_lo;
    int ld_out, ld_t;
    const __nv_bfloat16* relu_src;     // kEpiDgrad: the forward activation (hi part) this gradient flows through, [M, ld_out]
    // kEpiWgrad: fp32 output [M, ld_c], columns < n_c; column n_c (the value head of the fused head tile) -> c2[row]
    float* c;
    float* c2;
    int ld_c, n_c;
    // kEpiHead
    float *logits, *value;
};

template <int BN, bool SPLIT_A, bool SPLIT_W>
struct GemmCfg {
    static constexpr uint32_t kABytes = kBM * kBK * 2, kWBytes = BN * kBK * 2;
    static constexpr uint32_t kStageBytes = kABytes * (SPLIT_A ? 2 : 1) + kWBytes * (SPLIT_W ? 2 : 1);
    static constexpr int kStages = (kSmemBudget / kStageBytes) > 8 ? 8 : (int)(kSmemBudget / kStageBytes);
    static constexpr uint32_t kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
};

// forward / dgrad rows: x -> bf16 hi / lo, 128-bit row-major stores + 16-bit batch-major stores (a warp's 32 rows of one
// column are contiguous)
template <int EPI>
__device__ __forceinline__ void epilogue_act_row(uint32_t t_row, int n_cols, int n0, int row, bool row_ok, const GemmArgs& a) {
#pragma unroll 1
    for (int c0 = 0; c0 < n_cols; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(t_row + (uint32_t)c0, r);
        if (!row_ok) continue;
        const size_t ro = (size_t)row * a.ld_out + n0 + c0;
        uint32_t keep[16];
        if (EPI == kEpiDgrad) {
            const uint4* ps = reinterpret_cast<const uint4*>(a.relu_src + ro);
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const uint4 m = ps[v];
                keep[4 * v] = m.x; keep[4 * v + 1] = m.y; keep[4 * v + 2] = m.z; keep[4 * v + 3] = m.w;
            }
        }
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) {
            float x0 = __uint_as_float(r[2 * jj]), x1 = __uint_as_float(r[2 * jj + 1]);
            if (EPI == kEpiFwd) {
                const float2 bb = __ldg(reinterpret_cast<const float2*>(a.bias + n0 + c0) + jj);  // warp-uniform
                x0 = fmaxf(x0 + bb.x, 0.0f);
                x1 = fmaxf(x1 + bb.y, 0.0f);
            } else {  // relu'(h) = [h > 0]; h >= 0 always, so "non-zero" is "positive"
                if ((keep[jj] & 0x00007FFFu) == 0u) x0 = 0.0f;
                if ((keep[jj] & 0x7FFF0000u) == 0u) x1 = 0.0f;
            }
            __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
            hi[jj] = *reinterpret_cast<uint32_t*>(&h);
            lo[jj] = pack_bf16x2(x0 - __low2float(h), x1 - __high2float(h));
        }
        uint4* ph = reinterpret_cast<uint4*>(a.out_hi + ro);
        uint4* pl = reinterpret_cast<uint4*>(a.out_lo + ro);
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            ph[v] = make_uint4(hi[4 * v], hi[4 * v + 1], hi[4 * v + 2], hi[4 * v + 3]);
            pl[v] = make_uint4(lo[4 * v], lo[4 * v + 1], lo[4 * v + 2], lo[4 * v + 3]);
        }
        uint16_t* th = reinterpret_cast<uint16_t*>(a.out_t_hi) + (size_t)(n0 + c0) * a.ld_t + row;
        uint16_t* tl = reinterpret_cast<uint16_t*>(a.out_t_lo) + (size_t)(n0 + c0) * a.ld_t + row;
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) {
            th[(size_t)(2 * jj) * a.ld_t] = (uint16_t)(hi[jj] & 0xFFFFu);
            th[(size_t)(2 * jj + 1) * a.ld_t] = (uint16_t)(hi[jj] >> 16);
            tl[(size_t)(2 * jj) * a.ld_t] = (uint16_t)(lo[jj] & 0xFFFFu);
            tl[(size_t)(2 * jj + 1) * a.ld_t] = (uint16_t)(lo[jj] >> 16);
        }
    }
}

// wgrad rows: the accumulator IS the gradient block; thread = one input feature (row of W)
template <int BN>
__device__ __forceinline__ void epilogue_wgrad_row(uint32_t t_row, int n0, int row, bool row_ok, const GemmArgs& a) {
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(t_row + (uint32_t)c0, r);
        if (!row_ok) continue;
        if (a.c2 != nullptr) {  // fused head tile: 38 policy columns (row pitch 38 floats: 8-byte aligned) + the value column
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                const int col = n0 + c0 + j;
                if (col + 1 < a.n_c)
                    *reinterpret_cast<float2*>(a.c + (size_t)row * a.ld_c + col) = make_float2(__uint_as_float(r[j]), __uint_as_float(r[j + 1]));
                else if (col == a.n_c)
                    a.c2[row] = __uint_as_float(r[j]);
            }
        } else {
            float4* pc = reinterpret_cast<float4*>(a.c + (size_t)row * a.ld_c + n0 + c0);
#pragma unroll
            for (int v = 0; v < 8; ++v)
                pc[v] = make_float4(__uint_as_float(r[4 * v]), __uint_as_float(r[4 * v + 1]), __uint_as_float(r[4 * v + 2]),
                                    __uint_as_float(r[4 * v + 3]));
        }
    }
}

template <int BN, bool SPLIT_A, bool SPLIT_W, int EPI>
__global__ void __launch_bounds__(kMlpThreads, 1)
k_gemm_tc(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
          const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo, const GemmArgs a) {
    // Same pipeline as k_mlp_layer (brl_mlp.cu): persistent CTAs walk tiles blockIdx.x, + gridDim.x, ... (n-tile fastest);
    // the shared-memory stage ring runs across tile boundaries; two TMEM accumulators overlap epilogue and main loop.
    using Cfg = GemmCfg<BN, SPLIT_A, SPLIT_W>;
    constexpr int S = Cfg::kStages;
    constexpr uint32_t kTmemCols = 2 * BN;
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t base = (smem_addr(smem_dyn) + 1023u) & ~1023u;
    const uint32_t bar_base = base + S * Cfg::kStageBytes;  // full[S], empty[S], tmem_full[2], tmem_empty[2]
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (S + s); };
    auto tmem_full_bar = [&](int b) { return bar_base + 8u * (2 * S + b); };
    auto tmem_empty_bar = [&](int b) { return bar_base + 8u * (2 * S + 2 + b); };
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles_n = a.n_tiles_n, n_tiles = a.n_tiles;

    // Programmatic dependent launch: the GEMMs of one update form a chain of short kernels, so the next one may be
    // scheduled now and run its prologue (barrier init, tensor-memory allocation, descriptor prefetch) on free SMs /
    // behind this one's tail; it touches no global data before griddepcontrol.wait below.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
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
    asm volatile("griddepcontrol.wait;" ::: "memory");  // everything the preceding kernels wrote is visible from here on

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
                mbar_wait(tmem_empty_bar(buf), ((j >> 1) & 1u) ^ 1u);
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
                        umma_bf16(acc, da_hi, dw_hi, idesc, (kb | k) != 0);
                        if (SPLIT_A) umma_bf16(acc, umma_desc_sw128(sa_lo + koff), dw_hi, idesc, 1u);
                        if (SPLIT_W) umma_bf16(acc, da_hi, umma_desc_sw128(sw_lo + koff), idesc, 1u);
                    }
                    umma_commit(empty_bar(s));
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
            if (EPI == kEpiFwd || EPI == kEpiDgrad) epilogue_act_row<EPI>(t_row, BN, n0, row, row < a.M, a);
            else if (EPI == kEpiWgrad) epilogue_wgrad_row<BN>(t_row, n0, row, row < a.M, a);
            else epilogue_head_row(t_row, a.bias + n0, row < a.M, a.logits + (size_t)row * 38, a.value + row);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tmem_empty_bar(buf)) : "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_acc, kTmemCols);
}

// A [m_rows, k_cols] (pitch lda), Bt [n_rows, k_cols] (pitch ldb), both bf16 K-major; hi / lo pairs
template <int BN, bool SPLIT_A, bool SPLIT_W, int EPI>
static int32_t launch_gemm(cudaStream_t s, const void* a_hi, const void* a_lo, int m_rows, int lda, const void* w_hi, const void* w_lo,
                           int n_rows, int ldb, int k_cols, GemmArgs args) {
    using Cfg = GemmCfg<BN, SPLIT_A, SPLIT_W>;
    CUtensorMap ta_hi, ta_lo, tw_hi, tw_lo;
    bool ok = make_map(&ta_hi, a_hi, (uint64_t)m_rows, (uint64_t)k_cols, (uint64_t)lda, kBM) &&
              make_map(&tw_hi, w_hi, (uint64_t)n_rows, (uint64_t)k_cols, (uint64_t)ldb, BN);
    ta_lo = ta_hi;
    tw_lo = tw_hi;
    if (ok && SPLIT_A) ok = make_map(&ta_lo, a_lo, (uint64_t)m_rows, (uint64_t)k_cols, (uint64_t)lda, kBM);
    if (ok && SPLIT_W) ok = make_map(&tw_lo, w_lo, (uint64_t)n_rows, (uint64_t)k_cols, (uint64_t)ldb, BN);
    if (!ok) return fail(BRL_E_LAUNCH, "brl_ppo_grad: cuTensorMapEncodeTiled failed");
    auto kern = k_gemm_tc<BN, SPLIT_A, SPLIT_W, EPI>;
    static bool attr_set = false;  // idempotent; a race only repeats the call
    if (!attr_set) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmemBytes) != cudaSuccess)
            return fail(BRL_E_LAUNCH, "brl_ppo_grad: cannot reserve %u bytes of shared memory", Cfg::kSmemBytes);
        attr_set = true;
    }
    args.M = m_rows;
    args.k_blocks = (k_cols + kBK - 1) / kBK;
    args.n_tiles_n = (n_rows + BN - 1) / BN;
    args.n_tiles = args.n_tiles_n * ((m_rows + kBM - 1) / kBM);
    static int n_sm = 0;
    if (n_sm == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sm <= 0) n_sm = 148;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(args.n_tiles < n_sm ? args.n_tiles : n_sm));
    cfg.blockDim = dim3(kMlpThreads);
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cfg.stream = s;
    cudaLaunchAttribute at{};
    at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &at;
    cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, kern, ta_hi, ta_lo, tw_hi, tw_lo, args) != cudaSuccess) return check_launch("brl_ppo_grad (GEMM launch)");
    return BRL_OK;
}

// hidden-layer GEMMs with the tile width picked at run time (narrow tiles fill more of the 148 SMs at minibatch sizes)
template <bool SPLIT_A, int EPI>
static int32_t launch_hidden(bool narrow, cudaStream_t s, const void* a_hi, const void* a_lo, int m_rows, int lda, const void* w_hi,
                             const void* w_lo, int n_rows, int ldb, int k_cols, const GemmArgs& args) {
    return narrow ? launch_gemm<64, SPLIT_A, true, EPI>(s, a_hi, a_lo, m_rows, lda, w_hi, w_lo, n_rows, ldb, k_cols, args)
                  : launch_gemm<128, SPLIT_A, true, EPI>(s, a_hi, a_lo, m_rows, lda, w_hi, w_lo, n_rows, ldb, k_cols, args);
}

}  // namespace brl

using namespace brl;

extern "C" {

int64_t brl_mlp_num_params(void) { return (int64_t)flat_layout().total; }
int64_t brl_mlp_train_blob_bytes(void) { return (int64_t)train_blob().total; }
int64_t brl_mlp_train_scratch_bytes(int64_t batch) { return batch > 0 ? (int64_t)train_scratch(batch).total : 0; }

int32_t brl_mlp_pack_train(brl_stream_t stream, void** b, const void* opaque, size_t len) {
    int32_t rc;
    const BrlParams* p = get_params(opaque, len, &rc);
    if (!p) return rc;
    BRL_REQUIRE(b[0], "params");
    BRL_REQUIRE(b[1], "blob");
    PackArgs a{static_cast<const float*>(b[0]), static_cast<unsigned char*>(b[1]), mlp_layout(), train_blob(), flat_layout()};
    k_pack_train<<<dim3(kHidden / 32, kHidden / 32, kNumLayers), 256, 0, (cudaStream_t)stream>>>(a);
    return check_launch("brl_mlp_pack_train");
}

int32_t brl_ppo_grad(brl_stream_t stream, void** b, const void* opaque, size_t len) {
    if (opaque == nullptr || len != sizeof(BrlPpoParams))
        return fail(BRL_E_OPAQUE, "brl_ppo_grad: opaque must be one BrlPpoParams (%zu bytes), got %zu", sizeof(BrlPpoParams), len);
    const BrlPpoParams* p = static_cast<const BrlPpoParams*>(opaque);
    if (p->batch <= 0 || p->batch > (1 << 24)) return fail(BRL_E_OPAQUE, "brl_ppo_grad: batch must be in [1, 2^24]");
    static const char* names[] = {"obs", "blob", "scratch", "index", "mask", "action", "old_log_prob", "old_value", "advantages",
                                  "targets", "grads", "stats", "acc"};
    for (int k = 0; k < 13; ++k)
        if (b[k] == nullptr && k != 3) return fail(BRL_E_BUFFER, "brl_ppo_grad: buffer '%s' is NULL", names[k]);
    for (int k : {0, 1, 2, 10})
        if ((reinterpret_cast<uintptr_t>(b[k]) & 15u) != 0) return fail(BRL_E_BUFFER, "brl_ppo_grad: buffer '%s' is not 16-byte aligned", names[k]);
    if (encode_fn() == nullptr) return fail(BRL_E_LAUNCH, "brl_ppo_grad: cuTensorMapEncodeTiled not available from the driver");
    cudaStream_t s = (cudaStream_t)stream;
    const int B = (int)p->batch;
    const MlpLayout L = mlp_layout();
    const TrainBlob T = train_blob();
    const FlatLayout F = flat_layout();
    const TrainScratch S = train_scratch(B);
    const unsigned char* blob = static_cast<const unsigned char*>(b[1]);
    unsigned char* sc = static_cast<unsigned char*>(b[2]);
    float* grads = static_cast<float*>(b[10]);
    const bool narrow = (p->reserved & 1) != 0;
    auto bf = [&](size_t off) { return reinterpret_cast<__nv_bfloat16*>(sc + off); };
    const int ldt = S.ldt;
    int32_t rc = BRL_OK;

    // 1. minibatch gather of the observation (src/update.py:194-199, cast of src/update.py:95)
    {
        const dim3 grid((unsigned)((B + 31) / 32), kObsDimM / kGatherCols);
        const int32_t* index = static_cast<const int32_t*>(b[3]);
        uint16_t* o_r = reinterpret_cast<uint16_t*>(sc + S.obs_r);
        uint16_t* o_t = reinterpret_cast<uint16_t*>(sc + S.obs_t);
        if (p->flags & BRL_PPO_OBS_BF16) k_gather_obs<__nv_bfloat16><<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(b[0]), index, B, ldt, o_r, o_t);
        else if (p->flags & BRL_PPO_OBS_U8) k_gather_obs<uint8_t><<<grid, 256, 0, s>>>(static_cast<const uint8_t*>(b[0]), index, B, ldt, o_r, o_t);
        else k_gather_obs<float><<<grid, 256, 0, s>>>(static_cast<const float*>(b[0]), index, B, ldt, o_r, o_t);
    }
    // 2. forward, activations kept (row-major for the next layer / the ReLU mask, batch-major for the wgrad)
    for (int l = 0; l < 4 && rc == BRL_OK; ++l) {
        GemmArgs a{};
        a.bias = reinterpret_cast<const float*>(blob + L.bias[l]);
        a.out_hi = bf(S.h_r_hi[l]); a.out_lo = bf(S.h_r_lo[l]);
        a.out_t_hi = bf(S.h_t_hi[l]); a.out_t_lo = bf(S.h_t_lo[l]);
        a.ld_out = kHidden; a.ld_t = ldt;
        if (l == 0) rc = launch_hidden<false, kEpiFwd>(narrow, s, sc + S.obs_r, nullptr, B, kObsDimM, blob + L.w_hi[0], blob + L.w_lo[0], kHidden, kObsDimM, kObsDimM, a);
        else rc = launch_hidden<true, kEpiFwd>(narrow, s, sc + S.h_r_hi[l - 1], sc + S.h_r_lo[l - 1], B, kHidden, blob + L.w_hi[l], blob + L.w_lo[l], kHidden, kHidden, kHidden, a);
    }
    if (rc != BRL_OK) return rc;
    {
        GemmArgs a{};
        a.bias = reinterpret_cast<const float*>(blob + L.bias[4]);
        a.logits = reinterpret_cast<float*>(sc + S.logits);
        a.value = reinterpret_cast<float*>(sc + S.value);
        rc = launch_gemm<kHeadPad, true, true, kEpiHead>(s, sc + S.h_r_hi[3], sc + S.h_r_lo[3], B, kHidden, blob + L.w_hi[4], blob + L.w_lo[4], kHeadPad, kHidden, kHidden, a);
        if (rc != BRL_OK) return rc;
    }
    if ((rc = check_launch("brl_ppo_grad (forward)")) != BRL_OK) return rc;
    // 3. loss head + its backward (src/update.py:97-162)
    {
        void* lb[13] = {sc + S.logits, sc + S.value, b[3], b[4], b[5], b[6], b[7], b[8], b[9], sc + S.dlogits, sc + S.dvalue, b[11], b[12]};
        BrlPpoParams lp = *p;
        lp.flags &= (BRL_PPO_VALUE_CLIPPING | BRL_PPO_REWARD_SCALING | BRL_PPO_UNMASKED_POLICY);
        lp.reserved = 0;
        if ((rc = brl_ppo_loss(stream, lb, &lp, sizeof lp)) != BRL_OK) return rc;
        k_head_grad_pack<<<(unsigned)(((int64_t)B * kHeadPad + 255) / 256), 256, 0, s>>>(
            reinterpret_cast<const float*>(sc + S.dlogits), reinterpret_cast<const float*>(sc + S.dvalue), B, ldt, bf(S.dz5_r_hi), bf(S.dz5_r_lo),
            bf(S.dz5_t_hi), bf(S.dz5_t_lo));
    }
    // 4. backward: head, then hidden layers 4..1.  dz of layer l lives in dz_r[(4 - l) & 1] / dz_t[l - 1] (l = 1..4).
    {   // head wgrad: [1024, 39] = h4^T . dz5 -> w4 grads (38 columns) + w5 grads (the value column)
        GemmArgs a{};
        a.c = grads + F.w[4]; a.ld_c = 38; a.n_c = 38; a.c2 = grads + F.w[5];
        rc = launch_gemm<kHeadPad, true, true, kEpiWgrad>(s, sc + S.h_t_hi[3], sc + S.h_t_lo[3], kHidden, ldt, sc + S.dz5_t_hi, sc + S.dz5_t_lo, kHeadPad, ldt, B, a);
        if (rc != BRL_OK) return rc;
    }
    for (int l = 4; l >= 1 && rc == BRL_OK; --l) {
        // dgrad into layer l's pre-activation: dz_l = (dz_{l+1} . W_{l+1}^T) * [h_l > 0]
        const int k_out = l == 4 ? kHeadPad : kHidden;  // width of dz_{l+1}
        const void* up_hi = l == 4 ? (const void*)(sc + S.dz5_r_hi) : (const void*)(sc + S.dz_r_hi[(4 - (l + 1)) & 1]);
        const void* up_lo = l == 4 ? (const void*)(sc + S.dz5_r_lo) : (const void*)(sc + S.dz_r_lo[(4 - (l + 1)) & 1]);
        GemmArgs a{};
        a.out_hi = bf(S.dz_r_hi[(4 - l) & 1]); a.out_lo = bf(S.dz_r_lo[(4 - l) & 1]);
        a.out_t_hi = bf(S.dz_t_hi[l - 1]); a.out_t_lo = bf(S.dz_t_lo[l - 1]);
        a.ld_out = kHidden; a.ld_t = ldt;
        a.relu_src = bf(S.h_r_hi[l - 1]);
        rc = launch_hidden<true, kEpiDgrad>(narrow, s, up_hi, up_lo, B, k_out, blob + T.wn_hi[l], blob + T.wn_lo[l], kHidden, k_out, k_out, a);
        if (rc != BRL_OK) break;
        // wgrad of layer l (0-based parameter index l - 1): dW = h_{l-1}^T . dz_l
        GemmArgs w{};
        w.c = grads + F.w[l - 1]; w.ld_c = kHidden; w.n_c = kHidden;
        if (l == 1) rc = launch_hidden<false, kEpiWgrad>(narrow, s, sc + S.obs_t, nullptr, kObsDimM, ldt, sc + S.dz_t_hi[0], sc + S.dz_t_lo[0], kHidden, ldt, B, w);
        else rc = launch_hidden<true, kEpiWgrad>(narrow, s, sc + S.h_t_hi[l - 2], sc + S.h_t_lo[l - 2], kHidden, ldt, sc + S.dz_t_hi[l - 1], sc + S.dz_t_lo[l - 1], kHidden, ldt, B, w);
    }
    if (rc != BRL_OK) return rc;
    // 5. bias gradients
    {
        BiasArgs a{};
        for (int l = 0; l < 4; ++l) { a.t_hi[l] = bf(S.dz_t_hi[l]); a.t_lo[l] = bf(S.dz_t_lo[l]); }
        a.dlogits = reinterpret_cast<const float*>(sc + S.dlogits);
        a.dvalue = reinterpret_cast<const float*>(sc + S.dvalue);
        a.grads = grads; a.B = B; a.ldt = ldt; a.F = F;
        const int n_warps = 4 * kHidden + kHeadValid;
        k_bias_grad<<<(unsigned)((n_warps * 32 + 255) / 256), 256, 0, s>>>(a);
    }
    return check_launch("brl_ppo_grad");
}

}  // extern "C"
