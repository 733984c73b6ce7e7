// common.h -- error plumbing shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include "../../include/brl_b200.h"

namespace brl {

char* last_error_buffer();  // thread-local, 512 bytes

inline int32_t fail(int32_t code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(last_error_buffer(), 512, fmt, ap);
    va_end(ap);
    return code;
}

inline int32_t check_launch(const char* what) {
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return fail(BRL_E_LAUNCH, "%s: %s", what, cudaGetErrorString(err));
    return BRL_OK;
}

inline const BrlParams* get_params(const void* opaque, size_t opaque_len, int32_t* rc) {
    if (opaque == nullptr || opaque_len != sizeof(BrlParams)) {
        *rc = fail(BRL_E_OPAQUE, "opaque must be one BrlParams (%zu bytes), got %zu", sizeof(BrlParams), opaque_len);
        return nullptr;
    }
    const BrlParams* p = static_cast<const BrlParams*>(opaque);
    if (p->n_envs < 0) {
        *rc = fail(BRL_E_OPAQUE, "n_envs < 0");
        return nullptr;
    }
    *rc = BRL_OK;
    return p;
}

// Per-device slot of a function-local cache (SM count, "kernel attribute already set", co-resident CTA count): these are
// per-device facts, and a process may drive several GPUs.  Races only repeat idempotent driver calls.
constexpr int kMaxDevices = 64;
inline int device_slot() {
    int dev = 0;
    cudaGetDevice(&dev);
    return dev >= 0 && dev < kMaxDevices ? dev : 0;
}

// masked categorical over f32 logits[n,38] (brl_algo.cu); shared by brl_categorical and the unfused brl_policy_act path
int32_t launch_categorical(cudaStream_t stream, const float* logits, const unsigned char* mask, int32_t* action, float* log_prob,
                           int64_t n, int sample, uint64_t seed, int64_t env_offset, uint32_t step, const int32_t* row_index = nullptr,
                           float* logits_by_env = nullptr);

// brl_ppo_loss's body (brl_ppo.cu), optionally also writing d loss / d (logits, value) as bf16 hi / lo rows [B, 64] for brl_ppo_grad
// acc_zeroed: the caller already cleared the f64[16] scratch on this stream (keeps a memset out of a PDL kernel chain)
// defer_illegal (honoured only when illegal_l2_coef == 0): only the partial Gram matrices are formed before the loss kernel;
// the caller runs gram_tail() itself later (brl_ppo_grad: a spare block of k_bias_grad) with gram_args(buffers, p, &stats[6])
// and gram_blocks(batch, 1024) partials
int32_t launch_ppo_loss(cudaStream_t stream, void** buffers, const BrlPpoParams* p, void* dz_hi, void* dz_lo, bool acc_zeroed,
                        bool defer_illegal);
struct GramArgs;
GramArgs gram_args(void** loss_buffers, const BrlPpoParams* p, float* stat);
int gram_blocks(int64_t batch, int threads);

// Programmatic dependent launch for the short kernels of a dependent chain (the PPO optimizer step): the kernel may be
// scheduled while its predecessor in the stream drains; it must execute pdl_wait() before touching global memory and
// should call pdl_trigger() first so that ITS successor can be scheduled early too.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at{};
    at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

#define BRL_REQUIRE(ptr, name)                                                            \
    do {                                                                                  \
        if ((ptr) == nullptr) return brl::fail(BRL_E_BUFFER, "%s: buffer '%s' is NULL", __func__, name); \
        if ((reinterpret_cast<uintptr_t>(ptr) & 15u) != 0)                                \
            return brl::fail(BRL_E_BUFFER, "%s: buffer '%s' is not 16-byte aligned", __func__, name);    \
    } while (0)

}  // namespace brl
