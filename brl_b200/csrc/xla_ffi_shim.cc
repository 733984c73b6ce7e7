// xla_ffi_shim.cc -- adapters that expose the C-ABI ops (include/brl_b200.h) to a jitted
// JAX program, so that brl's `ppo.py` / `eval.py` loops can call the CUDA env as a drop-in
// for `pgx.bridge_bidding` (src/roll_out.py:51, src/duplicate.py:149, src/gae.py:32-38).
//
// STATUS: this image has neither jax/jaxlib nor the XLA FFI headers, so this file is
// compiled ONLY when `xla/ffi/api/ffi.h` is on the include path (build.py probes
// `jaxlib.include` / `jax.ffi.include_dir()`); it is NOT built or exercised in this
// environment -- INTEGRATION.md says so.  The product path here is driven through the
// same symbols from Python/ctypes with torch-owned buffers.
//
// Two conventions, both thin because the core ABI already has the custom-call shape:
//   (1) legacy GPU custom call, API_VERSION_STATUS_RETURNING (what jax 0.4.23 -- the
//       version brl pins, requirements.txt:25-26 -- offers):
//         void f(cudaStream_t, void** buffers, const char* opaque, size_t len, XlaCustomCallStatus*)
//       `buffers` = operands then results, exactly the order each op documents;
//   (2) typed FFI (jax >= 0.4.31): XLA_FFI_DEFINE_HANDLER_SYMBOL over ffi::Buffer args.
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define BRL_HAVE_XLA_FFI 1
#endif
#endif

#include <cstring>

#include "../../include/brl_b200.h"

// ---- (1) legacy custom calls: always compilable, no XLA headers needed ----------------
// XlaCustomCallStatus is opaque; failure is reported through the symbol XLA provides.
extern "C" {
struct XlaCustomCallStatus_;
typedef struct XlaCustomCallStatus_ XlaCustomCallStatus;
#if defined(BRL_HAVE_XLA_FFI) || defined(BRL_LINK_XLA_STATUS)
void XlaCustomCallStatusSetFailure(XlaCustomCallStatus*, const char*, size_t);
#else
static void XlaCustomCallStatusSetFailure(XlaCustomCallStatus*, const char*, size_t) {}
#endif

#define BRL_LEGACY_CUSTOM_CALL(op)                                                              \
    void op##_xla(brl_stream_t stream, void** buffers, const char* opaque, size_t opaque_len,   \
                  XlaCustomCallStatus* status) {                                                \
        if (op(stream, buffers, opaque, opaque_len) != BRL_OK) {                                \
            const char* msg = brl_last_error();                                                 \
            XlaCustomCallStatusSetFailure(status, msg, std::strlen(msg));                       \
        }                                                                                       \
    }

BRL_LEGACY_CUSTOM_CALL(brl_make_keys)
BRL_LEGACY_CUSTOM_CALL(brl_init)
BRL_LEGACY_CUSTOM_CALL(brl_reset_fields)
BRL_LEGACY_CUSTOM_CALL(brl_step)
BRL_LEGACY_CUSTOM_CALL(brl_duplicate_step)
BRL_LEGACY_CUSTOM_CALL(brl_duplicate_init)
BRL_LEGACY_CUSTOM_CALL(brl_observe)
BRL_LEGACY_CUSTOM_CALL(brl_legal_mask)
BRL_LEGACY_CUSTOM_CALL(brl_rollout_random)
BRL_LEGACY_CUSTOM_CALL(brl_imp_reward)
BRL_LEGACY_CUSTOM_CALL(brl_gae)
BRL_LEGACY_CUSTOM_CALL(brl_categorical)
BRL_LEGACY_CUSTOM_CALL(brl_match_stats)
BRL_LEGACY_CUSTOM_CALL(brl_state_fields)
BRL_LEGACY_CUSTOM_CALL(brl_gather_reward)
BRL_LEGACY_CUSTOM_CALL(brl_mlp_pack)
BRL_LEGACY_CUSTOM_CALL(brl_obs_to_bf16)
BRL_LEGACY_CUSTOM_CALL(brl_mlp_forward)
BRL_LEGACY_CUSTOM_CALL(brl_policy_act)
BRL_LEGACY_CUSTOM_CALL(brl_ppo_loss)
BRL_LEGACY_CUSTOM_CALL(brl_adam_clip)
BRL_LEGACY_CUSTOM_CALL(brl_adam_apply)
BRL_LEGACY_CUSTOM_CALL(brl_gather_rows)
BRL_LEGACY_CUSTOM_CALL(brl_mlp_pack_train)
BRL_LEGACY_CUSTOM_CALL(brl_mlp_adam_step)
BRL_LEGACY_CUSTOM_CALL(brl_ppo_grad)
BRL_LEGACY_CUSTOM_CALL(brl_eval_act_log)
BRL_LEGACY_CUSTOM_CALL(brl_eval_summary)
}  // extern "C"

// ---- (2) typed FFI handlers --------------------------------------------------------------
#ifdef BRL_HAVE_XLA_FFI
#include "xla/ffi/api/ffi.h"
namespace ffi = xla::ffi;

namespace {
ffi::Error to_error(int32_t rc) {
    return rc == BRL_OK ? ffi::Error::Success() : ffi::Error(ffi::ErrorCode::kInvalidArgument, brl_last_error());
}

BrlParams params(int64_t n, int64_t stride, int32_t n_deals, int32_t flags, uint64_t seed, uint32_t step,
                 int32_t k_steps, int64_t env_offset) {
    BrlParams p{};
    p.n_envs = n; p.state_stride = stride; p.n_deals = n_deals; p.flags = flags; p.seed = seed; p.step = step;
    p.k_steps = k_steps; p.env_offset = env_offset; p.illegal_penalty = -1.0f; p.illegal_bonus = 1.0f;
    return p;
}

// env.step(state, action): state/outputs are natively batched, so jax.vmap folds into N
// (ffi_call(..., vmap_method="broadcast_all") with the batch axis leading).
ffi::Error StepImpl(cudaStream_t stream, ffi::AnyBuffer state, ffi::Buffer<ffi::S32> action, ffi::AnyBuffer table,
                    ffi::Result<ffi::AnyBuffer> state_out, ffi::Result<ffi::AnyBuffer> obs,
                    ffi::Result<ffi::AnyBuffer> mask, ffi::Result<ffi::AnyBuffer> rewards,
                    ffi::Result<ffi::AnyBuffer> terminated, ffi::Result<ffi::AnyBuffer> current_player, int32_t flags) {
    const int64_t n = action.element_count();
    void* b[10] = {state.untyped_data(), action.untyped_data(), table.untyped_data(), state_out->untyped_data(),
                   obs->untyped_data(), mask->untyped_data(), rewards->untyped_data(), terminated->untyped_data(),
                   current_player->untyped_data(), nullptr};
    BrlParams p = params(n, n, (int32_t)(table.element_count() / BRL_DEAL_ROW_BYTES), flags, 0, 0, 0, 0);
    return to_error(brl_step((brl_stream_t)stream, b, &p, sizeof(p)));
}

ffi::Error GaeImpl(cudaStream_t stream, ffi::AnyBuffer done, ffi::Buffer<ffi::F32> value, ffi::Buffer<ffi::F32> reward,
                   ffi::Buffer<ffi::F32> last_val, ffi::Result<ffi::Buffer<ffi::F32>> adv,
                   ffi::Result<ffi::Buffer<ffi::F32>> targets, float gamma, float gae_lambda) {
    const int64_t n = last_val.element_count();
    void* b[6] = {done.untyped_data(), value.untyped_data(), reward.untyped_data(), last_val.untyped_data(),
                  adv->untyped_data(), targets->untyped_data()};
    BrlParams p = params(n, n, 0, 0, 0, 0, (int32_t)(value.element_count() / n), 0);
    p.gamma = gamma; p.gae_lambda = gae_lambda;
    return to_error(brl_gae((brl_stream_t)stream, b, &p, sizeof(p)));
}
}  // namespace

XLA_FFI_DEFINE_HANDLER_SYMBOL(brl_step_ffi, StepImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>().Arg<ffi::Buffer<ffi::S32>>().Arg<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>().Ret<ffi::AnyBuffer>().Ret<ffi::AnyBuffer>().Ret<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>().Ret<ffi::AnyBuffer>().Attr<int32_t>("flags"));
XLA_FFI_DEFINE_HANDLER_SYMBOL(brl_gae_ffi, GaeImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>().Arg<ffi::Buffer<ffi::F32>>().Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>().Ret<ffi::Buffer<ffi::F32>>()
                                  .Attr<float>("gamma").Attr<float>("gae_lambda"));
#endif  // BRL_HAVE_XLA_FFI
