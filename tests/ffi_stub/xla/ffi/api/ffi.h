// TEST INFRASTRUCTURE ONLY -- a minimal stand-in for the part of XLA's typed-FFI C++ API (xla/ffi/api/ffi.h) that
// brl_b200/csrc/xla_ffi_shim.cc uses, written from the public API's documented shape.  The real headers ship with jaxlib,
// which this image does not have; this stub lets tests/test_abi_and_layout.py at least COMPILE the typed-FFI half of
// the shim (g++ -fsyntax-only) so that it cannot rot unseen.  It proves nothing about behaviour inside XLA.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <utility>

namespace xla {
namespace ffi {

enum class ErrorCode { kOk = 0, kInvalidArgument = 3 };

class Error {
 public:
    Error() = default;
    Error(ErrorCode code, std::string message) : code_(code), message_(std::move(message)) {}
    static Error Success() { return Error(); }
    bool failure() const { return code_ != ErrorCode::kOk; }
 private:
    ErrorCode code_ = ErrorCode::kOk;
    std::string message_;
};

template <class T>
class ErrorOr {
 public:
    explicit ErrorOr(T v) : v_(std::move(v)) {}
    bool has_value() const { return true; }
    T& value() { return v_; }
    T* operator->() { return &v_; }
 private:
    T v_;
};

class AnyBuffer {
 public:
    explicit AnyBuffer(void* p = nullptr) : p_(p) {}
    void* untyped_data() const { return p_; }
 private:
    void* p_;
};

template <class T>
class Result {
 public:
    explicit Result(T v) : v_(std::move(v)) {}
    T* operator->() { return &v_; }
    T& operator*() { return v_; }
 private:
    T v_;
};

template <class T>
class Span {
 public:
    Span(T* data, std::size_t size) : data_(data), size_(size) {}
    T* begin() const { return data_; }
    T* end() const { return data_ + size_; }
    std::size_t size() const { return size_; }
 private:
    T* data_;
    std::size_t size_;
};

class RemainingArgs {
 public:
    std::size_t size() const { return 0; }
    template <class T> ErrorOr<T> get(std::size_t) const { return ErrorOr<T>(T()); }
};
class RemainingRets {
 public:
    std::size_t size() const { return 0; }
    template <class T> ErrorOr<Result<T>> get(std::size_t) const { return ErrorOr<Result<T>>(Result<T>(T())); }
};

template <class S> struct PlatformStream {};

struct Binding {
    template <class T> Binding& Ctx() { return *this; }
    Binding& RemainingArgs() { return *this; }
    Binding& RemainingRets() { return *this; }
    template <class T> Binding& Attr(const char*) { return *this; }
};
struct Ffi {
    static Binding Bind() { return Binding(); }
};

}  // namespace ffi
}  // namespace xla

// the real macro defines an exported XLA_FFI_Error* symbol(XLA_FFI_CallFrame*) that decodes the frame per the binding and
// calls `impl`; here: take the address of `impl` with the decoded signature so that signature mismatches fail to compile
#define XLA_FFI_DEFINE_HANDLER_SYMBOL(symbol, impl, binding)                                                    \
    extern "C" void* symbol() {                                                                                  \
        (void)(binding);                                                                                         \
        xla::ffi::Error (*fn)(struct CUstream_st*, xla::ffi::RemainingArgs, xla::ffi::RemainingRets,             \
                              xla::ffi::Span<const uint8_t>) = impl;                                             \
        return reinterpret_cast<void*>(fn);                                                                      \
    }
