// TEST INFRASTRUCTURE ONLY.  A minimal stand-in for XLA's xla/ffi/api/ffi.h, written from the public documentation of
// the XLA FFI binding DSL.  It exists so that jax-cpfem_b200/csrc/cpfem_ffi.cc can be parsed and TYPE-CHECKED with g++ in
// a container that has no JAX (tests/test_abi.py::test_ffi_adapter_type_checks): the Bind() chain accumulates the
// parameter types in binding order and XLA_FFI_DEFINE_HANDLER_SYMBOL static_asserts that the handler is callable with
// exactly those types and returns ffi::Error - the property a real build enforces.  Nothing here executes; a real build
// uses the headers of `jax.ffi.include_dir()` instead (cpfem_b200/jax_ffi.py).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <type_traits>
#include <utility>

#include "xla/ffi/api/c_api.h"

namespace xla {
namespace ffi {

enum DataType { F64, S64, S32 };
template <DataType> struct NativeOf;
template <> struct NativeOf<F64> { using type = double; };
template <> struct NativeOf<S64> { using type = int64_t; };
template <> struct NativeOf<S32> { using type = int32_t; };

enum class ErrorCode { kOk, kInvalidArgument, kInternal };
class Error {
 public:
    Error() = default;
    Error(ErrorCode c, std::string m) : code_(c), msg_(std::move(m)) {}
    static Error Success() { return Error(); }
    bool failure() const { return code_ != ErrorCode::kOk; }
    bool success() const { return !failure(); }
 private:
    ErrorCode code_ = ErrorCode::kOk;
    std::string msg_;
};

template <DataType dtype>
class Buffer {
 public:
    using T = typename NativeOf<dtype>::type;
    T* typed_data() const { return data_; }
    size_t element_count() const { return n_; }
    size_t size_bytes() const { return n_ * sizeof(T); }
 private:
    T* data_ = nullptr;
    size_t n_ = 0;
};
template <typename T>
class Result {
 public:
    T* operator->() { return &v_; }
    T& operator*() { return v_; }
 private:
    T v_;
};
template <DataType dtype> using ResultBuffer = Result<Buffer<dtype>>;

template <typename T>
class ErrorOr {
 public:
    bool has_value() const { return ok_; }
    T& value() { return v_; }
 private:
    bool ok_ = true;
    T v_;
};
class RemainingArgs {
 public:
    size_t size() const { return 0; }
    template <typename T> ErrorOr<T> get(size_t) const { return ErrorOr<T>(); }
};

template <typename T> struct PlatformStream {};
template <typename T> struct StructMember { explicit StructMember(const char*) {} };

template <typename... Ts> struct TypeList {};
template <typename T> struct CtxArg { using type = T; };
template <typename T> struct CtxArg<PlatformStream<T>> { using type = T; };
template <typename T> struct RetArg { using type = Result<T>; };

template <typename... Ts>
struct Binding {
    template <typename T> constexpr Binding<Ts..., typename CtxArg<T>::type> Ctx() const { return {}; }
    template <typename T> constexpr Binding<Ts..., T> Attr(const char*) const { return {}; }
    template <typename T> constexpr Binding<Ts..., T> Arg() const { return {}; }
    template <typename T> constexpr Binding<Ts..., typename RetArg<T>::type> Ret() const { return {}; }
    constexpr Binding<Ts..., ffi::RemainingArgs> RemainingArgs() const { return {}; }
    template <typename Fn>
    static constexpr bool Matches() { return std::is_invocable_r<Error, Fn, Ts...>::value; }
};
struct Ffi {
    static constexpr Binding<> Bind() { return {}; }
};

}  // namespace ffi
}  // namespace xla

#define XLA_FFI_REGISTER_STRUCT_ATTR_DECODING(T, ...) \
    static const int xla_ffi_struct_decoding_##T = ((void)std::initializer_list<int>{((void)(__VA_ARGS__), 0)}, 0)
#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, impl, binding)                                                              \
    static_assert(decltype(binding)::template Matches<decltype(&impl)>(),                                              \
                  #name ": handler signature does not match its binding (order: Ctx, Attr, Arg, Ret as bound)");       \
    extern "C" XLA_FFI_Error* name(XLA_FFI_CallFrame*) { return nullptr; }
