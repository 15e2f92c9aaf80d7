// STUB of jaxlib's xla/ffi/api/ffi.h - NOT the real header.  jaxlib is not installed in this repository's image, so
// integration/nbm_xla_ffi.cc could never be compiled here; this file declares just the few xla::ffi names the shim uses,
// with the real header's shapes (Buffer<T>::typed_data(), ResultBuffer<T> = Result<Buffer<T>> with operator->,
// Error(ErrorCode, std::string), Ffi::Bind().Ctx<>().Arg<>().Attr<>(name).Ret<>()), so that `__graft_entry__.build()`
// can type-check the shim: the handler's C++ signature must be invocable with exactly the argument list its binding
// declares (what the real XLA_FFI_DEFINE_HANDLER_SYMBOL enforces through Handler<...>), and every nbm_* call in it must
// match include/nbm_b200.h.  A maintainer builds the shim against jaxlib's own header; nothing in
// the product loads this.
#pragma once
#include <cstdint>
#include <string>
#include <tuple>
#include <type_traits>
#include <utility>

namespace xla {
namespace ffi {

enum class DataType { F32 };
constexpr DataType F32 = DataType::F32;
template <DataType> struct NativeOf;
template <> struct NativeOf<DataType::F32> { using type = float; };

template <DataType dt>
class Buffer {
 public:
  using T = typename NativeOf<dt>::type;
  T* typed_data() const { return data_; }
  std::size_t element_count() const { return n_; }
 private:
  T* data_ = nullptr;
  std::size_t n_ = 0;
};

template <typename T>
class Result {
 public:
  T* operator->() { return &v_; }
  T& operator*() { return v_; }
 private:
  T v_;
};
template <DataType dt> using ResultBuffer = Result<Buffer<dt>>;

enum class ErrorCode { kOk, kInternal, kInvalidArgument };
class Error {
 public:
  Error() = default;
  Error(ErrorCode c, std::string m) : code_(c), msg_(std::move(m)) {}
  static Error Success() { return Error(); }
  bool success() const { return code_ == ErrorCode::kOk; }
 private:
  ErrorCode code_ = ErrorCode::kOk;
  std::string msg_;
};

template <typename T> struct PlatformStream { using type = T; };

// what the handler receives for each binding entry
template <typename T> struct CtxArg;
template <typename T> struct CtxArg<PlatformStream<T>> { using type = T; };
template <typename T> struct RetArg { using type = Result<T>; };

template <typename... Ts>
struct Binding {
  template <typename C> Binding<Ts..., typename CtxArg<C>::type> Ctx() const { return {}; }
  template <typename A> Binding<Ts..., A> Arg() const { return {}; }
  template <typename A> Binding<Ts..., A> Attr(const char*) const { return {}; }
  template <typename R> Binding<Ts..., typename RetArg<R>::type> Ret() const { return {}; }
  template <typename F>
  static constexpr bool Accepts() { return std::is_invocable_r<Error, F, Ts...>::value; }
};

struct Ffi {
  static Binding<> Bind() { return {}; }
};

}  // namespace ffi
}  // namespace xla

struct XLA_FFI_CallFrame;
struct XLA_FFI_Error;
#define XLA_FFI_DEFINE_HANDLER_SYMBOL(sym, impl, binding)                                                   \
  static_assert(decltype(binding)::template Accepts<decltype(&impl)>(),                                     \
                "handler signature does not match its binding");                                            \
  extern "C" XLA_FFI_Error* sym(XLA_FFI_CallFrame*) { return nullptr; }
