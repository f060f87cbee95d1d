// COMPILE-CHECK MOCK, TEST INFRASTRUCTURE ONLY -- not XLA, never linked into anything that runs.
//
// jax / XLA are not installable in the build image, so csrc/jic_xla_ffi.cc cannot be compiled against the real
// xla/ffi/api/ffi.h here.  This header declares just the names that file uses, with the semantics of the real API that matter
// for a type check: a binding collects (Ctx | Arg | Ret | Attr) in order, and the implementation must be invocable with exactly
// that parameter list (Ctx<PlatformStream<T>> -> T, Arg<X> -> X, Ret<X> -> Result<X>, Attr<X> -> X) and return Error.
// tests/test_abi.py::test_xla_ffi_glue_type_checks compiles the glue against it with -fsyntax-only: this catches typos and a
// binding whose order disagrees with the handler's signature; it says nothing about XLA's ABI (the real header generates that).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <string_view>
#include <type_traits>

struct XLA_FFI_Error;
struct XLA_FFI_CallFrame;

namespace xla::ffi {

enum class DataType { PRED, U8, S32, S64, F32, F64 };
inline constexpr DataType F32 = DataType::F32;
inline constexpr DataType F64 = DataType::F64;
inline constexpr DataType U8 = DataType::U8;

template <class T>
class Span {
 public:
  size_t size() const { return n_; }
  const T& operator[](size_t i) const { return p_[i]; }
  const T* begin() const { return p_; }
  const T* end() const { return p_ + n_; }

 private:
  const T* p_ = nullptr;
  size_t n_ = 0;
};

class AnyBuffer {
 public:
  DataType element_type() const { return DataType::F64; }
  Span<const int64_t> dimensions() const { return {}; }
  void* untyped_data() const { return nullptr; }
  size_t element_count() const { return 0; }
  size_t size_bytes() const { return 0; }
};

template <DataType dtype>
class Buffer {
 public:
  using Native = std::conditional_t<dtype == DataType::F32, float, std::conditional_t<dtype == DataType::U8, uint8_t, double>>;
  Span<const int64_t> dimensions() const { return {}; }
  Native* typed_data() const { return nullptr; }
  void* untyped_data() const { return nullptr; }
  size_t element_count() const { return 0; }
};

template <class T>
class Result {
 public:
  T* operator->() { return &v_; }
  T& operator*() { return v_; }

 private:
  T v_;
};
template <DataType dtype>
using ResultBuffer = Result<Buffer<dtype>>;

class Error {
 public:
  static Error Success() { return {}; }
  static Error InvalidArgument(std::string) { return {}; }
  static Error Internal(std::string) { return {}; }
  bool failure() const { return false; }
  bool success() const { return true; }
};

template <class T>
struct PlatformStream {};

template <class... Params>
struct Handler {
  XLA_FFI_Error* Call(XLA_FFI_CallFrame*) const { return nullptr; }
};

template <class... Params>
struct Binding {
  template <class T>
  struct CtxParam { using type = T; };
  template <class T>
  struct CtxParam<PlatformStream<T>> { using type = T; };

  template <class T>
  Binding<Params..., typename CtxParam<T>::type> Ctx() const { return {}; }
  template <class T>
  Binding<Params..., T> Arg() const { return {}; }
  template <class T>
  Binding<Params..., Result<T>> Ret() const { return {}; }
  template <class T>
  Binding<Params..., T> Attr(std::string_view) const { return {}; }

  template <class F>
  Handler<Params...> To(F&&) const {
    static_assert(std::is_invocable_r_v<Error, F, Params...>, "handler signature does not match the binding (order and types)");
    return {};
  }
};

struct Ffi {
  static Binding<> Bind() { return {}; }
};

}  // namespace xla::ffi

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(symbol, impl, binding)          \
  extern "C" XLA_FFI_Error* symbol(XLA_FFI_CallFrame* call_frame) {   \
    static auto handler = (binding).To(impl);                         \
    return handler.Call(call_frame);                                  \
  }
