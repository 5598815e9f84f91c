// STUB of the XLA FFI C++ API (xla/ffi/api/ffi.h ships inside jaxlib, which is not installed in this
// image).  It declares just the subset ffi/jax_ffi_shim.cc uses, with the SAME binding discipline:
// Ffi::Bind().Ctx<>().Arg<>().Attr<>().Ret<>() accumulates the handler's parameter types in order and
// `.To(fn)` refuses to compile unless `fn` is callable with exactly those types -- so a handler whose
// Bind() and implementation disagree in arity or type (round 1's JpsPaint) fails here at compile time,
// as it would fail at registration time with the real header.  Nothing in it runs.
// Used by tests/test_ffi_shim.py:  g++ -fsyntax-only -DJPS_WITH_JAX_FFI -I ffi/stub -I include ...
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <type_traits>
#include <utility>

namespace xla {
namespace ffi {

enum DataType { F32, F64, U8, S32, S64, C64 };
template <DataType dt> struct NativeTypeOf;
template <> struct NativeTypeOf<F32> { using type = float; };
template <> struct NativeTypeOf<F64> { using type = double; };
template <> struct NativeTypeOf<U8> { using type = uint8_t; };
template <> struct NativeTypeOf<S32> { using type = int32_t; };
template <> struct NativeTypeOf<S64> { using type = int64_t; };

template <typename T>
class Span {
 public:
  const T* begin() const { return p_; }
  const T* end() const { return p_ + n_; }
  size_t size() const { return n_; }
  const T& operator[](size_t i) const { return p_[i]; }

 private:
  const T* p_ = nullptr;
  size_t n_ = 0;
};

template <DataType dt>
class Buffer {
 public:
  using T = typename NativeTypeOf<dt>::type;
  T* typed_data() const { return nullptr; }
  void* untyped_data() const { return nullptr; }
  size_t element_count() const { return 0; }
  size_t size_bytes() const { return 0; }
  Span<const int64_t> dimensions() const { return {}; }
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

enum class ErrorCode { kOk, kInvalidArgument, kInternal };
class Error {
 public:
  Error() = default;
  Error(ErrorCode, std::string) {}
  static Error Success() { return Error(); }
};

template <typename T> struct PlatformStream {};
template <typename T> struct CtxType;
template <typename T> struct CtxType<PlatformStream<T>> { using type = T; };

template <typename... Ts>
struct Binding {
  template <typename T> Binding<Ts..., typename CtxType<T>::type> Ctx() const { return {}; }
  template <typename T> Binding<Ts..., T> Arg() const { return {}; }
  template <typename T> Binding<Ts..., T> Attr(const char*) const { return {}; }
  template <typename T> Binding<Ts..., Result<T>> Ret() const { return {}; }
  template <typename Fn>
  int To(Fn&&) const {
    static_assert(std::is_invocable_r<Error, Fn, Ts...>::value,
                  "XLA FFI binding and handler implementation disagree in arity or argument types");
    return 0;
  }
};

struct Ffi {
  static Binding<> Bind() { return {}; }
};

}  // namespace ffi
}  // namespace xla

struct XLA_FFI_CallFrame;
struct XLA_FFI_Error;
#define XLA_FFI_DEFINE_HANDLER_SYMBOL(symbol, impl, binding)          \
  static const int symbol##_binding_check = (binding).To(impl);       \
  extern "C" XLA_FFI_Error* symbol(XLA_FFI_CallFrame*) { return nullptr; }
