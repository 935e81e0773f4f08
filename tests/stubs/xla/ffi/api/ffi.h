// SYNTAX-ONLY stand-in for jaxlib's xla/ffi/api/ffi.h (which does not exist in the build image): just enough of the
// binding vocabulary for tests/test_capi_and_host.py to keep autopdex_b200/csrc/xla/apdx_b200_xla.cc compiling and to
// check that every handler's parameter list matches its Bind() chain.  Test infrastructure; never shipped or linked.
#ifndef APDX_TEST_STUB_XLA_FFI_H
#define APDX_TEST_STUB_XLA_FFI_H
#include <cstddef>
#include <cstdint>
#include <string>
#include <type_traits>

namespace xla {
namespace ffi {

enum class DataType { F64 };
inline constexpr DataType F64 = DataType::F64;

template <DataType dt>
class Buffer {
 public:
  double *typed_data() const { return nullptr; }
  size_t element_count() const { return 0; }
  size_t size_bytes() const { return 0; }
};
template <typename T>
class Result {
  T value_;

 public:
  T *operator->() { return &value_; }
};
template <DataType dt>
using ResultBuffer = Result<Buffer<dt>>;

template <typename T>
struct PlatformStream {
  using type = T;
};

class Error {
  bool ok_;

 public:
  explicit Error(bool ok) : ok_(ok) {}
  static Error Success() { return Error(true); }
  static Error Internal(const std::string &) { return Error(false); }
  static Error InvalidArgument(const std::string &) { return Error(false); }
  bool failure() const { return !ok_; }
  bool success() const { return ok_; }
};

template <typename... Ts>
struct Binding {
  template <typename C>
  Binding<Ts..., typename C::type> Ctx() const { return {}; }
  template <typename T>
  Binding<Ts..., T> Attr(const char *) const { return {}; }
  template <typename T>
  Binding<Ts..., T> Arg() const { return {}; }
  template <typename T>
  Binding<Ts..., Result<T>> Ret() const { return {}; }
  template <typename Fn>
  int To(Fn) const {
    static_assert(std::is_invocable_r_v<Error, Fn, Ts...>, "handler signature does not match its Bind() chain");
    return 0;
  }
};
struct Ffi {
  static Binding<> Bind() { return {}; }
};

}  // namespace ffi
}  // namespace xla

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, impl, binding) \
  extern "C" void *name(void *) {                          \
    static int bound = (binding).To(impl);                 \
    (void)bound;                                           \
    return nullptr;                                        \
  }
#endif
