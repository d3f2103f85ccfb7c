// Minimal stand-in of jaxlib's xla/ffi/api/ffi.h -- TEST INFRASTRUCTURE ONLY (tests/test_abi_cpu.py).
//
// jaxlib (and with it the real header) is not installable in this image, so e_alphazero_b200/csrc/xla_ffi_shim.cc cannot be built
// against the real typed-FFI API here.  This file declares the slice of that API the shim uses -- Buffer / AnyBuffer / Result /
// RemainingArgs / RemainingRets / Error / ErrorOr / PlatformStream / Ffi::Bind() / XLA_FFI_DEFINE_HANDLER_SYMBOL -- with the same
// names and call shapes, backed by a trivial in-process call frame, so that (a) the shim is compile-checked on every CPU test run
// and (b) its handlers can be driven from a test without XLA.  It is NOT a re-implementation of XLA's FFI and never ships.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>

namespace xla {
namespace ffi {

enum DataType { PRED = 1, S8, S16, S32, S64, U8, U16, U32, U64, F16, F32, F64, BF16 };
enum class ErrorCode { kOk = 0, kCancelled, kUnknown, kInvalidArgument, kDeadlineExceeded, kNotFound, kAlreadyExists, kPermissionDenied,
                       kResourceExhausted, kFailedPrecondition, kAborted, kOutOfRange, kUnimplemented, kInternal, kUnavailable, kDataLoss };

class Error {
 public:
  Error() = default;
  Error(ErrorCode code, std::string message) : code_(code), message_(std::move(message)) {}
  static Error Success() { return Error(); }
  static Error InvalidArgument(std::string m) { return Error(ErrorCode::kInvalidArgument, std::move(m)); }
  static Error Internal(std::string m) { return Error(ErrorCode::kInternal, std::move(m)); }
  bool success() const { return code_ == ErrorCode::kOk; }
  bool failure() const { return !success(); }
  ErrorCode errc() const { return code_; }
  const std::string& message() const { return message_; }

 private:
  ErrorCode code_ = ErrorCode::kOk;
  std::string message_;
};

template <typename T>
class ErrorOr {
 public:
  ErrorOr(T v) : has_(true), value_(std::move(v)) {}  // NOLINT
  ErrorOr(Error e) : has_(false), error_(std::move(e)) {}  // NOLINT
  bool has_value() const { return has_; }
  T& value() { return value_; }
  T* operator->() { return &value_; }
  T& operator*() { return value_; }
  const Error& error() const { return error_; }

 private:
  bool has_;
  T value_{};
  Error error_;
};

template <typename T>
class Span {
 public:
  Span() = default;
  Span(const T* d, size_t n) : d_(d), n_(n) {}
  size_t size() const { return n_; }
  const T& operator[](size_t i) const { return d_[i]; }
  const T* begin() const { return d_; }
  const T* end() const { return d_ + n_; }

 private:
  const T* d_ = nullptr;
  size_t n_ = 0;
};

inline size_t ByteWidth(DataType t) {
  switch (t) {
    case PRED: case S8: case U8: return 1;
    case S16: case U16: case F16: case BF16: return 2;
    case S32: case U32: case F32: return 4;
    default: return 8;
  }
}
template <DataType dt> struct NativeOf { using type = uint8_t; };
template <> struct NativeOf<PRED> { using type = bool; };
template <> struct NativeOf<S32> { using type = int32_t; };
template <> struct NativeOf<U32> { using type = uint32_t; };
template <> struct NativeOf<F32> { using type = float; };
template <> struct NativeOf<U8> { using type = uint8_t; };

// what a test (or XLA) hands over for one array
struct RawBuffer {
  DataType dtype = U8;
  void* data = nullptr;
  std::vector<int64_t> dims;
};

class AnyBuffer {
 public:
  AnyBuffer() = default;
  explicit AnyBuffer(const RawBuffer* r) : r_(r) {}
  DataType element_type() const { return r_->dtype; }
  void* untyped_data() const { return r_->data; }
  template <typename T> T* typed_data() const { return static_cast<T*>(r_->data); }
  Span<int64_t> dimensions() const { return Span<int64_t>(r_->dims.data(), r_->dims.size()); }
  size_t element_count() const {
    size_t n = 1;
    for (int64_t d : r_->dims) n *= (size_t)d;
    return n;
  }
  size_t size_bytes() const { return element_count() * ByteWidth(r_->dtype); }

 protected:
  const RawBuffer* r_ = nullptr;
};

template <DataType dt>
class Buffer : public AnyBuffer {
 public:
  using AnyBuffer::AnyBuffer;
  typename NativeOf<dt>::type* typed_data() const { return static_cast<typename NativeOf<dt>::type*>(r_->data); }
  static constexpr DataType kType = dt;
};

template <typename T>
class Result {
 public:
  Result() = default;
  explicit Result(T v) : v_(v) {}
  T* operator->() { return &v_; }
  T& operator*() { return v_; }

 private:
  T v_;
};
template <DataType dt>
using ResultBuffer = Result<Buffer<dt>>;

template <typename T> struct IsTyped : std::false_type {};
template <DataType dt> struct IsTyped<Buffer<dt>> : std::true_type {};
template <typename T>
ErrorOr<T> Decode(const RawBuffer* r) {
  if constexpr (IsTyped<T>::value) {
    if (r->dtype != T::kType) return Error(ErrorCode::kInvalidArgument, "buffer dtype mismatch");
  }
  return T(r);
}

class RemainingArgs {
 public:
  RemainingArgs() = default;
  RemainingArgs(const RawBuffer* const* b, size_t n) : b_(b), n_(n) {}
  size_t size() const { return n_; }
  bool empty() const { return n_ == 0; }
  template <typename T>
  ErrorOr<T> get(size_t i) const {
    if (i >= n_) return Error(ErrorCode::kInvalidArgument, "argument index out of range");
    return Decode<T>(b_[i]);
  }

 private:
  const RawBuffer* const* b_ = nullptr;
  size_t n_ = 0;
};
class RemainingRets {
 public:
  RemainingRets() = default;
  RemainingRets(const RawBuffer* const* b, size_t n) : b_(b), n_(n) {}
  size_t size() const { return n_; }
  bool empty() const { return n_ == 0; }
  template <typename T>
  ErrorOr<Result<T>> get(size_t i) const {
    if (i >= n_) return Error(ErrorCode::kInvalidArgument, "result index out of range");
    auto d = Decode<T>(b_[i]);
    if (!d.has_value()) return d.error();
    return Result<T>(*d);
  }

 private:
  const RawBuffer* const* b_ = nullptr;
  size_t n_ = 0;
};

template <typename T>
struct PlatformStream {};

// ---- the in-process call frame of this stand-in
struct AttrValue {
  std::string name;
  bool is_float = false;
  int64_t i = 0;
  double f = 0;
};
struct CallFrame {
  void* stream = nullptr;
  std::vector<const RawBuffer*> args, rets;
  std::vector<AttrValue> attrs;
};

namespace internal {
struct CtxTag {};
template <typename T> struct ArgTag {};
template <typename T> struct RetTag {};
struct RemainingArgsTag {};
struct RemainingRetsTag {};
template <typename T> struct AttrTag { std::string name; };

struct DecodeState {
  const CallFrame* f;
  size_t arg = 0, ret = 0;
  Error err;
};
template <typename Tag> struct Decoder;
template <> struct Decoder<CtxTag> {
  static void* run(DecodeState& s, const CtxTag&) { return s.f->stream; }
};
template <typename T> struct Decoder<ArgTag<T>> {
  static T run(DecodeState& s, const ArgTag<T>&) {
    if (s.arg >= s.f->args.size()) { s.err = Error(ErrorCode::kInvalidArgument, "too few arguments"); return T(); }
    auto d = Decode<T>(s.f->args[s.arg++]);
    if (!d.has_value()) { s.err = d.error(); return T(); }
    return *d;
  }
};
template <typename T> struct Decoder<RetTag<T>> {
  static Result<T> run(DecodeState& s, const RetTag<T>&) {
    if (s.ret >= s.f->rets.size()) { s.err = Error(ErrorCode::kInvalidArgument, "too few results"); return Result<T>(); }
    auto d = Decode<T>(s.f->rets[s.ret++]);
    if (!d.has_value()) { s.err = d.error(); return Result<T>(); }
    return Result<T>(*d);
  }
};
template <> struct Decoder<RemainingArgsTag> {
  static RemainingArgs run(DecodeState& s, const RemainingArgsTag&) {
    RemainingArgs r(s.f->args.data() + s.arg, s.f->args.size() - s.arg);
    s.arg = s.f->args.size();
    return r;
  }
};
template <> struct Decoder<RemainingRetsTag> {
  static RemainingRets run(DecodeState& s, const RemainingRetsTag&) {
    RemainingRets r(s.f->rets.data() + s.ret, s.f->rets.size() - s.ret);
    s.ret = s.f->rets.size();
    return r;
  }
};
template <typename T> struct Decoder<AttrTag<T>> {
  static T run(DecodeState& s, const AttrTag<T>& t) {
    for (const AttrValue& a : s.f->attrs)
      if (a.name == t.name) {
        if (std::is_floating_point<T>::value != a.is_float) { s.err = Error(ErrorCode::kInvalidArgument, "attribute type mismatch: " + t.name); return T(); }
        return a.is_float ? (T)a.f : (T)a.i;
      }
    s.err = Error(ErrorCode::kInvalidArgument, "missing attribute: " + t.name);
    return T();
  }
};
}  // namespace internal

template <typename... Tags>
class Binding {
 public:
  explicit Binding(std::tuple<Tags...> t) : tags_(std::move(t)) {}
  template <typename T> auto Ctx() && { return Append(internal::CtxTag{}); }
  template <typename T> auto Arg() && { return Append(internal::ArgTag<T>{}); }
  template <typename T> auto Ret() && { return Append(internal::RetTag<T>{}); }
  auto RemainingArgs() && { return Append(internal::RemainingArgsTag{}); }
  auto RemainingRets() && { return Append(internal::RemainingRetsTag{}); }
  template <typename T> auto Attr(std::string name) && { return Append(internal::AttrTag<T>{std::move(name)}); }

  template <typename Fn>
  Error Call(Fn fn, const CallFrame* frame) const {
    internal::DecodeState s{frame};
    return CallImpl(fn, s, std::index_sequence_for<Tags...>{});
  }

 private:
  template <typename Tag>
  auto Append(Tag t) { return Binding<Tags..., Tag>(std::tuple_cat(std::move(tags_), std::make_tuple(std::move(t)))); }
  template <typename Fn, size_t... I>
  Error CallImpl(Fn fn, internal::DecodeState& s, std::index_sequence<I...>) const {
    // braced initialisation: decoding runs left to right, as the argument order requires
    std::tuple<decltype(internal::Decoder<Tags>::run(s, std::get<I>(tags_)))...> decoded{internal::Decoder<Tags>::run(s, std::get<I>(tags_))...};
    if (s.err.failure()) return s.err;
    return std::apply([&](auto&... a) { return fn(Cast<I>(a)...); }, decoded);
  }
  template <size_t I, typename A>
  static decltype(auto) Cast(A& a) {
    using Tag = std::tuple_element_t<I, std::tuple<Tags...>>;
    if constexpr (std::is_same<Tag, internal::CtxTag>::value) return static_cast<struct CUstream_st*>(a);
    else return (a);
  }
  std::tuple<Tags...> tags_;
};

struct Ffi {
  static Binding<> Bind() { return Binding<>(std::tuple<>()); }
};

}  // namespace ffi
}  // namespace xla

// The real macro defines `extern "C" XLA_FFI_Error* name(XLA_FFI_CallFrame*)`; here: `extern "C" int name(const CallFrame*, std::string* message)`
#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, impl, binding)                                       \
  extern "C" int name(const ::xla::ffi::CallFrame* frame, std::string* message) {                \
    static const auto* kBinding = new auto(binding);                                             \
    ::xla::ffi::Error e = kBinding->Call(impl, frame);                                           \
    if (message) *message = e.message();                                                         \
    return static_cast<int>(e.errc());                                                           \
  }
