// Minimal stand-in for xla/ffi/api/ffi.h (NOT the real header): just enough of the typed-FFI surface
// that klujax_b200/csrc/klujax_cuda_ffi.cc uses, so that the shim can be syntax- and type-checked in
// an image without XLA (tests/test_ffi_shim.py).  Names and shapes follow the XLA FFI API as used by
// /root/reference/klujax.cpp (Buffer, Result, Error, Ffi::Bind, XLA_FFI_DEFINE_HANDLER_SYMBOL).
#pragma once
#include <cstddef>
#include <cstdint>
#include <optional>
#include <string>
#include <vector>
namespace xla { namespace ffi {
enum class DataType { S32, U64, F64, C128 };
template <DataType> struct NativeOf;
template <> struct NativeOf<DataType::S32> { using type = int32_t; };
template <> struct NativeOf<DataType::U64> { using type = uint64_t; };
template <> struct NativeOf<DataType::F64> { using type = double; };
struct c128 { double re, im; };
template <> struct NativeOf<DataType::C128> { using type = c128; };
template <DataType D> struct Buffer {
    using T = typename NativeOf<D>::type;
    T* data = nullptr;
    std::vector<int64_t> dims;
    T* typed_data() const { return data; }
    const std::vector<int64_t>& dimensions() const { return dims; }
    size_t element_count() const { size_t n = 1; for (auto d : dims) n *= d; return n; }
};
template <typename B> struct Result { B b; B* operator->() { return &b; } };
struct Error {
    bool fail = false; std::string msg;
    static Error Success() { return {}; }
    static Error InvalidArgument(std::string m) { return {true, std::move(m)}; }
    static Error Internal(std::string m) { return {true, std::move(m)}; }
    bool failure() const { return fail; }
};
struct ScratchAllocator { std::optional<void*> Allocate(size_t, size_t = 1) { return std::nullopt; } };
template <typename S> struct PlatformStream {};
struct Binding {
    template <typename T> Binding Ctx() { return *this; }
    template <typename T> Binding Arg() { return *this; }
    template <typename T> Binding Ret() { return *this; }
    template <typename T> Binding Attr(const char*) { return *this; }
};
struct Ffi { static Binding Bind() { return {}; } };
}}  // namespace xla::ffi
#define XLA_FFI_CAT_(a, b) a##b
#define XLA_FFI_DEFINE_HANDLER_SYMBOL(sym, fn, binding) \
    static auto XLA_FFI_CAT_(sym, _binding) = (binding); \
    extern "C" void* sym = reinterpret_cast<void*>(+fn);
