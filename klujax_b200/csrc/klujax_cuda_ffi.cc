// klujax_cuda_ffi.cc — XLA typed-FFI handlers for platform "CUDA": the twins of the 21 CPU handlers
// of /root/reference/klujax.cpp (handler symbols :219-1383, capsule module :1386-1429), each a
// thin call into the C-ABI of include/klujax_cuda.h.  Same target names, same operand order, same
// result shapes as the CPU handlers, so that klujax.py only has to register them with
// platform="CUDA" (patches/klujax_py_cuda.patch).
//
// Compile-gated: the XLA FFI headers (and pybind11 for the capsule module) are not in this image,
// so here the file compiles to the availability stub at the bottom; where
// "xla/ffi/api/ffi.h" is on the include path it compiles to the real handlers.  Build line for a
// klujax checkout (setup.py adds it next to klujax.cpp):
//   g++ -std=c++17 -fPIC -shared -I<xla>/ -I<pybind11>/include -I/usr/local/cuda/include \
//       klujax_cuda_ffi.cc -L. -lklujax_cuda -lcudart -o klujax_cuda_cpp$(python3-config --extension-suffix)
//
// What differs from the CPU handlers, and why:
//  * every handler takes the platform stream and only enqueues; the exceptions are stated per
//    handler (analyze and the one-shot solve read Ai/Aj on the host; handles that arrive as
//    device buffers are read back: 8 bytes per handle + one stream synchronisation);
//  * errors: the CPU handlers return InvalidArgument for the first failing system
//    (klujax.cpp:311-314, :782).  The `_ex` entry points reproduce that (they synchronise, re-seed
//    the pivot sequence for systems that need it, and report the first failure).
#include "../../include/klujax_cuda.h"

#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define KLUJAX_HAVE_XLA_FFI 1
#endif
#endif

#ifdef KLUJAX_HAVE_XLA_FFI
#include <cuda_runtime_api.h>

#include <string>
#include <vector>

#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;
using S32 = ffi::Buffer<ffi::DataType::S32>;
using U64 = ffi::Buffer<ffi::DataType::U64>;
using F64 = ffi::Buffer<ffi::DataType::F64>;
using C128 = ffi::Buffer<ffi::DataType::C128>;

namespace {

ffi::Error status_of(int rc, const char* what) {
    if (rc == 0) return ffi::Error::Success();
    return ffi::Error::InvalidArgument(std::string(what) + ": " + klujax_cuda_last_error());
}
// a handle that arrives as a device buffer (it is a traced jit argument, klujax.py:354-357)
ffi::Error fetch_u64(cudaStream_t st, const U64& buf, std::vector<uint64_t>& out) {
    out.resize(buf.element_count());
    if (out.empty()) return ffi::Error::Success();
    if (cudaMemcpyAsync(out.data(), buf.typed_data(), out.size() * sizeof(uint64_t), cudaMemcpyDeviceToHost, st) !=
            cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess)
        return ffi::Error::Internal("reading a handle buffer back failed");
    return ffi::Error::Success();
}
ffi::Error fetch_i32(cudaStream_t st, const S32& buf, std::vector<int32_t>& out) {
    out.resize(buf.element_count());
    if (out.empty()) return ffi::Error::Success();
    if (cudaMemcpyAsync(out.data(), buf.typed_data(), out.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, st) !=
            cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess)
        return ffi::Error::Internal("reading an index buffer back failed");
    return ffi::Error::Success();
}
template <typename B>
const double* dptr(const B& b) {
    return reinterpret_cast<const double*>(b.typed_data());
}
template <typename B>
double* dptr(ffi::Result<B>& b) {
    return reinterpret_cast<double*>(b->typed_data());
}
// canonical shapes (klujax.py:1748-1849 produces them before binding)
struct Dims {
    int n_lhs, n_nz, n_col, n_rhs, n_lhs_A, n_lhs_b;
};
template <typename BA, typename BB>
ffi::Error canonical(const BA& Ax, const BB& b, Dims& d) {
    auto da = Ax.dimensions();
    auto db = b.dimensions();
    if (da.size() != 2) return ffi::Error::InvalidArgument("Ax is not 2D");
    if (db.size() != 3) return ffi::Error::InvalidArgument("b is not 3D");
    d.n_lhs_A = static_cast<int>(da[0]);
    d.n_nz = static_cast<int>(da[1]);
    d.n_lhs_b = static_cast<int>(db[0]);
    d.n_col = static_cast<int>(db[1]);
    d.n_rhs = static_cast<int>(db[2]);
    d.n_lhs = d.n_lhs_A > d.n_lhs_b ? d.n_lhs_A : d.n_lhs_b;
    return ffi::Error::Success();
}

// per-call scratch for the per-system status / growth arrays (XLA's allocator: no cudaMalloc)
struct Scratch {
    int32_t* status = nullptr;
    double* growth = nullptr;
};
ffi::Error get_scratch(ffi::ScratchAllocator& alloc, int n_lhs, Scratch& s) {
    auto a = alloc.Allocate(sizeof(int32_t) * static_cast<size_t>(n_lhs), 16);
    auto g = alloc.Allocate(sizeof(double) * static_cast<size_t>(n_lhs), 16);
    if (!a.has_value() || !g.has_value()) return ffi::Error::Internal("scratch allocation failed");
    s.status = static_cast<int32_t*>(*a);
    s.growth = static_cast<double*>(*g);
    return ffi::Error::Success();
}

}  // namespace

// ---- dot (klujax.cpp:165-249) -------------------------------------------------------------
template <typename B, int CX>
ffi::Error dot_cuda(cudaStream_t st, S32 Ai, S32 Aj, B Ax, B x, ffi::Result<B> b) {
    Dims d;
    if (auto e = canonical(Ax, x, d); e.failure()) return e;
    auto f = CX ? klujax_cuda_dot_c128 : klujax_cuda_dot_f64;
    return status_of(f(d.n_lhs, d.n_col, d.n_rhs, d.n_nz, Ai.typed_data(), Aj.typed_data(), dptr(Ax), dptr(x), dptr(b), st),
                     "dot");
}
// ---- solve (klujax.cpp:251-380): analyze + factor + solve + free; the analysis runs on the host,
// so Ai / Aj are read back (one synchronisation; the pattern cache of the C-ABI makes the rest of
// the host stage a lookup)
template <typename B, int CX>
ffi::Error solve_cuda(cudaStream_t st, ffi::ScratchAllocator alloc, S32 Ai, S32 Aj, B Ax, B b, ffi::Result<B> x) {
    Dims d;
    if (auto e = canonical(Ax, b, d); e.failure()) return e;
    std::vector<int32_t> hi, hj;
    if (auto e = fetch_i32(st, Ai, hi); e.failure()) return e;
    if (auto e = fetch_i32(st, Aj, hj); e.failure()) return e;
    Scratch s;
    if (auto e = get_scratch(alloc, d.n_lhs, s); e.failure()) return e;
    auto f = CX ? klujax_cuda_solve_c128 : klujax_cuda_solve_f64;
    int rc = f(d.n_lhs, d.n_col, d.n_rhs, d.n_nz, hi.data(), hj.data(), dptr(Ax), dptr(b), dptr(x), s.status, s.growth, st);
    if (rc) return status_of(rc, "klu_analyze / klu_factor");
    auto g = CX ? klujax_cuda_reseed_failed_c128 : klujax_cuda_reseed_failed_f64;
    int32_t fb = -1, nr = 0;
    rc = g(0, d.n_lhs, d.n_lhs, d.n_lhs, d.n_col, d.n_rhs, d.n_nz, Ai.typed_data(), Aj.typed_data(), dptr(Ax), dptr(b),
           dptr(x), s.status, s.growth, &fb, &nr, st);
    return status_of(rc, "klu_factor failed (singular matrix?)");
}
// ---- solve_with_symbol / tsolve_with_symbol (klujax.cpp:382-636)
template <typename B, int CX, int TR>
ffi::Error solve_with_symbol_cuda(cudaStream_t st, ffi::ScratchAllocator alloc, S32 Ai, S32 Aj, B Ax, B b, U64 symbolic,
                                  ffi::Result<B> x) {
    Dims d;
    if (auto e = canonical(Ax, b, d); e.failure()) return e;
    std::vector<uint64_t> h;
    if (auto e = fetch_u64(st, symbolic, h); e.failure()) return e;
    if (h.size() != 1 || h[0] == 0) return ffi::Error::InvalidArgument("symbolic pointer is null");
    Scratch s;
    if (auto e = get_scratch(alloc, d.n_lhs, s); e.failure()) return e;
    auto f = TR ? (CX ? klujax_cuda_tsolve_with_symbol_ex_c128 : klujax_cuda_tsolve_with_symbol_ex_f64)
                : (CX ? klujax_cuda_solve_with_symbol_ex_c128 : klujax_cuda_solve_with_symbol_ex_f64);
    int32_t fb = -1, nr = 0;
    return status_of(f(h[0], d.n_lhs, d.n_lhs_A, d.n_lhs_b, d.n_col, d.n_rhs, d.n_nz, Ai.typed_data(), Aj.typed_data(),
                       dptr(Ax), dptr(b), dptr(x), s.status, s.growth, /*check=*/1, &fb, &nr, st),
                     "klu_factor failed (singular matrix?)");
}
// ---- factor (klujax.cpp:638-731): numeric handles leave as a u64[n_lhs] device buffer
template <typename B, int CX>
ffi::Error factor_cuda(cudaStream_t st, ffi::ScratchAllocator alloc, S32 Ai, S32 Aj, B Ax, U64 symbolic,
                       ffi::Result<U64> numeric) {
    auto da = Ax.dimensions();
    if (da.size() != 2) return ffi::Error::InvalidArgument("Ax is not 2D");
    std::vector<uint64_t> h;
    if (auto e = fetch_u64(st, symbolic, h); e.failure()) return e;
    if (h.size() != 1 || h[0] == 0) return ffi::Error::InvalidArgument("symbolic pointer is null");
    Scratch s;
    if (auto e = get_scratch(alloc, static_cast<int>(da[0]), s); e.failure()) return e;
    auto f = CX ? klujax_cuda_factor_dev_c128 : klujax_cuda_factor_dev_f64;
    int32_t fb = -1;
    return status_of(f(h[0], static_cast<int>(da[0]), static_cast<int>(da[1]), Ai.typed_data(), Aj.typed_data(), dptr(Ax),
                       numeric->typed_data(), s.status, s.growth, /*check=*/1, &fb, st),
                     "klu_factor failed (singular matrix?)");
}
// ---- refactor / refactor_and_solve (klujax.cpp:733-993): output handles = input handles
template <typename B, int CX>
ffi::Error refactor_cuda(cudaStream_t st, ffi::ScratchAllocator alloc, S32 Ai, S32 Aj, B Ax, U64 symbolic, U64 numeric,
                         ffi::Result<U64> out_numeric) {
    auto da = Ax.dimensions();
    if (da.size() != 2) return ffi::Error::InvalidArgument("Ax is not 2D");
    std::vector<uint64_t> h;
    if (auto e = fetch_u64(st, symbolic, h); e.failure()) return e;
    if (h.size() != 1 || h[0] == 0) return ffi::Error::InvalidArgument("symbolic pointer is null");
    Scratch s;
    if (auto e = get_scratch(alloc, static_cast<int>(da[0]), s); e.failure()) return e;
    auto f = CX ? klujax_cuda_refactor_dev_c128 : klujax_cuda_refactor_dev_f64;
    int rc = f(h[0], static_cast<int>(da[0]), 0, 0, static_cast<int>(da[1]), Ai.typed_data(), Aj.typed_data(), dptr(Ax),
               nullptr, numeric.typed_data(), nullptr, out_numeric->typed_data(), s.status, s.growth, st);
    if (rc) return status_of(rc, "klu_refactor");
    // klu_refactor fails on an exactly zero pivot (klujax.cpp:782): first failing system aborts
    std::vector<int32_t> hs(da[0]);
    cudaMemcpyAsync(hs.data(), s.status, sizeof(int32_t) * hs.size(), cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    for (int32_t v : hs)
        if (v != 0) return ffi::Error::InvalidArgument("klu_refactor failed (singular matrix?)");
    return ffi::Error::Success();
}
template <typename B, int CX>
ffi::Error refactor_and_solve_cuda(cudaStream_t st, ffi::ScratchAllocator alloc, S32 Ai, S32 Aj, B Ax, B b, U64 symbolic,
                                   U64 numeric, ffi::Result<B> x, ffi::Result<U64> out_numeric) {
    Dims d;
    if (auto e = canonical(Ax, b, d); e.failure()) return e;
    std::vector<uint64_t> h;
    if (auto e = fetch_u64(st, symbolic, h); e.failure()) return e;
    if (h.size() != 1 || h[0] == 0) return ffi::Error::InvalidArgument("symbolic pointer is null");
    Scratch s;
    if (auto e = get_scratch(alloc, d.n_lhs, s); e.failure()) return e;
    auto f = CX ? klujax_cuda_refactor_and_solve_dev_c128 : klujax_cuda_refactor_and_solve_dev_f64;
    int rc = f(h[0], d.n_lhs, d.n_col, d.n_rhs, d.n_nz, Ai.typed_data(), Aj.typed_data(), dptr(Ax), dptr(b),
               numeric.typed_data(), dptr(x), out_numeric->typed_data(), s.status, s.growth, st);
    if (rc) return status_of(rc, "klu_refactor");
    std::vector<int32_t> hs(d.n_lhs);
    cudaMemcpyAsync(hs.data(), s.status, sizeof(int32_t) * hs.size(), cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    for (int32_t v : hs)
        if (v != 0) return ffi::Error::InvalidArgument("klu_refactor failed (singular matrix?)");
    return ffi::Error::Success();
}
// ---- solve_with_numeric / tsolve_with_numeric (klujax.cpp:995-1271): b is 1-D, 2-D or 3-D; the
// rank parsing and the broadcast decisions are the CPU handler's (klujax.cpp:1011-1059)
template <typename B, int CX, int TR>
ffi::Error solve_with_numeric_cuda(cudaStream_t st, U64 symbolic, U64 numeric, B b, ffi::Result<B> x) {
    std::vector<uint64_t> h;
    if (auto e = fetch_u64(st, symbolic, h); e.failure()) return e;
    if (h.size() != 1 || h[0] == 0) return ffi::Error::InvalidArgument("symbolic pointer is null");
    const int n_numeric = static_cast<int>(numeric.element_count());
    auto db = b.dimensions();
    int n_lhs_b = 1, n_col = 0, n_rhs = 1;
    if (db.size() == 1) {
        n_col = static_cast<int>(db[0]);
    } else if (db.size() == 2) {
        if (n_numeric > 1 && db[0] == n_numeric) {  // (n_lhs, n_col)
            n_lhs_b = static_cast<int>(db[0]);
            n_col = static_cast<int>(db[1]);
        } else {  // (n_col, n_rhs)
            n_col = static_cast<int>(db[0]);
            n_rhs = static_cast<int>(db[1]);
        }
    } else if (db.size() == 3) {
        n_lhs_b = static_cast<int>(db[0]);
        n_col = static_cast<int>(db[1]);
        n_rhs = static_cast<int>(db[2]);
    } else {
        return ffi::Error::InvalidArgument("b must be 1D, 2D or 3D");
    }
    auto f = TR ? (CX ? klujax_cuda_tsolve_with_numeric_dev_c128 : klujax_cuda_tsolve_with_numeric_dev_f64)
                : (CX ? klujax_cuda_solve_with_numeric_dev_c128 : klujax_cuda_solve_with_numeric_dev_f64);
    return status_of(f(h[0], n_numeric, numeric.typed_data(), n_lhs_b, n_col, n_rhs, dptr(b), dptr(x), st), "klu_solve");
}
// ---- free_numeric / free_symbolic (klujax.cpp:1273-1329)
ffi::Error free_numeric_cuda(cudaStream_t st, U64 numeric, ffi::Result<S32> status) {
    int32_t freed = 0;
    int rc = klujax_cuda_free_numeric_dev(static_cast<int>(numeric.element_count()), numeric.typed_data(), &freed, st);
    if (rc) return status_of(rc, "free_numeric");
    cudaMemcpyAsync(status->typed_data(), &freed, sizeof(int32_t), cudaMemcpyHostToDevice, st);
    cudaStreamSynchronize(st);  // `freed` lives on this stack frame
    return ffi::Error::Success();
}
ffi::Error free_symbolic_cuda(cudaStream_t st, U64 symbolic, ffi::Result<S32> status) {
    std::vector<uint64_t> h;
    if (auto e = fetch_u64(st, symbolic, h); e.failure()) return e;
    int32_t freed = 0;
    cudaStreamSynchronize(st);  // work that still uses the handle's device schedules
    int rc = klujax_cuda_free_symbolic(h.empty() ? 0 : h[0], &freed);
    if (rc) return status_of(rc, "free_symbolic");
    cudaMemcpyAsync(status->typed_data(), &freed, sizeof(int32_t), cudaMemcpyHostToDevice, st);
    cudaStreamSynchronize(st);
    return ffi::Error::Success();
}
// ---- analyze (klujax.cpp:1331-1383): host stage; Ai, Aj, n_col are read back
ffi::Error analyze_cuda(cudaStream_t st, S32 Ai, S32 Aj, S32 n_col_buf, ffi::Result<U64> symbolic) {
    if (n_col_buf.element_count() != 1) return ffi::Error::InvalidArgument("n_col must be a scalar.");
    std::vector<int32_t> hi, hj, hn;
    if (auto e = fetch_i32(st, Ai, hi); e.failure()) return e;
    if (auto e = fetch_i32(st, Aj, hj); e.failure()) return e;
    if (auto e = fetch_i32(st, n_col_buf, hn); e.failure()) return e;
    uint64_t h = 0;
    int rc = klujax_cuda_analyze(hn[0], static_cast<int>(hi.size()), hi.data(), hj.data(), &h);
    if (rc) return status_of(rc, "klu_analyze");
    cudaMemcpyAsync(symbolic->typed_data(), &h, sizeof(uint64_t), cudaMemcpyHostToDevice, st);
    cudaStreamSynchronize(st);
    return ffi::Error::Success();
}

// ---- handler symbols: same names as the CPU ones + _cuda ------------------------------------------
#define CTX ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
#define CTXS CTX.Ctx<ffi::ScratchAllocator>()
#define DEF2(name, fn, bind)                                                      \
    XLA_FFI_DEFINE_HANDLER_SYMBOL(name##_f64_cuda_handler, (fn<F64, 0>), bind(F64)) \
    XLA_FFI_DEFINE_HANDLER_SYMBOL(name##_c128_cuda_handler, (fn<C128, 1>), bind(C128))
#define BIND_DOT(T) CTX.Arg<S32>().Arg<S32>().Arg<T>().Arg<T>().Ret<T>()
#define BIND_SOLVE(T) CTXS.Arg<S32>().Arg<S32>().Arg<T>().Arg<T>().Ret<T>()
#define BIND_SWS(T) CTXS.Arg<S32>().Arg<S32>().Arg<T>().Arg<T>().Arg<U64>().Ret<T>()
#define BIND_FACTOR(T) CTXS.Arg<S32>().Arg<S32>().Arg<T>().Arg<U64>().Ret<U64>()
#define BIND_REFACTOR(T) CTXS.Arg<S32>().Arg<S32>().Arg<T>().Arg<U64>().Arg<U64>().Ret<U64>()
#define BIND_RAS(T) CTXS.Arg<S32>().Arg<S32>().Arg<T>().Arg<T>().Arg<U64>().Arg<U64>().Ret<T>().Ret<U64>()
#define BIND_SWN(T) CTX.Arg<U64>().Arg<U64>().Arg<T>().Ret<T>()
DEF2(dot, dot_cuda, BIND_DOT)
DEF2(solve, solve_cuda, BIND_SOLVE)
DEF2(factor, factor_cuda, BIND_FACTOR)
DEF2(refactor, refactor_cuda, BIND_REFACTOR)
DEF2(refactor_and_solve, refactor_and_solve_cuda, BIND_RAS)
XLA_FFI_DEFINE_HANDLER_SYMBOL(solve_with_symbol_f64_cuda_handler, (solve_with_symbol_cuda<F64, 0, 0>), BIND_SWS(F64))
XLA_FFI_DEFINE_HANDLER_SYMBOL(solve_with_symbol_c128_cuda_handler, (solve_with_symbol_cuda<C128, 1, 0>), BIND_SWS(C128))
XLA_FFI_DEFINE_HANDLER_SYMBOL(tsolve_with_symbol_f64_cuda_handler, (solve_with_symbol_cuda<F64, 0, 1>), BIND_SWS(F64))
XLA_FFI_DEFINE_HANDLER_SYMBOL(tsolve_with_symbol_c128_cuda_handler, (solve_with_symbol_cuda<C128, 1, 1>), BIND_SWS(C128))
XLA_FFI_DEFINE_HANDLER_SYMBOL(solve_with_numeric_f64_cuda_handler, (solve_with_numeric_cuda<F64, 0, 0>), BIND_SWN(F64))
XLA_FFI_DEFINE_HANDLER_SYMBOL(solve_with_numeric_c128_cuda_handler, (solve_with_numeric_cuda<C128, 1, 0>), BIND_SWN(C128))
XLA_FFI_DEFINE_HANDLER_SYMBOL(tsolve_with_numeric_f64_cuda_handler, (solve_with_numeric_cuda<F64, 0, 1>), BIND_SWN(F64))
XLA_FFI_DEFINE_HANDLER_SYMBOL(tsolve_with_numeric_c128_cuda_handler, (solve_with_numeric_cuda<C128, 1, 1>), BIND_SWN(C128))
XLA_FFI_DEFINE_HANDLER_SYMBOL(free_numeric_cuda_handler, free_numeric_cuda, CTX.Arg<U64>().Ret<S32>())
XLA_FFI_DEFINE_HANDLER_SYMBOL(free_symbolic_cuda_handler, free_symbolic_cuda, CTX.Arg<U64>().Ret<S32>())
XLA_FFI_DEFINE_HANDLER_SYMBOL(analyze_cuda_handler, analyze_cuda, CTX.Arg<S32>().Arg<S32>().Arg<S32>().Ret<U64>())

#if __has_include(<pybind11/pybind11.h>) && defined(KLUJAX_CUDA_PYMODULE)
// the capsule module of klujax.cpp:1386-1429, one zero-argument function per target
#include <pybind11/pybind11.h>
namespace py = pybind11;
#define CAP(name) m.def(#name, []() { return py::capsule((void*)&name##_cuda_handler); });
PYBIND11_MODULE(klujax_cuda_cpp, m) {
    CAP(dot_f64) CAP(dot_c128) CAP(solve_f64) CAP(solve_c128)
    CAP(solve_with_symbol_f64) CAP(solve_with_symbol_c128) CAP(tsolve_with_symbol_f64) CAP(tsolve_with_symbol_c128)
    CAP(factor_f64) CAP(factor_c128) CAP(refactor_f64) CAP(refactor_c128)
    CAP(refactor_and_solve_f64) CAP(refactor_and_solve_c128)
    CAP(solve_with_numeric_f64) CAP(solve_with_numeric_c128) CAP(tsolve_with_numeric_f64) CAP(tsolve_with_numeric_c128)
    CAP(free_numeric) CAP(free_symbolic) CAP(analyze)
}
#endif

extern "C" int klujax_cuda_ffi_available(void) { return 1; }
#else   // no XLA FFI headers on this machine: the C-ABI is the boundary, exercised through ctypes
extern "C" int klujax_cuda_ffi_available(void) { return 0; }
#endif
