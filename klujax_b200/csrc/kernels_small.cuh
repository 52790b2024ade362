// kernels_small.cuh — shared-memory regime (K1+K2+K4+K5+K7 fused): one warp per
// system keeps the system's whole value array (L, U, pivots, BTF off-diagonal
// entries, right-hand sides) in shared memory and interprets the packed static
// schedule (plan.hpp: PackedProgram).  sm_100a, FP64 CUDA cores.
//
// Replaces, per system, the loop body of /root/reference/klujax.cpp:431-451
// (gather Bx, klu_factor, klu_solve, klu_free_numeric) and its siblings
// (:560-580 tsolve, :672-691 factor, :770-787 refactor, :897-925
// refactor_and_solve, :1081-1092 / :1213-1224 solves with a numeric handle).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "plan.hpp"

namespace kb {

// ---- scalar traits: double, or complex stored as double2 (16-byte LDS/LDG) ----
struct zc {
    double x, y;
};
template <typename T>
struct Sc;
template <>
struct Sc<double> {
    using V = double;
    __device__ static __forceinline__ V zero() { return 0.0; }
    __device__ static __forceinline__ V ld(const V* p) { return *p; }
    __device__ static __forceinline__ void st(V* p, V v) { *p = v; }
    __device__ static __forceinline__ void fma_acc(V& acc, V a, V b) { acc = fma(a, b, acc); }
    __device__ static __forceinline__ V sub(V a, V b) { return a - b; }
    __device__ static __forceinline__ V add(V a, V b) { return a + b; }
    __device__ static __forceinline__ V mul(V a, V b) { return a * b; }
    __device__ static __forceinline__ V recip(V a) { return 1.0 / a; }
    __device__ static __forceinline__ V rdiv(V a, double s) { return a / s; }
    __device__ static __forceinline__ double mag(V a) { return fabs(a); }
    __device__ static __forceinline__ double mag2(V a) { return a * a; }
    __device__ static __forceinline__ bool bad_pivot(V a) { return a == 0.0 || !isfinite(a); }
    __device__ static __forceinline__ V shfl_down(V a, int o, int w) {
        return __shfl_down_sync(0xffffffffu, a, o, w);
    }
};
template <>
struct Sc<zc> {
    using V = double2;
    __device__ static __forceinline__ V zero() { return make_double2(0.0, 0.0); }
    __device__ static __forceinline__ V ld(const V* p) { return *p; }
    __device__ static __forceinline__ void st(V* p, V v) { *p = v; }
    __device__ static __forceinline__ void fma_acc(V& acc, V a, V b) {
        acc.x = fma(a.x, b.x, acc.x);
        acc.y = fma(a.x, b.y, acc.y);
        acc.x = fma(-a.y, b.y, acc.x);
        acc.y = fma(a.y, b.x, acc.y);
    }
    __device__ static __forceinline__ V sub(V a, V b) { return make_double2(a.x - b.x, a.y - b.y); }
    __device__ static __forceinline__ V add(V a, V b) { return make_double2(a.x + b.x, a.y + b.y); }
    __device__ static __forceinline__ V mul(V a, V b) {
        return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
    }
    // Smith's form: no overflow in |a|^2
    __device__ static __forceinline__ V recip(V a) {
        if (fabs(a.x) >= fabs(a.y)) {
            const double r = a.y / a.x, den = a.x + r * a.y;
            return make_double2(1.0 / den, -r / den);
        }
        const double r = a.x / a.y, den = r * a.x + a.y;
        return make_double2(r / den, -1.0 / den);
    }
    __device__ static __forceinline__ V rdiv(V a, double s) { return make_double2(a.x / s, a.y / s); }
    __device__ static __forceinline__ double mag(V a) { return hypot(a.x, a.y); }
    __device__ static __forceinline__ double mag2(V a) { return a.x * a.x + a.y * a.y; }
    __device__ static __forceinline__ bool bad_pivot(V a) {
        return (a.x == 0.0 && a.y == 0.0) || !isfinite(a.x) || !isfinite(a.y);
    }
    __device__ static __forceinline__ V shfl_down(V a, int o, int w) {
        return make_double2(__shfl_down_sync(0xffffffffu, a.x, o, w),
                            __shfl_down_sync(0xffffffffu, a.y, o, w));
    }
};

enum SmallFlags : int {
    SF_FACTOR = 1,         // scatter Ax, row-scale, run the refactor part of the program
    SF_SOLVE = 2,          // load b, run the solve part, store x
    SF_STORE_NUMERIC = 4,  // write V (LU) and Rs behind the numeric handle
    SF_LOAD_NUMERIC = 8,   // read V (LU) and Rs from the numeric handle
    SF_SCALE_IN = 16,      // X = b[perm] / Rs[perm]       (klu_solve)
    SF_SCALE_OUT = 32,     // x[perm] = X / Rs[perm]       (klu_tsolve)
};

struct SmallArgs {
    // packed programs: [0] first right-hand-side chunk (may include the refactor),
    // [1] further chunks (solve only)
    const uint32_t* ctrl[2];
    const uint32_t* stream[2];
    int nbatch[2];
    int n, n_val, n_slots, zero_slot, rc, nz, flags;
    const int* coo_dst;      // [nz] slot of each COO entry of this call, -1 = not in pattern
    const int* pattern_bad;  // device flag set by the COO map
    const int* rs_ptr;       // [n+1] row lists (original row order)
    const int* rs_slot;
    const int* in_perm;      // [n]
    const int* out_perm;     // [n]
    const void* Ax;
    long long ax_stride;     // elements between systems (0 = broadcast)
    const void* b;
    long long b_stride;
    void* x;
    int r;
    void* Vstore;            // [L_store][n_val]
    double* Rsstore;         // [L_store][n]
    const int* sysidx;       // optional: system m uses numeric element sysidx[m]
    int* status;
    double* growth;
    int L;
};

// One batch of the packed program: `run` operand lines, an optional lane-group
// reduction, then every leading lane finishes its task.
template <typename T>
__device__ __forceinline__ void run_program(typename Sc<T>::V* __restrict__ V,
                                            const uint32_t* __restrict__ ctrl,
                                            const uint32_t* __restrict__ stream, int nbatch,
                                            int lane, bool& bad, double& gmax) {
    using S = Sc<T>;
    using V_t = typename S::V;
    const uint32_t* sp = stream + lane;
    uint32_t c_next = nbatch > 0 ? __ldg(ctrl) : 0u;
    uint32_t h_next = nbatch > 0 ? __ldg(sp) : 0u;
    for (int bi = 0; bi < nbatch; ++bi) {
        const uint32_t c = c_next, h = h_next;
        const int run = c & 0xFFF, lg = (c >> 12) & 7;
        const uint32_t* ops = sp + 32;
        sp = ops + 32 * run;
        if (bi + 1 < nbatch) {  // prefetch the next batch header
            c_next = __ldg(ctrl + bi + 1);
            h_next = __ldg(sp);
        }
        const int kind = h >> 28, out = (h >> 14) & 0x3FFF, dv = h & 0x3FFF;
        V_t vo = S::zero(), vd = S::zero();
        if (kind) vo = S::ld(V + out);
        if (kind >= 3) vd = S::ld(V + dv);
        V_t acc = S::zero();
        int q = 0;
        for (; q + 4 <= run; q += 4) {
            const uint32_t w0 = __ldg(ops + 32 * q), w1 = __ldg(ops + 32 * (q + 1)),
                           w2 = __ldg(ops + 32 * (q + 2)), w3 = __ldg(ops + 32 * (q + 3));
            const V_t a0 = S::ld(V + (w0 >> 16)), b0 = S::ld(V + (w0 & 0xFFFF));
            const V_t a1 = S::ld(V + (w1 >> 16)), b1 = S::ld(V + (w1 & 0xFFFF));
            const V_t a2 = S::ld(V + (w2 >> 16)), b2 = S::ld(V + (w2 & 0xFFFF));
            const V_t a3 = S::ld(V + (w3 >> 16)), b3 = S::ld(V + (w3 & 0xFFFF));
            S::fma_acc(acc, a0, b0);
            S::fma_acc(acc, a1, b1);
            S::fma_acc(acc, a2, b2);
            S::fma_acc(acc, a3, b3);
        }
        for (; q < run; ++q) {
            const uint32_t w = __ldg(ops + 32 * q);
            S::fma_acc(acc, S::ld(V + (w >> 16)), S::ld(V + (w & 0xFFFF)));
        }
        if (lg) {
            const int g = 1 << lg;
            for (int o = g >> 1; o > 0; o >>= 1) acc = S::add(acc, S::shfl_down(acc, o, g));
        }
        if (kind) {
            V_t t = S::sub(vo, acc);
            if (kind == TK_RECIP) {
                bad |= S::bad_pivot(t);
                t = S::recip(t);
            } else if (kind >= TK_MUL) {
                t = S::mul(t, vd);
                if (kind == TK_MUL_TRACK) gmax = fmax(gmax, S::mag2(t));
            }
            S::st(V + out, t);
        }
        if (c & 0x8000u) __syncwarp();
    }
}

template <typename T>
__global__ void __launch_bounds__(32) small_kernel(const SmallArgs a) {
    using S = Sc<T>;
    using V_t = typename S::V;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    V_t* V = reinterpret_cast<V_t*>(smem_raw);
    double* Rs = reinterpret_cast<double*>(V + a.n_slots);
    const int lane = threadIdx.x;
    const V_t* Ax = static_cast<const V_t*>(a.Ax);
    const V_t* B = static_cast<const V_t*>(a.b);
    V_t* Xo = static_cast<V_t*>(a.x);
    V_t* Vst = static_cast<V_t*>(a.Vstore);
    const int n = a.n, rc = a.rc, r = a.r;

    for (int m = blockIdx.x; m < a.L; m += gridDim.x) {
        const int sm = a.sysidx ? a.sysidx[m] : m;
        bool bad = false;
        double gmax = 0.0;
        if (a.flags & SF_FACTOR) {
            for (int s = lane; s < a.n_val; s += 32) S::st(V + s, S::zero());
            __syncwarp();
            const V_t* Am = Ax + (long long)m * a.ax_stride;
            for (int e = lane; e < a.nz; e += 32) {
                const int d = __ldg(a.coo_dst + e);
                if (d >= 0) S::st(V + d, Am[e]);
            }
            __syncwarp();
            // max-abs row scaling (KLU scale = 2): Rs in original row order
            for (int i = lane; i < n; i += 32) {
                const int pb = __ldg(a.rs_ptr + i), pe = __ldg(a.rs_ptr + i + 1);
                double mx = 0.0;
                for (int p = pb; p < pe; ++p) mx = fmax(mx, S::mag(S::ld(V + __ldg(a.rs_slot + p))));
                if (mx == 0.0) mx = 1.0;
                Rs[i] = mx;
                for (int p = pb; p < pe; ++p) {
                    V_t* q = V + __ldg(a.rs_slot + p);
                    S::st(q, S::rdiv(S::ld(q), mx));
                }
            }
        } else if (a.flags & SF_LOAD_NUMERIC) {
            const V_t* src = Vst + (long long)sm * a.n_val;
            for (int s = lane; s < a.n_val; s += 32) S::st(V + s, src[s]);
            const double* rsrc = a.Rsstore + (long long)sm * n;
            for (int i = lane; i < n; i += 32) Rs[i] = rsrc[i];
        }
        if (lane == 0) S::st(V + a.zero_slot, S::zero());
        __syncwarp();

        const int nchunk = (a.flags & SF_SOLVE) ? (r + rc - 1) / rc : 1;
        for (int ch = 0; ch < nchunk; ++ch) {
            const int c0 = ch * rc;
            if (a.flags & SF_SOLVE) {
                const V_t* Bm = B + (long long)m * a.b_stride;
                for (int idx = lane; idx < n * rc; idx += 32) {
                    const int k = idx / rc, c = c0 + idx - k * rc;
                    V_t v = S::zero();
                    if (c < r) {
                        const int src = __ldg(a.in_perm + k);
                        v = Bm[(long long)src * r + c];
                        if (a.flags & SF_SCALE_IN) v = S::rdiv(v, Rs[src]);
                    }
                    S::st(V + a.n_val + idx, v);
                }
                __syncwarp();
            }
            const int pi = ch == 0 ? 0 : 1;
            run_program<T>(V, a.ctrl[pi], a.stream[pi], a.nbatch[pi], lane, bad, gmax);
            __syncwarp();
            if (a.flags & SF_SOLVE) {
                V_t* Xm = Xo + (long long)m * n * r;
                for (int idx = lane; idx < n * rc; idx += 32) {
                    const int k = idx / rc, c = c0 + idx - k * rc;
                    if (c < r) {
                        const int dst = __ldg(a.out_perm + k);
                        V_t v = S::ld(V + a.n_val + idx);
                        if (a.flags & SF_SCALE_OUT) v = S::rdiv(v, Rs[dst]);
                        Xm[(long long)dst * r + c] = v;
                    }
                }
            }
        }
        if (a.flags & SF_STORE_NUMERIC) {
            V_t* dst = Vst + (long long)sm * a.n_val;
            for (int s = lane; s < a.n_val; s += 32) dst[s] = S::ld(V + s);
            double* rdst = a.Rsstore + (long long)sm * n;
            for (int i = lane; i < n; i += 32) rdst[i] = Rs[i];
        }
        if (a.flags & SF_FACTOR) {
            const bool anybad = __any_sync(0xffffffffu, bad);
            for (int o = 16; o > 0; o >>= 1) gmax = fmax(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
            if (lane == 0) {
                int st = anybad ? 1 : 0;
                if (a.pattern_bad && *a.pattern_bad) st = -3;
                if (a.status) a.status[m] = st;
                if (a.growth) a.growth[m] = sqrt(gmax);
            }
        }
        __syncwarp();
    }
}

// COO (this call's order) -> value slot, by binary search in the analysed
// pattern; also the bounds validation the reference does on the host on every
// call (/root/reference/klujax.cpp:66-82).
__global__ void coo_map_kernel(int nz, int n, const int* __restrict__ Ai,
                               const int* __restrict__ Aj, const int* __restrict__ colptr,
                               const int* __restrict__ srow, const int* __restrict__ sslot,
                               int* __restrict__ dst, int* __restrict__ bad) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nz) return;
    const int i = Ai[e], j = Aj[e];
    int d = -1;
    if (i >= 0 && i < n && j >= 0 && j < n) {
        int lo = colptr[j], hi = colptr[j + 1] - 1;
        while (lo <= hi) {
            const int mid = (lo + hi) >> 1;
            const int v = srow[mid];
            if (v == i) {
                d = sslot[mid];
                break;
            }
            if (v < i)
                lo = mid + 1;
            else
                hi = mid - 1;
        }
    }
    dst[e] = d;
    if (d < 0) atomicExch(bad, 1);
}

}  // namespace kb
