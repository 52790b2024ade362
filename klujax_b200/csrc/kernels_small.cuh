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
    __device__ static __forceinline__ void fms_acc(V& acc, V a, V b) { acc = fma(-a, b, acc); }  // acc -= a * b
    __device__ static __forceinline__ V sub(V a, V b) { return a - b; }
    __device__ static __forceinline__ V add(V a, V b) { return a + b; }
    __device__ static __forceinline__ V mul(V a, V b) { return a * b; }
    __device__ static __forceinline__ V recip(V a) { return 1.0 / a; }
    __device__ static __forceinline__ V fast_recip(V a) { return __drcp_rn(a); }
    __device__ static __forceinline__ V rdiv(V a, double s) { return a / s; }
    __device__ static __forceinline__ double mag(V a) { return fabs(a); }
    __device__ static __forceinline__ double mag2(V a) { return a * a; }
    __device__ static __forceinline__ bool bad_pivot(V a) { return a == 0.0 || !isfinite(a); }
    __device__ static __forceinline__ V shfl_down(V a, int o, int w) {
        return __shfl_down_sync(0xffffffffu, a, o, w);
    }
    __device__ static __forceinline__ V one() { return 1.0; }
    __device__ static __forceinline__ V shfl_idx(V a, int src) { return __shfl_sync(0xffffffffu, a, src); }
    // split accumulators: independent FMA chains, combined once per batch
    struct Acc {
        double p, q;
    };
    __device__ static __forceinline__ Acc acc0() { return {0.0, 0.0}; }
    __device__ static __forceinline__ void mac(Acc& c, V a, V b, int which) {
        if (which & 1) c.q = fma(a, b, c.q); else c.p = fma(a, b, c.p);
    }
    __device__ static __forceinline__ V fold(const Acc& c) { return c.p + c.q; }
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
    __device__ static __forceinline__ void fms_acc(V& acc, V a, V b) {  // acc -= a * b
        acc.x = fma(-a.x, b.x, acc.x);
        acc.y = fma(-a.x, b.y, acc.y);
        acc.x = fma(a.y, b.y, acc.x);
        acc.y = fma(-a.y, b.x, acc.y);
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
    // conj(a) / |a|^2 with one reciprocal; Smith's form outside the range where |a|^2 is safe
    __device__ static __forceinline__ V fast_recip(V a) {
        const double m = fma(a.x, a.x, a.y * a.y);
        if (m > 1e-280 && m < 1e280) {
            const double inv = __drcp_rn(m);
            return make_double2(a.x * inv, -a.y * inv);
        }
        return recip(a);
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
    __device__ static __forceinline__ V one() { return make_double2(1.0, 0.0); }
    __device__ static __forceinline__ V shfl_idx(V a, int src) {
        return make_double2(__shfl_sync(0xffffffffu, a.x, src), __shfl_sync(0xffffffffu, a.y, src));
    }
    // four independent FMA chains (xx, yy, xy, yx), combined once per batch
    struct Acc {
        double xx, yy, xy, yx;
    };
    __device__ static __forceinline__ Acc acc0() { return {0.0, 0.0, 0.0, 0.0}; }
    __device__ static __forceinline__ void mac(Acc& c, V a, V b, int) {
        c.xx = fma(a.x, b.x, c.xx);
        c.yy = fma(a.y, b.y, c.yy);
        c.xy = fma(a.x, b.y, c.xy);
        c.yx = fma(a.y, b.x, c.yx);
    }
    __device__ static __forceinline__ V fold(const Acc& c) { return make_double2(c.xx - c.yy, c.xy + c.yx); }
};

enum SmallFlags : int {
    SF_FACTOR = 1,         // scatter Ax, row-scale, run the refactor part of the program
    SF_SOLVE = 2,          // load b, run the solve part, store x
    SF_STORE_NUMERIC = 4,  // write V (LU) and Rs behind the numeric handle
    SF_LOAD_NUMERIC = 8,   // read V (LU) and Rs from the numeric handle
    SF_SCALE_IN = 16,      // X = b[perm] / Rs[perm]       (klu_solve)
    SF_SCALE_OUT = 32,     // x[perm] = X / Rs[perm]       (klu_tsolve)
    SF_CERTIFY = 64,       // a failed pivot-parity certificate is status 2 (factor; klu_refactor never re-pivots)
};

struct SmallArgs {
    // packed programs: [0] first right-hand-side chunk (may include the refactor),
    // [1] further right-hand-side chunks (solve only)
    const uint32_t* stream[2];
    int nchunks[2];
    int nbatch[2];                          // batches (one-warp form) / levels (multi-warp form)
    uint32_t ctrl[2][kMaxBatches];          // warp-uniform control words (constant bank)
    uint16_t chunk_ptr[2][kMaxChunks + 1];  // first batch of each chunk
    int W;                   // systems (= consumer warps) per CTA
    int n, n_val, n_slots, zero_slot, rc, nz, flags;
    const uint16_t* tail_tab;  // dense tail in registers (kernels_wide.cuh): slots of the tail block, or null
    int tail_t0;
    const int* coo_dst;      // [nz] slot of each COO entry of this call, -1 = not in pattern
    const int* remap;        // compacted programs (plan.hpp: TaskProgram::remap): layout slot -> slot, else null
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
    unsigned long long* prof;  // optional [8] cycle counters (KLUJAX_CUDA_PROFILE=1)
};

constexpr int kRingBufs = 3;                                   // chunks in flight per CTA
constexpr int kChunkBytes = kChunkWords * 4;
constexpr int kRingBytes = kRingBufs * kChunkBytes;
constexpr int kSmallHeaderBytes = kRingBytes + 128;            // ring + mbarriers

// ---- mbarrier / TMA bulk copy (sm_90+ PTX; the schedule ring is fed by cp.async.bulk) ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// Operand fetch: slot fields of the stream words are pre-multiplied by 16.
template <typename T>
__device__ __forceinline__ typename Sc<T>::V* slot_hi(unsigned char* Vb, uint32_t w) {
    constexpr int SH = (sizeof(typename Sc<T>::V) == 16) ? 0 : 1;
    return reinterpret_cast<typename Sc<T>::V*>(Vb + ((w >> (14 + SH)) & (0x3FFF0u >> SH)));
}
template <typename T>
__device__ __forceinline__ typename Sc<T>::V* slot_lo(unsigned char* Vb, uint32_t w) {
    constexpr int SH = (sizeof(typename Sc<T>::V) == 16) ? 0 : 1;
    return reinterpret_cast<typename Sc<T>::V*>(Vb + ((w >> SH) & (0x3FFF0u >> SH)));
}

// Runs the batches [b0, b1) of one chunk out of the shared-memory ring.  Per batch and lane:
// acc = sum a*b over 2*trips operand lines; tree-reduce over g lanes; then
// V[out] = (V[out] - acc) [* V[pivot reciprocal]], or the reciprocal for a lane that finishes
// a pivot.  Everything that steers control flow comes from the control words in the
// parameter space, i.e. from uniform registers.
template <typename T>
__device__ __forceinline__ void run_chunk(const SmallArgs& a, int pi, int b0, int b1,
                                          unsigned char* __restrict__ Vb,
                                          const uint32_t* __restrict__ p, bool& bad, double& gmax) {
    using S = Sc<T>;
    using V_t = typename S::V;
    for (int b = b0; b < b1; ++b) {
        const uint32_t cw = a.ctrl[pi][b];
        int trips = cw & 0xFF;
        const int lg = (cw >> 8) & 7;
        const uint32_t h = p[0];
        const uint32_t* ops = p + 32;
        p = ops + trips * 64;
        V_t* po = slot_hi<T>(Vb, h);
        const V_t vo = *po;
        typename S::Acc c = S::acc0();
#pragma unroll 1
        do {
            const uint32_t w0 = ops[0], w1 = ops[32];
            ops += 64;
            const V_t a0 = *slot_hi<T>(Vb, w0), b0v = *slot_lo<T>(Vb, w0);
            const V_t a1 = *slot_hi<T>(Vb, w1), b1v = *slot_lo<T>(Vb, w1);
            S::mac(c, a0, b0v, 0);
            S::mac(c, a1, b1v, 1);
        } while (--trips);
        V_t acc = S::fold(c);
        // lane-group tree reduction; a group leader only ever reads lanes of its own group
        for (int s = lg; s-- > 0;) acc = S::add(acc, S::shfl_down(acc, 1 << s, 32));
        const V_t d = S::sub(vo, acc);
        V_t t = d;
        if (cw & kCwMul) t = S::mul(d, *slot_lo<T>(Vb, h));
        if (cw & kCwTrack) {  // entry of L: pivot-parity certificate, weighted by class (plan.hpp)
            const uint32_t cls = (h >> 2) & 3u;
            gmax = fmax(gmax, (h & 1u) ? S::mag2(t) * KB_TRACK_W2(cls) : 0.0);
        }
        if (cw & kCwPivot) {
            if (h & 2u) {
                bad |= S::bad_pivot(d);
                t = S::recip(d);
            }
        }
        *po = t;
        __syncwarp();
    }
}

// CTA = W consumer warps (one system each, private value array in shared memory) + one
// producer warp that streams the packed schedule through a ring of kRingBufs chunks with TMA
// bulk copies; full/empty mbarriers pace it.  All consumer warps of a CTA walk the same
// program, so the schedule is read from L2 once per CTA, not once per system.
template <typename T>
__global__ void __launch_bounds__(544, 1) small_kernel(const __grid_constant__ SmallArgs a) {
    using S = Sc<T>;
    using V_t = typename S::V;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint32_t* ring = reinterpret_cast<uint32_t*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + kRingBytes);
    uint64_t* empty = full + kRingBufs;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = a.W;
    const int n = a.n, rc = a.rc, r = a.r;
    const size_t per_sys = (static_cast<size_t>(a.n_slots) * sizeof(V_t) + static_cast<size_t>(n) * 8 + 15) & ~size_t(15);

    if (threadIdx.x == 0) {
        for (int i = 0; i < kRingBufs; ++i) {
            mbar_init(full + i, 1);
            mbar_init(empty + i, W);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int per_round = gridDim.x * W;
    const int rounds = (a.L + per_round - 1) / per_round;
    const int nrhs_chunks = (a.flags & SF_SOLVE) ? (r + rc - 1) / rc : 1;

    if (warp == W) {
        // ---------------- producer warp: one elected lane feeds the ring ----------------
        if (lane == 0) {
            uint32_t cc = 0;
            for (int rd = 0; rd < rounds; ++rd)
                for (int ch = 0; ch < nrhs_chunks; ++ch) {
                    const int pi = ch == 0 ? 0 : 1;
                    const uint32_t* src = a.stream[pi];
                    for (int c = 0; c < a.nchunks[pi]; ++c, ++cc) {
                        const uint32_t buf = cc % kRingBufs, use = cc / kRingBufs;
                        if (use > 0) mbar_wait(empty + buf, (use - 1) & 1);
                        mbar_expect_tx(full + buf, kChunkBytes);
                        bulk_g2s(ring + buf * kChunkWords, src + static_cast<size_t>(c) * kChunkWords,
                                 kChunkBytes, full + buf);
                    }
                }
        }
        return;
    }

    // ---------------- consumer warps ----------------
    V_t* V = reinterpret_cast<V_t*>(smem_raw + kSmallHeaderBytes + per_sys * warp);
    double* Rs = reinterpret_cast<double*>(V + a.n_slots);
    const V_t* Ax = static_cast<const V_t*>(a.Ax);
    const V_t* B = static_cast<const V_t*>(a.b);
    V_t* Xo = static_cast<V_t*>(a.x);
    V_t* Vst = static_cast<V_t*>(a.Vstore);
    uint32_t cc = 0;
    long long t_load = 0, t_b = 0, t_wait = 0, t_run = 0, t_store = 0, t0 = 0;
    const bool prof = a.prof != nullptr;

    for (int rd = 0; rd < rounds; ++rd) {
        const int m = rd * per_round + blockIdx.x * W + warp;
        const bool active = m < a.L;
        const int sm = (active && a.sysidx) ? a.sysidx[m] : m;
        bool bad = false;
        double gmax = 0.0;
        if (prof) t0 = clock64();
        if (active) {
            if (a.flags & SF_FACTOR) {
                for (int s = lane; s < a.n_val; s += 32) S::st(V + s, S::zero());
                __syncwarp();
                const V_t* Am = Ax + (long long)m * a.ax_stride;
                {   // pull the next round's inputs of this warp towards L2 while this system runs
                    const long long mn = (long long)m + per_round;
                    if (mn < a.L) {
                        const char* nx = reinterpret_cast<const char*>(Ax + mn * a.ax_stride);
                        for (int o = lane * 128; o < a.nz * (int)sizeof(V_t); o += 32 * 128)
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + o));
                        if (a.flags & SF_SOLVE) {
                            const char* nb = reinterpret_cast<const char*>(B + mn * a.b_stride);
                            for (int o = lane * 128; o < n * r * (int)sizeof(V_t); o += 32 * 128)
                                asm volatile("prefetch.global.L2 [%0];" ::"l"(nb + o));
                        }
                    }
                }
                int e = lane;
                for (; e + 96 < a.nz; e += 128) {
                    const int d0 = __ldg(a.coo_dst + e), d1 = __ldg(a.coo_dst + e + 32),
                              d2 = __ldg(a.coo_dst + e + 64), d3 = __ldg(a.coo_dst + e + 96);
                    const V_t v0 = Am[e], v1 = Am[e + 32], v2 = Am[e + 64], v3 = Am[e + 96];
                    if (d0 >= 0) S::st(V + d0, v0);
                    if (d1 >= 0) S::st(V + d1, v1);
                    if (d2 >= 0) S::st(V + d2, v2);
                    if (d3 >= 0) S::st(V + d3, v3);
                }
                for (; e < a.nz; e += 32) {
                    const int d = __ldg(a.coo_dst + e);
                    if (d >= 0) S::st(V + d, Am[e]);
                }
                __syncwarp();
                // max-abs row scaling (KLU scale = 2): Rs in original row order
                for (int i = lane; i < n; i += 32) {
                    const int pb = __ldg(a.rs_ptr + i), pe = __ldg(a.rs_ptr + i + 1);
                    // |z| = sqrt(re^2 + im^2) from the largest squared modulus of the row; the
                    // overflow-safe hypot form (what KLU uses) only when a square leaves the
                    // safe range
                    double q2 = 0.0;
                    for (int p = pb; p < pe; ++p) q2 = fmax(q2, S::mag2(S::ld(V + __ldg(a.rs_slot + p))));
                    double mx;
                    if (q2 > 1e-200 && q2 < 1e200) {
                        mx = sqrt(q2);
                    } else {
                        mx = 0.0;
                        for (int p = pb; p < pe; ++p) mx = fmax(mx, S::mag(S::ld(V + __ldg(a.rs_slot + p))));
                    }
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
        }
        if (lane == 0) {
            S::st(V + a.zero_slot, S::zero());
            S::st(V + a.zero_slot + 1, S::one());
            S::st(V + a.zero_slot + 2, S::zero());
        }
        __syncwarp();
        if (prof) { const long long t1 = clock64(); t_load += t1 - t0; t0 = t1; }

        for (int ch = 0; ch < nrhs_chunks; ++ch) {
            const int c0 = ch * rc;
            if (active && (a.flags & SF_SOLVE)) {
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
            if (prof) { const long long t1 = clock64(); t_b += t1 - t0; t0 = t1; }
            // ---- interpret the program, chunk by chunk, out of the ring ----
            const int pi = ch == 0 ? 0 : 1;
            for (int c = 0; c < a.nchunks[pi]; ++c, ++cc) {
                const uint32_t buf = cc % kRingBufs, use = cc / kRingBufs;
                mbar_wait(full + buf, use & 1);
                if (prof) { const long long t1 = clock64(); t_wait += t1 - t0; t0 = t1; }
                const uint32_t* base = ring + buf * kChunkWords;
                // every consumer warp runs the program (idle ones on scratch values)
                run_chunk<T>(a, pi, a.chunk_ptr[pi][c], a.chunk_ptr[pi][c + 1],
                             reinterpret_cast<unsigned char*>(V), base + lane, bad, gmax);
                __syncwarp();
                if (lane == 0) mbar_arrive(empty + buf);
                if (prof) { const long long t1 = clock64(); t_run += t1 - t0; t0 = t1; }
            }
            __syncwarp();
            if (active && (a.flags & SF_SOLVE)) {
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
        if (active) {
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
                    int st = anybad ? 1 : ((gmax <= kGrowthLimit2 || !(a.flags & SF_CERTIFY)) ? 0 : KLUJAX_PIVOT_DRIFT);
                    if (a.pattern_bad && *a.pattern_bad) st = -3;
                    if (a.status) a.status[m] = st;
                    if (a.growth) a.growth[m] = sqrt(gmax);
                }
            }
        }
        __syncwarp();
        if (prof) { const long long t1 = clock64(); t_store += t1 - t0; t0 = t1; }
    }
    if (prof && lane == 0) {
        atomicAdd(a.prof + 0, (unsigned long long)t_load);
        atomicAdd(a.prof + 1, (unsigned long long)t_b);
        atomicAdd(a.prof + 2, (unsigned long long)t_wait);
        atomicAdd(a.prof + 3, (unsigned long long)t_run);
        atomicAdd(a.prof + 4, (unsigned long long)t_store);
        atomicAdd(a.prof + 5, 1ull);
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
