// kernels_wide.cuh — shared-memory regime, second kernel: FOUR warps per system.
//
// Same residency as kernels_small.cuh (a system's whole value array lives in shared memory,
// the schedule arrives through the TMA-fed ring, control words are kernel parameters), but
// the schedule is in "as soon as possible" form: a partial update  V[out] -= sum V[a]*V[b]
// (by default a single multiply-add) is placed at the level where its operands become final
// (plan.cpp, split = 100 + 10*min + max).  A level is then a list of independent tasks, one
// per lane, spread over the 128 lanes of the system's four warps: no dot-product chains, no
// cross-lane reductions, one named barrier (bar.sync id, 128) per level.
//
// Replaces the same reference loops as kernels_small.cuh (/root/reference/klujax.cpp:431-451
// and siblings).  Stream: per 128-lane step and lane one header word
// (out<<18 | pivot-or-ONE<<4 | fresh<<3 | kind, kind 0 = idle lane) and ONE operand word
// (a<<18 | b<<4; ZERO, ZERO for a lane without a pair), stored as a pair (one 8-byte load);
// ctrl[step] = 1 | level-ends<<4.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "kernels_small.cuh"

namespace kb {

constexpr int kMaxWideSystems = 6;  // systems per CTA: 25 warps, 72 registers per thread (no spills)
constexpr int kWarpsPerSystem = 4;
constexpr int kLanesPerSystem = 32 * kWarpsPerSystem;
constexpr int kStepBytes = kLanesPerSystem * 8;           // one step of one system's lanes
constexpr int kStepsPerChunk = kChunkBytes / kStepBytes;  // 4

enum WideKind : uint32_t { WK_IDLE = 0, WK_SUB = 1, WK_MUL = 2, WK_MUL_TRACK = 3, WK_RECIP = 4, WK_MUL_TRACK_OFF = 5, WK_MUL_TRACK_DIAG = 6 };

__device__ __forceinline__ void named_barrier(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------------------------------------
// Dense tail in registers.  The last kTail = 32 pivots of the big block form a (nearly) dense
// trailing submatrix that carries ~60 % of the multiply-adds of the SAX systems (cfg 5: 10.4 k of
// 17.4 k) and the longest dependent chain (one pivot after the other).  Interpreted, every one of
// those multiply-adds costs three 16-byte shared-memory loads, one store and a schedule word.
// Here the four warps of the system take the Schur complement the interpreted steps left in
// shared memory into registers - warp w owns columns 8w..8w+7, lane i owns row i - and eliminate
// in lock step: the pivot row reaches the lanes of a warp by SHFL.IDX (its entries sit in lane k of
// the same warp), the multipliers L(:,k) of the warp that owns column k reach the other three
// through a 512-byte double-buffered exchange in shared memory (one named barrier per pivot);
// lanes at or above the pivot row multiply by zero instead of being predicated off.  The
// right-hand side rides along as a 33rd column (warp 0), U and the reciprocal pivots go back to
// their slots, warp 0 does the back substitution of the 32 tail rows by columns (x(j) by shuffle,
// U(:,j) from its slots) and leaves x in the right-hand-side slots for the interpreted back
// substitution of the rows above.  The loop over the pivots is 4 (owner warp) x 8 (unrolled
// column): ~1.5 k instructions, resident in the instruction cache (unlike the straight-line
// kernels of rowreg.cpp, profiles/rowreg_r02.md).
// Replaces, for those pivots, the column loop of klu_refactor and the tail of klu_solve's
// L- and U-solves (/root/reference/klujax.cpp:441-445 through KLU).
// The packer starts a new schedule chunk at the stage, so that it is entered from the chunk loop and
// not from the step loop, and the stage is a real function with its own register frame (inlined
// into the step loop its registers pushed the interpreter into spills: 4.9 -> 7.5 ms).  The loop
// over the 32 pivots is NOT unrolled and has no warp-specific code paths (a first version unrolled
// per owner warp and column: 10 k instructions, fetch-bound like the straight-line kernels):
// columns at or left of the pivot multiply by zero, warps whose columns are all done skip the step.
struct TailResult {
    double gmax;
    int bad;
};
template <typename T>
__device__ __forceinline__ TailResult tail_stage(const uint16_t* __restrict__ tab, unsigned char* Vb, double* Rs,
                                              typename Sc<T>::V* Xs, int wq, int lane, int bar_id) {
    using S = Sc<T>;
    using V_t = typename S::V;
    V_t* V = reinterpret_cast<V_t*>(Vb);
    TailResult res{0.0, 0};
    V_t acol[8];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
        const uint32_t s = __ldg(tab + (wq * 8 + jj) * kTail + lane);
        acol[jj] = s != 0xFFFFu ? V[s] : S::zero();
    }
    V_t y = wq == 0 ? Xs[0] : S::zero();
    V_t* lbuf = reinterpret_cast<V_t*>(Rs);  // Rs is dead after the right-hand side was loaded (solve mode)
#pragma unroll 1
    for (int k = 0; k < kTail; ++k) {
        const int kb = k >> 3, c = k & 7;
        V_t l = S::zero();
        if (wq == kb) {  // the warp that owns column k: pivot, multipliers, certificate
            V_t ac = acol[0];
#pragma unroll
            for (int jj = 1; jj < 8; ++jj)
                if (c == jj) ac = acol[jj];
            const V_t d = S::shfl_idx(ac, k);
            res.bad |= S::bad_pivot(d) ? 1 : 0;
            const V_t inv = S::fast_recip(d);
            if (lane > k) l = S::mul(ac, inv);
            res.gmax = fmax(res.gmax, S::mag2(l));
            if (lane == k) {  // the diagonal slot holds the reciprocal pivot
#pragma unroll
                for (int jj = 0; jj < 8; ++jj)
                    if (c == jj) acol[jj] = inv;
            }
            lbuf[(k & 1) * kTail + lane] = l;
        }
        named_barrier(bar_id, kLanesPerSystem);
        if (wq * 8 + 7 > k || wq == 0) {  // warps whose columns are all at or left of the pivot are done
            if (wq != kb) l = lbuf[(k & 1) * kTail + lane];
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const V_t u = S::shfl_idx(acol[jj], k);
                V_t lj = l;
                if (wq * 8 + jj <= k) lj = S::zero();  // finished column: multiply by zero
                S::fms_acc(acol[jj], lj, u);
            }
            if (wq == 0) S::fms_acc(y, l, S::shfl_idx(y, k));  // forward solve, fused
        }
    }
    // U and the reciprocal pivots back to their slots (rows at or above the diagonal)
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
        const uint32_t s = __ldg(tab + (wq * 8 + jj) * kTail + lane);
        if (s != 0xFFFFu && lane <= wq * 8 + jj) V[s] = acol[jj];
    }
    named_barrier(bar_id, kLanesPerSystem);
    if (wq == 0) {  // back substitution of the tail rows, by columns
        const V_t inv_me = V[__ldg(tab + lane * kTail + lane)];
#pragma unroll 1
        for (int j = kTail - 1; j >= 0; --j) {
            const V_t t = S::mul(y, inv_me);
            const V_t xj = S::shfl_idx(t, j);
            if (lane == j) y = t;
            const uint32_t s = __ldg(tab + j * kTail + lane);
            if (lane < j && s != 0xFFFFu) S::fms_acc(y, V[s], xj);
        }
        Xs[0] = y;
    }
    named_barrier(bar_id, kLanesPerSystem);
    return res;
}

// TAIL = true: the variant with the register stage of the dense tail; it runs four systems per CTA
// (17 warps, 5 per SM sub-partition: up to 96 registers per thread) instead of six (25 warps, 7 per
// sub-partition: 72 registers): the stage holds 8 complex columns per lane, and in the 72-register
// kernel it spilled into the interpreter loop.  Opt-in (KLUJAX_CUDA_TAIL=1), see klujax_cuda.cu.
constexpr int kMaxWideSystemsTail = 4;
template <typename T, bool TAIL>
__global__ void __launch_bounds__(32 * (kWarpsPerSystem * (TAIL ? kMaxWideSystemsTail : kMaxWideSystems) + 1), 1)
    wide_kernel(const __grid_constant__ SmallArgs a) {
    using S = Sc<T>;
    using V_t = typename S::V;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* ring = smem_raw;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + kRingBytes);
    uint64_t* empty = full + kRingBufs;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = a.W;
    const int n = a.n, rc = a.rc, r = a.r;
    const size_t per_sys =
        (static_cast<size_t>(a.n_slots) * sizeof(V_t) + static_cast<size_t>(n) * 8 + 16 + 15) & ~size_t(15);

    if (threadIdx.x == 0) {
        for (int i = 0; i < kRingBufs; ++i) {
            mbar_init(full + i, 1);
            mbar_init(empty + i, W * kWarpsPerSystem);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int per_round = gridDim.x * W;
    const int rounds = (a.L + per_round - 1) / per_round;
    const int nrhs_chunks = (a.flags & SF_SOLVE) ? (r + rc - 1) / rc : 1;

    if (warp == W * kWarpsPerSystem) {
        // ---------------- producer warp: one elected lane feeds the ring ----------------
        if (lane == 0) {
            uint32_t cc = 0;
            for (int rd = 0; rd < rounds; ++rd)
                for (int ch = 0; ch < nrhs_chunks; ++ch) {
                    const int pi = ch == 0 ? 0 : 1;
                    const unsigned char* src = reinterpret_cast<const unsigned char*>(a.stream[pi]);
                    for (int c = 0; c < a.nchunks[pi]; ++c, ++cc) {
                        const uint32_t buf = cc % kRingBufs, use = cc / kRingBufs;
                        if (use > 0) mbar_wait(empty + buf, (use - 1) & 1);
                        mbar_expect_tx(full + buf, kChunkBytes);
                        bulk_g2s(ring + buf * kChunkBytes, src + static_cast<size_t>(c) * kChunkBytes,
                                 kChunkBytes, full + buf);
                    }
                }
        }
        return;
    }

    // ---------------- consumer warps: kWarpsPerSystem per system ----------------
    const int slot = warp / kWarpsPerSystem, wq = warp % kWarpsPerSystem;
    const int tid = wq * 32 + lane;  // 0..127 inside the system's thread group
    unsigned char* Vb = smem_raw + kSmallHeaderBytes + per_sys * slot;
    V_t* V = reinterpret_cast<V_t*>(Vb);
    double* Rs = reinterpret_cast<double*>(V + a.n_slots);
    unsigned long long* flags = reinterpret_cast<unsigned long long*>(Rs + n);  // [0] bad, [1] max |L|^2
    const V_t* Ax = static_cast<const V_t*>(a.Ax);
    const V_t* B = static_cast<const V_t*>(a.b);
    V_t* Xo = static_cast<V_t*>(a.x);
    V_t* Vst = static_cast<V_t*>(a.Vstore);
    const int bar_id = 1 + slot;
    uint32_t cc = 0;  // running chunk index: the ring position of this warp
    long long t_load = 0, t_b = 0, t_wait = 0, t_run = 0, t_store = 0, t0 = 0;
    const bool prof = a.prof != nullptr;

    for (int rd = 0; rd < rounds; ++rd) {
        const int m = rd * per_round + blockIdx.x * W + slot;
        const bool active = m < a.L;
        const int sm = (active && a.sysidx) ? a.sysidx[m] : m;
        bool bad = false;
        double gmax = 0.0;
        if (prof) t0 = clock64();
        if (tid < 2) flags[tid] = 0ull;
        if (active) {
            if (a.flags & SF_FACTOR) {
                for (int s = tid; s < a.n_val; s += kLanesPerSystem) S::st(V + s, S::zero());
                named_barrier(bar_id, kLanesPerSystem);
                const V_t* Am = Ax + (long long)m * a.ax_stride;
                {   // pull the next round's inputs towards L2 while this system runs
                    const long long mn = (long long)m + per_round;
                    if (mn < a.L) {
                        const char* nx = reinterpret_cast<const char*>(Ax + mn * a.ax_stride);
                        for (int o = tid * 128; o < a.nz * (int)sizeof(V_t); o += kLanesPerSystem * 128)
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + o));
                        if (a.flags & SF_SOLVE) {
                            const char* nb = reinterpret_cast<const char*>(B + mn * a.b_stride);
                            for (int o = tid * 128; o < n * r * (int)sizeof(V_t); o += kLanesPerSystem * 128)
                                asm volatile("prefetch.global.L2 [%0];" ::"l"(nb + o));
                        }
                    }
                }
                for (int e = tid; e < a.nz; e += kLanesPerSystem) {
                    int d = __ldg(a.coo_dst + e);
                    if (d >= 0) {
                        if (a.remap) d = __ldg(a.remap + d);
                        S::st(V + d, Am[e]);
                    }
                }
                named_barrier(bar_id, kLanesPerSystem);
                // max-abs row scaling (KLU scale = 2): Rs in original row order
                for (int i = tid; i < n; i += kLanesPerSystem) {
                    const int pb = __ldg(a.rs_ptr + i), pe = __ldg(a.rs_ptr + i + 1);
                    double q2 = 0.0;
                    for (int p = pb; p < pe; ++p) q2 = fmax(q2, S::mag2(S::ld(V + __ldg(a.rs_slot + p))));
                    double mx;
                    if (q2 > 1e-200 && q2 < 1e200) {
                        mx = sqrt(q2);
                    } else {  // overflow-safe modulus (what KLU uses) outside the safe range
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
                for (int s = tid; s < a.n_val; s += kLanesPerSystem) S::st(V + s, src[s]);
                const double* rsrc = a.Rsstore + (long long)sm * n;
                for (int i = tid; i < n; i += kLanesPerSystem) Rs[i] = rsrc[i];
            }
        }
        if (tid == 0) {
            S::st(V + a.zero_slot, S::zero());
            S::st(V + a.zero_slot + 1, S::one());
            S::st(V + a.zero_slot + 2, S::zero());
        }
        named_barrier(bar_id, kLanesPerSystem);
        if (prof) { const long long t1 = clock64(); t_load += t1 - t0; t0 = t1; }

        for (int ch = 0; ch < nrhs_chunks; ++ch) {
            const int c0 = ch * rc;
            if (active && (a.flags & SF_SOLVE)) {
                const V_t* Bm = B + (long long)m * a.b_stride;
                for (int idx = tid; idx < n * rc; idx += kLanesPerSystem) {
                    const int k = idx / rc, c = c0 + idx - k * rc;
                    V_t v = S::zero();
                    if (c < r) {
                        const int src = __ldg(a.in_perm + k);
                        v = Bm[(long long)src * r + c];
                        if (a.flags & SF_SCALE_IN) v = S::rdiv(v, Rs[src]);
                    }
                    S::st(V + a.n_val + idx, v);
                }
            }
            named_barrier(bar_id, kLanesPerSystem);
            if (prof) { const long long t1 = clock64(); t_b += t1 - t0; t0 = t1; }

            // ---- walk the steps, chunk by chunk, out of the ring ----
            const int pi = ch == 0 ? 0 : 1;
            constexpr int SHR = (sizeof(V_t) == 16) ? 0 : 1;  // slot fields are pre-multiplied by 16
            constexpr uint32_t MSK = 0x3FFF0u >> SHR;
            for (int c = 0; c < a.nchunks[pi]; ++c, ++cc) {
                const uint32_t buf = cc % kRingBufs;
                mbar_wait(full + buf, (cc / kRingBufs) & 1);
                // a step = 128 (header word, operand word) pairs, one 8-byte load per lane
                const uint2* p = reinterpret_cast<const uint2*>(ring + buf * kChunkBytes) + tid;
                const int s0 = a.chunk_ptr[pi][c], s1 = a.chunk_ptr[pi][c + 1];
                uint32_t cw = a.ctrl[pi][s0];
                if (TAIL && (cw & 32u)) {  // first step of a chunk, after a level end: the dense tail in registers
                    const TailResult tr = tail_stage<T>(a.tail_tab, Vb, Rs, V + a.n_val + (a.tail_t0 + lane) * a.rc, wq,
                                                        lane, bar_id);
                    bad |= tr.bad != 0;
                    gmax = fmax(gmax, tr.gmax);
                }
                for (int s = s0; s < s1; ++s) {
                    const bool level_end = (cw & 16u) != 0;
                    const uint2 hw = p[0];
                    const uint32_t h = hw.x, w1 = hw.y;  // header word, operand word
                    cw = a.ctrl[pi][s + 1];                  // next step's control word, off the critical path
                    const uint32_t kind = h & 0x7u;
                    if (kind) {
                        V_t* po = reinterpret_cast<V_t*>(Vb + ((h >> (14 + SHR)) & MSK));
                        V_t t = S::zero();
                        if (!(h & 8u)) t = *po;  // 8: first write of a recycled slot starts from 0
                        {   // lanes without a pair multiply ZERO by ZERO: no lane-dependent branch
                            const V_t va = *reinterpret_cast<const V_t*>(Vb + ((w1 >> (14 + SHR)) & MSK));
                            const V_t vb = *reinterpret_cast<const V_t*>(Vb + ((w1 >> SHR) & MSK));
                            S::fms_acc(t, va, vb);
                        }
                        if (kind >= WK_MUL) {
                            if (kind == WK_RECIP) {
                                bad |= S::bad_pivot(t);
                                t = S::fast_recip(t);
                            } else {
                                t = S::mul(t, *reinterpret_cast<const V_t*>(Vb + ((h >> SHR) & MSK)));
                                if (kind != WK_MUL)  // entry of L: pivot-parity certificate (plan.hpp)
                                    gmax = fmax(gmax, S::mag2(t) * (kind == WK_MUL_TRACK ? 1.0 : kind == WK_MUL_TRACK_OFF ? 1e6 : 1e12));
                            }
                        }
                        *po = t;
                    }
                    p += kLanesPerSystem;
                    if (level_end) named_barrier(bar_id, kLanesPerSystem);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(empty + buf);
            }
            if (prof) { const long long t1 = clock64(); t_run += t1 - t0; t0 = t1; }

            if (active && (a.flags & SF_SOLVE)) {
                V_t* Xm = Xo + (long long)m * n * r;
                for (int idx = tid; idx < n * rc; idx += kLanesPerSystem) {
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
                for (int s = tid; s < a.n_val; s += kLanesPerSystem) dst[s] = S::ld(V + s);
                double* rdst = a.Rsstore + (long long)sm * n;
                for (int i = tid; i < n; i += kLanesPerSystem) rdst[i] = Rs[i];
            }
            if (a.flags & SF_FACTOR) {
                const bool anybad = __any_sync(0xffffffffu, bad);
                for (int o = 16; o > 0; o >>= 1) gmax = fmax(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
                if (lane == 0) {
                    if (anybad) atomicOr(flags, 1ull);
                    atomicMax(flags + 1, (unsigned long long)__double_as_longlong(gmax));
                }
            }
        }
        named_barrier(bar_id, kLanesPerSystem);
        if (active && (a.flags & SF_FACTOR) && tid == 0) {
            const double g2 = __longlong_as_double((long long)flags[1]);
            int st = flags[0] ? 1 : ((g2 <= kGrowthLimit2 || !(a.flags & SF_CERTIFY)) ? 0 : KLUJAX_PIVOT_DRIFT);
            if (a.pattern_bad && *a.pattern_bad) st = -3;
            if (a.status) a.status[m] = st;
            if (a.growth) a.growth[m] = sqrt(__longlong_as_double((long long)flags[1]));
        }
        named_barrier(bar_id, kLanesPerSystem);
        if (prof) { const long long t1 = clock64(); t_store += t1 - t0; t0 = t1; }
    }
    if (prof && tid == 0) {
        atomicAdd(a.prof + 0, (unsigned long long)t_load);
        atomicAdd(a.prof + 1, (unsigned long long)t_b);
        atomicAdd(a.prof + 2, (unsigned long long)t_wait);
        atomicAdd(a.prof + 3, (unsigned long long)t_run);
        atomicAdd(a.prof + 4, (unsigned long long)t_store);
        atomicAdd(a.prof + 5, 1ull);
    }
}

}  // namespace kb
