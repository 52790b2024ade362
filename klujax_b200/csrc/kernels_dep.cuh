// kernels_dep.cuh — dependency-driven regime: the large-matrix kernels for batches that are not
// larger than the machine (one system, or one right-hand-side column, per CTA).
//
// The level-scheduled kernels (kernels_level.cuh) pay a CTA barrier and a memory round trip per
// level; a system whose elimination DAG is a long chain (cfg 4: circuit matrices, 5 215 levels to
// factor, 6 205 per triangular solve) is then latency-bound at 1-2 us per level.  The kernels here
// have no level barriers: the work runs in a fixed order that respects the dependencies, and a
// unit waits only for the producers of its own operands.
//
//   dep_factor  (opt-in, KLUJAX_CUDA_DEP_FACTOR=1: correct, but slower than lv_factor on cfg 4 -
//               profiles/dep_r02.md)  left-looking refactorization, a warp per column (the shape of
//               klu_refactor's own loop, reached from /root/reference/klujax.cpp:781, 911): the
//               column lives in the warp's shared-memory buffer, every U(j,k) waits for the ready
//               flag of column j (device-side dependency flags in shared memory), the 32 lanes
//               subtract L(:,j)*U(j,k) - the records of a column go in chunks of 32 whose
//               destination positions and (for final columns) entries of L are staged with
//               cp.async - then pivot reciprocal, scaling of L(:,k), the pivot-parity certificate,
//               one coalesced store of the column, flag.
//   panel_solve klu_solve / klu_tsolve (klujax.cpp:1087, 1220) as one sequential list of row
//               tasks in panels of 32 (plan.hpp: PanelProgram): the right-hand side lives in shared
//               memory, lane r owns task r of the panel, operands from earlier panels are plain
//               shared-memory reads, operands produced inside the panel travel by warp shuffle
//               in a loop over the producing lanes.  A few warps take the panels in turn (one
//               hand-over counter); while a warp waits for its turn it loads the static part of its
//               panel (indices, entries of L / U, pivots), so the turn itself is shared-memory and
//               register work only.
//
// Deadlock freedom of dep_factor: columns are dealt round-robin and a warp handles its columns in
// increasing order, so the smallest unfinished column never waits.  All workers are warps of ONE
// resident CTA.
//
// Measured on the way (profiles/dep_r02.md): a flag hand-over between two warps of a CTA costs
// ~350 cycles on B200 (tools/ubench_flag.cu), and every additional warp that polls shared memory
// slows the warp everybody waits for (a first version of the solves - row tasks dealt to 32 warps
// with a flag per row - was 2x SLOWER than the level kernels): hence few waiting warps, and
// shuffles instead of flags wherever the dependency stays inside 32 consecutive tasks.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "kernels_small.cuh"
#include "plan.hpp"

namespace kb {

struct DepColProg {
    int n = 0, maxcol = 0, depth = 0;
    int64_t n_mac = 0;
    int4* d_colinfo = nullptr;
    int* d_ccand = nullptr;
    int4* d_urec = nullptr;
    int* d_dest = nullptr;
    ~DepColProg() {
        if (d_colinfo) cudaFree(d_colinfo);
        if (d_ccand) cudaFree(d_ccand);
        if (d_urec) cudaFree(d_urec);
        if (d_dest) cudaFree(d_dest);
    }
    bool upload(const ColProgram& cp) {
        n = cp.n;
        maxcol = cp.maxcol;
        depth = cp.depth;
        n_mac = cp.n_mac;
        auto up = [&](void** d, const void* src, size_t bytes) {
            if (cudaMalloc(d, std::max<size_t>(bytes, 16)) != cudaSuccess) return false;
            return bytes == 0 || cudaMemcpy(*d, src, bytes, cudaMemcpyHostToDevice) == cudaSuccess;
        };
        return up(reinterpret_cast<void**>(&d_colinfo), cp.colinfo.data(), cp.colinfo.size() * sizeof(int)) &&
               up(reinterpret_cast<void**>(&d_ccand), cp.ccand.data(), cp.ccand.size() * sizeof(int)) &&
               up(reinterpret_cast<void**>(&d_urec), cp.urec.data(), cp.urec.size() * sizeof(int)) &&
               up(reinterpret_cast<void**>(&d_dest), cp.dest.data(), cp.dest.size() * sizeof(int));
    }
};

struct DepPanelProg {
    int n = 0, npanels = 0, ntemps = 0, maxI = 0, maxE = 0;
    int64_t n_mac = 0;
    int4 *d_hdr = nullptr, *d_lane = nullptr;
    int *d_ext_slot = nullptr, *d_ext_idx = nullptr, *d_in_slot = nullptr;
    ~DepPanelProg() {
        if (d_hdr) cudaFree(d_hdr);
        if (d_lane) cudaFree(d_lane);
        if (d_ext_slot) cudaFree(d_ext_slot);
        if (d_ext_idx) cudaFree(d_ext_idx);
        if (d_in_slot) cudaFree(d_in_slot);
    }
    bool upload(const PanelProgram& pp) {
        n = pp.n;
        npanels = pp.npanels;
        ntemps = pp.ntemps;
        maxI = pp.maxI;
        maxE = pp.maxE;
        n_mac = pp.n_mac;
        auto up = [&](void** d, const std::vector<int>& v) {
            if (cudaMalloc(d, std::max<size_t>(v.size() * sizeof(int), 16)) != cudaSuccess) return false;
            return v.empty() || cudaMemcpy(*d, v.data(), v.size() * sizeof(int), cudaMemcpyHostToDevice) == cudaSuccess;
        };
        return up(reinterpret_cast<void**>(&d_hdr), pp.hdr) && up(reinterpret_cast<void**>(&d_lane), pp.lane) &&
               up(reinterpret_cast<void**>(&d_ext_slot), pp.ext_slot) && up(reinterpret_cast<void**>(&d_ext_idx), pp.ext_idx) &&
               up(reinterpret_cast<void**>(&d_in_slot), pp.in_slot);
    }
};

struct DepArgs {
    // refactorization
    const int4* colinfo;
    const int* ccand;
    const int4* urec;
    const int* dest;
    int scratch_elems;  // values per warp buffer; longer columns are updated in place in global memory
    int stage_elems;    // staged entries of L (and destinations) per warp and chunk of records
    // solves
    const int4 *hdr, *lane;
    const int *ext_slot, *ext_idx, *in_slot;
    int npanels, ntemps, stage_lines, r;   // stage_lines: in-panel lines a warp's staging buffer holds
    void* X;                 // the caller's x[system][row][column]
    const void* B;           // right-hand sides b[system * b_stride + row * r + column]
    long long b_stride;
    const int* perm_in;      // row of b behind pivot-order row k
    const int* perm_out;     // row of x that pivot-order row k goes to
    const double* Rs;        // Rs[row * rlp + system * rss]
    long long rlp, rss;
    int scale_in, scale_out;
    // common
    int n, L;
    void* V;                 // V[slot * lpad + system * vss]
    long long lpad, vss;
    const int* sysidx;
    int* bad;
    unsigned long long* growth2;
    unsigned sleep_ns;       // back-off between two polls of a flag
    unsigned long long* prof;  // KLUJAX_CUDA_DEP_PROF=1: cycle counters per phase (summed over warps)
};

template <int BYTES>
__device__ __forceinline__ void cp_async(void* smem_dst, const void* gsrc) {
    const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(d), "l"(gsrc), "n"(BYTES) : "memory");
}

// A waiting warp must not spin: 32 warps polling shared memory take the issue slots and the
// shared-memory pipe from the one warp everybody is waiting for (measured on cfg 4: the solve ran
// 2x SLOWER than the level kernel with plain spinning).  __nanosleep takes the warp off the
// scheduler between two polls.
__device__ __forceinline__ void dep_wait_flag(const volatile unsigned char* f, unsigned ns) {
    // back-off: the warp the chain is waiting for has waited briefly, the ones far ahead of the frontier poll rarely
    const unsigned cap = 8 * ns + 32;
    while (*f == 0) {
        __nanosleep(ns);
        ns = min(2 * ns + 8, cap);
    }
}

template <typename T>
__global__ void __launch_bounds__(512, 1) dep_factor(const DepArgs a) {
    using S = Sc<T>;
    using V_t = typename S::V;
    extern __shared__ __align__(16) unsigned char dep_smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int fl_bytes = (a.n + 15) & ~15;
    volatile unsigned char* flags = dep_smem;
    // per warp: column buffer [scratch_elems], staged entries of L [stage_elems], their destinations [stage_elems]
    const size_t warp_bytes = (static_cast<size_t>(a.scratch_elems) + a.stage_elems) * sizeof(V_t) + static_cast<size_t>(a.stage_elems) * sizeof(int);
    unsigned char* wb = dep_smem + fl_bytes + static_cast<size_t>(w) * warp_bytes;
    V_t* scratch = reinterpret_cast<V_t*>(wb);
    V_t* st_l = scratch + a.scratch_elems;
    int* st_d = reinterpret_cast<int*>(st_l + a.stage_elems);
    const long long ls = a.lpad;
    for (int m = blockIdx.x; m < a.L; m += gridDim.x) {
        for (int i = threadIdx.x; i < fl_bytes / 4; i += blockDim.x) reinterpret_cast<volatile unsigned*>(dep_smem)[i] = 0u;
        __syncthreads();
        const long long sm = a.sysidx ? static_cast<long long>(a.sysidx[m]) : static_cast<long long>(m);
        V_t* Vs = static_cast<V_t*>(a.V) + sm * a.vss;
        double gmax = 0.0;
        bool bad = false;
        long long tp[6] = {0, 0, 0, 0, 0, 0}, tc = 0;
        const bool prof = a.prof != nullptr;
#define DEP_TICK(i) if (prof) { const long long tn = clock64(); tp[i] += tn - tc; tc = tn; }
        if (prof) tc = clock64();
        for (int k = w; k < a.n; k += W) {
            const int4 ci = __ldg(a.colinfo + k);
            const int cb = ci.x, nu = ci.y, nc = ci.z;
            const int cand = __ldg(a.ccand + k);
            const bool in_smem = nc <= a.scratch_elems;  // warp-uniform
            V_t* Vc = Vs + static_cast<long long>(cb) * ls;
            auto cget = [&](int p) -> V_t { return in_smem ? scratch[p] : Vc[static_cast<long long>(p) * ls]; };
            auto cset = [&](int p, V_t v) {
                if (in_smem) scratch[p] = v;
                else Vc[static_cast<long long>(p) * ls] = v;
            };
            if (in_smem)
                for (int p = lane; p < nc; p += 32) scratch[p] = Vc[static_cast<long long>(p) * ls];
            __syncwarp();
            DEP_TICK(0)  // column info + column load
            // Update records in chunks of up to 32: lane r holds record u0 + r.  For the columns j that are
            // final already, the entries of L(:,j) and their destination positions of the WHOLE chunk are
            // brought into the warp's staging area with cp.async (one memory round trip per chunk; a
            // dense-ish column has hundreds of records), then the records are applied in order out of
            // shared memory.  A record whose column is not final yet is waited for when its turn comes.
            for (int u0 = 0; u0 < nu;) {
                const int ntry = min(32, nu - u0);
                int4 rc = make_int4(0, 0, 0, 0);
                if (lane < ntry) rc = __ldg(a.urec + ci.w + u0 + lane);
                const bool ready = lane < ntry && flags[rc.x] != 0;
                int cnt = lane < ntry ? min(rc.z, 32) : 0;  // staged entries of my record (destinations always, entries of L if final)
                int off = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, off, o);
                    if (lane >= o) off += v;
                }
                off -= cnt;
                const unsigned fits = __ballot_sync(0xffffffffu, off + cnt <= a.stage_elems);
                const unsigned rdy = __ballot_sync(0xffffffffu, ready);
                const int nrec = min(ntry, (~fits) ? __ffs(~fits) - 1 : 32);
                __threadfence_block();
                DEP_TICK(1)  // records + flags of the chunk
                for (int r = 0; r < nrec; ++r) {
                    const int lc = __shfl_sync(0xffffffffu, cnt, r), lb = __shfl_sync(0xffffffffu, rc.y, r),
                              dp = __shfl_sync(0xffffffffu, rc.w, r), of = __shfl_sync(0xffffffffu, off, r);
                    if (lane < lc) {
                        if ((rdy >> r) & 1u) cp_async<sizeof(V_t)>(st_l + of + lane, Vs + static_cast<long long>(lb + lane) * ls);
                        cp_async<4>(st_d + of + lane, a.dest + dp + lane);
                    }
                }
                asm volatile("cp.async.wait_all;" ::: "memory");
                __syncwarp();
                for (int r = 0; r < nrec; ++r) {
                    const int j = __shfl_sync(0xffffffffu, rc.x, r), lbeg = __shfl_sync(0xffffffffu, rc.y, r),
                              lcnt = __shfl_sync(0xffffffffu, rc.z, r), dp = __shfl_sync(0xffffffffu, rc.w, r),
                              of = __shfl_sync(0xffffffffu, off, r);
                    const bool staged = (rdy >> r) & 1u;
                    if (!staged) {
                        dep_wait_flag(flags + j, a.sleep_ns);
                        __threadfence_block();
                        DEP_TICK(2)  // waiting for a column
                    }
                    const V_t uv = cget(u0 + r);
                    if (lane < lcnt) {
                        const int d = st_d[of + lane];
                        const V_t l = staged ? st_l[of + lane] : Vs[static_cast<long long>(lbeg + lane) * ls];
                        V_t t = cget(d);
                        S::fms_acc(t, l, uv);
                        cset(d, t);
                    }
                    for (int q = lane + 32; q < lcnt; q += 32) {  // rare: more than 32 entries in L(:,j)
                        const int d = __ldg(a.dest + dp + q);
                        const V_t l = Vs[static_cast<long long>(lbeg + q) * ls];
                        V_t t = cget(d);
                        S::fms_acc(t, l, uv);
                        cset(d, t);
                    }
                    __syncwarp();
                }
                DEP_TICK(3)  // applying the chunk
                u0 += nrec;
            }
            const V_t pv = cget(nu);
            bad |= S::bad_pivot(pv);
            const V_t inv = S::recip(pv);
            __syncwarp();  // every lane has read the pivot before it is replaced by its reciprocal
            for (int p = nu + 1 + lane; p < nc; p += 32) {
                const V_t v = S::mul(cget(p), inv);
                // pivot-parity certificate, weighted by class (plan.hpp: TK_MUL_TRACK*)
                gmax = fmax(gmax, S::mag2(v) * (cand == -2 ? 1.0 : (p == cand ? 1e12 : 1e6)));
                cset(p, v);
            }
            if (lane == 0) cset(nu, inv);
            __syncwarp();
            DEP_TICK(4)  // pivot + scaling
            if (in_smem)
                for (int p = lane; p < nc; p += 32) Vc[static_cast<long long>(p) * ls] = scratch[p];
            __threadfence_block();
            __syncwarp();
            if (lane == 0) flags[k] = 1;
            DEP_TICK(5)  // store + fence + flag
        }
        if (prof && lane == 0)
            for (int i = 0; i < 6; ++i) atomicAdd(a.prof + i, static_cast<unsigned long long>(tp[i]));
        const bool anybad = __any_sync(0xffffffffu, bad);
        for (int o = 16; o > 0; o >>= 1) gmax = fmax(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
        if (lane == 0) {
            if (anybad) a.bad[m] = 1;
            atomicMax(a.growth2 + m, static_cast<unsigned long long>(__double_as_longlong(gmax)));
        }
        __syncthreads();  // the flags are cleared for the next system
    }
}

// Panels of 32 row tasks, taken in turn by the warps of the CTA (see the header).
constexpr int kPanelWarps = 8;
template <typename T>
__global__ void __launch_bounds__(32 * kPanelWarps, 1) panel_solve(const DepArgs a) {
    using S = Sc<T>;
    using V_t = typename S::V;
    extern __shared__ __align__(16) unsigned char dep_smem[];
    // external operands a lane holds in registers before its turn: a whole task (plan.hpp: kPanelChunk)
    // for real values, half of one for complex ones; the in-panel coefficients are staged in batches
    constexpr int kPanelEC = sizeof(V_t) == 8 ? 24 : 12, kStageB = sizeof(V_t) == 8 ? 16 : 8;
    const int n = a.n, r = a.r, nx = n + a.ntemps;
    V_t* Xs = reinterpret_cast<V_t*>(dep_smem);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, NW = blockDim.x >> 5;
    V_t* stage = Xs + nx + static_cast<size_t>(w) * a.stage_lines * 32;
    volatile int* done = reinterpret_cast<volatile int*>(Xs + nx + static_cast<size_t>(NW) * a.stage_lines * 32);
    const long long ls = a.lpad;
    const long long Q = static_cast<long long>(a.L) * r;
    for (long long q = blockIdx.x; q < Q; q += gridDim.x) {
        const int m = static_cast<int>(q / r), c = static_cast<int>(q - static_cast<long long>(m) * r);
        const long long sm = a.sysidx ? static_cast<long long>(a.sysidx[m]) : static_cast<long long>(m);
        const V_t* Vs = static_cast<const V_t*>(a.V) + sm * a.vss;
        const V_t* Bm = static_cast<const V_t*>(a.B) + static_cast<long long>(m) * a.b_stride + c;
        V_t* Xm = static_cast<V_t*>(a.X) + static_cast<long long>(m) * n * r + c;
        if (threadIdx.x == 0) *done = 0;
        for (int k = threadIdx.x; k < n; k += blockDim.x) {
            const int src = __ldg(a.perm_in + k);
            V_t v = Bm[static_cast<long long>(src) * r];
            if (a.scale_in) v = S::rdiv(v, a.Rs[static_cast<long long>(src) * a.rlp + sm * a.rss]);
            Xs[k] = v;
        }
        __syncthreads();
        auto coef = [&](int slot) -> V_t {  // entry of L / U, the constant -1 of a partial sum, or nothing
            if (slot >= 0) return __ldg(Vs + static_cast<long long>(slot) * ls);
            V_t v = S::zero();
            if (slot == -2) v = S::sub(v, S::one());
            return v;
        };
        long long tp[6] = {0, 0, 0, 0, 0, 0}, tc = 0;
        const bool prof = a.prof != nullptr;
        if (prof) tc = clock64();
        for (int p = w; p < a.npanels; p += NW) {
            // ---- static part: needs nothing from earlier panels ----
            const int4 h = __ldg(a.hdr + p);
            const int E = h.x, I = h.y;
            const int4 li = __ldg(a.lane + static_cast<size_t>(p) * 32 + lane);
            const bool mulp = (li.w & 4) != 0;
            V_t dv = S::one();
            if (li.y >= 0) dv = __ldg(Vs + static_cast<long long>(li.y) * ls);
            const int* isl = a.in_slot + static_cast<size_t>(h.w) * 32 + lane;
            // In-panel coefficients, staged in shared memory.  The producing lane's pivot reciprocal is
            // folded in here (x_s = acc_s * d_s, so acc_r -= (c * d_s) * acc_s): the shuffle loop below
            // then carries acc itself and its dependent chain is shuffle + one multiply-add per step.
            const V_t dve = mulp ? dv : S::one();
            unsigned rest = static_cast<unsigned>(li.z);
            for (int t0 = 0; t0 < I; t0 += kStageB) {  // all index loads, then all value loads, then the stores
                int sl[kStageB];
                V_t sv[kStageB];
#pragma unroll
                for (int u = 0; u < kStageB; ++u) sl[u] = t0 + u < I ? __ldg(isl + (t0 + u) * 32) : -1;
#pragma unroll
                for (int u = 0; u < kStageB; ++u) sv[u] = coef(sl[u]);
#pragma unroll
                for (int u = 0; u < kStageB; ++u) {
                    const int src = rest ? __ffs(rest) - 1 : lane;  // my (t0+u)-th producing lane
                    rest &= rest - 1;
                    const V_t ds = S::shfl_idx(dve, src);
                    if (t0 + u < I) stage[(t0 + u) * 32 + lane] = S::sub(S::zero(), S::mul(sv[u], ds));  // negated: the loop adds
                }
            }
            const int* esl = a.ext_slot + static_cast<size_t>(h.z) * 32 + lane;
            const int* eix = a.ext_idx + static_cast<size_t>(h.z) * 32 + lane;
            int ei[kPanelEC];
            V_t ec[kPanelEC];
            auto fetch = [&](int e0) {
                int es[kPanelEC];
#pragma unroll
                for (int u = 0; u < kPanelEC; ++u) {
                    const bool on = e0 + u < E;
                    ei[u] = on ? __ldg(eix + (e0 + u) * 32) : -1;
                    es[u] = on ? __ldg(esl + (e0 + u) * 32) : -1;
                }
#pragma unroll
                for (int u = 0; u < kPanelEC; ++u) ec[u] = coef(ei[u] >= 0 ? es[u] : -1);
            };
            fetch(0);
            DEP_TICK(0)  // static part (issue; the loads complete while the warp waits)
            // ---- my turn ----
            while (*done < p) __nanosleep(a.sleep_ns);
            __threadfence_block();
            DEP_TICK(1)  // waiting for the turn
            V_t acc = (li.w & 1) ? S::zero() : Xs[li.x];
            for (int e0 = 0; e0 < E; e0 += kPanelEC) {
                int ci[kPanelEC];
                V_t cc[kPanelEC];
#pragma unroll
                for (int u = 0; u < kPanelEC; ++u) ci[u] = ei[u], cc[u] = ec[u];
                if (e0 + kPanelEC < E) fetch(e0 + kPanelEC);  // next chunk in flight while this one is applied
#pragma unroll
                for (int u = 0; u < kPanelEC; ++u)
                    if (ci[u] >= 0) S::fms_acc(acc, cc[u], Xs[ci[u]]);
            }
            DEP_TICK(2)  // operands from earlier panels
            // operands produced inside the panel: one shuffle step per producing lane
            const unsigned any = __reduce_or_sync(0xffffffffu, static_cast<unsigned>(li.z));
            int t = 0;
            const int tmax = I > 0 ? I - 1 : 0;
            V_t cn = stage[lane];  // next (negated) coefficient of this lane, always one step ahead
            // unrolled over the 32 possible producing lanes: the shuffle source is an immediate and a
            // lane nobody consumes costs one uniform branch (a loop over the set bits put a
            // find-first-set in front of every shuffle: ~100 cycles per step instead of ~45)
#pragma unroll
            for (int s = 0; s < 31; ++s) {
                if (any & (1u << s)) {
                    const V_t xs = S::shfl_idx(acc, s);
                    const bool mine = (static_cast<unsigned>(li.z) >> s) & 1u;
                    V_t c = cn;
                    if (!mine) c = S::zero();
                    S::fma_acc(acc, c, xs);
                    t += mine ? 1 : 0;
                    cn = stage[min(t, tmax) * 32 + lane];
                }
            }
            DEP_TICK(3)  // shuffle steps
            if (!(li.w & 2)) Xs[li.x] = mulp ? S::mul(acc, dv) : acc;
            __threadfence_block();
            __syncwarp();
            if (lane == 0) *done = p + 1;
            DEP_TICK(4)  // store + hand-over
        }
        if (prof && lane == 0)
            for (int i = 0; i < 6; ++i) atomicAdd(a.prof + i, static_cast<unsigned long long>(tp[i]));
        __syncthreads();
        for (int k = threadIdx.x; k < n; k += blockDim.x) {
            const int dst = __ldg(a.perm_out + k);
            V_t v = Xs[k];
            if (a.scale_out) v = S::rdiv(v, a.Rs[static_cast<long long>(dst) * a.rlp + sm * a.rss]);
            Xm[static_cast<long long>(dst) * r] = v;
        }
        __syncthreads();
    }
}

// ---- host side -------------------------------------------------------------------------------
constexpr size_t kDepSmemMax = 227 * 1024;

static int dep_env_int_early(const char* name, int dflt) {
    const char* e = std::getenv(name);
    return e ? std::atoi(e) : dflt;
}

static bool dep_enabled() {
    if (const char* e = std::getenv("KLUJAX_CUDA_DEP")) return std::atoi(e) != 0;
    return true;
}

struct DepFactorPlan {
    bool ok = false;
    int scratch_elems = 0, stage_elems = 0, warps = 0;
    size_t smem = 0;
};
static DepFactorPlan dep_factor_plan(int n, int maxcol, size_t vsz) {
    DepFactorPlan p;
    const size_t fl = (static_cast<size_t>(n) + 15) & ~size_t(15);
    p.warps = std::min(16, std::max(1, dep_env_int_early("KLUJAX_CUDA_DEP_FACTOR_WARPS", 16)));
    p.stage_elems = 512;
    size_t budget = 200 * 1024;
    if (fl + static_cast<size_t>(p.warps) * (64 * vsz + p.stage_elems * (vsz + 4)) > budget) return p;
    size_t se = (budget - fl) / p.warps;
    se = (se - p.stage_elems * (vsz + 4)) / vsz;
    se = std::min<size_t>(se, static_cast<size_t>((maxcol + 1) & ~1));
    se &= ~size_t(1);
    p.scratch_elems = static_cast<int>(se);
    p.smem = fl + static_cast<size_t>(p.warps) * ((se + p.stage_elems) * vsz + p.stage_elems * 4);
    p.ok = true;
    return p;
}

// shared memory of panel_solve: right-hand side + temporaries, one staging buffer per warp, the counter
struct PanelPlan {
    bool ok = false;
    int warps = 0;
    size_t smem = 0;
};
static PanelPlan panel_plan(int n, int ntemps, int maxI, size_t vsz) {
    PanelPlan p;
    const int lines = std::max(1, maxI);
    for (int nw = kPanelWarps; nw >= 1; nw >>= 1) {
        const size_t need = (static_cast<size_t>(n) + ntemps + static_cast<size_t>(nw) * lines * 32) * vsz + 16;
        if (need <= kDepSmemMax) {
            p.ok = true;
            p.warps = nw;
            p.smem = need;
            break;
        }
    }
    return p;
}

static int dep_env_int(const char* name, int dflt) {
    const char* e = std::getenv(name);
    return e ? std::atoi(e) : dflt;
}

template <typename K>
static cudaError_t dep_launch(K kernel, int grid, int threads, size_t smem, cudaStream_t st, DepArgs a, int sleep_default) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kDepSmemMax));
    if (e != cudaSuccess) return e;
    a.sleep_ns = static_cast<unsigned>(std::max(0, dep_env_int("KLUJAX_CUDA_DEP_SLEEP", sleep_default)));
    if (dep_env_int("KLUJAX_CUDA_DEP_PROF", 0)) {  // diagnostic: synchronises and prints the phase counters
        unsigned long long* pr = nullptr;
        if (cudaMallocManaged(&pr, 8 * sizeof(unsigned long long)) != cudaSuccess) return cudaGetLastError();
        for (int i = 0; i < 8; ++i) pr[i] = 0;
        a.prof = pr;
        kernel<<<grid, threads, smem, st>>>(a);
        cudaStreamSynchronize(st);
        const double warps = static_cast<double>(grid) * (threads / 32);
        std::fprintf(stderr, "[DEP_PROF] grid %d threads %d: cycles per warp by phase:", grid, threads);
        for (int i = 0; i < 6; ++i) std::fprintf(stderr, " %.0f", pr[i] / warps);
        std::fprintf(stderr, "\n");
        cudaFree(pr);
        return cudaGetLastError();
    }
    kernel<<<grid, threads, smem, st>>>(a);
    return cudaGetLastError();
}

}  // namespace kb
