// kernels_level.cuh — level-scheduled regime (K1, K3, K4, K5, K6, K7) for
// matrices whose factors do not fit in shared memory.
//
// Layout in HBM: values are batch-interleaved, V[slot][system] (and
// X[row][system*n_rhs + column] for the solves), so consecutive threads handle
// consecutive systems / right-hand sides and every load of a value is
// coalesced.  Systems (and right-hand-side columns) never interact, so a CTA owns a
// slice of the batch axis and walks ALL level sets of the static task DAG
// (plan.hpp) on it alone, with __syncthreads() between levels - no grid barrier.
// There is no sparsity logic on the device, only index lists shared by the batch.
//
// Replaces klu_refactor / klu_solve / klu_tsolve as called per system at
// /root/reference/klujax.cpp:781, :911, :1087, :1220 and the b/x transposes at
// :1069-1076, :1095-1101 (there is no transposition pass here: the
// permutations P, Q and the row scaling are folded into the load and the store
// of X).  Also holds the batched COO SpMM twin of dot_* (klujax.cpp:165-249).
//
// Round 2: system-major value layout with a gather fill (lv_prepare_sysmajor) when a CTA owns one
// system; the solves run on the caller's x; lv_solve_grouped blocks the rows of a supernode in
// registers for many right-hand sides (plan.hpp: GroupProgram); when a CTA owns a single
// right-hand-side column the solves go to panel_solve (kernels_dep.cuh).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include <cub/device/device_radix_sort.cuh>

#include "kernels_dep.cuh"
#include "kernels_small.cuh"
#include "plan.hpp"

namespace kb {

struct LevelProg {
    int ntasks = 0, nlevels = 0;
    int4* d_task = nullptr;   // per task 2 x int4: {out, pivot slot (-1 none), index of 2nd pair, plen | kind << 24}
                              //                       {a0, b0 (first operand pair, inline), -, -}
    int2* d_pair = nullptr;   // operand slots
    int* d_level_ptr = nullptr;
    int max_level_tasks = 0;
    int64_t n_mac = 0;
    ~LevelProg() {
        if (d_task) cudaFree(d_task);
        if (d_pair) cudaFree(d_pair);
        if (d_level_ptr) cudaFree(d_level_ptr);
    }
    // xrow (optional, [n]): new name of X row k - the native layout names the rows of X by their
    // position in the caller's x (the output permutation is folded into the program)
    bool upload(const SlotLayout& lay, const TaskProgram& tp, std::string& err, const int* xrow = nullptr) {
        // Solve programs address X by row: the records hold the row itself (index - n_val, renamed
        // by xrow), the kernels never subtract n_val.
        const int nv = lay.n_val;
        const bool solve_prog = tp.mode == PM_SOLVE || tp.mode == PM_TSOLVE;
        auto nm = [&](int idx) {
            if (!solve_prog) return idx;
            return xrow ? xrow[idx - nv] : idx - nv;
        };
        ntasks = static_cast<int>(tp.tasks.size());
        nlevels = tp.nlevels;
        n_mac = tp.n_mac;
        std::vector<int4> rec(2 * static_cast<size_t>(std::max(ntasks, 1)));
        for (int t = 0; t < ntasks; ++t) {
            const Task& k = tp.tasks[t];
            if (k.plen >= (1 << 24)) {
                err = "dot product too long for the level program";
                return false;
            }
            rec[2 * t] = make_int4(nm(k.out), k.div, k.pbeg + 1, k.plen | (static_cast<int>(k.kind) << 24));
            rec[2 * t + 1] = k.plen ? make_int4(tp.pa[k.pbeg], nm(tp.pb[k.pbeg]), 0, 0) : make_int4(0, 0, 0, 0);
        }
        std::vector<int2> pr(std::max<size_t>(tp.pa.size(), 1));
        for (size_t q = 0; q < tp.pa.size(); ++q) pr[q] = make_int2(tp.pa[q], nm(tp.pb[q]));
        for (int l = 0; l < nlevels; ++l)
            max_level_tasks = std::max(max_level_tasks, tp.level_ptr[l + 1] - tp.level_ptr[l]);
        auto up = [&](void** d, const void* src, size_t bytes) {
            if (cudaMalloc(d, std::max<size_t>(bytes, 16)) != cudaSuccess) return false;
            return bytes == 0 || cudaMemcpy(*d, src, bytes, cudaMemcpyHostToDevice) == cudaSuccess;
        };
        if (!up(reinterpret_cast<void**>(&d_task), rec.data(), rec.size() * sizeof(int4)) ||
            !up(reinterpret_cast<void**>(&d_pair), pr.data(), pr.size() * sizeof(int2)) ||
            !up(reinterpret_cast<void**>(&d_level_ptr), tp.level_ptr.data(), tp.level_ptr.size() * sizeof(int))) {
            err = "level program upload failed (out of device memory?)";
            return false;
        }
        return true;
    }
};

struct LevelCall {
    int mode = 0, L = 0, r = 0, nz = 0, n = 0, n_val = 0, is_complex = 0, num_sms = 148, Lstore = 0;
    const void* Ax = nullptr;
    long long ax_stride = 0;
    const void* b = nullptr;
    long long b_stride = 0;
    void* x = nullptr;
    const int *coo_dst = nullptr, *pattern_bad = nullptr, *rs_ptr = nullptr, *rs_slot = nullptr;
    const int *pnum = nullptr, *q = nullptr, *sysidx = nullptr;
    const int *src_solve = nullptr, *src_tsolve = nullptr;  // native layout: row of b behind row j of x
    int32_t* status = nullptr;
    double* growth = nullptr;
    void* V = nullptr;   // [n_val][lpad] batch-interleaved, or [L][n_val] system-major
    double* Rs = nullptr;  // [n][lpad], or [L][n]
    long long lpad = 0;
    int sysmajor = 0;    // layout of V / Rs (see level_run_t)
    int certify = 1;     // a failed pivot-parity certificate is status 2
    cudaStream_t st = nullptr;
    // dependency-driven kernels (kernels_dep.cuh) instead of the level kernels, when set
    const DepColProg* dep_col = nullptr;
    const DepPanelProg* dep_panel = nullptr;
    const struct GroupProg* grp = nullptr;  // multi-right-hand-side solve with register blocking over rows
    // per-stream scratch of the caller for the status words (no allocation on the call path when it is large enough)
    void* aux = nullptr;
    size_t aux_bytes = 0;
    int* slot_src = nullptr;  // per-stream scratch [n_val]: inverse of coo_dst (system-major fill)
};

// scatter Ax[m][e] -> V[slot][sys]; V has been zeroed
template <typename T>
__global__ void lv_scatter(int L, int nz, const typename Sc<T>::V* __restrict__ Ax,
                           long long ax_stride, const int* __restrict__ coo_dst,
                           const int* __restrict__ sysidx, typename Sc<T>::V* __restrict__ V,
                           long long lpad, long long ss) {
    // 32 x 32 tile through shared memory: read along e, write along the system axis
    __shared__ typename Sc<T>::V tile[32][33];
    const int e0 = blockIdx.x * 32, m0 = blockIdx.y * 32;
    for (int dm = threadIdx.y; dm < 32; dm += blockDim.y) {
        const int e = e0 + threadIdx.x, m = m0 + dm;
        if (e < nz && m < L) tile[dm][threadIdx.x] = Ax[(long long)m * ax_stride + e];
    }
    __syncthreads();
    for (int de = threadIdx.y; de < 32; de += blockDim.y) {
        const int e = e0 + de, m = m0 + threadIdx.x;
        if (e < nz && m < L) {
            const int d = __ldg(coo_dst + e);
            const int sm = sysidx ? sysidx[m] : m;
            if (d >= 0) V[(long long)d * lpad + (long long)sm * ss] = tile[threadIdx.x][de];
        }
    }
}

template <typename T>
__global__ void lv_zero_slots(int L, int n_val, const int* __restrict__ sysidx,
                              typename Sc<T>::V* __restrict__ V, long long lpad, long long ss) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)n_val * L;
    for (long long i = tid; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int m = static_cast<int>(i % L);
        const long long s = i / L;
        V[s * lpad + (long long)(sysidx ? sysidx[m] : m) * ss] = Sc<T>::zero();
    }
}

// max-abs row scale factors and scaling, thread per (row, system)
template <typename T>
__global__ void lv_rowscale(int L, int n, const int* __restrict__ rs_ptr,
                            const int* __restrict__ rs_slot, const int* __restrict__ sysidx,
                            typename Sc<T>::V* __restrict__ V, double* __restrict__ Rs,
                            long long lpad, long long ss, long long rlp, long long rss) {
    using S = Sc<T>;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (long long)n * L) return;
    const int m = static_cast<int>(tid % L), i = static_cast<int>(tid / L);
    const int sm = sysidx ? sysidx[m] : m;
    const int pb = rs_ptr[i], pe = rs_ptr[i + 1];
    double mx = 0.0;
    for (int p = pb; p < pe; ++p) mx = fmax(mx, S::mag(V[(long long)rs_slot[p] * lpad + (long long)sm * ss]));
    if (mx == 0.0) mx = 1.0;
    Rs[(long long)i * rlp + (long long)sm * rss] = mx;
    for (int p = pb; p < pe; ++p) {
        typename S::V* q = V + (long long)rs_slot[p] * lpad + (long long)sm * ss;
        *q = S::rdiv(*q, mx);
    }
}

// System-major layout (one CTA per system): fill the system's contiguous value array and row-scale it
// in one kernel.  The fill is a GATHER - V[slot] = Ax[entry behind the slot] or 0 - so every value is
// written exactly once, coalesced along the slot axis; zero-then-scatter wrote the fill-in slots and
// then hit them again with 8-byte partial writes after they had left L2 (1.0 ms per call on cfg 4,
// 1.44 ms as three batch-interleaved kernels).  slot_src is the inverse of the call's COO map.
__global__ void coo_invert_kernel(int nz, const int* __restrict__ coo_dst, int* __restrict__ slot_src) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nz) return;
    const int d = coo_dst[e];
    if (d >= 0) slot_src[d] = e;
}

template <typename T>
__global__ void __launch_bounds__(1024) lv_prepare_sysmajor(int L, int n, int nz, int n_val,
                                                            const typename Sc<T>::V* __restrict__ Ax, long long ax_stride,
                                                            const int* __restrict__ slot_src, const int* __restrict__ rs_ptr,
                                                            const int* __restrict__ rs_slot, const int* __restrict__ sysidx,
                                                            typename Sc<T>::V* __restrict__ V, double* __restrict__ Rs) {
    using S = Sc<T>;
    using V_t = typename S::V;
    for (int m = blockIdx.x; m < L; m += gridDim.x) {
        const long long sm = sysidx ? static_cast<long long>(sysidx[m]) : static_cast<long long>(m);
        V_t* Vs = V + sm * n_val;
        double* Rm = Rs + sm * n;
        const V_t* Am = Ax + static_cast<long long>(m) * ax_stride;
        for (int s0 = threadIdx.x; s0 < n_val; s0 += 4 * blockDim.x) {  // four independent gathers in flight
            int e[4];
            V_t v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) e[u] = s0 + u * blockDim.x < n_val ? __ldg(slot_src + s0 + u * blockDim.x) : -1;
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = e[u] >= 0 ? Am[e[u]] : S::zero();
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (s0 + u * blockDim.x < n_val) Vs[s0 + u * blockDim.x] = v[u];
        }
        __syncthreads();
        // row scaling: the slots and the values of a row are loaded in batches of 8 (independent
        // loads, one memory round trip each) and short rows are scaled out of registers
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int pb = __ldg(rs_ptr + i), pe = __ldg(rs_ptr + i + 1);
            constexpr int RB = 8;
            int sl[RB];
            V_t vv[RB];
#pragma unroll
            for (int u = 0; u < RB; ++u) sl[u] = pb + u < pe ? __ldg(rs_slot + pb + u) : -1;
#pragma unroll
            for (int u = 0; u < RB; ++u) vv[u] = sl[u] >= 0 ? Vs[sl[u]] : S::zero();
            double mx = 0.0;
#pragma unroll
            for (int u = 0; u < RB; ++u) mx = fmax(mx, S::mag(vv[u]));
            for (int p = pb + RB; p < pe; ++p) mx = fmax(mx, S::mag(Vs[__ldg(rs_slot + p)]));
            if (mx == 0.0) mx = 1.0;
            Rm[i] = mx;
#pragma unroll
            for (int u = 0; u < RB; ++u)
                if (sl[u] >= 0) Vs[sl[u]] = S::rdiv(vv[u], mx);
            for (int p = pb + RB; p < pe; ++p) {
                V_t* q = Vs + __ldg(rs_slot + p);
                *q = S::rdiv(*q, mx);
            }
        }
        __syncthreads();
    }
}

struct LvArgs {
    const int4* task;
    const int2* pair;
    const int* level_ptr;
    int nlevels, n_val, r, slice;   // slice: batch columns owned by one CTA
    long long L, lpad, Q;           // Q = L * r right-hand-side columns in the solve
    long long vss;                  // V[slot * lpad + system * vss]
    long long xrs, xcs;             // X[row * xrs + column * xcs]
    void* V;
    void* X;
    const int* sysidx;
    int* bad;                       // [L] OR of singular pivots
    unsigned long long* growth2;    // [L] max |L|^2 as ordered bits
    int gsz;                        // cooperative kernels: CTAs that share one batch column
    unsigned* bar;                  // cooperative kernels: [columns] arrival counters (zeroed per launch)
    // native layout of the solve (X = the caller's x[system][row][column], no relayout pass):
    int native, n, lp_smem;         // lp_smem: the level table fits the dynamic shared memory of the launch
    long long xss;                  // X[system * xss + row * xrs + column]
    const void* B;                  // right-hand sides b[system * b_stride + row * r + column]
    long long b_stride;
    const int* src_of;              // row of b that output row j starts from (both permutations composed)
    const double* Rs;               // row scale factors Rs[row * rlp + system * rss]
    long long rlp, rss;
    int scale_in, scale_out;        // solve: b / Rs on the way in; transpose solve: x / Rs on the way out
};

// ---- cooperative variants: several CTAs (SMs) per batch column --------------------------
// With fewer batch columns than SMs (cfg 1: ONE system) a column gets gsz CTAs; the items of a
// level are dealt over all their threads and the barrier between levels is an arrival counter
// in global memory (the kernels are launched with cudaLaunchCooperativeKernel, so all CTAs
// are resident).  Values written by another CTA are read with ld.global.cg (L2, never a
// stale L1 line).
__device__ __forceinline__ void group_barrier(unsigned* ctr, unsigned nblk, unsigned& target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += nblk;
        __threadfence();
        atomicAdd(ctr, 1u);
        while (*reinterpret_cast<volatile unsigned*>(ctr) < target) __nanosleep(20);
        __threadfence();
    }
    __syncthreads();
}

template <typename T>
__global__ void __launch_bounds__(1024) lv_factor_coop(const LvArgs a) {
    using S = Sc<T>;
    using V_t = typename S::V;
    V_t* V = static_cast<V_t*>(a.V);
    const int grp = blockIdx.x / a.gsz, rank = blockIdx.x % a.gsz;
    const int m = grp;  // one system per group of CTAs
    if (m >= a.L) return;
    const long long sm = a.sysidx ? a.sysidx[m] : m;
    unsigned target = 0;
    unsigned* ctr = a.bar + grp;
    const int stride = a.gsz * blockDim.x, first = rank * blockDim.x + threadIdx.x;
    int te = __ldg(a.level_ptr);
    double gmax = 0.0;
    for (int l = 0; l < a.nlevels; ++l) {
        const int tb = te;
        te = __ldg(a.level_ptr + l + 1);
        for (int t = tb + first; t < te; t += stride) {
            const int4 rec = __ldg(a.task + 2 * t), fp = __ldg(a.task + 2 * t + 1);
            const int plen = rec.w & 0xFFFFFF, kind = rec.w >> 24;
            const int2* pr = a.pair + rec.z;
            V_t* o = V + (long long)rec.x * a.lpad + sm * a.vss;
            const V_t o0 = __ldcg(o);
            V_t acc = S::zero();
            if (plen)
                S::fma_acc(acc, __ldcg(V + (long long)fp.x * a.lpad + sm * a.vss),
                           __ldcg(V + (long long)fp.y * a.lpad + sm * a.vss));
#pragma unroll 4
            for (int p = 0; p < plen - 1; ++p) {
                const int2 ab = __ldg(pr + p);
                S::fma_acc(acc, __ldcg(V + (long long)ab.x * a.lpad + sm * a.vss),
                           __ldcg(V + (long long)ab.y * a.lpad + sm * a.vss));
            }
            V_t v = S::sub(o0, acc);
            if (kind == TK_RECIP) {
                if (S::bad_pivot(v)) a.bad[m] = 1;
                v = S::recip(v);
            } else if (kind >= TK_MUL) {
                v = S::mul(v, __ldcg(V + (long long)rec.y * a.lpad + sm * a.vss));
                if (kind >= TK_MUL_TRACK)  // pivot-parity certificate, weighted by class (plan.hpp)
                    gmax = fmax(gmax, S::mag2(v) * KB_TRACK_W2(kind - TK_MUL_TRACK));
            }
            *o = v;
        }
        group_barrier(ctr, a.gsz, target);
    }
    if (gmax > 0.0) atomicMax(a.growth2 + m, (unsigned long long)__double_as_longlong(gmax));  // one per thread, not per entry
}

template <typename T>
__global__ void __launch_bounds__(1024) lv_solve_coop(const LvArgs a) {
    using S = Sc<T>;
    using V_t = typename S::V;
    const int grp = blockIdx.x / a.gsz, rank = blockIdx.x % a.gsz;
    const long long q = grp;  // one right-hand-side column per group of CTAs
    if (q >= a.Q) return;
    const int m = static_cast<int>(q / a.r);
    const long long sm = a.sysidx ? a.sysidx[m] : m;
    const V_t* vb = static_cast<const V_t*>(a.V) + sm * a.vss;
    V_t* xb = static_cast<V_t*>(a.X) + (a.native ? (long long)m * a.xss + (q - (long long)m * a.r) : q * a.xcs);
    unsigned target = 0;
    unsigned* ctr = a.bar + grp;
    const int stride = a.gsz * blockDim.x, first = rank * blockDim.x + threadIdx.x;
    int te = __ldg(a.level_ptr);
    for (int l = 0; l < a.nlevels; ++l) {
        const int tb = te;
        te = __ldg(a.level_ptr + l + 1);
        for (int t = tb + first; t < te; t += stride) {
            const int4 rec = __ldg(a.task + 2 * t), fp = __ldg(a.task + 2 * t + 1);
            const int plen = rec.w & 0xFFFFFF, kind = rec.w >> 24;
            const int2* pr = a.pair + rec.z;
            V_t* o = xb + (long long)rec.x * a.xrs;
            const V_t o0 = __ldcg(o);
            V_t acc = S::zero();
            if (plen)
                S::fma_acc(acc, __ldg(vb + (long long)fp.x * a.lpad), __ldcg(xb + (long long)fp.y * a.xrs));
#pragma unroll 4
            for (int p = 0; p < plen - 1; ++p) {
                const int2 ab = __ldg(pr + p);
                S::fma_acc(acc, __ldg(vb + (long long)ab.x * a.lpad), __ldcg(xb + (long long)ab.y * a.xrs));
            }
            V_t v = S::sub(o0, acc);
            if (kind >= TK_MUL) v = S::mul(v, __ldg(vb + (long long)rec.y * a.lpad));
            *o = v;
        }
        group_barrier(ctr, a.gsz, target);
    }
}

// The level table in shared memory (one 30-cycle load per level instead of a global round trip
// at the top of every level) when it fits the launch's dynamic shared memory.
__device__ __forceinline__ const int* stage_level_table(const LvArgs& a, int* smem) {
    if (!a.lp_smem) return a.level_ptr;
    for (int i = threadIdx.x; i <= a.nlevels; i += blockDim.x) smem[i] = __ldg(a.level_ptr + i);
    __syncthreads();
    return smem;
}

// Refactorization.  A CTA owns a slice of `ncol` systems and walks ALL levels on it alone
// (systems never interact: the barrier between levels is __syncthreads(), not a grid barrier,
// and a CTA only ever reads values it wrote itself - same SM, coherent L1).  Inside the CTA a
// thread is (column lane, task group): its system, and with it the base address of its values,
// is fixed for the whole slice, the groups deal the tasks of a level between them - no division
// and no 64-bit index arithmetic in the loops.
template <typename T>
__global__ void __launch_bounds__(1024) lv_factor(const LvArgs a) {
    using S = Sc<T>;
    using V_t = typename S::V;
    extern __shared__ int lv_smem[];
    const int* lp = stage_level_table(a, lv_smem);
    const int lanes = a.slice, G = blockDim.x / lanes;
    const int g = threadIdx.x / lanes, cl = threadIdx.x - g * lanes;
    for (long long s0 = (long long)blockIdx.x * a.slice; s0 < a.L; s0 += (long long)gridDim.x * a.slice) {
        const bool on = g < G && s0 + cl < a.L;
        const int m = on ? static_cast<int>(s0) + cl : 0;
        V_t* vb = static_cast<V_t*>(a.V) + (a.sysidx ? (long long)a.sysidx[m] : (long long)m) * a.vss;
        const unsigned lpb = static_cast<unsigned>(a.lpad * sizeof(V_t));  // see lv_solve
        auto vp = [&](int slot) {
            return reinterpret_cast<V_t*>(reinterpret_cast<char*>(vb) + (unsigned long long)static_cast<unsigned>(slot) * lpb);
        };
        // The task record of a thread's first item of the NEXT level is fetched before the
        // barrier, so a level costs one memory round trip (the operands), not two.
        int te = lp[0], te_next = lp[1];
        int4 pre_rec = make_int4(0, 0, 0, 0), pre_first = pre_rec;
        if (on && te + g < te_next) {
            pre_rec = __ldg(a.task + 2 * (te + g));
            pre_first = __ldg(a.task + 2 * (te + g) + 1);
        }
        double gmax = 0.0;  // running maximum of the certificate
        for (int l = 0; l < a.nlevels; ++l) {
            const int tb = te;
            te = te_next;
            te_next = (l + 1 < a.nlevels) ? lp[l + 2] : te;
            const int4 cur_rec = pre_rec, cur_first = pre_first;
            if (on && te + g < te_next) {
                pre_rec = __ldg(a.task + 2 * (te + g));
                pre_first = __ldg(a.task + 2 * (te + g) + 1);
            }
            if (on)
                for (int t = tb + g; t < te; t += G) {
                    const bool pre = t == tb + g;
                    const int4 rec = pre ? cur_rec : __ldg(a.task + 2 * t);
                    const int4 first = pre ? cur_first : __ldg(a.task + 2 * t + 1);
                    const int plen = rec.w & 0xFFFFFF, kind = rec.w >> 24;
                    const int2* pr = a.pair + rec.z;
                    V_t* o = vp(rec.x);
                    const V_t o0 = *o;
                    V_t acc = S::zero();
                    if (plen) S::fma_acc(acc, *vp(first.x), *vp(first.y));
#pragma unroll 4
                    for (int p = 0; p < plen - 1; ++p) {
                        const int2 ab = __ldg(pr + p);
                        const V_t x = *vp(ab.x);
                        const V_t y = *vp(ab.y);
                        S::fma_acc(acc, x, y);
                    }
                    V_t v = S::sub(o0, acc);
                    if (kind == TK_RECIP) {
                        if (S::bad_pivot(v)) a.bad[m] = 1;
                        v = S::recip(v);
                    } else if (kind >= TK_MUL) {
                        v = S::mul(v, *vp(rec.y));
                        if (kind >= TK_MUL_TRACK)  // pivot-parity certificate, weighted by class (plan.hpp)
                            gmax = fmax(gmax, S::mag2(v) * KB_TRACK_W2(kind - TK_MUL_TRACK));
                    }
                    *o = v;
                }
            __syncthreads();
        }
        // ONE atomic per thread and system: an atomicMax per entry of L (155 k per system on cfg 4, all
        // to one address) cost a third of the kernel (refactor 9.6 -> 6.1 ms)
        if (on && gmax > 0.0) atomicMax(a.growth2 + m, (unsigned long long)__double_as_longlong(gmax));
    }
}

// C adjacent values as 16-byte vector accesses (C * sizeof(V) is a multiple of 16 when C > 1)
template <typename V_t, int C>
__device__ __forceinline__ void ld_cols(const V_t* p, V_t (&v)[C]) {
    if constexpr (C * sizeof(V_t) % 16 == 0) {
        constexpr int NV = C * sizeof(V_t) / 16;
        double2 t[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) t[i] = reinterpret_cast<const double2*>(p)[i];
        memcpy(v, t, sizeof(t));
    } else {
#pragma unroll
        for (int j = 0; j < C; ++j) v[j] = p[j];
    }
}
template <typename V_t, int C>
__device__ __forceinline__ void st_cols(V_t* p, const V_t (&v)[C]) {
    if constexpr (C * sizeof(V_t) % 16 == 0) {
        constexpr int NV = C * sizeof(V_t) / 16;
        double2 t[NV];
        memcpy(t, v, sizeof(t));
#pragma unroll
        for (int i = 0; i < NV; ++i) reinterpret_cast<double2*>(p)[i] = t[i];
    } else {
#pragma unroll
        for (int j = 0; j < C; ++j) p[j] = v[j];
    }
}

// Triangular solves.  Right-hand-side columns (system m, column c: q = m*r + c) never interact
// either: a CTA owns a slice of q and walks all levels on it, __syncthreads() between levels.
// A thread is (column lane, task group) and owns C ADJACENT columns (C divides r, so they belong
// to one system and are one aligned vector access): the task record, the operand indices and the
// entry of L or U are loaded once and used for C right-hand sides.  X is either the
// batch-interleaved scratch (X[row][q], many systems with few right-hand sides each; C = 1) or -
// native - the caller's x itself: the load b -> x with both permutations composed (src_of) and the
// row scaling is the prologue of the slice, the scaling of the transpose solve its epilogue, and
// the rows of X are named by their position in x, so nothing is relaid out and no scratch is
// allocated.  PRE: prefetch of the next level's record as in lv_factor (one column per CTA; with
// wide slices a thread runs many items per level and the extra live registers cost more than
// they save: measured 3x slower on the multi-RHS solve of cfg 3).
template <typename T, int C, bool PRE>
__global__ void __launch_bounds__(C >= 8 ? 768 : 1024) lv_solve(const LvArgs a) {
    using S = Sc<T>;
    using V_t = typename S::V;
    extern __shared__ int lv_smem[];
    const int* lp = stage_level_table(a, lv_smem);
    const int lanes = a.slice / C, G = blockDim.x / lanes;
    const int g = threadIdx.x / lanes, cl = threadIdx.x - g * lanes;
    for (long long s0 = (long long)blockIdx.x * a.slice; s0 < a.Q; s0 += (long long)gridDim.x * a.slice) {
        const long long q0 = s0 + (long long)cl * C;
        const bool ok = g < G && q0 < a.Q;  // Q is a multiple of C: all C columns or none
        const int m0 = static_cast<int>((ok ? q0 : 0) / a.r);
        const int c0 = ok ? static_cast<int>(q0 - (long long)m0 * a.r) : 0;
        const long long sm = a.sysidx ? (long long)a.sysidx[m0] : (long long)m0;
        const V_t* vb = static_cast<const V_t*>(a.V) + sm * a.vss;
        V_t* xb = static_cast<V_t*>(a.X) + (a.native ? (long long)m0 * a.xss + c0 : (ok ? q0 : 0) * a.xcs);
        const V_t* bb = static_cast<const V_t*>(a.B) + (long long)m0 * a.b_stride + c0;
        // addresses = 64-bit base + 32-bit index x 32-bit byte stride: one IMAD.WIDE.U32 each (the
        // host checks that the strides fit; a 64-bit index product costs ~10 instructions)
        const unsigned xrb = static_cast<unsigned>(a.xrs * sizeof(V_t)), lpb = static_cast<unsigned>(a.lpad * sizeof(V_t));
        auto xp = [&](int row) {
            return reinterpret_cast<V_t*>(reinterpret_cast<char*>(xb) + (unsigned long long)static_cast<unsigned>(row) * xrb);
        };
        auto vp = [&](int slot) {
            return reinterpret_cast<const V_t*>(reinterpret_cast<const char*>(vb) +
                                                (unsigned long long)static_cast<unsigned>(slot) * lpb);
        };
        if (a.native) {
            if (ok)
                for (int row = g; row < a.n; row += G) {
                    const int src = __ldg(a.src_of + row);
                    V_t v[C];
                    ld_cols<V_t, C>(bb + (long long)src * a.r, v);
                    if (a.scale_in) {
                        const double sc = a.Rs[(long long)src * a.rlp + sm * a.rss];
#pragma unroll
                        for (int j = 0; j < C; ++j) v[j] = S::rdiv(v[j], sc);
                    }
                    st_cols<V_t, C>(xp(row), v);
                }
            __syncthreads();
        }
        int te = lp[0], te_next = a.nlevels > 0 ? lp[1] : te;
        int4 pre_rec = make_int4(0, 0, 0, 0), pre_first = pre_rec;
        if (PRE && ok && te + g < te_next) {
            pre_rec = __ldg(a.task + 2 * (te + g));
            pre_first = __ldg(a.task + 2 * (te + g) + 1);
        }
        for (int l = 0; l < a.nlevels; ++l) {
            const int tb = te;
            te = te_next;
            te_next = (l + 1 < a.nlevels) ? lp[l + 2] : te;
            const int4 cur_rec = pre_rec, cur_first = pre_first;
            if (PRE && ok && te + g < te_next) {
                pre_rec = __ldg(a.task + 2 * (te + g));
                pre_first = __ldg(a.task + 2 * (te + g) + 1);
            }
            if (ok)
                for (int t = tb + g; t < te; t += G) {
                    const bool pre = PRE && t == tb + g;
                    const int4 rec = pre ? cur_rec : __ldg(a.task + 2 * t);
                    const int4 first = pre ? cur_first : __ldg(a.task + 2 * t + 1);
                    const int plen = rec.w & 0xFFFFFF, kind = rec.w >> 24;
                    const int2* pr = a.pair + rec.z;
                    V_t* o = xp(rec.x);
                    V_t acc[C], xv[C];
#pragma unroll
                    for (int j = 0; j < C; ++j) acc[j] = S::zero();
                    if (plen) {
                        const V_t lu = __ldg(vp(first.x));
                        ld_cols<V_t, C>(xp(first.y), xv);
#pragma unroll
                        for (int j = 0; j < C; ++j) S::fma_acc(acc[j], lu, xv[j]);
                    }
#pragma unroll 4
                    for (int p = 0; p < plen - 1; ++p) {
                        const int2 ab = __ldg(pr + p);
                        const V_t lu = __ldg(vp(ab.x));
                        ld_cols<V_t, C>(xp(ab.y), xv);
#pragma unroll
                        for (int j = 0; j < C; ++j) S::fma_acc(acc[j], lu, xv[j]);
                    }
                    V_t dv = S::one();
                    if (kind >= TK_MUL) dv = __ldg(vp(rec.y));
                    ld_cols<V_t, C>(o, xv);
#pragma unroll
                    for (int j = 0; j < C; ++j) {
                        xv[j] = S::sub(xv[j], acc[j]);
                        if (kind >= TK_MUL) xv[j] = S::mul(xv[j], dv);
                    }
                    st_cols<V_t, C>(o, xv);
                }
            __syncthreads();
        }
        if (a.native && a.scale_out && ok) {
            for (int row = g; row < a.n; row += G) {
                const double sc = a.Rs[(long long)row * a.rlp + sm * a.rss];
                V_t v[C];
                ld_cols<V_t, C>(xp(row), v);
#pragma unroll
                for (int j = 0; j < C; ++j) v[j] = S::rdiv(v[j], sc);
                st_cols<V_t, C>(xp(row), v);
            }
        }
    }
}

// ---- multi-right-hand-side solves with register blocking over rows (plan.hpp: GroupProgram) ----
struct GroupProg {
    int n = 0, ngroups = 0, nlevels = 0;
    int64_t n_mac = 0, n_ksteps = 0;
    int4 *d_ghdr = nullptr, *d_krec = nullptr;
    int *d_gaux = nullptr, *d_level_ptr = nullptr;
    ~GroupProg() {
        if (d_ghdr) cudaFree(d_ghdr);
        if (d_krec) cudaFree(d_krec);
        if (d_gaux) cudaFree(d_gaux);
        if (d_level_ptr) cudaFree(d_level_ptr);
    }
    bool upload(const GroupProgram& gp) {
        n = gp.n;
        ngroups = gp.ngroups;
        nlevels = gp.nlevels;
        n_mac = gp.n_mac;
        n_ksteps = gp.n_ksteps;
        auto up = [&](void** d, const std::vector<int>& v) {
            if (cudaMalloc(d, std::max<size_t>(v.size() * sizeof(int), 16)) != cudaSuccess) return false;
            return v.empty() || cudaMemcpy(*d, v.data(), v.size() * sizeof(int), cudaMemcpyHostToDevice) == cudaSuccess;
        };
        return up(reinterpret_cast<void**>(&d_ghdr), gp.ghdr) && up(reinterpret_cast<void**>(&d_krec), gp.krec) &&
               up(reinterpret_cast<void**>(&d_gaux), gp.gaux) && up(reinterpret_cast<void**>(&d_level_ptr), gp.level_ptr);
    }
};

constexpr int kGroupKC = 16;  // k-steps staged per chunk
template <typename V_t, int C>
constexpr size_t grouped_warp_bytes() {
    return (static_cast<size_t>(kGroupKC) * 32 * C + static_cast<size_t>(kGroupKC) * kGroupRows + 36) * sizeof(V_t) +
           (static_cast<size_t>(kGroupKC) * 12 + 16) * sizeof(int);
}
struct GrpArgs {
    const int4* ghdr;
    const int4* krec;   // 3 per k-step
    const int* gaux;    // kGroupAux per group
    const int* level_ptr;
    int nlevels, n, r;
    long long Q, lpad, vss, xss, b_stride, rlp, rss;
    const void* V;
    void* X;
    const void* B;
    const int* sysidx;
    const int* src_of;
    const double* Rs;
    int scale_in, scale_out;
};

// A CTA owns 32*C adjacent right-hand-side columns of ONE system (C divides r, 32*C divides r or the
// tail lanes idle); a warp is one group task: lane = C adjacent columns, the accumulators of the up to
// 8 rows of the group live in registers, a row of X is read once per k-step and applied to every row
// of the group that has an entry there (warp-uniform branches: the slots are the same for all lanes).
// The k-loop is software-pipelined: the loads of step k+1 (row of X, up to 8 entries of L / U) are in
// flight while step k is applied, the record of step k+2 while those are issued.
template <typename T, int C>
__global__ void __launch_bounds__(C * sizeof(typename Sc<T>::V) >= 32 ? 320 : 640) lv_solve_grouped(const GrpArgs a) {
    using S = Sc<T>;
    using V_t = typename S::V;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, NW = blockDim.x >> 5;
    constexpr int KC = kGroupKC;
    extern __shared__ __align__(16) unsigned char grp_smem[];
    constexpr size_t kWarpBytes = grouped_warp_bytes<V_t, C>();
    unsigned char* wbase = grp_smem + static_cast<size_t>(w) * kWarpBytes;
    V_t* sx = reinterpret_cast<V_t*>(wbase);                                         // [KC][32 * C]
    V_t* scf = sx + KC * 32 * C;                                                      // [KC][8]
    V_t* sin = scf + KC * kGroupRows;                                                 // [36]: pivots [8], in-group entries [28]
    int* srec = reinterpret_cast<int*>(sin + 36);                                     // [KC][12]
    int* srow = srec + KC * 12;                                                       // [8]: row names of the task (keeps them out of registers)
    const int cols_per_cta = 32 * C;
    const long long slices_per_sys = (a.r + cols_per_cta - 1) / cols_per_cta;
    const long long nslices = static_cast<long long>(a.Q / a.r) * slices_per_sys;
    for (long long sl = blockIdx.x; sl < nslices; sl += gridDim.x) {
        const int m0 = static_cast<int>(sl / slices_per_sys);
        const int c0 = static_cast<int>(sl - static_cast<long long>(m0) * slices_per_sys) * cols_per_cta + lane * C;
        const bool ok = c0 < a.r;  // r is a multiple of C: all C columns or none
        const long long sm = a.sysidx ? static_cast<long long>(a.sysidx[m0]) : static_cast<long long>(m0);
        const V_t* vb = static_cast<const V_t*>(a.V) + sm * a.vss;
        V_t* xb = static_cast<V_t*>(a.X) + static_cast<long long>(m0) * a.xss + (ok ? c0 : 0);
        const V_t* bb = static_cast<const V_t*>(a.B) + static_cast<long long>(m0) * a.b_stride + (ok ? c0 : 0);
        auto xp = [&](int row) { return xb + static_cast<long long>(row) * a.r; };
        auto vp = [&](int slot) { return vb + static_cast<long long>(slot) * a.lpad; };
        if (ok)
            for (int row = w; row < a.n; row += NW) {
                const int src = __ldg(a.src_of + row);
                V_t v[C];
                ld_cols<V_t, C>(bb + static_cast<long long>(src) * a.r, v);
                if (a.scale_in) {
                    const double sc = a.Rs[static_cast<long long>(src) * a.rlp + sm * a.rss];
#pragma unroll
                    for (int j = 0; j < C; ++j) v[j] = S::rdiv(v[j], sc);
                }
                st_cols<V_t, C>(xp(row), v);
            }
        __syncthreads();
        int te = __ldg(a.level_ptr);
        for (int l = 0; l < a.nlevels; ++l) {
            const int tb = te;
            te = __ldg(a.level_ptr + l + 1);
            {   // ALL lanes walk the tasks (the staging uses warp-wide synchronisation): lanes beyond the last
                // column of the system compute on column 0 and never store
                // header + row names of a task are three 16-byte words; those of the warp's NEXT task of
                // the level are fetched while the current one runs
                int4 hn0 = make_int4(0, 0, 0, 0), hn1 = hn0, hn2 = hn0;
                if (tb + w < te) hn0 = __ldg(a.ghdr + 3 * (tb + w)), hn1 = __ldg(a.ghdr + 3 * (tb + w) + 1), hn2 = __ldg(a.ghdr + 3 * (tb + w) + 2);
                for (int t = tb + w; t < te; t += NW) {
                    const int4 h = hn0, q1 = hn1, q2 = hn2;
                    if (t + NW < te) hn0 = __ldg(a.ghdr + 3 * (t + NW)), hn1 = __ldg(a.ghdr + 3 * (t + NW) + 1), hn2 = __ldg(a.ghdr + 3 * (t + NW) + 2);
                    const int nk = h.x, nt = h.z & 255, mulbits = (h.z >> 8) & 255;
                    const bool partial = (h.z >> 30) & 1;
                    const int4* kr = a.krec + 3 * static_cast<long long>(h.y);
                    const int* aux = a.gaux + static_cast<long long>(kGroupAux) * h.w;
                    const int rows[kGroupRows] = {q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
                    __syncwarp();
                    if (lane == 0) {
                        *reinterpret_cast<int4*>(srow) = q1;
                        *reinterpret_cast<int4*>(srow + 4) = q2;
                    }
                    V_t acc[kGroupRows][C];
#pragma unroll
                    for (int i = 0; i < kGroupRows; ++i) {
                        if (i < nt) ld_cols<V_t, C>(xp(rows[i]), acc[i]);
                    }
                    if (!partial) {  // pivots and in-group entries of the final task: in flight with the first chunk
                        __syncwarp();
                        for (int q = lane; q < 36; q += 32) {
                            const int sl1 = __ldg(aux + 8 + q);
                            if (sl1 >= 0) {
                                if constexpr (sizeof(V_t) == 16) cp_async<16>(sin + q, vp(sl1));
                                else cp_async<8>(sin + q, vp(sl1));
                            } else {
                                sin[q] = q < 8 ? S::one() : S::zero();
                            }
                        }
                        if (nk == 0) {
                            asm volatile("cp.async.wait_all;" ::: "memory");
                            __syncwarp();
                        }
                    }
                    // k-steps in chunks of KC through the warp's shared-memory staging area: the records
                    // (coalesced), then ALL rows of X and entries of L / U of the chunk with cp.async - one
                    // memory round trip per chunk instead of one per k-step
                    for (int k0 = 0; k0 < nk; k0 += KC) {
                        const int kc = min(KC, nk - k0);
                        __syncwarp();
                        for (int q = lane; q < 3 * kc; q += 32) reinterpret_cast<int4*>(srec)[q] = __ldg(kr + 3 * k0 + q);
                        __syncwarp();
                        for (int k = 0; k < kc; ++k) {
                            const V_t* src = xp(srec[12 * k]);
                            V_t* dst = sx + (k * 32 + lane) * C;
                            if constexpr (C * sizeof(V_t) % 16 == 0) {
#pragma unroll
                                for (int q = 0; q < static_cast<int>(C * sizeof(V_t) / 16); ++q)
                                    cp_async<16>(reinterpret_cast<char*>(dst) + 16 * q, reinterpret_cast<const char*>(src) + 16 * q);
                            } else {
                                cp_async<8>(dst, src);
                            }
                        }
                        for (int q = lane; q < kGroupRows * kc; q += 32) {
                            const int sl1 = srec[12 * (q >> 3) + 1 + (q & 7)];
                            if (sl1 >= 0) {
                                if constexpr (sizeof(V_t) == 16) cp_async<16>(scf + q, vp(sl1));
                                else cp_async<8>(scf + q, vp(sl1));
                            }
                        }
                        asm volatile("cp.async.wait_all;" ::: "memory");
                        __syncwarp();
                        for (int k = 0; k < kc; ++k) {
                            const int mask = srec[12 * k + 9];
                            V_t xv[C];
                            ld_cols<V_t, C>(sx + (k * 32 + lane) * C, xv);
#pragma unroll
                            for (int i = 0; i < kGroupRows; ++i)
                                if (mask & (1 << i)) {
                                    const V_t cfi = scf[k * kGroupRows + i];
#pragma unroll
                                    for (int j = 0; j < C; ++j) S::fms_acc(acc[i][j], cfi, xv[j]);
                                }
                        }
                    }
                    if (partial) {  // a chunk of the group's k-steps applied early: X[rows] -= sum, in place
#pragma unroll
                        for (int i = 0; i < kGroupRows; ++i)
                            if (i < nt && ok) st_cols<V_t, C>(xp(srow[i]), acc[i]);
                        continue;
                    }
                    // rows of the group, in order: finish, store, feed the later rows of the group
#pragma unroll
                    for (int i = 0; i < kGroupRows; ++i) {
                        if (i < nt) {
                            if ((mulbits >> i) & 1) {
                                const V_t dv = sin[i];
#pragma unroll
                                for (int j = 0; j < C; ++j) acc[i][j] = S::mul(acc[i][j], dv);
                            }
                            if (ok) st_cols<V_t, C>(xp(srow[i]), acc[i]);
#pragma unroll
                            for (int jr = i + 1; jr < kGroupRows; ++jr) {
                                if (jr < nt) {
                                    const V_t lu = sin[8 + jr * (jr - 1) / 2 + i];  // 0 where row jr has no entry at row i
#pragma unroll
                                    for (int j = 0; j < C; ++j) S::fms_acc(acc[jr][j], lu, acc[i][j]);
                                }
                            }
                        }
                    }
                }
            }
            __syncthreads();
        }
        if (a.scale_out && ok) {
            for (int row = w; row < a.n; row += NW) {
                const double sc = a.Rs[static_cast<long long>(row) * a.rlp + sm * a.rss];
                V_t v[C];
                ld_cols<V_t, C>(xp(row), v);
#pragma unroll
                for (int j = 0; j < C; ++j) v[j] = S::rdiv(v[j], sc);
                st_cols<V_t, C>(xp(row), v);
            }
        }
        __syncthreads();
    }
}

// native layout for the cooperative solve (several CTAs per column: the prologue cannot live in
// the kernel without one more group barrier): x[m][j][c] = b[m][src_of[j]][c] (/ Rs), and the
// scaling of the transpose solve on the way out
template <typename T>
__global__ void lv_load_native(int L, int n, int r, const typename Sc<T>::V* __restrict__ b, long long b_stride,
                               const int* __restrict__ src_of, int scale, const double* __restrict__ Rs,
                               long long rlp, long long rss, const int* __restrict__ sysidx,
                               typename Sc<T>::V* __restrict__ x) {
    using S = Sc<T>;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (long long)L * n * r) return;
    const int c = static_cast<int>(tid % r);
    const long long t2 = tid / r;
    const int j = static_cast<int>(t2 % n), m = static_cast<int>(t2 / n);
    const int src = src_of[j];
    typename S::V v = b[(long long)m * b_stride + (long long)src * r + c];
    if (scale) v = S::rdiv(v, Rs[(long long)src * rlp + (long long)(sysidx ? sysidx[m] : m) * rss]);
    x[tid] = v;
}
template <typename T>
__global__ void lv_scale_native(int L, int n, int r, const double* __restrict__ Rs, long long rlp, long long rss,
                                const int* __restrict__ sysidx, typename Sc<T>::V* __restrict__ x) {
    using S = Sc<T>;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (long long)L * n * r) return;
    const long long t2 = tid / r;
    const int j = static_cast<int>(t2 % n), m = static_cast<int>(t2 / n);
    x[tid] = S::rdiv(x[tid], Rs[(long long)j * rlp + (long long)(sysidx ? sysidx[m] : m) * rss]);
}

// X[k][m*r+c] = b[m][perm[k]][c] (/ Rs[perm[k]][m])
template <typename T>
__global__ void lv_load_x(int L, int n, int r, const typename Sc<T>::V* __restrict__ b,
                          long long b_stride, const int* __restrict__ perm, int scale,
                          const double* __restrict__ Rs, long long rlp, long long rss,
                          const int* __restrict__ sysidx, typename Sc<T>::V* __restrict__ X,
                          long long xrs, long long xcs) {
    using S = Sc<T>;
    const long long Q = (long long)L * r;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= Q * n) return;
    const long long q = tid % Q;
    const int k = static_cast<int>(tid / Q);
    const int m = static_cast<int>(q / r), c = static_cast<int>(q % r);
    const int src = perm[k];
    typename S::V v = b[(long long)m * b_stride + (long long)src * r + c];
    if (scale) v = S::rdiv(v, Rs[(long long)src * rlp + (long long)(sysidx ? sysidx[m] : m) * rss]);
    X[(long long)k * xrs + q * xcs] = v;
}

template <typename T>
__global__ void lv_store_x(int L, int n, int r, typename Sc<T>::V* __restrict__ x,
                           const int* __restrict__ perm, int scale, const double* __restrict__ Rs,
                           long long rlp, long long rss, const int* __restrict__ sysidx,
                           const typename Sc<T>::V* __restrict__ X, long long xrs, long long xcs) {
    using S = Sc<T>;
    const long long Q = (long long)L * r;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= Q * n) return;
    const long long q = tid % Q;
    const int k = static_cast<int>(tid / Q);
    const int m = static_cast<int>(q / r), c = static_cast<int>(q % r);
    const int dst = perm[k];
    typename S::V v = X[(long long)k * xrs + q * xcs];
    if (scale) v = S::rdiv(v, Rs[(long long)dst * rlp + (long long)(sysidx ? sysidx[m] : m) * rss]);
    x[((long long)m * n + dst) * r + c] = v;
}

__global__ void lv_finish_status(int L, const int* __restrict__ bad,
                                 const unsigned long long* __restrict__ growth2,
                                 const int* __restrict__ pattern_bad, int32_t* status,
                                 double* growth, int certify) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= L) return;
    const double g2 = __longlong_as_double((long long)growth2[m]);
    int st = bad[m] ? 1 : ((g2 <= kGrowthLimit2 || !certify) ? 0 : KLUJAX_PIVOT_DRIFT);
    if (pattern_bad && *pattern_bad) st = -3;
    if (status) status[m] = st;
    if (growth) growth[m] = sqrt(g2);
}

// Batch columns per CTA: one column each while the batch is smaller than the machine
// (latency-bound: maximise the number of independent CTAs), else full warps of columns.
static int pick_slice(long long ncols, int num_sms) {
    if (const char* e = std::getenv("KLUJAX_CUDA_LEVEL_SLICE")) return std::max(1, std::atoi(e));
    if (ncols <= num_sms) return 1;
    const long long per = (ncols + num_sms - 1) / num_sms;
    if (per < 32) return static_cast<int>(per);
    // 64 columns per CTA: measured best on the multi-RHS solve of cfg 3 (273 ms vs 288 at 32 and
    // 369 at 128) - narrower slices keep more rows of X of a CTA in L1/L2
    return static_cast<int>(std::min<long long>(64, (per + 31) / 32 * 32));
}

// System-major V (one system per CTA) when the batch is not larger than the machine.
static int want_sysmajor(long long L, int num_sms) {
    if (const char* e = std::getenv("KLUJAX_CUDA_SYSMAJOR")) return std::atoi(e) != 0;
    return L <= num_sms;
}

// CTAs per batch column for the cooperative kernels: 0 = use the ordinary kernels.
template <typename K>
static int coop_group_size(K kernel, long long ncols, int slice, int num_sms, long long ntasks, int nlevels) {
    if (const char* e = std::getenv("KLUJAX_CUDA_COOP"))
        if (std::atoi(e) == 0) return 0;
    if (slice != 1 || ncols * 2 > num_sms) return 0;
    // the global-memory barrier costs ~1.6 us per level against ~0.5 us for __syncthreads():
    // worth it only when an average level is wider than one CTA (cfg 1: the factorization
    // 13.7 -> 7.8 ms, while its narrow solve levels would go 0.9 -> 2.0 ms)
    if (ntasks < 1024LL * std::max(nlevels, 1)) return 0;
    int dev = 0, coop = 0, occ = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    if (!coop || cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, 1024, 0) != cudaSuccess || occ < 1) return 0;
    long long g = std::min<long long>(32, (static_cast<long long>(occ) * num_sms) / ncols);
    if (const char* e = std::getenv("KLUJAX_CUDA_COOP_G")) g = std::min<long long>(g, std::max(1, std::atoi(e)));
    return g >= 2 ? static_cast<int>(g) : 0;
}

// Threads per CTA of the level kernels: enough for the widest level of the slice, at most 1024.
static int level_threads(int max_level_tasks, int slice) {
    (void)max_level_tasks;
    // at least one task group: a thread is (column lane, task group) and a slice has <= 64 lanes
    if (const char* e = std::getenv("KLUJAX_CUDA_LEVEL_THREADS"))
        return std::max((slice + 31) / 32 * 32, std::max(32, std::min(1024, std::atoi(e))));
    return 1024;
}

// The level table of a launch lives in shared memory up to this size (16 k levels); a larger
// carve-out would take too much of the SM's L1, which caches the values and the rows of X.
constexpr size_t kLevelTableSmem = 64 * 1024;

// The solve runs on the caller's x (no scratch, no relayout) when a CTA owns one column, or when a
// system has enough right-hand sides for a coalesced slice of them (x[system][row][column] is
// contiguous along the columns); many systems with a few right-hand sides each go through the
// batch-interleaved scratch, where consecutive threads = consecutive systems stay coalesced.
static bool level_native(long long Q, int r, int num_sms) {
    if (const char* e = std::getenv("KLUJAX_CUDA_LEVEL_NATIVE")) return std::atoi(e) != 0;
    return pick_slice(Q, num_sms) == 1 || r >= 16;
}

// Columns per CTA of the native multi-right-hand-side solve.  A CTA of 1024 threads fills an SM
// (56-64 registers per thread), so the slices are sized for whole, balanced waves of num_sms
// CTAs: 16 384 columns on 148 SMs are 148 slices of 112 (one wave), not 256 slices of 64 (a full
// wave and a 73 % one).  At most 128 columns: wider slices lose the rows of X from L1.
static int native_slice(long long Q, int num_sms, int C) {
    int waves = 1;
    long long per = (Q + num_sms - 1) / num_sms;
    int cap = 128;
    if (const char* e = std::getenv("KLUJAX_CUDA_LEVEL_MAXSLICE")) cap = std::max(C, std::atoi(e));
    while (per > cap) {
        ++waves;
        per = (Q + (long long)num_sms * waves - 1) / ((long long)num_sms * waves);
    }
    return static_cast<int>(std::max<long long>(C, (per + C - 1) / C * C));
}

template <typename K>
static void launch_level(K kernel, int grid, int threads, size_t smem, cudaStream_t st, const LvArgs& a) {
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kLevelTableSmem));
    kernel<<<grid, threads, smem, st>>>(a);
}

template <typename T>
static int level_run_t(const LevelCall& c, LevelProg* pf, LevelProg* ps, std::string& err) {
    using V_t = typename Sc<T>::V;
    auto chk = [&](cudaError_t e, const char* what) {
        if (e != cudaSuccess) {
            err = std::string(what) + ": " + cudaGetErrorString(e);
            return false;
        }
        return true;
    };
    const bool own_storage = (c.V == nullptr);
    V_t* V = static_cast<V_t*>(c.V);
    double* Rs = c.Rs;
    long long lpad = c.lpad;
    // Layout of V / Rs.  Batch-interleaved (V[slot][system]) makes the accesses of consecutive
    // threads = consecutive systems coalesced; with at most one system per SM a CTA owns ONE
    // system and its threads walk different slots of it, so the system-major layout
    // (V[system][slot]) keeps a CTA's working set in contiguous lines of its own L1/L2 sectors.
    int sysmajor = c.sysmajor;
    void *tmpV = nullptr, *tmpRs = nullptr, *tmpX = nullptr, *tmpAux = nullptr;
    auto cleanup = [&]() {
        // stream-ordered frees: the buffers stay valid for the work enqueued above
        if (tmpV) cudaFreeAsync(tmpV, c.st);
        if (tmpRs) cudaFreeAsync(tmpRs, c.st);
        if (tmpX) cudaFreeAsync(tmpX, c.st);
        if (tmpAux) cudaFreeAsync(tmpAux, c.st);
    };
    if (own_storage) {
        lpad = (c.L + 31) / 32 * 32;
        sysmajor = want_sysmajor(c.L, c.num_sms);
        if (!chk(cudaMallocAsync(&tmpV, sizeof(V_t) * (size_t)c.n_val * lpad, c.st), "alloc V") ||
            !chk(cudaMallocAsync(&tmpRs, sizeof(double) * (size_t)c.n * lpad, c.st), "alloc Rs")) {
            cleanup();
            return KLU_OUT_OF_MEMORY;
        }
        V = static_cast<V_t*>(tmpV);
        Rs = static_cast<double*>(tmpRs);
    }
    // aux: bad[L], growth2[L], then the arrival counters of the cooperative kernels
    const size_t ncnt = static_cast<size_t>(std::max<long long>(c.L, (long long)c.L * std::max(c.r, 1))) ;
    const size_t aux_bytes = 64 + sizeof(int) * (size_t)c.L + sizeof(unsigned long long) * (size_t)c.L + 64 +
                             (ncnt <= 1024 ? sizeof(unsigned) * ncnt : 0) + 64;
    void* auxp = c.aux;
    if (!auxp || c.aux_bytes < aux_bytes) {
        if (!chk(cudaMallocAsync(&tmpAux, aux_bytes, c.st), "alloc aux")) {
            cleanup();
            return KLU_OUT_OF_MEMORY;
        }
        auxp = tmpAux;
    }
    if (!chk(cudaMemsetAsync(auxp, 0, aux_bytes, c.st), "memset aux")) {
        cleanup();
        return KLU_OUT_OF_MEMORY;
    }
    unsigned long long* growth2 = reinterpret_cast<unsigned long long*>(static_cast<char*>(auxp) + 64);
    int* bad = reinterpret_cast<int*>(growth2 + c.L);
    unsigned* coop_bar = ncnt <= 1024 ? reinterpret_cast<unsigned*>(bad + ((c.L + 15) / 16) * 16) : nullptr;

    const long long v_slot = sysmajor ? 1 : lpad, v_sys = sysmajor ? c.n_val : 1;
    if (v_slot * 16 >= (1LL << 32) || (long long)c.L * std::max(c.r, 1) * 16 >= (1LL << 32)) {
        err = "batch too large for the 32-bit strides of the level kernels (split the batch)";
        cleanup();
        return KLU_TOO_LARGE;
    }
    const long long r_row = sysmajor ? 1 : lpad, r_sys = sysmajor ? c.n : 1;
    const bool factor = (pf != nullptr) || (c.dep_col != nullptr);
    if (factor) {
        if (sysmajor && c.slot_src && !std::getenv("KLUJAX_CUDA_NO_PREPARE")) {
            if (!chk(cudaMemsetAsync(c.slot_src, 0xFF, sizeof(int) * static_cast<size_t>(c.n_val), c.st), "memset slot map")) {
                cleanup();
                return KLU_OUT_OF_MEMORY;
            }
            coo_invert_kernel<<<(c.nz + 255) / 256, 256, 0, c.st>>>(c.nz, c.coo_dst, c.slot_src);
            lv_prepare_sysmajor<T><<<static_cast<unsigned>(std::min<long long>(c.L, 4LL * c.num_sms)), 1024, 0, c.st>>>(
                c.L, c.n, c.nz, c.n_val, static_cast<const V_t*>(c.Ax), c.ax_stride, c.slot_src, c.rs_ptr, c.rs_slot, c.sysidx, V, Rs);
        } else {
        const long long tot = (long long)c.n_val * c.L;
        const int zb = static_cast<int>(std::min<long long>((tot + 255) / 256, 148LL * 16));
        lv_zero_slots<T><<<zb, 256, 0, c.st>>>(c.L, c.n_val, c.sysidx, V, v_slot, v_sys);
        dim3 tb(32, 8), tg((c.nz + 31) / 32, (c.L + 31) / 32);
        lv_scatter<T><<<tg, tb, 0, c.st>>>(c.L, c.nz, static_cast<const V_t*>(c.Ax), c.ax_stride,
                                          c.coo_dst, c.sysidx, V, v_slot, v_sys);
        const long long rt = (long long)c.n * c.L;
        lv_rowscale<T><<<static_cast<unsigned>((rt + 255) / 256), 256, 0, c.st>>>(
            c.L, c.n, c.rs_ptr, c.rs_slot, c.sysidx, V, Rs, v_slot, v_sys, r_row, r_sys);
        }
        if (c.dep_col) {
            const DepFactorPlan fp = dep_factor_plan(c.n, c.dep_col->maxcol, sizeof(V_t));
            DepArgs d = {};
            d.colinfo = c.dep_col->d_colinfo; d.ccand = c.dep_col->d_ccand; d.urec = c.dep_col->d_urec; d.dest = c.dep_col->d_dest;
            d.scratch_elems = fp.scratch_elems;
            d.stage_elems = fp.stage_elems;
            d.n = c.n; d.L = c.L; d.V = V; d.lpad = v_slot; d.vss = v_sys; d.sysidx = c.sysidx; d.bad = bad; d.growth2 = growth2;
            if (!chk(dep_launch(dep_factor<T>, static_cast<int>(std::min<long long>(c.L, 2LL * c.num_sms)), 32 * fp.warps, fp.smem, c.st, d, 40),
                     "dependency-driven factor")) {
                cleanup();
                return KLU_OUT_OF_MEMORY;
            }
            lv_finish_status<<<(c.L + 255) / 256, 256, 0, c.st>>>(c.L, bad, growth2, c.pattern_bad,
                                                                c.status, c.growth, c.certify);
        } else {
        LvArgs a = {};
        a.task = pf->d_task; a.pair = pf->d_pair; a.level_ptr = pf->d_level_ptr;
        a.nlevels = pf->nlevels; a.n_val = c.n_val; a.L = c.L; a.r = 1; a.lpad = v_slot; a.vss = v_sys; a.xrs = 0; a.xcs = 0; a.Q = c.L;
        a.V = V; a.X = nullptr; a.sysidx = c.sysidx; a.bad = bad; a.growth2 = growth2;
        a.slice = pick_slice(c.L, c.num_sms);
        const int grid = static_cast<int>(std::min<long long>((c.L + a.slice - 1) / a.slice, 2LL * c.num_sms));
        a.gsz = coop_bar ? coop_group_size(lv_factor_coop<T>, c.L, a.slice, c.num_sms, pf->ntasks, pf->nlevels) : 0;
        if (a.gsz >= 2) {
            a.bar = coop_bar;
            void* kargs[] = {&a};
            if (!chk(cudaMemsetAsync(coop_bar, 0, sizeof(unsigned) * c.L, c.st), "memset barrier") ||
                !chk(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(lv_factor_coop<T>), dim3(static_cast<unsigned>(c.L * a.gsz)),
                                                 dim3(1024), kargs, 0, c.st), "cooperative factor")) {
                cleanup();
                return KLU_OUT_OF_MEMORY;
            }
        } else {
            const size_t lp_bytes = sizeof(int) * (static_cast<size_t>(pf->nlevels) + 1);
            a.lp_smem = lp_bytes <= kLevelTableSmem ? 1 : 0;
            launch_level(lv_factor<T>, grid, level_threads(pf->max_level_tasks, a.slice), a.lp_smem ? lp_bytes : 0, c.st, a);
        }
        lv_finish_status<<<(c.L + 255) / 256, 256, 0, c.st>>>(c.L, bad, growth2, c.pattern_bad,
                                                            c.status, c.growth, c.certify);
        }
    }
    if (c.dep_panel) {
        const bool tr = (c.mode == PM_FACTOR_TSOLVE || c.mode == PM_TSOLVE);
        const PanelPlan pl = panel_plan(c.n, c.dep_panel->ntemps, c.dep_panel->maxI, sizeof(V_t));
        DepArgs d = {};
        d.hdr = c.dep_panel->d_hdr; d.lane = c.dep_panel->d_lane; d.ext_slot = c.dep_panel->d_ext_slot;
        d.ext_idx = c.dep_panel->d_ext_idx; d.in_slot = c.dep_panel->d_in_slot;
        d.npanels = c.dep_panel->npanels; d.ntemps = c.dep_panel->ntemps; d.stage_lines = std::max(1, c.dep_panel->maxI); d.r = c.r;
        d.X = c.x; d.B = c.b; d.b_stride = c.b_stride;
        d.perm_in = tr ? c.q : c.pnum; d.perm_out = tr ? c.pnum : c.q;
        d.Rs = Rs; d.rlp = r_row; d.rss = r_sys; d.scale_in = tr ? 0 : 1; d.scale_out = tr ? 1 : 0;
        d.n = c.n; d.L = c.L; d.V = V; d.lpad = v_slot; d.vss = v_sys; d.sysidx = c.sysidx; d.bad = bad; d.growth2 = growth2;
        const long long Qd = (long long)c.L * c.r;
        const int nw = std::min(pl.warps, std::max(1, dep_env_int("KLUJAX_CUDA_PANEL_WARPS", pl.warps)));
        if (!chk(dep_launch(panel_solve<T>, static_cast<int>(std::min<long long>(Qd, 4LL * c.num_sms)), 32 * nw, pl.smem, c.st, d, 20),
                 "panel solve")) {
            cleanup();
            return KLU_OUT_OF_MEMORY;
        }
    } else if (c.grp) {
        const bool tr = (c.mode == PM_FACTOR_TSOLVE || c.mode == PM_TSOLVE);
        GrpArgs g = {};
        g.ghdr = c.grp->d_ghdr; g.krec = c.grp->d_krec; g.gaux = c.grp->d_gaux; g.level_ptr = c.grp->d_level_ptr;
        g.nlevels = c.grp->nlevels; g.n = c.n; g.r = c.r; g.Q = (long long)c.L * c.r; g.lpad = v_slot; g.vss = v_sys;
        g.xss = (long long)c.n * c.r; g.b_stride = c.b_stride; g.rlp = r_row; g.rss = r_sys;
        g.V = V; g.X = c.x; g.B = c.b; g.sysidx = c.sysidx; g.src_of = tr ? c.src_tsolve : c.src_solve; g.Rs = Rs;
        g.scale_in = tr ? 0 : 1; g.scale_out = tr ? 1 : 0;
        // two columns per lane: 10 warps x 96 registers and 99 KB of staging per CTA, two CTAs per SM - the 256
        // slices of cfg 3 run in ONE wave (four columns per lane: 128 CTAs of 10 warps, 68 ms against 49)
        int C = c.r % 2 == 0 ? 2 : 1;
        if (const char* e = std::getenv("KLUJAX_CUDA_GROUP_C")) {
            const int want = std::atoi(e);
            if ((want == 1 || want == 2 || (want == 4 && !c.is_complex)) && c.r % want == 0) C = want;
        }
        const long long nsl = (long long)c.L * ((c.r + 32 * C - 1) / (32 * C));
        const int grid = static_cast<int>(std::min<long long>(nsl, 1 << 20));
        const int thmax = C * sizeof(V_t) >= 32 ? 320 : 640;
        const int th = std::min(thmax, std::max(32, dep_env_int("KLUJAX_CUDA_GROUP_THREADS", 320) / 32 * 32));
        auto launch_grouped = [&](auto kernel, size_t warp_bytes) {
            const size_t smem = warp_bytes * static_cast<size_t>(th / 32);
            cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kDepSmemMax));
            kernel<<<grid, th, smem, c.st>>>(g);
        };
        if (C == 4)
            launch_grouped(lv_solve_grouped<double, 4>, grouped_warp_bytes<double, 4>());
        else if (C == 2)
            launch_grouped(lv_solve_grouped<T, 2>, grouped_warp_bytes<V_t, 2>());
        else
            launch_grouped(lv_solve_grouped<T, 1>, grouped_warp_bytes<V_t, 1>());
    } else if (ps) {
        const bool tr = (c.mode == PM_FACTOR_TSOLVE || c.mode == PM_TSOLVE);
        const long long Q = (long long)c.L * c.r;
        const long long tot = Q * c.n;
        const int xslice = pick_slice(Q, c.num_sms);
        const bool native = level_native(Q, c.r, c.num_sms);
        V_t* X = static_cast<V_t*>(c.x);
        // X: the caller's x (native), else a scratch with the columns interleaved (X[row][column])
        long long x_row = c.r, x_col = 0;
        if (!native) {
            if (!chk(cudaMallocAsync(&tmpX, sizeof(V_t) * (size_t)c.n * Q, c.st), "alloc X")) {
                cleanup();
                return KLU_OUT_OF_MEMORY;
            }
            X = static_cast<V_t*>(tmpX);
            x_row = Q;
            x_col = 1;
            lv_load_x<T><<<static_cast<unsigned>((tot + 255) / 256), 256, 0, c.st>>>(
                c.L, c.n, c.r, static_cast<const V_t*>(c.b), c.b_stride, tr ? c.q : c.pnum, tr ? 0 : 1,
                Rs, r_row, r_sys, c.sysidx, X, x_row, x_col);
        }
        LvArgs a = {};
        a.task = ps->d_task; a.pair = ps->d_pair; a.level_ptr = ps->d_level_ptr;
        a.nlevels = ps->nlevels; a.n_val = c.n_val; a.L = c.L; a.r = c.r; a.lpad = v_slot; a.vss = v_sys; a.xrs = x_row; a.xcs = x_col; a.Q = Q;
        a.V = V; a.X = X; a.sysidx = c.sysidx; a.bad = bad; a.growth2 = growth2;
        a.slice = xslice;
        a.native = native ? 1 : 0; a.n = c.n; a.xss = (long long)c.n * c.r;
        a.B = c.b; a.b_stride = c.b_stride; a.src_of = tr ? c.src_tsolve : c.src_solve;
        a.Rs = Rs; a.rlp = r_row; a.rss = r_sys; a.scale_in = tr ? 0 : 1; a.scale_out = tr ? 1 : 0;
        const size_t lp_bytes = sizeof(int) * (static_cast<size_t>(ps->nlevels) + 1);
        a.lp_smem = lp_bytes <= kLevelTableSmem ? 1 : 0;
        const int grid = static_cast<int>(std::min<long long>((Q + a.slice - 1) / a.slice, 2LL * c.num_sms));
        a.gsz = coop_bar ? coop_group_size(lv_solve_coop<T>, Q, a.slice, c.num_sms, ps->ntasks, ps->nlevels) : 0;
        if (a.gsz >= 2) {
            a.bar = coop_bar;
            void* kargs[] = {&a};
            if (native)
                lv_load_native<T><<<static_cast<unsigned>((tot + 255) / 256), 256, 0, c.st>>>(
                    c.L, c.n, c.r, static_cast<const V_t*>(c.b), c.b_stride, a.src_of, a.scale_in, Rs, r_row, r_sys,
                    c.sysidx, X);
            if (!chk(cudaMemsetAsync(coop_bar, 0, sizeof(unsigned) * Q, c.st), "memset barrier") ||
                !chk(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(lv_solve_coop<T>), dim3(static_cast<unsigned>(Q * a.gsz)),
                                                 dim3(1024), kargs, 0, c.st), "cooperative solve")) {
                cleanup();
                return KLU_OUT_OF_MEMORY;
            }
            if (native && a.scale_out)
                lv_scale_native<T><<<static_cast<unsigned>((tot + 255) / 256), 256, 0, c.st>>>(
                    c.L, c.n, c.r, Rs, r_row, r_sys, c.sysidx, X);
        } else {
            const int threads = level_threads(ps->max_level_tasks, a.slice);
            const size_t sm_bytes = a.lp_smem ? lp_bytes : 0;
            // C adjacent right-hand sides per thread (native layout; C divides r and the slice)
            int C = 1;
            if (native && a.slice > 1) {
                C = (c.r % 4 == 0 && !c.is_complex) ? 4 : (c.r % 2 == 0 ? 2 : 1);
                if (const char* e = std::getenv("KLUJAX_CUDA_LEVEL_C")) {
                    const int want = std::atoi(e);
                    if ((want == 1 || want == 2 || want == 4 || (want == 8 && !c.is_complex)) && c.r % want == 0) C = want;
                }
                if (!std::getenv("KLUJAX_CUDA_LEVEL_SLICE")) a.slice = native_slice(Q, c.num_sms, C);
                a.slice = std::max(C, a.slice / C * C);
            }
            const int grid2 = static_cast<int>(std::min<long long>((Q + a.slice - 1) / a.slice, 2LL * c.num_sms));
            if (a.slice == 1)
                launch_level(lv_solve<T, 1, true>, grid2, threads, sm_bytes, c.st, a);
            else if (C == 8)
                launch_level(lv_solve<double, 8, false>, grid2, std::min(threads, 768), sm_bytes, c.st, a);
            else if (C == 4)
                launch_level(lv_solve<T, 4, false>, grid2, threads, sm_bytes, c.st, a);
            else if (C == 2)
                launch_level(lv_solve<T, 2, false>, grid2, threads, sm_bytes, c.st, a);
            else
                launch_level(lv_solve<T, 1, false>, grid2, threads, sm_bytes, c.st, a);
        }
        if (!native)
            lv_store_x<T><<<static_cast<unsigned>((tot + 255) / 256), 256, 0, c.st>>>(
                c.L, c.n, c.r, static_cast<V_t*>(c.x), tr ? c.pnum : c.q, tr ? 1 : 0, Rs, r_row, r_sys, c.sysidx, X,
                x_row, x_col);
    }
    if (!chk(cudaGetLastError(), "level kernels")) {
        cleanup();
        return KLU_OUT_OF_MEMORY;
    }
    cleanup();
    return 0;
}

static int level_run(const LevelCall& c, LevelProg* pf, LevelProg* ps, std::string& err) {
    return c.is_complex ? level_run_t<zc>(c, pf, ps, err) : level_run_t<double>(c, pf, ps, err);
}

// ---- dot_f64 / dot_c128: b[m,i,k] = sum_e Ax[m,e] * x[m,Aj[e],k] over Ai[e] = i ----
// Deterministic: the entries are grouped by row with a STABLE radix sort of the call's (Ai, e) pairs
// (cub), then a thread per (system, row, column) adds the products of its row in the order the
// entries arrive in the COO list - the summation order of the reference's serial loop
// (/root/reference/klujax.cpp:165-249) - and writes b once.  (Round 1 used a thread per product
// with atomicAdd: non-deterministic rounding, uncoalesced.)
__global__ void spmm_keys_kernel(int nz, int n, const int* __restrict__ Ai, const int* __restrict__ Aj,
                                 int* __restrict__ key, int* __restrict__ val) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nz) return;
    const int i = Ai[e], j = Aj[e];
    key[e] = (i < 0 || i >= n || j < 0 || j >= n) ? n : i;  // entries out of range form a last bucket nobody reads
    val[e] = e;
}
__global__ void spmm_rowptr_kernel(int nz, int n, const int* __restrict__ skey, int* __restrict__ rowptr) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    int lo = 0, hi = nz;  // first position whose key is >= i
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (skey[mid] < i) lo = mid + 1;
        else hi = mid;
    }
    rowptr[i] = lo;
}
template <typename T>
__global__ void spmm_kernel(int L, int n, int r, int nz, const int* __restrict__ rowptr, const int* __restrict__ sval,
                            const int* __restrict__ Aj, const typename Sc<T>::V* __restrict__ Ax,
                            const typename Sc<T>::V* __restrict__ x, typename Sc<T>::V* __restrict__ b) {
    using S = Sc<T>;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)L * n * r;
    if (tid >= total) return;
    const int k = static_cast<int>(tid % r);
    const long long t2 = tid / r;
    const int i = static_cast<int>(t2 % n);
    const long long m = t2 / n;
    typename S::V acc = S::zero();
    for (int p = rowptr[i]; p < rowptr[i + 1]; ++p) {
        const int e = sval[p];
        S::fma_acc(acc, Ax[m * nz + e], x[(m * n + Aj[e]) * r + k]);
    }
    b[tid] = acc;
}

// dAx[m][e] = sum_k ct[m][Ai[e]][k] * x[m][Aj[e]][k]: transpose of the product above with
// respect to the values (/root/reference/klujax.py:1712-1714).  Thread per (system, entry).
template <typename T>
__global__ void sddmm_kernel(int L, int n, int r, int nz, const int* __restrict__ Ai,
                             const int* __restrict__ Aj, const typename Sc<T>::V* __restrict__ ct,
                             const typename Sc<T>::V* __restrict__ x,
                             typename Sc<T>::V* __restrict__ dAx) {
    using S = Sc<T>;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (long long)L * nz) return;
    const int e = static_cast<int>(tid % nz);
    const long long m = tid / nz;
    const int i = Ai[e], j = Aj[e];
    typename S::V acc = S::zero();
    if (i >= 0 && i < n && j >= 0 && j < n) {
        const typename S::V* pc = ct + (m * n + i) * r;
        const typename S::V* px = x + (m * n + j) * r;
        for (int k = 0; k < r; ++k) S::fma_acc(acc, pc[k], px[k]);
    }
    dAx[tid] = acc;
}

static int coo_sddmm(int cx, int L, int n, int r, int nz, const int* Ai, const int* Aj,
                     const double* ct, const double* x, double* dAx, cudaStream_t st, std::string& err) {
    const long long total = (long long)L * nz;
    if (total == 0) return 0;
    const unsigned grid = static_cast<unsigned>((total + 255) / 256);
    if (cx)
        sddmm_kernel<zc><<<grid, 256, 0, st>>>(L, n, r, nz, Ai, Aj, reinterpret_cast<const double2*>(ct),
                                               reinterpret_cast<const double2*>(x),
                                               reinterpret_cast<double2*>(dAx));
    else
        sddmm_kernel<double><<<grid, 256, 0, st>>>(L, n, r, nz, Ai, Aj, ct, x, dAx);
    if (cudaGetLastError() != cudaSuccess) {
        err = "dot_transpose: launch failed";
        return KLU_INVALID;
    }
    return 0;
}

static int coo_spmm(int cx, int L, int n, int r, int nz, const int* Ai, const int* Aj,
                    const double* Ax, const double* x, double* b, cudaStream_t st, std::string& err) {
    const long long total = (long long)L * n * r;
    if (total == 0) return 0;
    auto bad = [&](const char* what) {
        err = std::string("dot: ") + what;
        cudaGetLastError();
        return KLU_INVALID;
    };
    // scratch (stream-ordered): keys / entry ids before and after the sort, row pointers, cub's own
    int* scratch = nullptr;
    const size_t nzp = static_cast<size_t>(std::max(nz, 1));
    if (cudaMallocAsync(reinterpret_cast<void**>(&scratch), sizeof(int) * (4 * nzp + n + 2), st) != cudaSuccess) return bad("out of device memory");
    int *key = scratch, *val = key + nzp, *skey = val + nzp, *sval = skey + nzp, *rowptr = sval + nzp;
    void* tmp = nullptr;
    size_t tmp_bytes = 0;
    int end_bit = 1;
    while ((1 << end_bit) <= n && end_bit < 31) ++end_bit;  // keys are 0..n
    int rc = 0;
    if (nz > 0) {
        spmm_keys_kernel<<<(nz + 255) / 256, 256, 0, st>>>(nz, n, Ai, Aj, key, val);
        if (cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, key, skey, val, sval, nz, 0, end_bit, st) != cudaSuccess ||
            cudaMallocAsync(&tmp, std::max<size_t>(tmp_bytes, 16), st) != cudaSuccess ||
            cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, key, skey, val, sval, nz, 0, end_bit, st) != cudaSuccess)
            rc = bad("sorting the entries by row failed");
    }
    if (!rc) {
        spmm_rowptr_kernel<<<(n + 1 + 255) / 256, 256, 0, st>>>(nz, n, skey, rowptr);
        const unsigned grid = static_cast<unsigned>((total + 255) / 256);
        if (cx)
            spmm_kernel<zc><<<grid, 256, 0, st>>>(L, n, r, nz, rowptr, sval, Aj, reinterpret_cast<const double2*>(Ax),
                                                  reinterpret_cast<const double2*>(x), reinterpret_cast<double2*>(b));
        else
            spmm_kernel<double><<<grid, 256, 0, st>>>(L, n, r, nz, rowptr, sval, Aj, Ax, x, b);
        if (cudaGetLastError() != cudaSuccess) rc = bad("launch failed");
    }
    if (tmp) cudaFreeAsync(tmp, st);
    cudaFreeAsync(scratch, st);
    return rc;
}

}  // namespace kb
