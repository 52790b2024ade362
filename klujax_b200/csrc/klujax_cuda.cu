// klujax_cuda.cu — C-ABI (include/klujax_cuda.h) and host-side handle objects of
// the B200 twin of klujax's numeric hot path.  Each extern "C" entry point stands
// where one XLA-FFI handler of /root/reference/klujax.cpp stands; the shape
// parsing those handlers do (klujax.cpp:17-84, :1011-1064) is the caller's job
// (klujax_b200/api.py mirrors klujax.py for it).
#include <cuda_runtime.h>

#include <algorithm>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/klujax_cuda.h"
#include "plan.hpp"
#include "rowreg.hpp"
#include "symbolic.hpp"
#include "kernels_small.cuh"
#include "kernels_level.cuh"
#include "kernels_wide.cuh"

namespace kb {

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
#define CUDA_OK(call)                                                                    \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess)                                                           \
            return fail(KLU_OUT_OF_MEMORY, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
    cudaError_t alloc(size_t nbytes) {
        release();
        bytes = nbytes;
        return cudaMalloc(&p, nbytes ? nbytes : 16);
    }
    template <typename U>
    cudaError_t upload(const std::vector<U>& v) {
        cudaError_t e = alloc(v.size() * sizeof(U));
        if (e != cudaSuccess || v.empty()) return e;
        return cudaMemcpy(p, v.data(), v.size() * sizeof(U), cudaMemcpyHostToDevice);
    }
    template <typename U>
    U* as() const {
        return static_cast<U*>(p);
    }
};

struct SmallProg {
    DevBuf stream, remap, rs_slot;  // remap / rs_slot: compacted programs only
    DevBuf tail_tab;                // register stage of the dense tail (plan.hpp), if the program uses it
    int tail_t0 = -1;
    std::vector<uint32_t> ctrl;
    std::vector<uint16_t> chunk_ptr;
    int nchunks = 0, nbatch = 0, n_slots = 0, zero_slot = 0, nlevels = 0, n_val = 0;
    int64_t n_mac = 0;
};

// Everything that depends on the pivot sequence (one per value type).
// A generated register-resident kernel (rowreg.hpp): plan, PTX assembled by the driver's JIT,
// the 16-bit tables on the device.  One per (transpose, right-hand-side width).
struct RRKernel {
    bool usable = false;
    std::string why;
    rr::Plan plan;
    rr::PtxKernel pk;
    cudaLibrary_t lib = nullptr;
    cudaKernel_t kern = nullptr;
    DevBuf tabs, splace;  // splace: slot of every entry of the (col,row)-sorted pattern (COO map of the call)
    int regs = 0, ctas_per_sm = 1;
    ~RRKernel() {
        if (lib) cudaLibraryUnload(lib);
    }
};

struct Seeded {
    bool ok = false;
    PivotPlan piv;
    SlotLayout lay;
    int regime = 0;  // 0 = shared-memory warp-per-system, 1 = level-scheduled global
    DevBuf d_colptr, d_srow, d_sslot, d_rs_ptr, d_rs_slot, d_pnum, d_q;
    DevBuf d_src_solve, d_src_tsolve;  // level solves on the caller's x: source row of b per row of x
    std::map<std::pair<int, int>, std::unique_ptr<SmallProg>> small;  // key (mode, rc + 1000 * wide)
    std::map<std::pair<int, int>, std::unique_ptr<LevelProg>> level;
    // dependency-driven regime (kernels_dep.cuh): column program of the refactorization, panel
    // programs of the solves
    std::unique_ptr<DepColProg> dep_col;
    bool dep_col_tried = false;
    std::map<int, std::unique_ptr<DepPanelProg>> dep_panel;  // key: transpose
    std::map<int, std::unique_ptr<GroupProg>> grp;            // key: transpose (multi-right-hand-side solves)
    std::map<std::pair<int, int>, std::unique_ptr<RRKernel>> rowreg;  // key (transpose, n_rhs)
    DevBuf d_scsc;  // CSC position of every entry of the (col,row)-sorted pattern (COO map of the rowreg kernels)
    int64_t info_levels = 0, info_mac = 0, info_nbatch = 0, info_stream_words = 0;
};

struct Symbolic {
    CscPattern A;
    Ordering ord;
    int device = -1;  // CUDA device that holds the schedules (set by the first numeric call)
    std::vector<int> Ai, Aj;  // COO of the analyze call
    Seeded seeded[2];
    // per-stream scratch of the COO map (slot of every COO entry of the call + pattern flag):
    // calls on different streams may overlap on the device
    struct Scratch {
        DevBuf coo_dst, coo_csc, bad;
        DevBuf aux;  // status words of the level regime (kernels_level.cuh: level_run_t), grown on demand
        size_t aux_bytes = 0;
        DevBuf slot_src;  // level regime, system-major fill: COO entry behind every value slot
    };
    std::map<cudaStream_t, std::unique_ptr<Scratch>> scratch;
    Scratch* cur = nullptr;  // scratch of the call in flight (under mu)
    std::mutex mu;
};

struct NumericBatch {
    uint32_t id = 0;
    Symbolic* sym = nullptr;     // pivot sequence + schedules of this batch
    Symbolic* parent = nullptr;  // the caller's symbolic handle (== sym unless the batch was re-seeded)
    std::unique_ptr<Symbolic> owned;  // re-seeded batches own their symbolic object
    int is_complex = 0, regime = 0, L = 0, live = 0;
    int64_t lpad = 0;  // level regime: padded batch stride
    int sysmajor = 0;  // level regime: V[system][slot] instead of V[slot][system] (kernels_level.cuh)
    DevBuf V, Rs;
    std::vector<uint8_t> freed;
};

static std::mutex g_reg_mu;
static std::unordered_map<uint32_t, NumericBatch*> g_batches;
static uint32_t g_next_batch = 1;
static int g_num_sms[64] = {0};

static int num_sms() {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!g_num_sms[dev]) {
        cudaDeviceGetAttribute(&g_num_sms[dev], cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms[dev] <= 0) g_num_sms[dev] = 148;
        // first numeric call on this device: the stream-ordered scratch of the fused solves stays in the
        // pool across synchronisations (the default threshold 0 hands it back to the driver at every
        // sync, and the next call pays a real allocation: 20-70 ms outliers on cfg 4 / cfg 1)
        cudaMemPool_t pool = nullptr;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess && pool) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    return g_num_sms[dev];
}

static inline Symbolic* sym_of(uint64_t h) { return reinterpret_cast<Symbolic*>(static_cast<uintptr_t>(h)); }

// ---------------------------------------------------------------------------
// seeding: one pivoting factorization on the host fixes pivots and patterns
// ---------------------------------------------------------------------------
static constexpr size_t kMaxDynSmem = 232448;  // 227 KB per CTA on sm_100

template <typename T>
static int seed_from_csc_values(Symbolic* S, int cx, const T* csc_values) {
    Seeded& sd = S->seeded[cx];
    if (!seed_factor<T>(S->A, S->ord, csc_values, sd.piv)) {
        const int st = sd.piv.status;
        return fail(st ? st : KLU_SINGULAR,
                    st == KLU_INVALID ? "duplicate entries in the pattern (coalesce first)"
                                      : "seed system is singular: cannot fix a pivot sequence");
    }
    build_layout(S->A, S->ord, sd.piv, sd.lay);
    // regime: does factor+solve with one right-hand side fit in shared memory?
    TaskProgram tp;
    PackedProgram pp;
    const size_t vsz = cx ? 16 : 8;
    sd.regime = 1;
    if (static_cast<size_t>(sd.lay.n_val + sd.lay.n + 3) * vsz + sd.lay.n * 8ull + 16 <= kMaxDynSmem - kSmallHeaderBytes &&
        sd.lay.n_val + sd.lay.n + 3 <= 16383) {
        build_program(S->ord, sd.piv, sd.lay, PM_FACTOR_SOLVE, 1, tp, 0);
        if (pack_program(sd.lay, tp, pp)) sd.regime = 0;
        sd.info_nbatch = pp.nbatch;
        sd.info_stream_words = static_cast<int64_t>(pp.stream.size());
        sd.info_levels = tp.nlevels;
    }
    if (const char* force = std::getenv("KLUJAX_CUDA_FORCE_REGIME")) sd.regime = std::atoi(force) ? 1 : 0;
    {
        // multiply-adds of one refactorization: every U(j,k) meets the whole column L(:,j)
        const SlotLayout& lay = sd.lay;
        int64_t mac = 0;
        for (int k = 0; k < lay.n; ++k)
            for (int su = lay.col_begin[k]; su < lay.diag_slot[k]; ++su) {
                const int j = lay.slot_row[su];
                mac += lay.col_begin[j + 1] - lay.diag_slot[j] - 1;
            }
        sd.info_mac = mac;
        if (sd.regime == 1) sd.info_levels = -1;  // known once the level program is built
    }
    sd.ok = true;
    return 0;
}

static int upload_seeded(Symbolic* S, int cx) {
    Seeded& sd = S->seeded[cx];
    if (sd.d_colptr.p) return 0;
    CUDA_OK(sd.d_colptr.upload(sd.lay.sorted_colptr));
    CUDA_OK(sd.d_srow.upload(sd.lay.sorted_row));
    CUDA_OK(sd.d_sslot.upload(sd.lay.sorted_slot));
    {
        std::vector<int> scsc(sd.lay.sorted_slot.size());
        for (size_t q = 0; q < scsc.size(); ++q) scsc[q] = sd.lay.slot_csc[sd.lay.sorted_slot[q]];
        CUDA_OK(sd.d_scsc.upload(scsc));
    }
    CUDA_OK(sd.d_rs_ptr.upload(sd.lay.rs_ptr));
    CUDA_OK(sd.d_rs_slot.upload(sd.lay.rs_slot));
    CUDA_OK(sd.d_pnum.upload(sd.piv.Pnum));
    CUDA_OK(sd.d_q.upload(S->ord.Q));
    {
        // solve: x[Q[k]] <- b[Pnum[k]]; transpose solve: x[Pnum[k]] <- b[Q[k]] (klu_solve / klu_tsolve)
        const int n = sd.lay.n;
        std::vector<int> ss(n), st(n);
        for (int k = 0; k < n; ++k) {
            ss[S->ord.Q[k]] = sd.piv.Pnum[k];
            st[sd.piv.Pnum[k]] = S->ord.Q[k];
        }
        CUDA_OK(sd.d_src_solve.upload(ss));
        CUDA_OK(sd.d_src_tsolve.upload(st));
    }
    return 0;
}

// values of one system given in some COO order -> the analysed CSC order
template <typename T>
static int seed_from_coo(Symbolic* S, int cx, int nz, const int* Ai, const int* Aj, const T* Ax0) {
    const CscPattern& A = S->A;
    if (nz != static_cast<int>(A.rowidx.size()))
        return fail(KLU_INVALID, "n_nz differs from the analysed pattern");
    // (col,row) -> csc position through a per-column sorted view
    std::vector<int> order(nz);
    for (int j = 0; j < A.n; ++j) {
        for (int p = A.colptr[j]; p < A.colptr[j + 1]; ++p) order[p] = p;
        std::sort(order.begin() + A.colptr[j], order.begin() + A.colptr[j + 1],
                  [&](int x, int y) { return A.rowidx[x] < A.rowidx[y]; });
    }
    std::vector<T> csc(nz, T(0));
    for (int e = 0; e < nz; ++e) {
        const int i = Ai[e], j = Aj[e];
        if (i < 0 || i >= A.n || j < 0 || j >= A.n) return fail(KLU_INVALID, "index out of range");
        auto b = order.begin() + A.colptr[j], en = order.begin() + A.colptr[j + 1];
        auto it = std::lower_bound(b, en, i, [&](int p, int row) { return A.rowidx[p] < row; });
        if (it == en || A.rowidx[*it] != i)
            return fail(KLU_INVALID, "entry outside the analysed sparsity pattern");
        csc[*it] = Ax0[e];
    }
    return seed_from_csc_values<T>(S, cx, csc.data());
}

static int ensure_seeded_dev(Symbolic* S, int cx, int nz, const int32_t* Ai_dev,
                             const int32_t* Aj_dev, const void* Ax_dev, cudaStream_t st) {
    Seeded& sd = S->seeded[cx];
    if (!sd.ok) {
        if (nz != static_cast<int>(S->A.rowidx.size()))
            return fail(KLU_INVALID, "n_nz differs from the analysed pattern");
        std::vector<int> Ai(nz), Aj(nz);
        std::vector<double> Ax(static_cast<size_t>(nz) * (cx ? 2 : 1));
        CUDA_OK(cudaMemcpyAsync(Ai.data(), Ai_dev, sizeof(int) * nz, cudaMemcpyDeviceToHost, st));
        CUDA_OK(cudaMemcpyAsync(Aj.data(), Aj_dev, sizeof(int) * nz, cudaMemcpyDeviceToHost, st));
        CUDA_OK(cudaMemcpyAsync(Ax.data(), Ax_dev, sizeof(double) * Ax.size(), cudaMemcpyDeviceToHost, st));
        CUDA_OK(cudaStreamSynchronize(st));
        int rc = cx ? seed_from_coo<std::complex<double>>(
                          S, 1, nz, Ai.data(), Aj.data(),
                          reinterpret_cast<const std::complex<double>*>(Ax.data()))
                    : seed_from_coo<double>(S, 0, nz, Ai.data(), Aj.data(), Ax.data());
        if (rc) return rc;
    }
    return upload_seeded(S, cx);
}

// ---------------------------------------------------------------------------
// program caches
// ---------------------------------------------------------------------------
static int get_small_prog(Symbolic* S, int cx, int mode, int rc, bool wide, bool compact, SmallProg** out,
                          bool tail = false) {
    Seeded& sd = S->seeded[cx];
    auto key = std::make_pair(mode, rc + (wide ? 1000 : 0) + (compact ? 100000 : 0) + (tail ? 10000000 : 0));
    auto it = sd.small.find(key);
    if (it == sd.small.end()) {
        TaskProgram tp;
        PackedProgram pp;
        int split = 0;  // measured: the dot form beats as-soon-as-possible solve updates here (220k vs 234k cycles)
        if (const char* e = std::getenv("KLUJAX_CUDA_SPLIT")) split = std::atoi(e);
        if (wide) {
            split = 111;  // one multiply-add per partial update (measured best: 11 < 14 < 24 < 47;
                          // the kernel's step format is fixed to it)
        }
        const int tail_t0 = (tail && wide) ? pick_tail(S->ord, sd.piv, sd.lay) : -1;
        build_program(S->ord, sd.piv, sd.lay, mode, rc, tp, split, compact ? (cx ? 8 : 16) : 0, tail_t0);
        if (!(wide ? pack_program_wide(sd.lay, tp, pp, cx ? 16 : 8) : pack_program(sd.lay, tp, pp))) return fail(KLU_TOO_LARGE, "program does not fit the shared-memory regime");
        auto sp = std::make_unique<SmallProg>();
        CUDA_OK(sp->stream.upload(pp.stream));
        sp->nchunks = pp.nchunks;
        sp->ctrl = pp.ctrl;
        sp->chunk_ptr = pp.chunk_ptr;
        sp->nbatch = pp.nbatch;
        sp->n_slots = pp.n_slots;
        sp->zero_slot = pp.zero_slot;
        sp->nlevels = tp.nlevels;
        sp->n_mac = tp.n_mac;
        sp->n_val = tp.n_val;
        if (std::getenv("KLUJAX_CUDA_VERBOSE"))
            std::fprintf(stderr, "[program] mode %d rc %d wide %d compact %d tail_t0 %d: %d steps/batches, %d levels (stage at %d), %lld multiply-adds, %d value slots, %d chunks\n",
                         mode, rc, int(wide), int(compact), tp.tail_t0, pp.nbatch, tp.nlevels, tp.tail_level,
                         static_cast<long long>(tp.n_mac), tp.n_val, pp.nchunks);
        if (tp.tail_t0 >= 0) {
            CUDA_OK(sp->tail_tab.upload(tp.tail_tab));
            sp->tail_t0 = tp.tail_t0;
        }
        if (!tp.remap.empty()) {
            CUDA_OK(sp->remap.upload(tp.remap));
            std::vector<int> rs(sd.lay.rs_slot.size());
            for (size_t q = 0; q < rs.size(); ++q) rs[q] = tp.remap[sd.lay.rs_slot[q]];
            CUDA_OK(sp->rs_slot.upload(rs));
        }
        it = sd.small.emplace(key, std::move(sp)).first;
    }
    *out = it->second.get();
    return 0;
}

static int get_level_prog(Symbolic* S, int cx, int mode, long long batch_cols, LevelProg** out, bool native = false) {
    Seeded& sd = S->seeded[cx];
    // Task granularity (plan.cpp: Leveler): with few batch columns the time per level is one
    // memory round trip, so every multiply-add is applied as soon as its operands are final
    // (split 2: cfg 4 refactor+solves 193 -> 40 ms per step); with thousands of columns the
    // kernels are traffic-bound and partial updates are grouped (>= 8 multiply-adds, split 3)
    // or not split at all for the factorization (split 0).  A factorization with millions of
    // multiply-adds (cfg 1: 1.65e7 for n = 1000, cfg 3: 2.6e8) is never expanded into one task
    // per multiply-add - the schedule alone would be hundreds of MB and seconds of host time
    // (cfg 1 `solve`: 3.2 s per call) - but into groups of >= 8 (cfg 1: factor + solve of one
    // system 39 -> 19 ms, schedule built in 0.2 s).
    int split;
    if (mode == PM_FACTOR) {
        if (batch_cols >= 2048) split = 0;
        else split = sd.info_mac >= 4000000 ? 3 : 2;
    } else {
        split = batch_cols >= 4096 ? 3 : 2;
    }
    if (const char* e = std::getenv("KLUJAX_CUDA_LEVEL_SPLIT")) split = std::atoi(e);
    auto key = std::make_pair(mode, split + (native ? 1000 : 0));
    auto it = sd.level.find(key);
    if (it == sd.level.end()) {
        TaskProgram tp;
        build_program(S->ord, sd.piv, sd.lay, mode, 1, tp, split);
        auto lp = std::make_unique<LevelProg>();
        std::string err;
        // native: the rows of X carry the names of their final positions in x
        const int* xrow = !native ? nullptr : (mode == PM_TSOLVE ? sd.piv.Pnum.data() : S->ord.Q.data());
        if (!lp->upload(sd.lay, tp, err, xrow)) return fail(KLU_OUT_OF_MEMORY, err);
        if (mode == PM_FACTOR) sd.info_levels = tp.nlevels;
        it = sd.level.emplace(key, std::move(lp)).first;
    }
    *out = it->second.get();
    return 0;
}

// Dependency-driven regime: programs are built on first use; nullptr = the pattern does not
// qualify (the level kernels take over).
static DepColProg* get_dep_col(Symbolic* S, int cx) {
    Seeded& sd = S->seeded[cx];
    if (!sd.dep_col_tried) {
        sd.dep_col_tried = true;
        int64_t cap = 64LL << 20;  // multiply-adds: the destination list is 4 bytes each
        if (const char* e = std::getenv("KLUJAX_CUDA_DEP_MAXMAC")) cap = std::atoll(e);
        if (sd.info_mac <= cap) {
            ColProgram cp;
            if (build_col_program(sd.piv, sd.lay, cp) && dep_factor_plan(sd.lay.n, cp.maxcol, cx ? 16 : 8).ok) {
                auto dp = std::make_unique<DepColProg>();
                if (dp->upload(cp)) sd.dep_col = std::move(dp);
                else cudaGetLastError();
            }
        }
        if (std::getenv("KLUJAX_CUDA_VERBOSE"))
            std::fprintf(stderr, "[dep] column program: %s (mac %lld)\n", sd.dep_col ? "built" : "not used", static_cast<long long>(sd.info_mac));
    }
    return sd.dep_col.get();
}

static DepPanelProg* get_dep_panel(Symbolic* S, int cx, bool transpose) {
    Seeded& sd = S->seeded[cx];
    auto it = sd.dep_panel.find(transpose ? 1 : 0);
    if (it == sd.dep_panel.end()) {
        std::unique_ptr<DepPanelProg> dp;
        PanelProgram pp;
        // the right-hand side has to fit in shared memory (checked before the program is built)
        if (static_cast<size_t>(sd.lay.n) * (cx ? 16 : 8) < kDepSmemMax && build_panel_program(S->ord, sd.piv, sd.lay, transpose, pp) &&
            panel_plan(sd.lay.n, pp.ntemps, pp.maxI, cx ? 16 : 8).ok) {
            dp = std::make_unique<DepPanelProg>();
            if (!dp->upload(pp)) {
                cudaGetLastError();
                dp.reset();
            }
        }
        if (std::getenv("KLUJAX_CUDA_VERBOSE"))
            std::fprintf(stderr, "[dep] panel program (transpose %d): %s, %d panels, %d temporaries, depth E %d I %d\n", int(transpose),
                         dp ? "built" : "not used", pp.npanels, pp.ntemps, pp.maxE, pp.maxI);
        it = sd.dep_panel.emplace(transpose ? 1 : 0, std::move(dp)).first;
    }
    return it->second.get();
}

static GroupProg* get_group_prog(Symbolic* S, int cx, bool transpose) {
    Seeded& sd = S->seeded[cx];
    auto it = sd.grp.find(transpose ? 1 : 0);
    if (it == sd.grp.end()) {
        std::unique_ptr<GroupProg> gp;
        GroupProgram hp;
        // rows carry the names of their final positions in the caller's x (the solve runs on it)
        if (build_group_program(S->ord, sd.piv, sd.lay, transpose, transpose ? sd.piv.Pnum.data() : S->ord.Q.data(), hp)) {
            gp = std::make_unique<GroupProg>();
            if (!gp->upload(hp)) {
                cudaGetLastError();
                gp.reset();
            }
        }
        if (std::getenv("KLUJAX_CUDA_VERBOSE"))
            std::fprintf(stderr, "[level] group program (transpose %d): %s, %d groups, %d levels, %.2f rows per read of X\n", int(transpose),
                         gp ? "built" : "not used", hp.ngroups, hp.nlevels, hp.n_ksteps ? double(hp.n_mac) / double(hp.n_ksteps) : 0.0);
        it = sd.grp.emplace(transpose ? 1 : 0, std::move(gp)).first;
    }
    return it->second.get();
}

// ---------------------------------------------------------------------------
// the one numeric driver behind every entry point
// ---------------------------------------------------------------------------
struct Op {
    int mode = PM_FACTOR_SOLVE;
    int L = 0, r = 0, nz = 0;
    const int32_t *Ai = nullptr, *Aj = nullptr;
    const void* Ax = nullptr;
    long long ax_stride = 0;
    const void* b = nullptr;
    long long b_stride = 0;
    void* x = nullptr;
    NumericBatch* batch = nullptr;  // store (factor modes) or load (solve modes)
    bool store_numeric = false;
    bool certify = true;  // false: klu_refactor semantics (the pivot order of the handle is kept, no certificate)
    const int* sysidx_dev = nullptr;
    int32_t* status = nullptr;
    double* growth = nullptr;
    cudaStream_t st = nullptr;
};

template <typename T>
static int launch_small(Symbolic* S, int cx, const Op& op) {
    Seeded& sd = S->seeded[cx];
    const int n = sd.lay.n;
    const bool factor = op.mode <= PM_FACTOR_TSOLVE;
    const bool solve = op.mode != PM_FACTOR;
    const bool transpose = (op.mode == PM_FACTOR_TSOLVE || op.mode == PM_TSOLVE);
    const size_t vsz = sizeof(typename Sc<T>::V);
    int rc = 0;
    SmallProg *p0 = nullptr, *p1 = nullptr;
    if (solve) {
        // widest right-hand-side chunk that still fits the slot fields / shared memory
        for (rc = std::min(op.r, 4); rc >= 1; --rc) {
            const size_t slots = static_cast<size_t>(sd.lay.n_val) + static_cast<size_t>(n) * rc + 3;
            if (slots <= 16383 && slots * vsz + n * 8ull + 16 <= kMaxDynSmem - kSmallHeaderBytes) break;
        }
        if (rc < 1) return fail(KLU_TOO_LARGE, "system does not fit the shared-memory regime");
    }
    // Two shared-memory kernels: one warp per system (kernels_small.cuh) or four warps per
    // system (kernels_wide.cuh).  The latter shortens the dependent chain of a system and wins
    // when shared memory holds few systems per SM (cfg 5: 5 systems, 8.2 vs 10.7 ms); with many
    // small systems per SM the one-warp form keeps more of them in flight (cfg 2: 15 systems,
    // 0.15 vs 0.29 ms).
    bool wide;
    {
        const size_t slots = static_cast<size_t>(sd.lay.n_val) + static_cast<size_t>(n) * rc + 3;
        const size_t fit = (kMaxDynSmem - kSmallHeaderBytes) / (slots * vsz + n * 8ull + 16);
        wide = fit <= 8;
        if (const char* ev = std::getenv("KLUJAX_CUDA_WIDE")) wide = std::atoi(ev) != 0;
    }
    // Fused factor + solve that keeps nothing: the program recycles dead value slots
    // (plan.hpp: TaskProgram::remap), so more systems fit in shared memory.
    bool compact = wide && factor && solve && !op.store_numeric && op.r <= rc;
    if (const char* ev = std::getenv("KLUJAX_CUDA_COMPACT")) compact = compact && std::atoi(ev) != 0;
    // dense tail in registers (kernels_wide.cuh): fused factor + solve with one right-hand side.
    // An experiment, OFF unless KLUJAX_CUDA_TAIL=1: it halves the interpreted steps of config 5 but
    // runs 7.4 ms against the interpreter's 4.97 ms (profiles/tail_stage_r02.md).
    bool tail = false;
    if (const char* ev = std::getenv("KLUJAX_CUDA_TAIL"))
        tail = std::atoi(ev) != 0 && wide && op.mode == PM_FACTOR_SOLVE && !op.store_numeric && rc == 1 && op.r == 1;
    int e = get_small_prog(S, cx, op.mode, rc, wide, compact, &p0, tail);
    if (!e && solve && op.r > rc) e = get_small_prog(S, cx, transpose ? PM_TSOLVE : PM_SOLVE, rc, wide, false, &p1);
    if (e == KLU_TOO_LARGE && wide) {
        // the single-multiply-add schedule of the multi-warp kernel has too many steps for the
        // control-word table (large n): the one-warp kernel's dot-form schedule packed at
        // seeding time, so it is the fallback inside this regime
        wide = false;
        p0 = p1 = nullptr;
        e = get_small_prog(S, cx, op.mode, rc, false, false, &p0);
        if (!e && solve && op.r > rc) e = get_small_prog(S, cx, transpose ? PM_TSOLVE : PM_SOLVE, rc, false, false, &p1);
    }
    if (e) return e;
    static thread_local SmallArgs a;  // ~14 KB of parameters: control words travel by value
    std::memset(&a, 0, sizeof(a));
    a.stream[0] = p0->stream.as<uint32_t>();
    a.nchunks[0] = p0->nchunks;
    a.nbatch[0] = p0->nbatch;
    std::copy(p0->ctrl.begin(), p0->ctrl.end(), a.ctrl[0]);
    std::copy(p0->chunk_ptr.begin(), p0->chunk_ptr.end(), a.chunk_ptr[0]);
    if (p1) {
        a.stream[1] = p1->stream.as<uint32_t>();
        a.nchunks[1] = p1->nchunks;
        a.nbatch[1] = p1->nbatch;
        std::copy(p1->ctrl.begin(), p1->ctrl.end(), a.ctrl[1]);
        std::copy(p1->chunk_ptr.begin(), p1->chunk_ptr.end(), a.chunk_ptr[1]);
    }
    a.n = n;
    a.n_val = p0->n_val;
    a.n_slots = p0->n_slots;
    a.zero_slot = p0->zero_slot;
    a.rc = rc;
    a.nz = op.nz;
    a.flags = (factor ? SF_FACTOR : SF_LOAD_NUMERIC) | (solve ? SF_SOLVE : 0) |
              (op.store_numeric ? SF_STORE_NUMERIC : 0) | (op.certify ? SF_CERTIFY : 0) |
              (solve ? (transpose ? SF_SCALE_OUT : SF_SCALE_IN) : 0);
    a.coo_dst = S->cur->coo_dst.as<int>();
    a.pattern_bad = S->cur->bad.as<int>();
    a.rs_ptr = sd.d_rs_ptr.as<int>();
    a.rs_slot = p0->rs_slot.p ? p0->rs_slot.as<int>() : sd.d_rs_slot.as<int>();
    a.tail_tab = p0->tail_t0 >= 0 ? p0->tail_tab.as<uint16_t>() : nullptr;
    a.tail_t0 = p0->tail_t0;
    a.remap = p0->remap.p ? p0->remap.as<int>() : nullptr;
    a.in_perm = transpose ? sd.d_q.as<int>() : sd.d_pnum.as<int>();
    a.out_perm = transpose ? sd.d_pnum.as<int>() : sd.d_q.as<int>();
    a.Ax = op.Ax;
    a.ax_stride = op.ax_stride;
    a.b = op.b;
    a.b_stride = op.b_stride;
    a.x = op.x;
    a.r = op.r;
    if (op.batch) {
        a.Vstore = op.batch->V.p;
        a.Rsstore = op.batch->Rs.as<double>();
    }
    a.sysidx = op.sysidx_dev;
    a.status = op.status;
    a.growth = op.growth;
    a.L = op.L;
    // CTA = W systems + the producer warp + the schedule ring.  One-warp form: one consumer warp
    // per system, as many systems as fit.  Multi-warp form: kWarpsPerSystem warps per system,
    // at most 5 systems per CTA (thread limit), several CTAs per SM when systems are small.
    const size_t per_sys = wide ? ((static_cast<size_t>(a.n_slots) * vsz + n * 8ull + 16 + 15) & ~size_t(15))
                                : ((static_cast<size_t>(a.n_slots) * vsz + n * 8ull + 15) & ~size_t(15));
    const bool tail_kernel = wide && p0->tail_t0 >= 0;
    int W = static_cast<int>(std::min<size_t>(wide ? (tail_kernel ? kMaxWideSystemsTail : kMaxWideSystems) : 16,
                                              (kMaxDynSmem - kSmallHeaderBytes) / per_sys));
    if (W < 1) return fail(KLU_TOO_LARGE, "system does not fit the shared-memory regime");
    const int sms = num_sms();
    while (W > 1 && (op.L + W - 1) / W < sms) --W;  // small batches: spread over all SMs first
    if (const char* fw = std::getenv("KLUJAX_CUDA_SMALL_W")) W = std::max(1, std::min(W, std::atoi(fw)));
    a.W = W;
    const size_t smem = kSmallHeaderBytes + per_sys * W;
    const int threads = wide ? 32 * (kWarpsPerSystem * W + 1) : 32 * (W + 1);
    auto kern = wide ? (tail_kernel ? wide_kernel<T, true> : wide_kernel<T, false>) : small_kernel<T>;
    CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    int occ = 0;
    CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
    if (occ < 1) return fail(KLU_TOO_LARGE, "shared-memory regime: zero occupancy");
    const long long ctas = (op.L + W - 1) / W;
    const int grid = static_cast<int>(std::min<long long>(ctas, static_cast<long long>(occ) * sms));
    static const bool do_prof = std::getenv("KLUJAX_CUDA_PROFILE") != nullptr;
    unsigned long long* dprof = nullptr;
    if (do_prof) {
        cudaMalloc(reinterpret_cast<void**>(&dprof), 64);
        cudaMemset(dprof, 0, 64);
        a.prof = dprof;
    }
    kern<<<grid, threads, smem, op.st>>>(a);
    if (do_prof) {
        unsigned long long h[8];
        cudaStreamSynchronize(op.st);
        cudaMemcpy(h, dprof, 64, cudaMemcpyDeviceToHost);
        cudaFree(dprof);
        const double nw = h[5] ? static_cast<double>(h[5]) : 1.0, per = static_cast<double>(op.L) / nw;
        std::fprintf(stderr, "[PHASES] L=%d W=%d grid=%d warps=%llu cycles/system: load+scale %.0f | rhs %.0f | ring-wait %.0f | run %.0f | store %.0f\n",
                     op.L, W, grid, h[5], h[0] / nw / per, h[1] / nw / per, h[2] / nw / per, h[3] / nw / per, h[4] / nw / per);
    }
    CUDA_OK(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------
// register-resident regime (rowreg.hpp): plan, generate PTX, JIT, launch
// ---------------------------------------------------------------------------
// Opt-in (KLUJAX_CUDA_ROWREG=1): measured on B200 the generated straight-line kernels are bound by
// instruction fetch (cfg 5: ~32 000 SASS instructions = 0.5 MB per system, ncu: 44 % of the stall
// samples are "no instruction", IPC 0.25 per sub-partition) and lose to the shared-memory
// interpreters (7.4 vs 5.0 ms), see profiles/rowreg_r02.md.
static bool rowreg_enabled() {
    const char* e = std::getenv("KLUJAX_CUDA_ROWREG");  // read per call: tests switch it
    return e ? std::atoi(e) != 0 : false;
}

static RRKernel* get_rowreg(Symbolic* S, int cx, bool transpose, int r) {
    Seeded& sd = S->seeded[cx];
    auto key = std::make_pair(transpose ? 1 : 0, r);
    auto it = sd.rowreg.find(key);
    if (it != sd.rowreg.end()) return it->second.get();
    auto k = std::make_unique<RRKernel>();
    RRKernel* K = k.get();
    sd.rowreg.emplace(key, std::move(k));
    rr::Options opt;
    opt.r = r;
    opt.transpose = transpose;
    opt.max_regs = cx ? 60 : 120;  // beyond this the assembler spills too much (cfg 5 needs 58)
    if (const char* e = std::getenv("KLUJAX_CUDA_RR_PINNED")) opt.pinned_rows = std::atoi(e);
    if (const char* e = std::getenv("KLUJAX_CUDA_RR_MAXREGS")) opt.max_regs = std::atoi(e);
    if (const char* e = std::getenv("KLUJAX_CUDA_RR_REBALANCE")) opt.rebalance = std::atoi(e) != 0;
    if (const char* e = std::getenv("KLUJAX_CUDA_RR_LEVEL_ORDER")) opt.level_order = std::atoi(e) != 0;
    if (const char* e = std::getenv("KLUJAX_CUDA_RR_STRIPE")) opt.stripe = std::atoi(e) != 0;
    if (const char* e = std::getenv("KLUJAX_CUDA_RR_STRIPE_MIN")) opt.stripe_min = std::atoi(e);
    if (!rr::build_plan(S->A, S->ord, sd.piv, sd.lay, cx != 0, opt, K->plan)) {
        K->why = K->plan.why;
        return K;
    }
    // warps per CTA: as many systems as shared memory holds, at most 8 (255 registers per thread)
    {
        rr::PtxKernel probe;
        rr::emit_ptx(K->plan, probe);
        const size_t per_warp = std::max<size_t>(probe.smem_per_warp, 1);
        int wpc = static_cast<int>(std::min<size_t>(8, kMaxDynSmem / per_warp));
        if (const char* e = std::getenv("KLUJAX_CUDA_RR_WPC")) wpc = std::max(1, std::min(wpc, std::atoi(e)));
        if (wpc < 1) {
            K->why = "shared memory of one system exceeds the CTA limit";
            return K;
        }
        K->pk.warps_per_cta = wpc;
        if (const char* e = std::getenv("KLUJAX_CUDA_RR_LOCKSTEP")) K->pk.lockstep = std::atoi(e) != 0;
    }
    if (!rr::emit_ptx(K->plan, K->pk)) {
        K->why = "PTX rendering failed";
        return K;
    }
    if (const char* dump = std::getenv("KLUJAX_CUDA_RR_DUMP_PTX")) {
        if (FILE* f = std::fopen(dump, "w")) {
            std::fwrite(K->pk.ptx.data(), 1, K->pk.ptx.size(), f);
            std::fclose(f);
        }
    }
    std::vector<char> elog(16384, 0);
    cudaJitOption jopt[2] = {cudaJitErrorLogBuffer, cudaJitErrorLogBufferSizeBytes};
    void* jval[2] = {elog.data(), reinterpret_cast<void*>(static_cast<uintptr_t>(elog.size()))};
    cudaError_t e = cudaLibraryLoadData(&K->lib, K->pk.ptx.c_str(), jopt, jval, 2, nullptr, nullptr, 0);
    if (e != cudaSuccess) {
        K->why = std::string("JIT of the generated PTX failed: ") + cudaGetErrorString(e) + " " + elog.data();
        cudaGetLastError();
        K->lib = nullptr;
        return K;
    }
    e = cudaLibraryGetKernel(&K->kern, K->lib, K->pk.name.c_str());
    if (e != cudaSuccess) {
        K->why = std::string("cudaLibraryGetKernel: ") + cudaGetErrorString(e);
        cudaGetLastError();
        return K;
    }
    const size_t smem = K->pk.smem_per_warp * K->pk.warps_per_cta;
    e = cudaFuncSetAttribute(reinterpret_cast<const void*>(K->kern), cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(smem));
    cudaFuncAttributes fa{};
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, reinterpret_cast<const void*>(K->kern));
    int occ = 0;
    if (e == cudaSuccess)
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, reinterpret_cast<const void*>(K->kern),
                                                          32 * K->pk.warps_per_cta, smem);
    if (e != cudaSuccess || occ < 1) {
        K->why = std::string("generated kernel cannot be launched: ") + cudaGetErrorString(e);
        cudaGetLastError();
        return K;
    }
    K->regs = fa.numRegs;
    K->ctas_per_sm = occ;
    {
        std::vector<int> sp(sd.lay.sorted_slot.size());
        for (size_t q = 0; q < sp.size(); ++q) sp[q] = K->plan.place[sd.lay.slot_csc[sd.lay.sorted_slot[q]]];
        if (K->splace.upload(sp) != cudaSuccess) {
            K->why = "out of device memory for the slot map";
            return K;
        }
    }
    if (K->tabs.upload(K->plan.tabs) != cudaSuccess) {
        K->why = "out of device memory for the tables";
        return K;
    }
    if (std::getenv("KLUJAX_CUDA_RR_VERBOSE"))
        std::fprintf(stderr, "[rowreg] n=%d r=%d: %d value registers, %d registers/thread, %zu B local, %d warps/CTA x %d CTA/SM, %zu B smem/warp, %zu instructions\n",
                     K->plan.n, r, K->plan.P, fa.numRegs, fa.localSizeBytes, K->pk.warps_per_cta, occ,
                     K->pk.smem_per_warp, K->plan.ins.size());
    K->usable = true;
    return K;
}

static int launch_rowreg(Symbolic* S, int cx, RRKernel* K, const Op& op) {
    Seeded& sd = S->seeded[cx];
    const int nz = op.nz;
    coo_map_kernel<<<(nz + 255) / 256, 256, 0, op.st>>>(nz, sd.lay.n, op.Ai, op.Aj, sd.d_colptr.as<int>(),
                                                       sd.d_srow.as<int>(), K->splace.as<int>(),
                                                       S->cur->coo_csc.as<int>(), S->cur->bad.as<int>());
    CUDA_OK(cudaGetLastError());
    const void* Ax = op.Ax;
    const void* b = op.b;
    void* x = op.x;
    const void* tabs = K->tabs.p;
    const void* dst = S->cur->coo_csc.p;
    void* status = op.status;
    void* growth = op.growth;
    const void* bad = S->cur->bad.p;
    unsigned long long axs = static_cast<unsigned long long>(op.ax_stride), bs = static_cast<unsigned long long>(op.b_stride);
    unsigned int L = static_cast<unsigned int>(op.L);
    static const bool do_prof = std::getenv("KLUJAX_CUDA_PROFILE") != nullptr;
    unsigned long long* dprof = nullptr;
    if (do_prof) {
        cudaMalloc(reinterpret_cast<void**>(&dprof), 64);
        cudaMemset(dprof, 0, 64);
    }
    void* args[] = {&Ax, &b, &x, &tabs, &dst, &status, &growth, &bad, &axs, &bs, &dprof, &L};
    const int wpc = K->pk.warps_per_cta;
    const long long ctas = (static_cast<long long>(op.L) + wpc - 1) / wpc;
    const int grid = static_cast<int>(std::min<long long>(ctas, static_cast<long long>(K->ctas_per_sm) * num_sms()));
    CUDA_OK(cudaLaunchKernel(reinterpret_cast<const void*>(K->kern), dim3(grid), dim3(32 * wpc), args,
                             K->pk.smem_per_warp * wpc, op.st));
    if (do_prof) {
        unsigned long long h[8];
        cudaStreamSynchronize(op.st);
        cudaMemcpy(h, dprof, 64, cudaMemcpyDeviceToHost);
        cudaFree(dprof);
        const double ns = h[4] ? static_cast<double>(h[4]) : 1.0;
        std::fprintf(stderr, "[PHASES rowreg] L=%d grid=%d x %d warps: cycles/system: stage %.0f | scale+load %.0f | eliminate %.0f | back-solve+store %.0f\n",
                     op.L, grid, wpc, h[0] / ns, h[1] / ns, h[2] / ns, h[3] / ns);
    }
    return 0;
}

static int run_numeric(Symbolic* S, int cx, const Op& op) {
    std::lock_guard<std::mutex> lock(S->mu);
    {
        // the device copies of the schedule live on ONE device: a handle is bound to the device of
        // its first numeric call (one process driving several GPUs analyzes once per device)
        int dev = 0;
        cudaGetDevice(&dev);
        if (S->device < 0) S->device = dev;
        if (S->device != dev)
            return fail(KLU_INVALID, "symbolic handle belongs to CUDA device " + std::to_string(S->device) +
                                         ", the call runs on device " + std::to_string(dev));
    }
    Seeded& sd = S->seeded[cx];
    const bool factor = op.mode <= PM_FACTOR_TSOLVE;
    if (op.L <= 0) return 0;
    {
        auto& slot = S->scratch[op.st];
        if (!slot) {
            slot = std::make_unique<Symbolic::Scratch>();
            CUDA_OK(slot->coo_dst.alloc(sizeof(int) * (S->A.rowidx.size() + 1)));
            CUDA_OK(slot->coo_csc.alloc(sizeof(int) * (S->A.rowidx.size() + 1)));
            CUDA_OK(slot->bad.alloc(sizeof(int)));
            CUDA_OK(cudaMemset(slot->bad.p, 0, sizeof(int)));
        }
        S->cur = slot.get();
    }
    if (factor && op.nz <= 0) return fail(KLU_SINGULAR, "matrix without entries is singular");
    if (factor) {
        int e = ensure_seeded_dev(S, cx, op.nz, op.Ai, op.Aj, op.Ax, op.st);
        if (e) return e;
        CUDA_OK(cudaMemsetAsync(S->cur->bad.p, 0, sizeof(int), op.st));
        const int nz = op.nz;
        coo_map_kernel<<<(nz + 255) / 256, 256, 0, op.st>>>(
            nz, sd.lay.n, op.Ai, op.Aj, sd.d_colptr.as<int>(), sd.d_srow.as<int>(),
            sd.d_sslot.as<int>(), S->cur->coo_dst.as<int>(), S->cur->bad.as<int>());
        CUDA_OK(cudaGetLastError());
    } else if (!sd.ok) {
        return fail(KLU_INVALID, "numeric handle without a factorization");
    }
    if (sd.regime == 0 && op.mode == PM_FACTOR_SOLVE && !op.store_numeric && op.r >= 1 && op.r <= 4 &&
        op.status && op.growth && rowreg_enabled()) {
        // fused factor + solve that keeps nothing: the register-resident kernel, if the pattern fits
        RRKernel* K = get_rowreg(S, cx, false, op.r);
        if (K->usable) return launch_rowreg(S, cx, K, op);
        if (std::getenv("KLUJAX_CUDA_ROWREG_REQUIRE")) return fail(KLU_TOO_LARGE, "register-resident kernel unavailable: " + K->why);
    }
    if (sd.regime == 0) return cx ? launch_small<zc>(S, 1, op) : launch_small<double>(S, 0, op);
    // level-scheduled regime
    LevelCall lc;
    lc.mode = op.mode;
    lc.L = op.L;
    lc.r = op.r;
    lc.nz = op.nz;
    lc.Ax = op.Ax;
    lc.ax_stride = op.ax_stride;
    lc.b = op.b;
    lc.b_stride = op.b_stride;
    lc.x = op.x;
    lc.coo_dst = S->cur->coo_dst.as<int>();
    lc.pattern_bad = S->cur->bad.as<int>();
    lc.rs_ptr = sd.d_rs_ptr.as<int>();
    lc.rs_slot = sd.d_rs_slot.as<int>();
    lc.pnum = sd.d_pnum.as<int>();
    lc.q = sd.d_q.as<int>();
    lc.src_solve = sd.d_src_solve.as<int>();
    lc.src_tsolve = sd.d_src_tsolve.as<int>();
    lc.sysidx = op.sysidx_dev;
    lc.status = op.status;
    lc.growth = op.growth;
    lc.certify = op.certify ? 1 : 0;
    lc.st = op.st;
    lc.n = sd.lay.n;
    lc.n_val = sd.lay.n_val;
    lc.is_complex = cx;
    lc.num_sms = num_sms();
    {
        // bad[L], growth2[L] and the arrival counters of the cooperative kernels, per stream
        const size_t cols = static_cast<size_t>(op.L) * static_cast<size_t>(std::max(op.r, 1));
        const size_t need = 256 + 12 * static_cast<size_t>(op.L) + (cols <= 1024 ? 4 * cols : 0) + 64;
        if (S->cur->aux_bytes < need) {
            const size_t grow = std::max(need, 2 * S->cur->aux_bytes);
            DevBuf nb;
            if (nb.alloc(grow) == cudaSuccess) {
                std::swap(S->cur->aux.p, nb.p);  // the old buffer is freed with nb (cudaFree waits for work in flight)
                std::swap(S->cur->aux.bytes, nb.bytes);
                S->cur->aux_bytes = grow;
            } else {
                cudaGetLastError();
            }
        }
        lc.aux = S->cur->aux.p;
        lc.aux_bytes = S->cur->aux_bytes;
        if (factor && !S->cur->slot_src.p && S->cur->slot_src.alloc(sizeof(int) * (static_cast<size_t>(sd.lay.n_val) + 1)) != cudaSuccess)
            cudaGetLastError();
        lc.slot_src = S->cur->slot_src.as<int>();
    }
    if (op.batch) {
        lc.V = op.batch->V.p;
        lc.Rs = op.batch->Rs.as<double>();
        lc.lpad = op.batch->lpad;
        lc.sysmajor = op.batch->sysmajor;
        lc.Lstore = op.batch->L;
    }
    LevelProg *pf = nullptr, *ps = nullptr;
    // Batches that are not larger than the machine: the dependency-driven kernels (a CTA per
    // system / right-hand-side column, no level barriers), if the pattern qualifies.
    const bool dep_on = dep_enabled();
    if (factor) {
        // the column kernel is an experiment that did not beat the level kernel (profiles/dep_r02.md): opt-in
        if (dep_on && dep_env_int("KLUJAX_CUDA_DEP_FACTOR", 0) && op.L <= lc.num_sms) lc.dep_col = get_dep_col(S, cx);
        if (lc.dep_col) {
            sd.info_levels = lc.dep_col->depth;
        } else {
            int e = get_level_prog(S, cx, PM_FACTOR, op.L, &pf);
            if (e) return e;
        }
    }
    if (op.mode != PM_FACTOR) {
        const bool tr = (op.mode == PM_FACTOR_TSOLVE || op.mode == PM_TSOLVE);
        const long long Q = static_cast<long long>(op.L) * op.r;
        if (dep_on && Q <= 2LL * lc.num_sms) lc.dep_panel = get_dep_panel(S, cx, tr);
        // many right-hand-side columns per system: the row-blocked solve on the caller's x
        if (!lc.dep_panel && dep_env_int("KLUJAX_CUDA_GROUPED", 1) && op.r >= 16 && Q >= 2048 && level_native(Q, op.r, lc.num_sms))
            lc.grp = get_group_prog(S, cx, tr);
        if (!lc.dep_panel && !lc.grp) {
            int e = get_level_prog(S, cx, tr ? PM_TSOLVE : PM_SOLVE, Q, &ps, level_native(Q, op.r, lc.num_sms));
            if (e) return e;
        }
    }
    std::string err;
    int rc = level_run(lc, pf, ps, err);
    if (rc) return fail(rc, err);
    return 0;
}

// ---------------------------------------------------------------------------
// numeric batches (device-resident counterpart of klu_numeric* arrays)
// ---------------------------------------------------------------------------
static int new_batch(Symbolic* S, int cx, int L, NumericBatch** out) {
    Seeded& sd = S->seeded[cx];
    auto nb = std::make_unique<NumericBatch>();
    nb->sym = S;
    nb->parent = S;
    nb->is_complex = cx;
    nb->regime = sd.regime;
    nb->L = L;
    nb->live = L;
    nb->freed.assign(L, 0);
    const size_t vsz = cx ? 16 : 8;
    nb->lpad = (L + 31) / 32 * 32;
    nb->sysmajor = (sd.regime != 0 && want_sysmajor(L, num_sms())) ? 1 : 0;
    const size_t nsys = sd.regime == 0 ? static_cast<size_t>(L) : static_cast<size_t>(nb->lpad);
    CUDA_OK(nb->V.alloc(nsys * sd.lay.n_val * vsz));
    CUDA_OK(nb->Rs.alloc(nsys * sd.lay.n * sizeof(double)));
    {
        std::lock_guard<std::mutex> g(g_reg_mu);
        nb->id = g_next_batch++;
        g_batches[nb->id] = nb.get();
    }
    *out = nb.release();
    return 0;
}

// A handle array may mix elements of several batches (systems that needed their own pivot
// sequence live in re-seeded child batches): a group = the positions of one batch.
struct NumGroup {
    NumericBatch* nb = nullptr;
    std::vector<int> pos, el;
    bool identity() const {  // positions 0..L-1 hold elements 0..L-1 of the whole batch
        if (static_cast<int>(pos.size()) != nb->L) return false;
        for (size_t q = 0; q < pos.size(); ++q)
            if (pos[q] != static_cast<int>(q) || el[q] != static_cast<int>(q)) return false;
        return true;
    }
};
static int resolve_groups(int n, const uint64_t* h, Symbolic* S, int cx, std::vector<NumGroup>& out) {
    if (n <= 0) return fail(KLU_INVALID, "empty numeric handle array");
    out.clear();
    std::lock_guard<std::mutex> g(g_reg_mu);
    for (int m = 0; m < n; ++m) {
        if (h[m] == 0) return fail(KLU_INVALID, "numeric pointer is null");
        const uint32_t id = static_cast<uint32_t>(h[m] >> 32);
        const int el = static_cast<int>(h[m] & 0xffffffffu) - 1;
        auto it = g_batches.find(id);
        if (it == g_batches.end()) return fail(KLU_INVALID, "stale numeric handle");
        NumericBatch* nb = it->second;
        if (el < 0 || el >= nb->L || nb->freed[el]) return fail(KLU_INVALID, "stale numeric handle");
        if (nb->parent != S || nb->is_complex != cx)
            return fail(KLU_INVALID, "numeric handle belongs to another symbolic handle / dtype");
        NumGroup* grp = nullptr;
        for (NumGroup& q : out)
            if (q.nb == nb) grp = &q;
        if (!grp) {
            out.emplace_back();
            grp = &out.back();
            grp->nb = nb;
        }
        grp->pos.push_back(m);
        grp->el.push_back(el);
    }
    return 0;
}

// index list on the device, stream-ordered: no cudaMalloc, no synchronisation (the copy out of
// pageable memory is staged before cudaMemcpyAsync returns)
struct StreamBuf {
    void* p = nullptr;
    cudaStream_t st = nullptr;
    ~StreamBuf() {
        if (p) cudaFreeAsync(p, st);
    }
    cudaError_t alloc(size_t bytes, cudaStream_t s) {
        st = s;
        return cudaMallocAsync(&p, bytes ? bytes : 16, s);
    }
    template <typename U>
    cudaError_t upload(const std::vector<U>& v, cudaStream_t s) {
        cudaError_t e = alloc(v.size() * sizeof(U), s);
        if (e != cudaSuccess || v.empty()) return e;
        return cudaMemcpyAsync(p, v.data(), v.size() * sizeof(U), cudaMemcpyHostToDevice, s);
    }
    template <typename U>
    U* as() const {
        return static_cast<U*>(p);
    }
};

}  // namespace kb

using namespace kb;

// ===========================================================================
// extern "C"
// ===========================================================================
extern "C" {

int klujax_cuda_abi_version(void) { return KLUJAX_CUDA_ABI_VERSION; }
const char* klujax_cuda_last_error(void) { return g_err.c_str(); }

int klujax_cuda_analyze(int n_col, int n_nz, const int32_t* Ai, const int32_t* Aj, uint64_t* out) {
    if (out) *out = 0;
    if (!out || n_col <= 0 || n_nz < 0 || (n_nz > 0 && (!Ai || !Aj)))
        return fail(KLU_INVALID, "analyze: bad arguments");
    for (int e = 0; e < n_nz; ++e) {
        if (Ai[e] < 0 || Ai[e] >= n_col) return fail(KLU_INVALID, "Ai index out of bounds.");
        if (Aj[e] < 0 || Aj[e] >= n_col) return fail(KLU_INVALID, "Aj index out of bounds.");
    }
    auto S = std::make_unique<Symbolic>();
    S->A = coo_to_csc(n_col, n_nz, Ai, Aj);
    S->Ai.assign(Ai, Ai + n_nz);
    S->Aj.assign(Aj, Aj + n_nz);
    if (!analyze_pattern(S->A, S->ord)) return fail(KLU_INVALID, "klu_analyze failed.");
    *out = static_cast<uint64_t>(reinterpret_cast<uintptr_t>(S.release()));
    return 0;
}

int klujax_cuda_free_symbolic(uint64_t symbolic, int32_t* freed) {
    if (symbolic == 0) {
        if (freed) *freed = 0;
        return 0;
    }
    delete sym_of(symbolic);
    if (freed) *freed = 1;
    return 0;
}

int klujax_cuda_seed_host(uint64_t symbolic, int is_complex, const double* Ax_host) {
    Symbolic* S = sym_of(symbolic);
    if (!S || !Ax_host) return fail(KLU_INVALID, "seed_host: null argument");
    std::lock_guard<std::mutex> lock(S->mu);
    const int nz = static_cast<int>(S->Ai.size());
    if (is_complex)
        return seed_from_coo<std::complex<double>>(S, 1, nz, S->Ai.data(), S->Aj.data(),
                                                   reinterpret_cast<const std::complex<double>*>(Ax_host));
    return seed_from_coo<double>(S, 0, nz, S->Ai.data(), S->Aj.data(), Ax_host);
}

int klujax_cuda_symbolic_info(uint64_t symbolic, int is_complex, int64_t* o) {
    Symbolic* S = sym_of(symbolic);
    if (!S || !o) return fail(KLU_INVALID, "symbolic_info: null argument");
    const Seeded& sd = S->seeded[is_complex ? 1 : 0];
    o[0] = S->ord.n;
    o[1] = S->ord.nblocks;
    o[2] = S->ord.maxblock;
    o[3] = S->ord.nzoff;
    o[4] = S->ord.structural_rank;
    o[5] = sd.ok ? sd.lay.nnz_l : -1;
    o[6] = sd.ok ? sd.lay.nnz_u : -1;
    o[7] = sd.ok ? 1 : 0;
    o[8] = sd.ok ? sd.regime : -1;
    o[9] = sd.ok ? sd.info_levels : -1;
    o[10] = sd.ok ? sd.info_mac : -1;
    o[11] = sd.ok ? sd.piv.noffdiag : -1;
    o[12] = sd.ok ? sd.info_nbatch : -1;
    o[13] = sd.ok ? sd.info_stream_words : -1;
    return 0;
}

int klujax_cuda_symbolic_perm(uint64_t symbolic, int32_t* P, int32_t* Q, int32_t* R) {
    Symbolic* S = sym_of(symbolic);
    if (!S) return fail(KLU_INVALID, "symbolic_perm: null handle");
    if (P) std::copy(S->ord.P.begin(), S->ord.P.end(), P);
    if (Q) std::copy(S->ord.Q.begin(), S->ord.Q.end(), Q);
    if (R) std::copy(S->ord.R.begin(), S->ord.R.begin() + S->ord.nblocks + 1, R);
    return 0;
}

int klujax_cuda_symbolic_pattern(uint64_t symbolic, int is_complex, int32_t* Pnum, int32_t* Lp,
                                 int32_t* Li, int32_t* Up, int32_t* Ui) {
    Symbolic* S = sym_of(symbolic);
    if (!S) return fail(KLU_INVALID, "symbolic_pattern: null handle");
    const Seeded& sd = S->seeded[is_complex ? 1 : 0];
    if (!sd.ok) return fail(KLU_INVALID, "symbolic_pattern: not seeded");
    const int n = sd.lay.n;
    if (Pnum) std::copy(sd.piv.Pnum.begin(), sd.piv.Pnum.end(), Pnum);
    std::vector<int> k1(n);
    for (int b = 0; b < S->ord.nblocks; ++b)
        for (int k = S->ord.R[b]; k < S->ord.R[b + 1]; ++k) k1[k] = S->ord.R[b];
    int lp = 0, up = 0;
    for (int k = 0; k < n; ++k) {
        if (Up) Up[k] = up;
        if (Lp) Lp[k] = lp;
        for (int s = sd.lay.col_begin[k]; s < sd.lay.diag_slot[k]; ++s, ++up)
            if (Ui) Ui[up] = sd.lay.slot_row[s] - k1[k];
        for (int s = sd.lay.diag_slot[k] + 1; s < sd.lay.col_begin[k + 1]; ++s, ++lp)
            if (Li) Li[lp] = sd.lay.slot_row[s] - k1[k];
    }
    if (Up) Up[n] = up;
    if (Lp) Lp[n] = lp;
    return 0;
}

int klujax_cuda_rowreg_selftest(uint64_t symbolic, int is_complex, int transpose, int n_rhs,
                                const double* Ax_host, const double* b_host, double* x_host,
                                int32_t* status_out, double* growth_out, int64_t* stats16,
                                const char* ptx_path) {
    Symbolic* S = sym_of(symbolic);
    if (!S) return fail(KLU_INVALID, "rowreg_selftest: null handle");
    const int cx = is_complex ? 1 : 0;
    const Seeded& sd = S->seeded[cx];
    if (!sd.ok) return fail(KLU_INVALID, "rowreg_selftest: not seeded");
    rr::Options opt;
    opt.r = n_rhs;
    opt.transpose = transpose != 0;
    if (const char* e = std::getenv("KLUJAX_CUDA_RR_PINNED")) opt.pinned_rows = std::atoi(e);
    if (const char* e = std::getenv("KLUJAX_CUDA_RR_MAXREGS")) opt.max_regs = std::atoi(e);
    if (const char* e = std::getenv("KLUJAX_CUDA_RR_REBALANCE")) opt.rebalance = std::atoi(e) != 0;
    if (const char* e = std::getenv("KLUJAX_CUDA_RR_LEVEL_ORDER")) opt.level_order = std::atoi(e) != 0;
    if (const char* e = std::getenv("KLUJAX_CUDA_RR_STRIPE")) opt.stripe = std::atoi(e) != 0;
    if (const char* e = std::getenv("KLUJAX_CUDA_RR_STRIPE_MIN")) opt.stripe_min = std::atoi(e);
    rr::Plan plan;
    const bool ok = rr::build_plan(S->A, S->ord, sd.piv, sd.lay, cx != 0, opt, plan);
    rr::PtxKernel pk;
    if (ok) rr::emit_ptx(plan, pk);
    if (stats16) {
        const rr::Stats& st = plan.st;
        const int64_t v[16] = {st.P, st.mac_ins, st.useful_mac, st.bcast, st.pivots, st.lmul, st.zero, st.dump,
                               st.mac_s, st.planes, st.epochs, st.pinned, st.loadv,
                               static_cast<int64_t>(plan.ins.size()), static_cast<int64_t>(plan.tabs.size() / 32),
                               static_cast<int64_t>(pk.smem_per_warp)};
        std::copy(v, v + 16, stats16);
    }
    if (!ok) return fail(KLU_TOO_LARGE, "rowreg plan: " + plan.why);
    if (ptx_path) {
        FILE* f = std::fopen(ptx_path, "w");
        if (!f) return fail(KLU_INVALID, "rowreg_selftest: cannot write the PTX file");
        std::fwrite(pk.ptx.data(), 1, pk.ptx.size(), f);
        std::fclose(f);
    }
    if (Ax_host && b_host && x_host) {
        const int nz = static_cast<int>(S->A.rowidx.size()), n = sd.lay.n;
        // analyze-call COO order -> analysed CSC order
        int st = 0;
        double g = 0.0;
        if (cx) {
            using Z = std::complex<double>;
            const Z* ax = reinterpret_cast<const Z*>(Ax_host);
            std::vector<Z> csc(nz);
            for (int p = 0; p < nz; ++p) csc[p] = ax[S->A.coo_of[p]];
            st = rr::emulate<Z>(plan, csc.data(), reinterpret_cast<const Z*>(b_host), reinterpret_cast<Z*>(x_host), &g);
        } else {
            std::vector<double> csc(nz);
            for (int p = 0; p < nz; ++p) csc[p] = Ax_host[S->A.coo_of[p]];
            st = rr::emulate<double>(plan, csc.data(), b_host, x_host, &g);
        }
        (void)n;
        if (status_out) *status_out = st;
        if (growth_out) *growth_out = g;
    }
    return 0;
}

// ---- fused factor + (t)solve ------------------------------------------------
static int do_solve_with_symbol_strided(int cx, int transpose, uint64_t symbolic, int n_lhs, int n_col, int n_rhs,
                                        int n_nz, const int32_t* Ai, const int32_t* Aj, const double* Ax,
                                        long long ax_stride, const double* b, long long b_stride, double* x,
                                        int32_t* status, double* growth, void* stream);
static int do_solve_with_symbol(int cx, int transpose, uint64_t symbolic, int n_lhs, int n_col,
                                int n_rhs, int n_nz, const int32_t* Ai, const int32_t* Aj,
                                const double* Ax, const double* b, double* x, int32_t* status,
                                double* growth, void* stream) {
    return do_solve_with_symbol_strided(cx, transpose, symbolic, n_lhs, n_col, n_rhs, n_nz, Ai, Aj, Ax, n_nz, b,
                                        static_cast<long long>(n_col) * n_rhs, x, status, growth, stream);
}

// ax_stride / b_stride in elements; 0 = one matrix / one right-hand side for the whole batch
static int do_solve_with_symbol_strided(int cx, int transpose, uint64_t symbolic, int n_lhs, int n_col, int n_rhs,
                                        int n_nz, const int32_t* Ai, const int32_t* Aj, const double* Ax,
                                        long long ax_stride, const double* b, long long b_stride, double* x,
                                        int32_t* status, double* growth, void* stream) {
    Symbolic* S = sym_of(symbolic);
    if (!S) return fail(KLU_INVALID, "symbolic pointer is null");
    if (n_col != S->ord.n) return fail(KLU_INVALID, "n_col differs from the analysed matrix");
    if (n_lhs <= 0 || n_rhs <= 0) return 0;
    Op op;
    op.mode = transpose ? PM_FACTOR_TSOLVE : PM_FACTOR_SOLVE;
    op.L = n_lhs;
    op.r = n_rhs;
    op.nz = n_nz;
    op.Ai = Ai;
    op.Aj = Aj;
    op.Ax = Ax;
    op.ax_stride = ax_stride;
    op.b = b;
    op.b_stride = b_stride;
    op.x = x;
    op.status = status;
    op.growth = growth;
    op.st = static_cast<cudaStream_t>(stream);
    return run_numeric(S, cx, op);
}

// ---- pivot parity, second line of defence: re-seed below the C-ABI ---------------------------
// The batch runs with ONE static pivot sequence (seeded from the first system the symbolic handle
// ever factorized); status 2 (KLUJAX_PIVOT_DRIFT) marks the systems for which KLU's own threshold
// pivoting (klu_factor per system, klujax.cpp:310/441) would have chosen other pivots, status 1 a
// pivot that is exactly zero with the static sequence.  Those systems are solved again, on the
// GPU, with a pivot sequence seeded from one of THEM (a temporary symbolic object), round after
// round.  This synchronises the stream - it is what `check` semantics cost anyway.
}  // extern "C" (templates follow)
namespace {
template <typename V>
__global__ void gather_rows_kernel(const V* __restrict__ src, long long src_stride, const int* __restrict__ idx,
                                   int rows, long long cols, V* __restrict__ dst) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= rows * cols) return;
    const long long rw = t / cols, c = t - rw * cols;
    dst[t] = src[(long long)idx[rw] * src_stride + c];
}
template <typename V>
__global__ void scatter_rows_kernel(const V* __restrict__ src, const int* __restrict__ idx, int rows, long long cols,
                                    V* __restrict__ dst) {
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= rows * cols) return;
    const long long rw = t / cols, c = t - rw * cols;
    dst[(long long)idx[rw] * cols + c] = src[t];
}
constexpr int kReseedRounds = 8;
}  // namespace

static int do_solve_with_symbol_strided(int cx, int transpose, uint64_t symbolic, int n_lhs, int n_col, int n_rhs,
                                        int n_nz, const int32_t* Ai, const int32_t* Aj, const double* Ax,
                                        long long ax_stride, const double* b, long long b_stride, double* x,
                                        int32_t* status, double* growth, void* stream);

// returns 0, or the status of the first system that still fails; *first_bad = its index
template <typename V>
static int reseed_failed(int cx, int transpose, int L, int n, int r, int nz, const int32_t* Ai_dev,
                         const int32_t* Aj_dev, const V* Ax, long long ax_stride, const V* b, long long b_stride,
                         V* x, int32_t* status, double* growth, cudaStream_t st, int* n_redone, int* first_bad) {
    if (n_redone) *n_redone = 0;
    if (first_bad) *first_bad = -1;
    std::vector<int32_t> hst(L);
    CUDA_OK(cudaMemcpyAsync(hst.data(), status, sizeof(int32_t) * L, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    std::vector<int> sub;
    for (int m = 0; m < L; ++m)
        if (hst[m] == KLUJAX_PIVOT_DRIFT || hst[m] == KLU_SINGULAR) sub.push_back(m);
    if (!sub.empty()) {
        std::vector<int> Ai(nz), Aj(nz);
        CUDA_OK(cudaMemcpyAsync(Ai.data(), Ai_dev, sizeof(int) * nz, cudaMemcpyDeviceToHost, st));
        CUDA_OK(cudaMemcpyAsync(Aj.data(), Aj_dev, sizeof(int) * nz, cudaMemcpyDeviceToHost, st));
        CUDA_OK(cudaStreamSynchronize(st));
        const long long xcols = static_cast<long long>(n) * r;
        for (int round = 0; round < kReseedRounds && !sub.empty(); ++round) {
            const int Ls = static_cast<int>(sub.size());
            DevBuf d_idx, d_ax, d_b, d_x, d_st, d_gr;
            CUDA_OK(d_idx.upload(sub));
            // a broadcast matrix (ax_stride 0) has ONE certificate: the whole call is redone
            const bool ax_b = ax_stride == 0, b_b = b_stride == 0;
            if (!ax_b) CUDA_OK(d_ax.alloc(sizeof(V) * static_cast<size_t>(Ls) * nz));
            if (!b_b) CUDA_OK(d_b.alloc(sizeof(V) * static_cast<size_t>(Ls) * xcols));
            CUDA_OK(d_x.alloc(sizeof(V) * static_cast<size_t>(Ls) * xcols));
            CUDA_OK(d_st.alloc(sizeof(int32_t) * Ls));
            CUDA_OK(d_gr.alloc(sizeof(double) * Ls));
            const int T = 256;
            if (!ax_b)
                gather_rows_kernel<V><<<static_cast<unsigned>((static_cast<long long>(Ls) * nz + T - 1) / T), T, 0, st>>>(
                    Ax, ax_stride, d_idx.as<int>(), Ls, nz, d_ax.as<V>());
            if (!b_b)
                gather_rows_kernel<V><<<static_cast<unsigned>((static_cast<long long>(Ls) * xcols + T - 1) / T), T, 0, st>>>(
                    b, b_stride, d_idx.as<int>(), Ls, xcols, d_b.as<V>());
            uint64_t h = 0;
            int rc = klujax_cuda_analyze(n, nz, Ai.data(), Aj.data(), &h);
            if (rc) return rc;
            rc = do_solve_with_symbol_strided(cx, transpose, h, Ls, n, r, nz, Ai_dev, Aj_dev,
                                              reinterpret_cast<const double*>(ax_b ? Ax : d_ax.as<V>()), ax_b ? 0 : nz,
                                              reinterpret_cast<const double*>(b_b ? b : d_b.as<V>()), b_b ? 0 : xcols,
                                              reinterpret_cast<double*>(d_x.as<V>()), d_st.as<int32_t>(),
                                              d_gr.as<double>(), st);
            std::vector<int32_t> st_s(Ls, 0);
            if (!rc) {
                scatter_rows_kernel<V><<<static_cast<unsigned>((static_cast<long long>(Ls) * xcols + T - 1) / T), T, 0, st>>>(
                    d_x.as<V>(), d_idx.as<int>(), Ls, xcols, x);
                scatter_rows_kernel<int32_t><<<(Ls + T - 1) / T, T, 0, st>>>(d_st.as<int32_t>(), d_idx.as<int>(), Ls, 1, status);
                if (growth)
                    scatter_rows_kernel<double><<<(Ls + T - 1) / T, T, 0, st>>>(d_gr.as<double>(), d_idx.as<int>(), Ls, 1, growth);
                cudaMemcpyAsync(st_s.data(), d_st.p, sizeof(int32_t) * Ls, cudaMemcpyDeviceToHost, st);
            }
            cudaStreamSynchronize(st);
            int32_t dummy = 0;
            klujax_cuda_free_symbolic(h, &dummy);
            if (rc == KLU_SINGULAR) {
                // the seed system of this round is singular under KLU's own pivoting: final
                hst[sub[0]] = KLU_SINGULAR;
                sub.erase(sub.begin());
                continue;
            }
            if (rc) return rc;
            if (n_redone) *n_redone += Ls;
            std::vector<int> again;
            for (int q = 0; q < Ls; ++q) {
                hst[sub[q]] = st_s[q];
                // the seed system (q == 0) carries its own pivot sequence: its verdict is final
                if (q > 0 && (st_s[q] == KLUJAX_PIVOT_DRIFT || st_s[q] == KLU_SINGULAR)) again.push_back(sub[q]);
            }
            sub.swap(again);
        }
    }
    for (int m = 0; m < L; ++m)
        if (hst[m] != 0) {
            if (first_bad) *first_bad = m;
            return hst[m];
        }
    return 0;
}
extern "C" {

#define DEF_SWS(name, cx, tr)                                                                       \
    int name(uint64_t symbolic, int n_lhs, int n_col, int n_rhs, int n_nz, const int32_t* Ai,       \
             const int32_t* Aj, const double* Ax, const double* b, double* x, int32_t* status,      \
             double* growth, void* stream) {                                                        \
        return do_solve_with_symbol(cx, tr, symbolic, n_lhs, n_col, n_rhs, n_nz, Ai, Aj, Ax, b, x,  \
                                    status, growth, stream);                                        \
    }
DEF_SWS(klujax_cuda_solve_with_symbol_f64, 0, 0)
DEF_SWS(klujax_cuda_solve_with_symbol_c128, 1, 0)
DEF_SWS(klujax_cuda_tsolve_with_symbol_f64, 0, 1)
DEF_SWS(klujax_cuda_tsolve_with_symbol_c128, 1, 1)

// The entry an XLA-FFI shim binds: broadcasting of a single matrix or a single
// right-hand side without materialising it (klujax.py:1844-1848), and - with check != 0 - the
// reference's error behaviour: synchronise, re-seed the systems whose pivot certificate failed,
// return the status of the first system that still fails (klujax.cpp:311-314 aborts the batch).
static int do_reseed_failed(int cx, int transpose, int n_lhs, int n_lhs_A, int n_lhs_b, int n_col, int n_rhs, int n_nz,
                            const int32_t* Ai, const int32_t* Aj, const double* Ax, const double* b, double* x,
                            int32_t* status, double* growth, int32_t* first_bad, int32_t* n_redone, void* stream);
static int do_solve_with_symbol_ex(int cx, int transpose, uint64_t symbolic, int n_lhs, int n_lhs_A, int n_lhs_b,
                                   int n_col, int n_rhs, int n_nz, const int32_t* Ai, const int32_t* Aj,
                                   const double* Ax, const double* b, double* x, int32_t* status, double* growth,
                                   int check, int32_t* first_bad, int32_t* n_redone, void* stream) {
    if (first_bad) *first_bad = -1;
    if (n_redone) *n_redone = 0;
    if ((n_lhs_A != n_lhs && n_lhs_A != 1) || (n_lhs_b != n_lhs && n_lhs_b != 1))
        return fail(KLU_INVALID, "n_lhs_A / n_lhs_b must be n_lhs or 1");
    const long long axs = (n_lhs_A == 1 && n_lhs > 1) ? 0 : n_nz;
    const long long bs = (n_lhs_b == 1 && n_lhs > 1) ? 0 : static_cast<long long>(n_col) * n_rhs;
    if (check && (!status || !growth)) return fail(KLU_INVALID, "check needs the status and growth arrays");
    int rc = do_solve_with_symbol_strided(cx, transpose, symbolic, n_lhs, n_col, n_rhs, n_nz, Ai, Aj, Ax, axs, b, bs,
                                          x, status, growth, stream);
    if (rc || !check || n_lhs <= 0 || n_rhs <= 0) return rc;
    return do_reseed_failed(cx, transpose, n_lhs, n_lhs_A, n_lhs_b, n_col, n_rhs, n_nz, Ai, Aj, Ax, b, x, status, growth,
                            first_bad, n_redone, stream);
}
#define DEF_SWS_EX(name, cx, tr)                                                                              \
    int name(uint64_t symbolic, int n_lhs, int n_lhs_A, int n_lhs_b, int n_col, int n_rhs, int n_nz,          \
             const int32_t* Ai, const int32_t* Aj, const double* Ax, const double* b, double* x,               \
             int32_t* status, double* growth, int check, int32_t* first_bad, int32_t* n_redone, void* stream) { \
        return do_solve_with_symbol_ex(cx, tr, symbolic, n_lhs, n_lhs_A, n_lhs_b, n_col, n_rhs, n_nz, Ai, Aj, Ax, b, \
                                       x, status, growth, check, first_bad, n_redone, stream);                 \
    }
// check + re-seed of a batch that has been enqueued already (host batches streamed in chunks)
static int do_reseed_failed(int cx, int transpose, int n_lhs, int n_lhs_A, int n_lhs_b, int n_col, int n_rhs, int n_nz,
                            const int32_t* Ai, const int32_t* Aj, const double* Ax, const double* b, double* x,
                            int32_t* status, double* growth, int32_t* first_bad, int32_t* n_redone, void* stream) {
    if (first_bad) *first_bad = -1;
    if (n_redone) *n_redone = 0;
    if (n_lhs <= 0 || n_rhs <= 0) return 0;
    if (!status) return fail(KLU_INVALID, "reseed_failed needs the status array");
    const long long axs = (n_lhs_A == 1 && n_lhs > 1) ? 0 : n_nz;
    const long long bs = (n_lhs_b == 1 && n_lhs > 1) ? 0 : static_cast<long long>(n_col) * n_rhs;
    int fb = -1, nr = 0, rc;
    if (cx)
        rc = reseed_failed<double2>(1, transpose, n_lhs, n_col, n_rhs, n_nz, Ai, Aj, reinterpret_cast<const double2*>(Ax),
                                    axs, reinterpret_cast<const double2*>(b), bs, reinterpret_cast<double2*>(x), status,
                                    growth, static_cast<cudaStream_t>(stream), &nr, &fb);
    else
        rc = reseed_failed<double>(0, transpose, n_lhs, n_col, n_rhs, n_nz, Ai, Aj, Ax, axs, b, bs, x, status, growth,
                                   static_cast<cudaStream_t>(stream), &nr, &fb);
    if (first_bad) *first_bad = fb;
    if (n_redone) *n_redone = nr;
    if (rc == KLUJAX_PIVOT_DRIFT) return fail(rc, "klu_factor: the pivot sequence could not be certified");
    if (rc == KLU_INVALID && fb >= 0) return fail(rc, "entry outside the analysed sparsity pattern or index out of range");
    if (rc > 0) return fail(rc, "klu_factor failed (singular matrix?)");
    return rc;
}
int klujax_cuda_reseed_failed_f64(int transpose, int n_lhs, int n_lhs_A, int n_lhs_b, int n_col, int n_rhs, int n_nz,
                                  const int32_t* Ai, const int32_t* Aj, const double* Ax, const double* b, double* x,
                                  int32_t* status, double* growth, int32_t* first_bad, int32_t* n_redone, void* stream) {
    return do_reseed_failed(0, transpose, n_lhs, n_lhs_A, n_lhs_b, n_col, n_rhs, n_nz, Ai, Aj, Ax, b, x, status, growth,
                            first_bad, n_redone, stream);
}
int klujax_cuda_reseed_failed_c128(int transpose, int n_lhs, int n_lhs_A, int n_lhs_b, int n_col, int n_rhs, int n_nz,
                                   const int32_t* Ai, const int32_t* Aj, const double* Ax, const double* b, double* x,
                                   int32_t* status, double* growth, int32_t* first_bad, int32_t* n_redone, void* stream) {
    return do_reseed_failed(1, transpose, n_lhs, n_lhs_A, n_lhs_b, n_col, n_rhs, n_nz, Ai, Aj, Ax, b, x, status, growth,
                            first_bad, n_redone, stream);
}
DEF_SWS_EX(klujax_cuda_solve_with_symbol_ex_f64, 0, 0)
DEF_SWS_EX(klujax_cuda_solve_with_symbol_ex_c128, 1, 0)
DEF_SWS_EX(klujax_cuda_tsolve_with_symbol_ex_f64, 0, 1)
DEF_SWS_EX(klujax_cuda_tsolve_with_symbol_ex_c128, 1, 1)

// One-shot `solve` (klujax.cpp:251-380: analyze + factor + solve + free, every call).  The
// analysis, the pivoting seed factorization and the schedule only depend on the pattern, so the
// last few patterns keep their symbolic object (with the device copies of (Ai, Aj) and of the
// schedule): a caller looping over `solve` with one pattern pays them once.  Eviction frees the
// object like klujax.cpp:336 does after every call.
namespace {
struct SolveCacheEntry {
    int n = 0, device = 0;  // the device copies of (Ai, Aj) and of the schedule live on ONE device
    std::vector<int32_t> Ai, Aj;
    uint64_t sym = 0;
    DevBuf dAi, dAj;
    uint64_t stamp = 0;
    int users = 0;  // calls in flight outside the mutex: such an entry is not evicted
};
std::mutex g_solve_cache_mu;
std::vector<std::unique_ptr<SolveCacheEntry>> g_solve_cache;
uint64_t g_solve_stamp = 0;
constexpr size_t kSolveCacheEntries = 4;
}  // namespace

static int do_solve(int cx, int n_lhs, int n_col, int n_rhs, int n_nz, const int32_t* Ai_host,
                    const int32_t* Aj_host, const double* Ax, const double* b, double* x,
                    int32_t* status, double* growth, void* stream) {
    if (n_nz < 0 || (n_nz > 0 && (!Ai_host || !Aj_host))) return fail(KLU_INVALID, "solve: null indices");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int cur_dev = 0;
    cudaGetDevice(&cur_dev);
    // the mutex covers the lookup / insertion / eviction only; the numeric call runs outside it, so
    // solve() calls on different patterns or devices do not serialise each other
    std::unique_lock<std::mutex> lock(g_solve_cache_mu);
    SolveCacheEntry* hit = nullptr;
    for (auto& e : g_solve_cache)
        if (e->n == n_col && e->device == cur_dev && static_cast<int>(e->Ai.size()) == n_nz &&
            std::equal(e->Ai.begin(), e->Ai.end(), Ai_host) && std::equal(e->Aj.begin(), e->Aj.end(), Aj_host)) {
            hit = e.get();
            break;
        }
    if (!hit) {
        auto ne = std::make_unique<SolveCacheEntry>();
        int rc = klujax_cuda_analyze(n_col, n_nz, Ai_host, Aj_host, &ne->sym);
        if (rc) return rc;
        ne->n = n_col;
        ne->device = cur_dev;
        ne->Ai.assign(Ai_host, Ai_host + n_nz);
        ne->Aj.assign(Aj_host, Aj_host + n_nz);
        cudaError_t e = ne->dAi.alloc(sizeof(int) * (n_nz + 1));
        if (e == cudaSuccess) e = ne->dAj.alloc(sizeof(int) * (n_nz + 1));
        if (e == cudaSuccess) e = cudaMemcpy(ne->dAi.p, Ai_host, sizeof(int) * n_nz, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(ne->dAj.p, Aj_host, sizeof(int) * n_nz, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            klujax_cuda_free_symbolic(ne->sym, nullptr);
            return fail(KLU_OUT_OF_MEMORY, cudaGetErrorString(e));
        }
        if (g_solve_cache.size() >= kSolveCacheEntries) {  // evict the least recently used pattern
            size_t victim = 0;
            for (size_t i = 1; i < g_solve_cache.size(); ++i)
                if (g_solve_cache[i]->users == 0 &&
                    (g_solve_cache[victim]->users > 0 || g_solve_cache[i]->stamp < g_solve_cache[victim]->stamp))
                    victim = i;
            if (g_solve_cache[victim]->users == 0) {
                cudaDeviceSynchronize();  // its kernels may still be in flight on another stream
                klujax_cuda_free_symbolic(g_solve_cache[victim]->sym, nullptr);
                g_solve_cache.erase(g_solve_cache.begin() + victim);
            }
        }
        g_solve_cache.push_back(std::move(ne));
        hit = g_solve_cache.back().get();
    }
    hit->stamp = ++g_solve_stamp;
    hit->users++;
    const uint64_t hsym = hit->sym;
    const int32_t *dAi = hit->dAi.as<int32_t>(), *dAj = hit->dAj.as<int32_t>();
    lock.unlock();
    int rc = do_solve_with_symbol(cx, 0, hsym, n_lhs, n_col, n_rhs, n_nz, dAi, dAj, Ax, b, x, status, growth, stream);
    lock.lock();
    hit->users--;
    if (rc && hit->users == 0) {  // a failed call (singular seed, ...) leaves nothing behind, like the reference
        cudaStreamSynchronize(st);
        for (size_t i = 0; i < g_solve_cache.size(); ++i)
            if (g_solve_cache[i].get() == hit) {
                klujax_cuda_free_symbolic(hit->sym, nullptr);
                g_solve_cache.erase(g_solve_cache.begin() + i);
                break;
            }
    }
    return rc;
}
int klujax_cuda_solve_f64(int n_lhs, int n_col, int n_rhs, int n_nz, const int32_t* Ai,
                          const int32_t* Aj, const double* Ax, const double* b, double* x,
                          int32_t* status, double* growth, void* stream) {
    return do_solve(0, n_lhs, n_col, n_rhs, n_nz, Ai, Aj, Ax, b, x, status, growth, stream);
}
int klujax_cuda_solve_c128(int n_lhs, int n_col, int n_rhs, int n_nz, const int32_t* Ai,
                           const int32_t* Aj, const double* Ax, const double* b, double* x,
                           int32_t* status, double* growth, void* stream) {
    return do_solve(1, n_lhs, n_col, n_rhs, n_nz, Ai, Aj, Ax, b, x, status, growth, stream);
}

// ---- factor / refactor / refactor_and_solve ---------------------------------
}  // extern "C" (templates follow)
// ---- numeric handles: factor / refactor / refactor_and_solve / solves ---------------------------
static uint64_t handle_of(const NumericBatch* nb, int el) {
    return (static_cast<uint64_t>(nb->id) << 32) | static_cast<uint64_t>(el + 1);
}
// drops one element of a batch (the batch goes when its last element goes)
static void drop_element(NumericBatch* nb, int el) {
    bool dead = false;
    {
        std::lock_guard<std::mutex> g(g_reg_mu);
        if (el < 0 || el >= nb->L || nb->freed[el]) return;
        nb->freed[el] = 1;
        if (--nb->live == 0) {
            g_batches.erase(nb->id);
            dead = true;
        }
    }
    if (dead) {
        cudaDeviceSynchronize();  // storage may still be read by enqueued work
        delete nb;
    }
}

static int factor_into(Symbolic* S, int cx, NumericBatch* nb, int L, int nz, const int32_t* Ai, const int32_t* Aj,
                       const void* Ax, long long ax_stride, int32_t* status, double* growth, cudaStream_t st) {
    Op op;
    op.mode = PM_FACTOR;
    op.L = L;
    op.nz = nz;
    op.Ai = Ai;
    op.Aj = Aj;
    op.Ax = Ax;
    op.ax_stride = ax_stride;
    op.batch = nb;
    op.store_numeric = true;
    op.status = status;
    op.growth = growth;
    op.st = st;
    return run_numeric(S, cx, op);
}

// factor_* (klujax.cpp:638-731).  check = 0: enqueue only; every system gets the pivot sequence of
// the symbolic handle, the certificate is left in status / growth.  check != 0: the reference's
// semantics - klu_factor pivots every system (klujax.cpp:679) - are restored: systems whose
// certificate failed are factorized again in a child batch with a pivot sequence seeded from one
// of THEM (round after round); the returned handles then point into several batches.  A system
// that is singular under its own pivoting fails the call; like the reference (klujax.cpp:684-687)
// everything factorized so far is freed.
template <typename V>
static int do_factor_t(int cx, uint64_t symbolic, int n_lhs, int n_nz, const int32_t* Ai, const int32_t* Aj,
                       const V* Ax, uint64_t* numeric_out, int32_t* status, double* growth, int check,
                       int32_t* first_bad, void* stream) {
    if (first_bad) *first_bad = -1;
    Symbolic* S = sym_of(symbolic);
    if (!S) return fail(KLU_INVALID, "symbolic pointer is null");
    if (n_lhs <= 0 || !numeric_out) return fail(KLU_INVALID, "factor: empty batch");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    {
        std::lock_guard<std::mutex> lock(S->mu);
        int e = ensure_seeded_dev(S, cx, n_nz, Ai, Aj, Ax, st);
        if (e) return e;
    }
    StreamBuf tmp_st, tmp_gr;
    if (check && !status) {
        CUDA_OK(tmp_st.alloc(sizeof(int32_t) * n_lhs, st));
        status = tmp_st.as<int32_t>();
    }
    NumericBatch* nb = nullptr;
    int e = new_batch(S, cx, n_lhs, &nb);
    if (e) return e;
    e = factor_into(S, cx, nb, n_lhs, n_nz, Ai, Aj, Ax, n_nz, status, growth, st);
    if (e) {
        for (int m = 0; m < n_lhs; ++m) drop_element(nb, m);
        return e;
    }
    for (int m = 0; m < n_lhs; ++m) numeric_out[m] = handle_of(nb, m);
    if (!check) return 0;

    std::vector<int32_t> hst(n_lhs);
    CUDA_OK(cudaMemcpyAsync(hst.data(), status, sizeof(int32_t) * n_lhs, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    std::vector<int> sub;
    for (int m = 0; m < n_lhs; ++m)
        if (hst[m] == KLUJAX_PIVOT_DRIFT || hst[m] == KLU_SINGULAR) sub.push_back(m);
    int rc_fatal = 0;
    if (!sub.empty()) {
        std::vector<int> hAi(n_nz), hAj(n_nz);
        CUDA_OK(cudaMemcpyAsync(hAi.data(), Ai, sizeof(int) * n_nz, cudaMemcpyDeviceToHost, st));
        CUDA_OK(cudaMemcpyAsync(hAj.data(), Aj, sizeof(int) * n_nz, cudaMemcpyDeviceToHost, st));
        CUDA_OK(cudaStreamSynchronize(st));
        for (int round = 0; round < kReseedRounds && !sub.empty() && !rc_fatal; ++round) {
            const int Ls = static_cast<int>(sub.size());
            uint64_t h = 0;
            int rc = klujax_cuda_analyze(S->ord.n, n_nz, hAi.data(), hAj.data(), &h);
            if (rc) {
                rc_fatal = rc;
                break;
            }
            std::unique_ptr<Symbolic> child(sym_of(h));
            StreamBuf d_idx, d_ax, d_st, d_gr;
            CUDA_OK(d_idx.upload(sub, st));
            CUDA_OK(d_ax.alloc(sizeof(V) * static_cast<size_t>(Ls) * n_nz, st));
            CUDA_OK(d_st.alloc(sizeof(int32_t) * Ls, st));
            CUDA_OK(d_gr.alloc(sizeof(double) * Ls, st));
            const int T = 256;
            gather_rows_kernel<V><<<static_cast<unsigned>((static_cast<long long>(Ls) * n_nz + T - 1) / T), T, 0, st>>>(
                Ax, n_nz, d_idx.as<int>(), Ls, n_nz, d_ax.as<V>());
            {
                std::lock_guard<std::mutex> lock(child->mu);
                rc = ensure_seeded_dev(child.get(), cx, n_nz, Ai, Aj, d_ax.p, st);
            }
            if (rc == KLU_SINGULAR) {  // the seed of this round is singular under KLU's own pivoting: final
                hst[sub[0]] = KLU_SINGULAR;
                sub.erase(sub.begin());
                continue;
            }
            if (rc) {
                rc_fatal = rc;
                break;
            }
            NumericBatch* cb = nullptr;
            rc = new_batch(child.get(), cx, Ls, &cb);
            if (rc) {
                rc_fatal = rc;
                break;
            }
            cb->parent = S;
            Symbolic* C = child.get();
            cb->owned = std::move(child);
            rc = factor_into(C, cx, cb, Ls, n_nz, Ai, Aj, d_ax.p, n_nz, d_st.as<int32_t>(), d_gr.as<double>(), st);
            std::vector<int32_t> st_s(Ls, 0);
            if (!rc) {
                scatter_rows_kernel<int32_t><<<(Ls + T - 1) / T, T, 0, st>>>(d_st.as<int32_t>(), d_idx.as<int>(), Ls, 1, status);
                if (growth)
                    scatter_rows_kernel<double><<<(Ls + T - 1) / T, T, 0, st>>>(d_gr.as<double>(), d_idx.as<int>(), Ls, 1, growth);
                cudaMemcpyAsync(st_s.data(), d_st.p, sizeof(int32_t) * Ls, cudaMemcpyDeviceToHost, st);
            }
            cudaStreamSynchronize(st);
            if (rc) {
                for (int q = 0; q < Ls; ++q) drop_element(cb, q);
                rc_fatal = rc;
                break;
            }
            std::vector<int> again;
            for (int q = 0; q < Ls; ++q) {
                const int m = sub[q];
                if (st_s[q] == 0) {
                    drop_element(nb, m);  // the parent batch no longer holds this system
                    numeric_out[m] = handle_of(cb, q);
                    hst[m] = 0;
                } else {
                    if (q == 0 || (st_s[q] != KLUJAX_PIVOT_DRIFT && st_s[q] != KLU_SINGULAR))
                        hst[m] = st_s[q];  // the seed carries its own pivots: its verdict is final
                    else
                        again.push_back(m);
                    // not kept in this child batch
                }
            }
            for (int q = 0; q < Ls; ++q)
                if (st_s[q] != 0) drop_element(cb, q);
            sub.swap(again);
        }
        for (int m : sub) (void)m;  // rounds exhausted: hst keeps the failing status
    }
    int bad = -1;
    for (int m = 0; m < n_lhs && bad < 0; ++m)
        if (hst[m] != 0) bad = m;
    if (rc_fatal || bad >= 0) {
        klujax_cuda_free_numeric(n_lhs, numeric_out, nullptr);
        for (int m = 0; m < n_lhs; ++m) numeric_out[m] = 0;
        if (rc_fatal) return rc_fatal;
        if (first_bad) *first_bad = bad;
        return fail(hst[bad], hst[bad] == KLUJAX_PIVOT_DRIFT ? "klu_factor: the pivot sequence could not be certified"
                              : hst[bad] == KLU_INVALID    ? "entry outside the analysed sparsity pattern or index out of range"
                                                           : "klu_factor failed (singular matrix?)");
    }
    return 0;
}
static int do_factor(int cx, uint64_t symbolic, int n_lhs, int n_nz, const int32_t* Ai, const int32_t* Aj,
                     const double* Ax, uint64_t* numeric_out, int32_t* status, double* growth, int check,
                     int32_t* first_bad, void* stream) {
    if (cx)
        return do_factor_t<double2>(1, symbolic, n_lhs, n_nz, Ai, Aj, reinterpret_cast<const double2*>(Ax), numeric_out,
                                    status, growth, check, first_bad, stream);
    return do_factor_t<double>(0, symbolic, n_lhs, n_nz, Ai, Aj, Ax, numeric_out, status, growth, check, first_bad, stream);
}

// refactor_* / refactor_and_solve_* (klujax.cpp:733-993).  klu_refactor keeps the pivot order of
// the handle and never re-pivots, so there is no certificate here (status 0 or 1); the handles may
// point into several batches (see factor).
template <typename V>
static int do_refactor_t(int cx, int with_solve, uint64_t symbolic, int n_lhs, int n_col, int n_rhs, int n_nz,
                         const int32_t* Ai, const int32_t* Aj, const V* Ax, const V* b, const uint64_t* numeric_in,
                         V* x, uint64_t* numeric_out, int32_t* status, double* growth, void* stream) {
    Symbolic* S = sym_of(symbolic);
    if (!S) return fail(KLU_INVALID, "symbolic pointer is null");
    if (!numeric_in) return fail(KLU_INVALID, "numeric pointer is null");
    std::vector<NumGroup> groups;
    int e = resolve_groups(n_lhs, numeric_in, S, cx, groups);
    if (e) return e;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long xcols = static_cast<long long>(n_col) * n_rhs;
    const int T = 256;
    for (NumGroup& g : groups) {
        const int Lg = static_cast<int>(g.pos.size());
        const bool direct = groups.size() == 1 && g.identity();
        StreamBuf d_pos, d_el, d_ax, d_b, d_x, d_st, d_gr;
        Op op;
        op.mode = with_solve ? PM_FACTOR_SOLVE : PM_FACTOR;
        op.L = Lg;
        op.r = with_solve ? n_rhs : 0;
        op.nz = n_nz;
        op.Ai = Ai;
        op.Aj = Aj;
        op.ax_stride = n_nz;
        op.b_stride = xcols;
        op.batch = g.nb;
        op.store_numeric = true;
        op.certify = false;
        op.st = st;
        if (direct) {
            op.Ax = Ax;
            op.b = b;
            op.x = x;
            op.status = status;
            op.growth = growth;
        } else {
            CUDA_OK(d_pos.upload(g.pos, st));
            CUDA_OK(d_el.upload(g.el, st));
            CUDA_OK(d_ax.alloc(sizeof(V) * static_cast<size_t>(Lg) * n_nz, st));
            CUDA_OK(d_st.alloc(sizeof(int32_t) * Lg, st));
            CUDA_OK(d_gr.alloc(sizeof(double) * Lg, st));
            gather_rows_kernel<V><<<static_cast<unsigned>((static_cast<long long>(Lg) * n_nz + T - 1) / T), T, 0, st>>>(
                Ax, n_nz, d_pos.as<int>(), Lg, n_nz, d_ax.as<V>());
            op.Ax = d_ax.p;
            op.sysidx_dev = d_el.as<int>();
            op.status = d_st.as<int32_t>();
            op.growth = d_gr.as<double>();
            if (with_solve) {
                CUDA_OK(d_b.alloc(sizeof(V) * static_cast<size_t>(Lg) * xcols, st));
                CUDA_OK(d_x.alloc(sizeof(V) * static_cast<size_t>(Lg) * xcols, st));
                gather_rows_kernel<V><<<static_cast<unsigned>((static_cast<long long>(Lg) * xcols + T - 1) / T), T, 0, st>>>(
                    b, xcols, d_pos.as<int>(), Lg, xcols, d_b.as<V>());
                op.b = d_b.p;
                op.x = d_x.p;
            }
        }
        e = run_numeric(g.nb->sym, cx, op);
        if (e) return e;
        if (!direct) {
            if (with_solve)
                scatter_rows_kernel<V><<<static_cast<unsigned>((static_cast<long long>(Lg) * xcols + T - 1) / T), T, 0, st>>>(
                    d_x.as<V>(), d_pos.as<int>(), Lg, xcols, x);
            if (status)
                scatter_rows_kernel<int32_t><<<(Lg + T - 1) / T, T, 0, st>>>(d_st.as<int32_t>(), d_pos.as<int>(), Lg, 1, status);
            if (growth)
                scatter_rows_kernel<double><<<(Lg + T - 1) / T, T, 0, st>>>(d_gr.as<double>(), d_pos.as<int>(), Lg, 1, growth);
        }
    }
    if (numeric_out)
        for (int m = 0; m < n_lhs; ++m) numeric_out[m] = numeric_in[m];  // klujax.cpp:785-786
    return 0;
}
static int do_refactor(int cx, int with_solve, uint64_t symbolic, int n_lhs, int n_col, int n_rhs,
                       int n_nz, const int32_t* Ai, const int32_t* Aj, const double* Ax,
                       const double* b, const uint64_t* numeric_in, double* x,
                       uint64_t* numeric_out, int32_t* status, double* growth, void* stream) {
    if (cx)
        return do_refactor_t<double2>(1, with_solve, symbolic, n_lhs, n_col, n_rhs, n_nz, Ai, Aj,
                                      reinterpret_cast<const double2*>(Ax), reinterpret_cast<const double2*>(b),
                                      numeric_in, reinterpret_cast<double2*>(x), numeric_out, status, growth, stream);
    return do_refactor_t<double>(0, with_solve, symbolic, n_lhs, n_col, n_rhs, n_nz, Ai, Aj, Ax, b, numeric_in, x,
                                 numeric_out, status, growth, stream);
}

// solve_with_numeric_* / tsolve_with_numeric_* (klujax.cpp:995-1271)
template <typename V>
static int do_solve_numeric_t(int cx, int transpose, uint64_t symbolic, int n_numeric, const uint64_t* numeric,
                              int n_lhs_b, int n_col, int n_rhs, const V* b, V* x, void* stream) {
    Symbolic* S = sym_of(symbolic);
    if (!S) return fail(KLU_INVALID, "symbolic pointer is null");
    if (n_col != S->ord.n) return fail(KLU_INVALID, "n_col differs from the analysed matrix");
    std::vector<NumGroup> groups;
    int e = resolve_groups(n_numeric, numeric, S, cx, groups);
    if (e) return e;
    // broadcast rules of klujax.cpp:1044-1059
    const bool bc_num = (n_numeric == 1), bc_b = (n_lhs_b == 1);
    int L;
    if (!bc_num && !bc_b) {
        if (n_numeric != n_lhs_b) return fail(KLU_INVALID, "numeric and b batch size mismatch");
        L = n_numeric;
    } else if (!bc_num) {
        L = n_numeric;
    } else if (!bc_b) {
        L = n_lhs_b;
    } else {
        L = 1;
    }
    if (n_rhs <= 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long xcols = static_cast<long long>(n_col) * n_rhs;
    const int T = 256;
    for (NumGroup& g : groups) {
        int Lg = static_cast<int>(g.pos.size());
        const bool direct = groups.size() == 1 && !bc_num && g.identity();
        StreamBuf d_pos, d_el, d_b, d_x;
        Op op;
        op.mode = transpose ? PM_TSOLVE : PM_SOLVE;
        op.r = n_rhs;
        op.b_stride = bc_b && L > 1 ? 0 : xcols;
        op.batch = g.nb;
        op.st = st;
        if (direct) {
            op.L = L;
            op.b = b;
            op.x = x;
        } else if (groups.size() == 1) {
            // one batch, permuted / partial / broadcast element list: systems stay in place
            std::vector<int> full(L);
            for (int m = 0; m < L; ++m) full[m] = bc_num ? g.el[0] : g.el[m];
            CUDA_OK(d_el.upload(full, st));
            op.L = L;
            op.b = b;
            op.x = x;
            op.sysidx_dev = d_el.as<int>();
        } else {
            CUDA_OK(d_pos.upload(g.pos, st));
            CUDA_OK(d_el.upload(g.el, st));
            CUDA_OK(d_x.alloc(sizeof(V) * static_cast<size_t>(Lg) * xcols, st));
            op.L = Lg;
            op.sysidx_dev = d_el.as<int>();
            op.x = d_x.p;
            if (bc_b) {
                op.b = b;
                op.b_stride = Lg > 1 ? 0 : xcols;
            } else {
                CUDA_OK(d_b.alloc(sizeof(V) * static_cast<size_t>(Lg) * xcols, st));
                gather_rows_kernel<V><<<static_cast<unsigned>((static_cast<long long>(Lg) * xcols + T - 1) / T), T, 0, st>>>(
                    b, xcols, d_pos.as<int>(), Lg, xcols, d_b.as<V>());
                op.b = d_b.p;
                op.b_stride = xcols;
            }
        }
        e = run_numeric(g.nb->sym, cx, op);
        if (e) return e;
        if (groups.size() > 1)
            scatter_rows_kernel<V><<<static_cast<unsigned>((static_cast<long long>(Lg) * xcols + T - 1) / T), T, 0, st>>>(
                d_x.as<V>(), d_pos.as<int>(), Lg, xcols, x);
    }
    return 0;
}
static int do_solve_numeric(int cx, int transpose, uint64_t symbolic, int n_numeric,
                            const uint64_t* numeric, int n_lhs_b, int n_col, int n_rhs,
                            const double* b, double* x, void* stream) {
    if (cx)
        return do_solve_numeric_t<double2>(1, transpose, symbolic, n_numeric, numeric, n_lhs_b, n_col, n_rhs,
                                           reinterpret_cast<const double2*>(b), reinterpret_cast<double2*>(x), stream);
    return do_solve_numeric_t<double>(0, transpose, symbolic, n_numeric, numeric, n_lhs_b, n_col, n_rhs, b, x, stream);
}
extern "C" {
int klujax_cuda_factor_f64(uint64_t symbolic, int n_lhs, int n_nz, const int32_t* Ai,
                           const int32_t* Aj, const double* Ax, uint64_t* numeric_out,
                           int32_t* status, double* growth, void* stream) {
    return do_factor(0, symbolic, n_lhs, n_nz, Ai, Aj, Ax, numeric_out, status, growth, 0, nullptr, stream);
}
int klujax_cuda_factor_c128(uint64_t symbolic, int n_lhs, int n_nz, const int32_t* Ai,
                            const int32_t* Aj, const double* Ax, uint64_t* numeric_out,
                            int32_t* status, double* growth, void* stream) {
    return do_factor(1, symbolic, n_lhs, n_nz, Ai, Aj, Ax, numeric_out, status, growth, 0, nullptr, stream);
}
int klujax_cuda_factor_ex_f64(uint64_t symbolic, int n_lhs, int n_nz, const int32_t* Ai, const int32_t* Aj,
                              const double* Ax, uint64_t* numeric_out, int32_t* status, double* growth, int check,
                              int32_t* first_bad, void* stream) {
    return do_factor(0, symbolic, n_lhs, n_nz, Ai, Aj, Ax, numeric_out, status, growth, check, first_bad, stream);
}
int klujax_cuda_factor_ex_c128(uint64_t symbolic, int n_lhs, int n_nz, const int32_t* Ai, const int32_t* Aj,
                               const double* Ax, uint64_t* numeric_out, int32_t* status, double* growth, int check,
                               int32_t* first_bad, void* stream) {
    return do_factor(1, symbolic, n_lhs, n_nz, Ai, Aj, Ax, numeric_out, status, growth, check, first_bad, stream);
}
int klujax_cuda_refactor_f64(uint64_t symbolic, int n_lhs, int n_nz, const int32_t* Ai,
                             const int32_t* Aj, const double* Ax, const uint64_t* nin,
                             uint64_t* nout, int32_t* status, double* growth, void* stream) {
    return do_refactor(0, 0, symbolic, n_lhs, 0, 0, n_nz, Ai, Aj, Ax, nullptr, nin, nullptr, nout,
                       status, growth, stream);
}
int klujax_cuda_refactor_c128(uint64_t symbolic, int n_lhs, int n_nz, const int32_t* Ai,
                              const int32_t* Aj, const double* Ax, const uint64_t* nin,
                              uint64_t* nout, int32_t* status, double* growth, void* stream) {
    return do_refactor(1, 0, symbolic, n_lhs, 0, 0, n_nz, Ai, Aj, Ax, nullptr, nin, nullptr, nout,
                       status, growth, stream);
}
int klujax_cuda_refactor_and_solve_f64(uint64_t symbolic, int n_lhs, int n_col, int n_rhs, int n_nz,
                                       const int32_t* Ai, const int32_t* Aj, const double* Ax,
                                       const double* b, const uint64_t* nin, double* x,
                                       uint64_t* nout, int32_t* status, double* growth,
                                       void* stream) {
    return do_refactor(0, 1, symbolic, n_lhs, n_col, n_rhs, n_nz, Ai, Aj, Ax, b, nin, x, nout,
                       status, growth, stream);
}
int klujax_cuda_refactor_and_solve_c128(uint64_t symbolic, int n_lhs, int n_col, int n_rhs,
                                        int n_nz, const int32_t* Ai, const int32_t* Aj,
                                        const double* Ax, const double* b, const uint64_t* nin,
                                        double* x, uint64_t* nout, int32_t* status, double* growth,
                                        void* stream) {
    return do_refactor(1, 1, symbolic, n_lhs, n_col, n_rhs, n_nz, Ai, Aj, Ax, b, nin, x, nout,
                       status, growth, stream);
}
#define DEF_SWN(name, cx, tr)                                                                     \
    int name(uint64_t symbolic, int n_numeric, const uint64_t* numeric, int n_lhs_b, int n_col,   \
             int n_rhs, const double* b, double* x, void* stream) {                               \
        return do_solve_numeric(cx, tr, symbolic, n_numeric, numeric, n_lhs_b, n_col, n_rhs, b, x, \
                                stream);                                                          \
    }
DEF_SWN(klujax_cuda_solve_with_numeric_f64, 0, 0)
DEF_SWN(klujax_cuda_solve_with_numeric_c128, 1, 0)
DEF_SWN(klujax_cuda_tsolve_with_numeric_f64, 0, 1)
DEF_SWN(klujax_cuda_tsolve_with_numeric_c128, 1, 1)

// ---- device-resident handle arrays (what XLA delivers on the CUDA platform: u64[n_lhs] buffers,
// klujax.cpp:392-395, :690, :1044) ------------------------------------------------------------------
// Output handles are written to the device stream-ordered (no synchronisation).  Input handles
// have to be known on the host to pick the batch and its schedule: they are copied back (8 bytes
// per system, one stream synchronisation per call - stated in INTEGRATION.md).
static int fetch_handles(const uint64_t* dev, int n, std::vector<uint64_t>& host, cudaStream_t st) {
    host.resize(n > 0 ? n : 0);
    if (n <= 0) return 0;
    CUDA_OK(cudaMemcpyAsync(host.data(), dev, sizeof(uint64_t) * n, cudaMemcpyDeviceToHost, st));
    CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}
static int push_handles(const std::vector<uint64_t>& host, uint64_t* dev, cudaStream_t st) {
    if (!dev || host.empty()) return 0;
    CUDA_OK(cudaMemcpyAsync(dev, host.data(), sizeof(uint64_t) * host.size(), cudaMemcpyHostToDevice, st));
    return 0;
}
#define DEF_FACTOR_DEV(name, cx)                                                                                      \
    int name(uint64_t symbolic, int n_lhs, int n_nz, const int32_t* Ai, const int32_t* Aj, const double* Ax,          \
             uint64_t* numeric_out_dev, int32_t* status, double* growth, int check, int32_t* first_bad, void* stream) { \
        std::vector<uint64_t> h(n_lhs > 0 ? n_lhs : 0, 0);                                                            \
        int rc = do_factor(cx, symbolic, n_lhs, n_nz, Ai, Aj, Ax, h.data(), status, growth, check, first_bad, stream); \
        if (rc) return rc;                                                                                            \
        return push_handles(h, numeric_out_dev, static_cast<cudaStream_t>(stream));                                   \
    }
DEF_FACTOR_DEV(klujax_cuda_factor_dev_f64, 0)
DEF_FACTOR_DEV(klujax_cuda_factor_dev_c128, 1)
#define DEF_REFACTOR_DEV(name, cx, ws)                                                                               \
    int name(uint64_t symbolic, int n_lhs, int n_col, int n_rhs, int n_nz, const int32_t* Ai, const int32_t* Aj,     \
             const double* Ax, const double* b, const uint64_t* numeric_in_dev, double* x, uint64_t* numeric_out_dev, \
             int32_t* status, double* growth, void* stream) {                                                        \
        std::vector<uint64_t> h;                                                                                     \
        int rc = fetch_handles(numeric_in_dev, n_lhs, h, static_cast<cudaStream_t>(stream));                         \
        if (rc) return rc;                                                                                           \
        rc = do_refactor(cx, ws, symbolic, n_lhs, n_col, n_rhs, n_nz, Ai, Aj, Ax, b, h.data(), x, nullptr, status,   \
                         growth, stream);                                                                            \
        if (rc) return rc;                                                                                           \
        if (numeric_out_dev && numeric_out_dev != numeric_in_dev)                                                    \
            CUDA_OK(cudaMemcpyAsync(numeric_out_dev, numeric_in_dev, sizeof(uint64_t) * n_lhs,                       \
                                    cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));                   \
        return 0;                                                                                                    \
    }
DEF_REFACTOR_DEV(klujax_cuda_refactor_dev_f64, 0, 0)
DEF_REFACTOR_DEV(klujax_cuda_refactor_dev_c128, 1, 0)
DEF_REFACTOR_DEV(klujax_cuda_refactor_and_solve_dev_f64, 0, 1)
DEF_REFACTOR_DEV(klujax_cuda_refactor_and_solve_dev_c128, 1, 1)
#define DEF_SWN_DEV(name, cx, tr)                                                                                    \
    int name(uint64_t symbolic, int n_numeric, const uint64_t* numeric_dev, int n_lhs_b, int n_col, int n_rhs,       \
             const double* b, double* x, void* stream) {                                                             \
        std::vector<uint64_t> h;                                                                                     \
        int rc = fetch_handles(numeric_dev, n_numeric, h, static_cast<cudaStream_t>(stream));                        \
        if (rc) return rc;                                                                                           \
        return do_solve_numeric(cx, tr, symbolic, n_numeric, h.data(), n_lhs_b, n_col, n_rhs, b, x, stream);         \
    }
DEF_SWN_DEV(klujax_cuda_solve_with_numeric_dev_f64, 0, 0)
DEF_SWN_DEV(klujax_cuda_solve_with_numeric_dev_c128, 1, 0)
DEF_SWN_DEV(klujax_cuda_tsolve_with_numeric_dev_f64, 0, 1)
DEF_SWN_DEV(klujax_cuda_tsolve_with_numeric_dev_c128, 1, 1)
int klujax_cuda_free_numeric_dev(int n, const uint64_t* numeric_dev, int32_t* freed, void* stream) {
    std::vector<uint64_t> h;
    int rc = fetch_handles(numeric_dev, n, h, static_cast<cudaStream_t>(stream));
    if (rc) return rc;
    return klujax_cuda_free_numeric(n, h.data(), freed);
}

int klujax_cuda_free_numeric(int n, const uint64_t* numeric, int32_t* freed) {
    if (freed) *freed = 1;  // the reference always reports 1 (klujax.cpp:1290)
    std::vector<NumericBatch*> dead;
    {
        std::lock_guard<std::mutex> g(g_reg_mu);
        for (int m = 0; m < n; ++m) {
            if (!numeric || numeric[m] == 0) continue;
            const uint32_t id = static_cast<uint32_t>(numeric[m] >> 32);
            const int el = static_cast<int>(numeric[m] & 0xffffffffu) - 1;
            auto it = g_batches.find(id);
            if (it == g_batches.end()) continue;
            NumericBatch* nb = it->second;
            if (el < 0 || el >= nb->L || nb->freed[el]) continue;
            nb->freed[el] = 1;
            if (--nb->live == 0) {
                g_batches.erase(it);
                dead.push_back(nb);
            }
        }
    }
    for (NumericBatch* nb : dead) {
        cudaDeviceSynchronize();  // storage may still be read by enqueued solves
        delete nb;
    }
    return 0;
}

// ---- dot ----------------------------------------------------------------------
int klujax_cuda_dot_f64(int n_lhs, int n_col, int n_rhs, int n_nz, const int32_t* Ai,
                        const int32_t* Aj, const double* Ax, const double* x, double* b,
                        void* stream) {
    std::string err;
    int rc = coo_spmm(0, n_lhs, n_col, n_rhs, n_nz, Ai, Aj, Ax, x, b, static_cast<cudaStream_t>(stream), err);
    return rc ? fail(rc, err) : 0;
}
int klujax_cuda_dot_c128(int n_lhs, int n_col, int n_rhs, int n_nz, const int32_t* Ai,
                         const int32_t* Aj, const double* Ax, const double* x, double* b,
                         void* stream) {
    std::string err;
    int rc = coo_spmm(1, n_lhs, n_col, n_rhs, n_nz, Ai, Aj, Ax, x, b, static_cast<cudaStream_t>(stream), err);
    return rc ? fail(rc, err) : 0;
}

// ---- transpose of dot with respect to the values -------------------------------
int klujax_cuda_dot_transpose_ax_f64(int n_lhs, int n_col, int n_rhs, int n_nz, const int32_t* Ai,
                                     const int32_t* Aj, const double* ct, const double* x,
                                     double* dAx, void* stream) {
    std::string err;
    int rc = coo_sddmm(0, n_lhs, n_col, n_rhs, n_nz, Ai, Aj, ct, x, dAx, static_cast<cudaStream_t>(stream), err);
    return rc ? fail(rc, err) : 0;
}
int klujax_cuda_dot_transpose_ax_c128(int n_lhs, int n_col, int n_rhs, int n_nz, const int32_t* Ai,
                                      const int32_t* Aj, const double* ct, const double* x,
                                      double* dAx, void* stream) {
    std::string err;
    int rc = coo_sddmm(1, n_lhs, n_col, n_rhs, n_nz, Ai, Aj, ct, x, dAx, static_cast<cudaStream_t>(stream), err);
    return rc ? fail(rc, err) : 0;
}

}  // extern "C"
