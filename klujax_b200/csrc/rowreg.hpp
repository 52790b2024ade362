// rowreg.hpp — the register-resident regime: ONE warp per system, every lane owns rows.
//
// Replaces the loop body of /root/reference/klujax.cpp:431-451 (klu_factor + klu_solve per
// system) for matrices whose live values fit the register file of a warp (cfg 2, cfg 5).
//
// Formulation: right-looking elimination with the static pivot sequence.  Row i
// of the permuted, row-scaled matrix lives in ONE lane; its entries sit in value registers whose
// NAMES are shared by all lanes, so that the update of pivot k,
//        A(i, j) -= L(i, k) * U(k, j)      for all rows i with L(i,k) != 0,
// is one predicated complex multiply-add for the whole warp: L(i,k) and A(i,j) are the lane's own
// registers, U(k,j) arrives by SHFL.IDX from the lane that owns row k.  No value of the
// factorization ever round-trips through shared memory, and there is no per-multiply-add
// schedule word: the program below is turned into straight-line PTX specialised to the
// pattern (rowreg_ptx.cpp) and assembled by the driver's JIT once per symbolic handle.
// The right-hand side travels as extra columns (the forward solve is fused into the
// elimination); rows that pivot early are parked in shared memory (U only — L is dead as soon
// as it has been applied) for the back substitution, rows that pivot last stay in registers.
//
// This header: the planner (host, once per symbolic handle and right-hand-side width), its
// instruction list, and a host interpreter of that list (tests/: the planner is checked against
// the oracle on the CPU; only the PTX rendering needs a GPU).
#pragma once
#include <complex>
#include <cstdint>
#include <string>
#include <vector>

#include "plan.hpp"
#include "symbolic.hpp"

namespace kb {
namespace rr {

constexpr uint16_t kNone = 0xFFFF;
constexpr int kLanes = 32;

enum OpKind : uint8_t {
    OP_ROWSCALE = 1,  // a = first table, b = tables per lane-row (max entries), c = lane-row q:
                      //   Rs[q*32+lane] = max |A(row,:)|, entries of the row divided by it (in Abuf)
    OP_LOADV,         // lanes in mask: R[a] = pool[t * 32 + lane]   (entries of A, first touched in this epoch)
    OP_LOADB,         // R[a] = b[tab[lane]] / Rs[tab2[lane]]   (tab2 none: unscaled; tab none: 0)
    OP_OFFMAC,        // R[a] -= Abuf[tab[lane]] * xbuf[tab2[lane]]  (BTF off-diagonal entries)
    OP_PIVOT,         // D = shfl(R[a], lane b); INV[t] = 1 / D; singular if D == 0 or not finite
    OP_STOREINV,      // R[a] = INV[t] in lane b
    OP_LMUL,          // lanes in mask: R[a] *= INV[t]; growth = max(growth, |R[a]|^2)
    OP_BCAST,         // U[t] = shfl(R[a], lane b)
    OP_ZERO,          // lanes in mask: R[a] = 0
    OP_LSPREAD,       // L[t] = shfl(R[a], lane tab[lane]): multipliers for entries moved to other lanes
    OP_MAC,           // lanes in mask: R[a] -= R[b] * U[t]   (b < 0: the multiplier is L[t2])
    OP_DUMP,          // lanes in mask: Ubuf[t * 32 + lane] = R[a]      (plane t: immediate address)
    OP_DUMPT,         // lanes in mask: Ubuf[tab[lane]] = R[a]          (entries moved to another lane)
    OP_XMUL,          // X[t] = R[a] * R[b]             (all lanes; row kept in registers)
    OP_XMUL_S,        // X[t] = R[a] * Ubuf[t2 * 32 + lane]  (lanes in mask; parked row)
    OP_SETX,          // lanes in mask: R[a] = X[t]
    OP_BCASTX,        // U[t2] = shfl(X[t], lane b)
    OP_MAC_S,         // lanes in mask: R[a] -= Ubuf[t2 * 32 + lane] * U[t]
    OP_STOREX,        // x[tab[lane]] = R[a] / Rs[tab2[lane]]  (tab2 none: unscaled; tab none: skip)
    OP_XBUF,          // xbuf[tab[lane]] = R[a]         (only when off-diagonal blocks exist)
    OP_SYNC,          // __syncwarp: shared-memory writes of the warp become visible
};

struct Ins {
    uint8_t op = 0;
    int16_t a = -1, b = -1, c = 0;
    int32_t t = -1, t2 = -1;     // temporaries (INV / U / X are separate name spaces)
    int32_t tab = -1, tab2 = -1; // index of a 32-entry table in Plan::tabs
    uint32_t mask = 0xffffffffu;
};

struct Stats {
    int P = 0;               // value registers (complex or real)
    int macs = 0, mac_ins = 0, bcast = 0, pivots = 0, lmul = 0, zero = 0, dump = 0, loadv = 0;
    int mac_s = 0, xmul = 0, planes = 0, epochs = 0, pinned = 0, spread = 0, moved = 0;
    long long useful_mac = 0;  // multiply-adds the pattern needs (lanes that do real work)
};

struct Plan {
    int n = 0, nz = 0, r = 1, P = 0;
    bool is_complex = true, transpose = false;
    int n_inv = 0, n_u = 0, n_x = 0, n_l = 0;  // temporaries
    int planes = 0;                       // shared-memory pool: planes of 32 values (entries of A, parked values)
    std::vector<uint16_t> place;          // CSC position of an entry of A -> slot of the pool
    bool need_xbuf = false;
    std::vector<Ins> ins;
    std::vector<uint16_t> tabs;           // 32 entries per table
    Stats st;
    bool ok = false;
    std::string why;                      // when !ok
};

struct Options {
    int max_regs = 46;      // value registers the kernel may use (complex: 4 x 32-bit each)
    int pinned_rows = 32;   // rows (per BTF block) whose U stays in registers until the back solve
    int r = 1;              // right-hand-side columns carried through the elimination (1..4)
    bool transpose = false;
    bool rebalance = true;  // move entries of U out of overloaded lanes
    bool lazy_load = true;  // load an entry of A when the epoch that first touches it starts
    bool stripe = true;     // stripe long parked rows over lane groups (see rowreg.cpp)
    int stripe_min = 6;
    bool level_order = false;  // eliminate level set by level set instead of in pivot order
};

// builds the program for fused factor + solve (transpose: A^T x = b by eliminating A^T with the
// same pivots: same element growth |L||U|)
bool build_plan(const CscPattern& A, const Ordering& ord, const PivotPlan& piv, const SlotLayout& lay,
                bool is_complex, const Options& opt, Plan& out);

// host interpreter of the instruction list (test infrastructure for the planner; not a product path)
// Ax: nz values in the analysed CSC order, b: n*r (row-major [n][r]), x: n*r.  Returns status
// (0 ok, 1 singular) and max |L|.
template <typename T>
int emulate(const Plan& p, const T* Ax_csc, const T* b, T* x, double* growth);

// PTX rendering (rowreg_ptx.cpp)
struct PtxKernel {
    std::string ptx, name;
    int warps_per_cta = 8;
    bool lockstep = true;  // CTA barrier at every pivot: the warps share the instruction stream
    size_t smem_per_warp = 0;
    int P = 0;
};
bool emit_ptx(const Plan& p, PtxKernel& out);

}  // namespace rr
}  // namespace kb
