// plan.hpp — the static elimination schedule derived once per symbolic handle.
//
// From the pivot plan (seed_lu.cpp) the host derives, ONCE, everything the
// numeric kernels need; no kernel ever touches a sparsity structure again:
//
//   * a slot layout: every entry of L, U, the pivots and the BTF off-diagonal
//     blocks gets one slot in a per-system value array V (fill-in included);
//   * task lists in "dot" form (Crout):  V[out] = (V[out] - sum V[a]*V[b]) [*V[d]]
//     for the refactorization (what klu_refactor does column by column,
//     /root/reference/klujax.cpp:781,911) and for the triangular solves
//     (klu_solve / klu_tsolve, klujax.cpp:315,445,573,917,1087,1220), including
//     the BTF block back-substitution;
//   * level sets of that task DAG (dependencies + anti-dependencies), which is
//     the schedule both kernel regimes execute.
#pragma once
#include <cstdint>
#include <vector>

#include "symbolic.hpp"

namespace kb {

enum TaskKind : uint8_t {
    TK_NOP = 0,
    TK_SUB = 1,        // V[out] -= acc
    TK_RECIP = 2,      // V[out]  = 1 / (V[out] - acc)        (pivot; singularity check)
    TK_MUL = 3,        // V[out]  = (V[out] - acc) * V[div]   (solve with a pivot)
    TK_MUL_TRACK = 4,  // same, and track max |V[out]|        (entry of L: growth certificate)
    // pivot-parity certificate of a column whose static pivot is NOT the diagonal candidate (KLU
    // chose the column maximum at seeding time because the diagonal failed the threshold test):
    // KLU makes the same choice iff no other candidate is larger, |L(i,k)| <= 1, and the diagonal
    // candidate still fails, |L(diag,k)| < tol.  The tracked magnitude is weighted so that ONE
    // number certifies every column: growth <= 1/tol = 1000  (weights 1, 1/tol, 1/tol^2).
    TK_MUL_TRACK_OFF = 5,   // entry of L in such a column:            |L| * 1e3
    TK_MUL_TRACK_DIAG = 6,  // the diagonal candidate in such a column: |L| * 1e6
};

struct Task {
    int out = 0, div = -1;
    int pbeg = 0, plen = 0;
    int level = 0;
    uint8_t kind = TK_SUB;
    uint8_t fresh = 0;  // first write of a recycled slot: start from 0, not from the old content
};

// Slot layout of one system's value array.
struct SlotLayout {
    int n = 0;
    int n_lu = 0;       // slots [0, n_lu): per pivot column k: U(:,k) | pivot | L(:,k)
    int n_off = 0;      // slots [n_lu, n_lu+n_off): BTF off-diagonal entries
    int n_val = 0;      // n_lu + n_off
    std::vector<int> col_begin;   // n+1: first slot of column k
    std::vector<int> diag_slot;   // n
    std::vector<int> slot_row;    // n_val: global pivot row of the entry
    std::vector<int> slot_col;    // n_val: global pivot column
    std::vector<int> slot_csc;    // n_val: csc position of A landing here, -1 = fill-in
    // A's pattern sorted by (col,row) -> slot, for the per-call COO map
    std::vector<int> sorted_colptr, sorted_row, sorted_slot;
    // row-scale lists: slots holding the entries of ORIGINAL row i
    std::vector<int> rs_ptr, rs_slot;
    int64_t nnz_l = 0, nnz_u = 0;  // strictly lower / strictly upper counts
};

enum ProgramMode : int {
    PM_FACTOR = 0,         // refactor only                    (factor / refactor)
    PM_FACTOR_SOLVE = 1,   // refactor + solve                 (solve, solve_with_symbol, refactor_and_solve)
    PM_FACTOR_TSOLVE = 2,  // refactor + transpose solve       (tsolve_with_symbol)
    PM_SOLVE = 3,          // solve from stored factors        (solve_with_numeric)
    PM_TSOLVE = 4,         // transpose solve, stored factors  (tsolve_with_numeric)
};

// Leveled task list over the index space [0, n_val) U [n_val, n_val + n*rc):
// value slots, then the right-hand-side slots X[k][c] = n_val + k*rc + c.
struct TaskProgram {
    int mode = 0, rc = 0, nlevels = 0;
    std::vector<Task> tasks;       // sorted by level
    std::vector<int> level_ptr;    // nlevels+1
    std::vector<int> pa, pb;       // operand pairs
    int64_t n_mac = 0;
    // Slot compaction (programs that do not keep the factors: fused factor + solve).  A value
    // slot is live from its first write (level 0 for the entries of A) to its last use; dead
    // slots are recycled for fill-in that is born later, so the value array a system needs in
    // shared memory shrinks to the peak number of live values.  n_val = value slots the
    // program addresses (X follows at n_val); remap[s] = physical slot of layout slot s
    // (empty = identity, n_val = lay.n_val).
    int n_val = 0;
    std::vector<int> remap;
    // Dense tail in registers (kernels_wide.cuh: tail stage).  The last kTail pivots of the
    // largest block form a (nearly) dense trailing submatrix that holds most of the multiply-adds;
    // with tail_t0 >= 0 the program leaves its factorization, its part of the forward solve and
    // its back substitution to the kernel's register stage: tasks fed by tail pivots are dropped,
    // the stage runs before the first step of level tail_level, tail_tab[col*kTail + row] is the
    // (physical) slot of entry (tail_t0 + row, tail_t0 + col) or 0xFFFF.
    int tail_t0 = -1, tail_level = 0;
    std::vector<uint16_t> tail_tab;
};
constexpr int kTail = 32;

void build_layout(const CscPattern& A, const Ordering& ord, const PivotPlan& piv, SlotLayout& lay);
// compact: 0 = keep the layout's slots; else recycle dead slots (TaskProgram::remap), keeping
// every slot in its shared-memory bank column modulo `compact` (8 for 16-byte, 16 for 8-byte values)
void build_program(const Ordering& ord, const PivotPlan& piv, const SlotLayout& lay, int mode,
                   int rc, TaskProgram& prog, int split = 0, int compact = 0, int tail_t0 = -1);
// first pivot of a dense tail the register stage can take (fused factor + solve), or -1
int pick_tail(const Ordering& ord, const PivotPlan& piv, const SlotLayout& lay);

// Packed schedule of the shared-memory (small-matrix) regime: one warp per system.
//
//  * control words, one per batch, warp-uniform; they travel as KERNEL PARAMETERS (constant
//    bank), so every branch of the interpreter is a uniform branch:
//        trips (bits 0-7) | log2 g (bits 8-10) | some lane finishes a pivot (bit 11)
//        | some lane multiplies by a pivot reciprocal (bit 12) | some lane is an L entry (bit 13)
//    g = lanes sharing one dot product; chunk_ptr[c] = first batch of chunk c.
//  * the lane stream, one 32-bit word per lane and line, cut into chunks of kChunkWords words
//    (the unit one TMA bulk copy brings into the shared-memory ring).  Per batch:
//        header line  : out<<18 | (pivot slot or ONE)<<4 | is-pivot<<1 | is-L-entry
//        2*trips operand lines: a<<18 | b<<4
//    Idle lanes multiply the ZERO slot and write to the TRASH slot: no lane-dependent branch.
constexpr int kChunkWords = 2048;
constexpr int kMaxBatches = 1536;  // per program (control words live in the parameter space)
constexpr int kMaxChunks = 255;
constexpr uint32_t kCwPivot = 1u << 11, kCwMul = 1u << 12, kCwTrack = 1u << 13;
// squared weight of a tracked |L|^2 by certificate class (0 diagonal pivot, 1 off-diagonal pivot
// column, 2 diagonal candidate of such a column); status code of a failed certificate
#define KB_TRACK_W2(cls) ((cls) == 0 ? 1.0 : (cls) == 1 ? 1e6 : 1e12)
constexpr int KLUJAX_PIVOT_DRIFT = 2;
constexpr double kGrowthLimit2 = 1000.0 * 1000.0 * (1.0 + 1e-9);
struct PackedProgram {
    std::vector<uint32_t> ctrl;       // nbatch
    std::vector<uint16_t> chunk_ptr;  // nchunks + 1
    std::vector<uint32_t> stream;     // nchunks * kChunkWords
    int nchunks = 0, nbatch = 0;
    int zero_slot = 0, n_slots = 0;  // slots: values | X | ZERO | ONE | TRASH
    int64_t lines_used = 0;
};
bool pack_program(const SlotLayout& lay, const TaskProgram& prog, PackedProgram& out);
// multi-warp form (kernels_wide.cuh): ctrl[step] = P | level-ends<<4 for 128-lane steps of one
// 128 (header word, operand word) pairs, one per lane; chunk_ptr[c] = first step of chunk c; nbatch = steps
bool pack_program_wide(const SlotLayout& lay, const TaskProgram& prog, PackedProgram& out, int elem_bytes = 16);


// ---------------------------------------------------------------------------------------------
// Dependency-driven regime (kernels_dep.cuh): no level barriers.  For batches that are not larger
// than the machine a CTA owns ONE system and the level-scheduled kernels pay a CTA barrier plus a
// memory round trip per level (cfg 4: 5 215 + 2 x 6 205 levels per step).  Here the work is dealt
// to the warps of the CTA in a fixed sequential order and every unit waits only for the units it
// really depends on, through ready flags in shared memory (refactorization), or runs in one
// sequential order with the dependencies of 32 consecutive tasks resolved by warp shuffles (solves).
//
// ColProgram - left-looking refactorization by columns, the shape of klu_refactor's own loop
// (/root/reference/klujax.cpp:781, 911 through KLU): a warp owns column k; for every U(j,k) in
// topological order it waits for column j and subtracts L(:,j) * U(j,k) from the column, then
// finishes the pivot and scales L(:,k).  The critical path is the depth of the column
// elimination DAG (cfg 4: 2 610 columns against 5 215 entry levels).
struct ColProgram {
    int n = 0, maxcol = 0, depth = 0;
    std::vector<int> colinfo;    // 4 per column: first slot, number of U entries, entries, first update record
    std::vector<int> ccand;      // per column: -2 diagonal pivot, -1 off-diagonal pivot, else position (inside the
                                 // column) of the diagonal candidate that failed the threshold test (plan.hpp: TK_MUL_TRACK_*)
    std::vector<int> urec;       // 4 per U entry, in slot order: pivot column j, first slot of L(:,j), its length, offset into dest
    std::vector<int> dest;       // per multiply-add: position of the destination inside column k
    int64_t n_mac = 0;
};
bool build_col_program(const PivotPlan& piv, const SlotLayout& lay, ColProgram& out);

// PanelProgram - klu_solve / klu_tsolve as ONE sequential list of row tasks (the order of
// klu_solve's own loops), cut into panels of 32 consecutive tasks: lane r of a warp owns task r of
// the panel.  Operands produced by earlier panels are "external" (read from the right-hand-side
// array, which is final for them because panels run strictly one after the other: no flags),
// operands produced by an earlier lane of the same panel are resolved inside the warp by a
// shuffle loop over the producing lanes (kernels_dep.cuh: panel_solve).  A chain of dependent
// rows - the shape of circuit and banded matrices - costs one shuffle step (~45 cycles) per row
// instead of one CTA barrier and memory round trip (~950 cycles) per row.
//   hdr       4 ints per panel: external depth E, in-panel depth I, first external line, first in-panel line
//   lane      4 ints per task : row (>= n: temporary of a split dot product), pivot slot or -1,
//                               mask of the producing lanes, bits (1 start from 0, 2 do not store, 4 multiply by the pivot)
//   ext_slot / ext_idx  E lines of 32 per panel: value slot (-2 = the constant -1) and row, -1 = none
//   in_slot             I lines of 32 per panel: value slot per producing lane, ascending (-2 = the constant -1)
// Dot products with more than `kPanelChunk` operands are split: partial tasks (start from 0, write
// a temporary row) run on other lanes, the final task subtracts their results.
constexpr int kPanelChunk = 24;
struct PanelProgram {
    int n = 0, npanels = 0, ntemps = 0, maxI = 0, maxE = 0;
    std::vector<int> hdr, lane, ext_slot, ext_idx, in_slot;
    int64_t n_mac = 0;
};
bool build_panel_program(const Ordering& ord, const PivotPlan& piv, const SlotLayout& lay, bool transpose,
                         PanelProgram& out);


// GroupProgram - the multi-right-hand-side solves with register blocking over rows (kernels_level.cuh:
// lv_solve_grouped).  The level kernel reads a row of X per entry of L / U (cfg 3: 508 GB of L2 / HBM
// traffic per step against 19 GB algorithmic).  Here up to kGroupRows consecutive row tasks of
// klu_solve's / klu_tsolve's own order form a group when they share operands (the rows of a supernode
// do): a thread owns a few right-hand-side columns and keeps the accumulators of ALL rows of the group
// in registers, so a row of X is read ONCE per group ("k-step") and applied to every row that has an
// entry there; dependencies between rows of one group are resolved thread-locally (every thread
// has all rows of its own columns).  Groups are levelled like tasks (fewer, wider levels: a chain of
// rows inside a supernode collapses into its groups).
//   ghdr  12 ints per task (sorted by level): k-steps, first k-step, rows | multiply bits << 8 (bit 30: a PARTIAL
//         task - X[rows] -= sum over its chunk of k-steps, nothing else), index of aux, row names [8] (one load
//         of three 16-byte words, prefetched for the warp's next task)
//   gaux  kGroupAux ints per group: row names [8], pivot slots [8] (-1 none), in-group slots [28]
//         (row j <- row i at j*(j-1)/2 + i, -1 none; -2: row j continues row i, coefficient -1)
//   krec  12 ints per k-step: row name of the operand, value slot per row of the group [8] (-1 none), mask of
//         the rows that have an entry, 2 pad
constexpr int kGroupRows = 8, kGroupAux = 44;
struct GroupProgram {
    int n = 0, ngroups = 0, nlevels = 0;
    std::vector<int> ghdr, gaux, krec, level_ptr;
    int64_t n_mac = 0, n_ksteps = 0;
};
// xrow (optional): name of pivot-order row k (its position in the caller's x); identity if null
bool build_group_program(const Ordering& ord, const PivotPlan& piv, const SlotLayout& lay, bool transpose,
                         const int* xrow, GroupProgram& out);

}  // namespace kb
