// symbolic.hpp — host-side symbolic stage of the B200 path.
//
// What stays on the CPU (BASELINE.json north_star): the COO->CSC map
// (/root/reference/klujax.cpp:86-125), KLU's analysis (BTF + AMD per block,
// reached from klujax.cpp:1366) and ONE pivoting factorization of a seed
// system (what klu_factor does per system at klujax.cpp:310/441/679), from
// which the static pivot sequence and the L/U patterns are taken.  Everything
// numeric after that runs on the GPU from the static plan (plan.hpp).
#pragma once
#include <cstdint>
#include <complex>
#include <string>
#include <vector>

namespace kb {

constexpr int KLU_OK = 0;
constexpr int KLU_SINGULAR = 1;
constexpr int KLU_OUT_OF_MEMORY = -2;
constexpr int KLU_INVALID = -3;
constexpr int KLU_TOO_LARGE = -4;

// Column-compressed pattern produced by the reference's stable bucket sort.
struct CscPattern {
    int n = 0;
    std::vector<int> colptr;  // n+1
    std::vector<int> rowidx;  // nz, in COO arrival order inside each column
    std::vector<int> coo_of;  // nz, gather map: csc slot -> COO position (Bk)
    bool has_duplicates = false;
};

CscPattern coo_to_csc(int n, int nz, const int* Ai, const int* Aj);

// Block triangular form + fill-reducing order (klu_analyze with klu_defaults).
struct Ordering {
    int n = 0, nblocks = 0, maxblock = 0, nzoff = 0, structural_rank = 0;
    std::vector<int> P, Q, R;      // row perm, col perm, block boundaries
    std::vector<double> lnz_est;   // per block (AMD_LNZ + nk)
};

int maximum_transversal(int n, const std::vector<int>& colptr,
                        const std::vector<int>& rowidx, std::vector<int>& match);
int strong_components(int n, const std::vector<int>& colptr,
                      const std::vector<int>& rowidx, std::vector<int>& Q,
                      std::vector<int>& P, std::vector<int>& R);
// approximate minimum degree on the pattern of C + C' (default AMD controls)
bool amd_order(int n, const std::vector<int>& Cp, const std::vector<int>& Ci,
               std::vector<int>& perm, double* lnz);
bool analyze_pattern(const CscPattern& A, Ordering& out);

// Result of the pivoting seed factorization: the static pivot plan.
struct PivotPlan {
    int n = 0;
    std::vector<int> Pnum;             // final row permutation (old row of pivot k)
    std::vector<int> Pinv;             // inverse of Pnum
    // per global pivot column k, block-local row indices in pivot order:
    std::vector<int> Lp, Li;           // strictly-lower pattern, Lp has n+1
    std::vector<int> Up, Ui;           // strictly-upper pattern in topological order
    std::vector<int> Offp, Offi;       // BTF off-diagonal entries per column (global pivot rows)
    std::vector<int> Offsrc;           // csc slot each off-diagonal entry comes from
    int noffdiag = 0;
    // pivot-parity certificate (plan.hpp): per pivot column, -2 if the pivot was the column's
    // diagonal candidate (KLU keeps track of which row "is the diagonal" as off-diagonal pivots
    // exchange the roles: klu_kernel's P / Pinv flips), else the pivot-order row of the diagonal
    // candidate that failed the threshold test, or -1 if it was pivotal already
    std::vector<int> diag_cand;
    int status = KLU_OK;
};

template <typename T>
bool seed_factor(const CscPattern& A, const Ordering& ord, const T* csc_values,
                 PivotPlan& plan);

}  // namespace kb
