/*
 * klu_oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C CPU restatement of the algorithm klujax delegates to: SuiteSparse
 * KLU + BTF + AMD (v7.5.0, pinned by /root/reference/justfile:53-56), driven
 * exactly the way /root/reference/klujax.cpp drives it.  SuiteSparse itself is
 * NOT vendored in /root/reference and cannot be built here (no network), so
 * this is a restatement of the *published* algorithm (Davis & Palamadai
 * Natarajan, "Algorithm 907: KLU", ACM TOMS 37(3) 2010; Amestoy/Davis/Duff,
 * "Algorithm 837: AMD", ACM TOMS 30(3) 2004; Duff's MC21 maxtrans; Tarjan SCC).
 *
 * PARITY STATUS: the permutations P/Q/R/Pnum and the L/U patterns are
 * "parity unpinned" against real SuiteSparse binaries (none available in this
 * container, and the reference's tests never observe them).  Solutions x ARE
 * pinned: against the reference's README known-answer system
 * (/root/reference/README.md:62-86), the docs' batched KATs
 * (/root/reference/docs/examples/batched.md) and dense LAPACK / SuperLU on the
 * reference's own test generators (tests/test_oracle_*.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product (klujax_b200/)
 * never links or imports anything under oracle/.
 */
#ifndef KLU_ORACLE_H
#define KLU_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* KLU status codes (klu.h) */
#define KO_OK 0
#define KO_SINGULAR 1
#define KO_OUT_OF_MEMORY (-2)
#define KO_INVALID (-3)
#define KO_TOO_LARGE (-4)

/* ---- BTF (btf_maxtrans / btf_strongcomp / btf_order) ------------------- */
int ko_btf_maxtrans(int nrow, int ncol, const int *Ap, const int *Ai,
                    int *Match, int *Work /* 5*ncol */);
int ko_btf_strongcomp(int n, const int *Ap, const int *Ai, int *Q, int *P,
                      int *R, int *Work /* 4*n */);
int ko_btf_order(int n, const int *Ap, const int *Ai, int *P, int *Q, int *R,
                 int *nmatch, int *Work /* 5*n */);

/* ---- AMD (amd_order with default Control) ------------------------------ */
/* returns 0 ok, 1 ok-but-jumbled, <0 error. lnz_out = AMD_LNZ (may be NULL) */
int ko_amd_order(int n, const int *Ap, const int *Ai, int *P, double *lnz_out);

/* ---- KLU symbolic ------------------------------------------------------- */
typedef struct {
    int n, nz, nblocks, maxblock, nzoff, structural_rank;
    int *P, *Q, *R;   /* P[n], Q[n], R[nblocks+1] (R has n+1 slots) */
    double *Lnz;      /* per block estimate */
} ko_symbolic;

ko_symbolic *ko_analyze(int n, const int *Ap, const int *Ai, int *status);
void ko_free_symbolic(ko_symbolic *S);

/* ---- KLU numeric (is_complex selects klu_ vs klu_z_ arithmetic) -------- */
typedef struct {
    int n, nblocks, nzoff, is_complex;
    int lnz, unz, noffdiag; /* incl. diagonals, like Numeric->lnz/unz */
    int *Pnum, *Pinv;
    int *Lp, *Llen, *Up, *Ulen; /* per column k (global index): start/len   */
    int *Li, *Ui;               /* row indices (pivot order, block-local)    */
    double *Lx, *Ux;            /* values; complex = interleaved (re,im)     */
    double *Udiag;              /* n entries                                  */
    double *Rs;                 /* n, pivot order after factor               */
    int *Offp, *Offi;
    double *Offx;
    double *Xwork;              /* n entries                                  */
    int64_t lcap, ucap;
} ko_numeric;

/* Ax: CSC values (complex interleaved). returns NULL on failure (status set) */
ko_numeric *ko_factor(const int *Ap, const int *Ai, const double *Ax,
                      const ko_symbolic *S, int is_complex, int *status);
int ko_refactor(const int *Ap, const int *Ai, const double *Ax,
                const ko_symbolic *S, ko_numeric *N, int *status);
/* B is column-major, leading dimension d, nrhs columns, solved in place */
int ko_solve(const ko_symbolic *S, ko_numeric *N, int d, int nrhs, double *B);
int ko_tsolve(const ko_symbolic *S, ko_numeric *N, int d, int nrhs, double *B);
void ko_free_numeric(ko_numeric *N);

/* ---- klujax.cpp batch drivers (klujax_oracle.c) ------------------------- */
/* All follow /root/reference/klujax.cpp; is_complex: 0 = *_f64, 1 = *_c128.
 * Return 0 on success, non-zero where the reference returns ffi::Error.     */
void kjo_coo_to_csc(int n_col, int n_nz, const int *Ai, const int *Aj,
                    int *Bi, int *Bp, int *Bk);
int kjo_analyze(int n_col, int n_nz, const int *Ai, const int *Aj,
                uint64_t *symbolic);
int kjo_free_symbolic(uint64_t symbolic);
int kjo_free_numeric(int n, const uint64_t *numeric);
int kjo_solve(int is_complex, int n_lhs, int n_col, int n_rhs, int n_nz,
              const int *Ai, const int *Aj, const double *Ax, const double *b,
              double *x);
int kjo_solve_with_symbol(int is_complex, int transpose, int n_lhs, int n_col,
                          int n_rhs, int n_nz, const int *Ai, const int *Aj,
                          const double *Ax, const double *b, uint64_t symbolic,
                          double *x);
int kjo_factor(int is_complex, int n_lhs, int n_nz, const int *Ai,
               const int *Aj, const double *Ax, uint64_t symbolic,
               uint64_t *numeric);
int kjo_refactor(int is_complex, int n_lhs, int n_nz, const int *Ai,
                 const int *Aj, const double *Ax, uint64_t symbolic,
                 const uint64_t *numeric_in, uint64_t *numeric_out);
int kjo_solve_with_numeric(int is_complex, int transpose, int n_numeric,
                           int n_lhs_b, int n_col, int n_rhs,
                           uint64_t symbolic, const uint64_t *numeric,
                           const double *b, double *x);
int kjo_refactor_and_solve(int is_complex, int n_lhs, int n_col, int n_rhs,
                           int n_nz, const int *Ai, const int *Aj,
                           const double *Ax, const double *b,
                           uint64_t symbolic, const uint64_t *numeric_in,
                           double *x, uint64_t *numeric_out);
/* Faithful CPU baseline with OpenMP-free threading left to the caller:
 * processes systems [lo,hi) of a solve_with_symbol batch (used by bench.py to
 * spread the reference's serial loop over host cores).                      */
int kjo_solve_with_symbol_range(int is_complex, int transpose, int lo, int hi,
                                int n_col, int n_rhs, int n_nz, const int *Ai,
                                const int *Aj, const double *Ax,
                                const double *b, uint64_t symbolic, double *x);

/* introspection for pattern-parity tests */
const ko_symbolic *kjo_symbolic_ptr(uint64_t h);
const ko_numeric *kjo_numeric_ptr(uint64_t h);

#ifdef __cplusplus
}
#endif
#endif
