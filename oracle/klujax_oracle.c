/*
 * klujax_oracle.c — TEST INFRASTRUCTURE ONLY (see klu_oracle.h).
 *
 * CPU restatement of the batch drivers in /root/reference/klujax.cpp, i.e.
 * the way klujax marshals a batch through KLU: per-call COO->CSC
 * (klujax.cpp:86-125), [L][n][r] <-> [L][r][n] transposes (:283-289,
 * :328-334), and a serial loop over n_lhs with a fresh klu_factor per system
 * (:300-322, :431-451).  The ko_* calls stand where klu_* stands there.
 * Sizes/offsets are size_t here (the reference uses int).
 */
#include <stdlib.h>
#include <string.h>
#include "klu_oracle.h"

/* klujax.cpp:86-125 — stable counting sort by column; Bk is the gather map */
void kjo_coo_to_csc(int n_col, int n_nz, const int *Ai, const int *Aj,
                    int *Bi, int *Bp, int *Bk) {
    int n, j, cumsum = 0, temp, last = 0;
    for (j = 0; j <= n_col; j++) Bp[j] = 0;
    for (n = 0; n < n_nz; n++) Bp[Aj[n]] += 1;
    for (j = 0; j <= n_col; j++) {
        temp = Bp[j];
        Bp[j] = cumsum;
        cumsum += temp;
    }
    for (n = 0; n < n_nz; n++) {
        int col = Aj[n], dest = Bp[col];
        Bi[dest] = Ai[n];
        Bk[dest] = n;
        Bp[col] += 1;
    }
    for (j = 0; j <= n_col; j++) {
        temp = Bp[j];
        Bp[j] = last;
        last = temp;
    }
}

static int check_idx(int n_col, int n_nz, const int *Ai, const int *Aj) {
    int n;
    for (n = 0; n < n_nz; n++)
        if (Ai[n] < 0 || Ai[n] >= n_col || Aj[n] < 0 || Aj[n] >= n_col)
            return 0;
    return 1;
}

/* klujax.cpp:1331-1374 */
int kjo_analyze(int n_col, int n_nz, const int *Ai, const int *Aj,
                uint64_t *symbolic) {
    int status = 0;
    *symbolic = 0;
    if (!check_idx(n_col, n_nz, Ai, Aj)) return 1;
    int *Bi = (int *)malloc(sizeof(int) * (size_t)(n_nz + 1));
    int *Bp = (int *)malloc(sizeof(int) * (size_t)(n_col + 2));
    int *Bk = (int *)malloc(sizeof(int) * (size_t)(n_nz + 1));
    kjo_coo_to_csc(n_col, n_nz, Ai, Aj, Bi, Bp, Bk);
    ko_symbolic *S = ko_analyze(n_col, Bp, Bi, &status);
    free(Bi);
    free(Bp);
    free(Bk);
    if (!S) return 2;
    *symbolic = (uint64_t)(uintptr_t)S;
    return 0;
}

/* klujax.cpp:1293-1315 */
int kjo_free_symbolic(uint64_t symbolic) {
    if (symbolic == 0) return 0;
    ko_free_symbolic((ko_symbolic *)(uintptr_t)symbolic);
    return 1;
}

/* klujax.cpp:1273-1291 */
int kjo_free_numeric(int n, const uint64_t *numeric) {
    int i;
    for (i = 0; i < n; i++)
        if (numeric[i] != 0) ko_free_numeric((ko_numeric *)(uintptr_t)numeric[i]);
    return 1;
}

const ko_symbolic *kjo_symbolic_ptr(uint64_t h) {
    return (const ko_symbolic *)(uintptr_t)h;
}
const ko_numeric *kjo_numeric_ptr(uint64_t h) {
    return (const ko_numeric *)(uintptr_t)h;
}

typedef struct {
    int *Bi, *Bp, *Bk;
    double *Bx;
} csc_t;

static void csc_make(csc_t *c, int es, int n_col, int n_nz, const int *Ai,
                     const int *Aj) {
    c->Bk = (int *)malloc(sizeof(int) * (size_t)(n_nz + 1));
    c->Bi = (int *)malloc(sizeof(int) * (size_t)(n_nz + 1));
    c->Bp = (int *)malloc(sizeof(int) * (size_t)(n_col + 2));
    c->Bx = (double *)malloc(sizeof(double) * (size_t)es * (size_t)(n_nz + 1));
    kjo_coo_to_csc(n_col, n_nz, Ai, Aj, c->Bi, c->Bp, c->Bk);
}
static void csc_free(csc_t *c) {
    free(c->Bk);
    free(c->Bi);
    free(c->Bp);
    free(c->Bx);
}
/* klujax.cpp:305-307 */
static void csc_gather(const csc_t *c, int es, int n_nz, const double *Ax_m) {
    int k;
    if (es == 1)
        for (k = 0; k < n_nz; k++) c->Bx[k] = Ax_m[c->Bk[k]];
    else
        for (k = 0; k < n_nz; k++) {
            c->Bx[2 * k] = Ax_m[2 * (size_t)c->Bk[k]];
            c->Bx[2 * k + 1] = Ax_m[2 * (size_t)c->Bk[k] + 1];
        }
}

/* [n_col][n_rhs] (row-major) -> [n_rhs][n_col]; klujax.cpp:283-289 */
static void tr_in(int es, int n_col, int n_rhs, const double *b, double *xt) {
    int n, p, e;
    for (n = 0; n < n_col; n++)
        for (p = 0; p < n_rhs; p++)
            for (e = 0; e < es; e++)
                xt[((size_t)p * n_col + n) * es + e] =
                    b[((size_t)n * n_rhs + p) * es + e];
}
/* klujax.cpp:328-334 */
static void tr_out(int es, int n_col, int n_rhs, const double *xt, double *x) {
    int n, p, e;
    for (n = 0; n < n_col; n++)
        for (p = 0; p < n_rhs; p++)
            for (e = 0; e < es; e++)
                x[((size_t)n * n_rhs + p) * es + e] =
                    xt[((size_t)p * n_col + n) * es + e];
}

/* the loop body of klujax.cpp:431-451 / :560-580 for systems [lo,hi) */
int kjo_solve_with_symbol_range(int is_complex, int transpose, int lo, int hi,
                                int n_col, int n_rhs, int n_nz, const int *Ai,
                                const int *Aj, const double *Ax,
                                const double *b, uint64_t symbolic, double *x) {
    int es = is_complex ? 2 : 1, i, status = 0, rc = 0;
    ko_symbolic *S = (ko_symbolic *)(uintptr_t)symbolic;
    if (!S) return 1;
    if (!check_idx(n_col, n_nz, Ai, Aj)) return 1;
    csc_t c;
    csc_make(&c, es, n_col, n_nz, Ai, Aj);
    size_t per = (size_t)n_col * (size_t)n_rhs * (size_t)es;
    double *xt = (double *)malloc(sizeof(double) * (per + 1));
    for (i = lo; i < hi; i++) {
        csc_gather(&c, es, n_nz, Ax + (size_t)i * n_nz * es);
        tr_in(es, n_col, n_rhs, b + (size_t)i * per, xt);
        ko_numeric *N = ko_factor(c.Bp, c.Bi, c.Bx, S, is_complex, &status);
        if (!N || status < KO_OK) {
            rc = 3; /* "klu_factor failed (singular matrix?)" */
            if (N) ko_free_numeric(N);
            break;
        }
        if (transpose)
            ko_tsolve(S, N, n_col, n_rhs, xt);
        else
            ko_solve(S, N, n_col, n_rhs, xt);
        ko_free_numeric(N);
        tr_out(es, n_col, n_rhs, xt, x + (size_t)i * per);
    }
    free(xt);
    csc_free(&c);
    return rc;
}

/* klujax.cpp:382-463 (solve) and :513-590 (tsolve) */
int kjo_solve_with_symbol(int is_complex, int transpose, int n_lhs, int n_col,
                          int n_rhs, int n_nz, const int *Ai, const int *Aj,
                          const double *Ax, const double *b, uint64_t symbolic,
                          double *x) {
    return kjo_solve_with_symbol_range(is_complex, transpose, 0, n_lhs, n_col,
                                       n_rhs, n_nz, Ai, Aj, Ax, b, symbolic, x);
}

/* klujax.cpp:251-338 */
int kjo_solve(int is_complex, int n_lhs, int n_col, int n_rhs, int n_nz,
              const int *Ai, const int *Aj, const double *Ax, const double *b,
              double *x) {
    uint64_t sym = 0;
    int rc = kjo_analyze(n_col, n_nz, Ai, Aj, &sym);
    if (rc) return rc;
    rc = kjo_solve_with_symbol(is_complex, 0, n_lhs, n_col, n_rhs, n_nz, Ai, Aj,
                               Ax, b, sym, x);
    kjo_free_symbolic(sym);
    return rc;
}

/* klujax.cpp:638-693 */
int kjo_factor(int is_complex, int n_lhs, int n_nz, const int *Ai,
               const int *Aj, const double *Ax, uint64_t symbolic,
               uint64_t *numeric) {
    int es = is_complex ? 2 : 1, i, j, status = 0;
    ko_symbolic *S = (ko_symbolic *)(uintptr_t)symbolic;
    if (!S) return 1;
    int n_col = S->n;
    if (!check_idx(n_col, n_nz, Ai, Aj)) return 1;
    csc_t c;
    csc_make(&c, es, n_col, n_nz, Ai, Aj);
    for (i = 0; i < n_lhs; i++) {
        csc_gather(&c, es, n_nz, Ax + (size_t)i * n_nz * es);
        ko_numeric *N = ko_factor(c.Bp, c.Bi, c.Bx, S, is_complex, &status);
        if (!N || status < KO_OK) {
            if (N) ko_free_numeric(N);
            for (j = 0; j < i; j++) { /* klujax.cpp:684-687 */
                ko_free_numeric((ko_numeric *)(uintptr_t)numeric[j]);
                numeric[j] = 0;
            }
            csc_free(&c);
            return 3;
        }
        numeric[i] = (uint64_t)(uintptr_t)N;
    }
    csc_free(&c);
    return 0;
}

/* klujax.cpp:733-789 */
int kjo_refactor(int is_complex, int n_lhs, int n_nz, const int *Ai,
                 const int *Aj, const double *Ax, uint64_t symbolic,
                 const uint64_t *numeric_in, uint64_t *numeric_out) {
    int es = is_complex ? 2 : 1, i, status = 0;
    ko_symbolic *S = (ko_symbolic *)(uintptr_t)symbolic;
    if (!S) return 1;
    int n_col = S->n;
    if (!check_idx(n_col, n_nz, Ai, Aj)) return 1;
    csc_t c;
    csc_make(&c, es, n_col, n_nz, Ai, Aj);
    for (i = 0; i < n_lhs; i++) {
        ko_numeric *N = (ko_numeric *)(uintptr_t)numeric_in[i];
        if (!N) {
            csc_free(&c);
            return 1;
        }
        csc_gather(&c, es, n_nz, Ax + (size_t)i * n_nz * es);
        if (!ko_refactor(c.Bp, c.Bi, c.Bx, S, N, &status)) {
            csc_free(&c);
            return 3;
        }
        numeric_out[i] = numeric_in[i];
    }
    csc_free(&c);
    return 0;
}

/* klujax.cpp:995-1103 (solve) and :1141-1235 (tsolve); the rank parsing of
 * b (1/2/3-D) is done by the caller, which passes n_lhs_b/n_col/n_rhs.      */
int kjo_solve_with_numeric(int is_complex, int transpose, int n_numeric,
                           int n_lhs_b, int n_col, int n_rhs,
                           uint64_t symbolic, const uint64_t *numeric,
                           const double *b, double *x) {
    int es = is_complex ? 2 : 1, m, n_lhs;
    ko_symbolic *S = (ko_symbolic *)(uintptr_t)symbolic;
    if (!S) return 1;
    int bc_num = (n_numeric == 1), bc_b = (n_lhs_b == 1);
    if (!bc_num && !bc_b) {
        if (n_numeric != n_lhs_b) return 1;
        n_lhs = n_numeric;
    } else if (!bc_num) {
        n_lhs = n_numeric;
    } else if (!bc_b) {
        n_lhs = n_lhs_b;
    } else {
        n_lhs = 1;
    }
    size_t per = (size_t)n_col * (size_t)n_rhs * (size_t)es;
    double *xt = (double *)malloc(sizeof(double) * (per + 1));
    for (m = 0; m < n_lhs; m++) {
        uint64_t h = bc_num ? numeric[0] : numeric[m];
        if (!h) {
            free(xt);
            return 1;
        }
        ko_numeric *N = (ko_numeric *)(uintptr_t)h;
        tr_in(es, n_col, n_rhs, b + (bc_b ? 0 : (size_t)m * per), xt);
        if (transpose)
            ko_tsolve(S, N, n_col, n_rhs, xt);
        else
            ko_solve(S, N, n_col, n_rhs, xt);
        tr_out(es, n_col, n_rhs, xt, x + (size_t)m * per);
    }
    free(xt);
    return 0;
}

/* klujax.cpp:839-936 */
int kjo_refactor_and_solve(int is_complex, int n_lhs, int n_col, int n_rhs,
                           int n_nz, const int *Ai, const int *Aj,
                           const double *Ax, const double *b,
                           uint64_t symbolic, const uint64_t *numeric_in,
                           double *x, uint64_t *numeric_out) {
    int rc = kjo_refactor(is_complex, n_lhs, n_nz, Ai, Aj, Ax, symbolic,
                          numeric_in, numeric_out);
    if (rc) return rc;
    return kjo_solve_with_numeric(is_complex, 0, n_lhs, n_lhs, n_col, n_rhs,
                                  symbolic, numeric_in, b, x);
}
