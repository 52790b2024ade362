/*
 * klu_oracle.c — TEST INFRASTRUCTURE ONLY (see klu_oracle.h).
 *
 * klu_analyze restated (KLU's analyze_worker with klu_defaults: btf = 1,
 * ordering = AMD; reached from /root/reference/klujax.cpp:296 and :1366),
 * plus the real/complex instantiations of klu_kernel.inc and the type
 * dispatch KluTraits<T> performs at /root/reference/klujax.cpp:129-163.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "klu_oracle.h"

#define EMPTY (-1)
#define FLIP(i) (-(i) - 2)
#define UNFLIP(i) (((i) < EMPTY) ? FLIP(i) : (i))

ko_symbolic *ko_analyze(int n, const int *Ap, const int *Ai, int *status) {
    int k, p, block, nblocks, maxblock = 1, nzoff = 0, nmatch = 0, ok = 1;
    *status = KO_OK;
    if (n <= 0 || Ap[0] != 0) {
        *status = KO_INVALID;
        return 0;
    }
    for (k = 0; k < n; k++) {
        if (Ap[k] > Ap[k + 1]) {
            *status = KO_INVALID;
            return 0;
        }
        for (p = Ap[k]; p < Ap[k + 1]; p++)
            if (Ai[p] < 0 || Ai[p] >= n) {
                *status = KO_INVALID;
                return 0;
            }
    }
    int nz = Ap[n];
    ko_symbolic *S = (ko_symbolic *)calloc(1, sizeof(ko_symbolic));
    S->n = n;
    S->nz = nz;
    S->P = (int *)malloc(sizeof(int) * (size_t)n);
    S->Q = (int *)malloc(sizeof(int) * (size_t)n);
    S->R = (int *)malloc(sizeof(int) * (size_t)(n + 1));
    S->Lnz = (double *)malloc(sizeof(double) * (size_t)n);
    int *Pbtf = (int *)malloc(sizeof(int) * (size_t)n);
    int *Qbtf = (int *)malloc(sizeof(int) * (size_t)n);
    int *Work = (int *)malloc(sizeof(int) * (size_t)(5 * n + 5));
    int *Pinv = (int *)malloc(sizeof(int) * (size_t)n);
    int *Cp = (int *)malloc(sizeof(int) * (size_t)(n + 1));
    int *Ci = (int *)malloc(sizeof(int) * (size_t)(nz + 1));
    int *Pblk = (int *)malloc(sizeof(int) * (size_t)n);

    nblocks = ko_btf_order(n, Ap, Ai, Pbtf, Qbtf, S->R, &nmatch, Work);
    S->structural_rank = nmatch;
    if (nmatch < n) {
        for (k = 0; k < n; k++) Qbtf[k] = UNFLIP(Qbtf[k]);
        *status = KO_SINGULAR; /* warning only; factor will hit a zero pivot */
    }
    for (block = 0; block < nblocks; block++) {
        int nk = S->R[block + 1] - S->R[block];
        if (nk > maxblock) maxblock = nk;
    }
    for (k = 0; k < n; k++) Pinv[Pbtf[k]] = k;

    for (block = 0; ok && block < nblocks; block++) {
        int k1 = S->R[block], k2 = S->R[block + 1], nk = k2 - k1, pc = 0;
        double lnz1;
        for (k = k1; k < k2; k++) {
            int newcol = k - k1, oldcol = Qbtf[k];
            Cp[newcol] = pc;
            for (p = Ap[oldcol]; p < Ap[oldcol + 1]; p++) {
                int newrow = Pinv[Ai[p]];
                if (newrow < k1) {
                    nzoff++;
                } else {
                    Ci[pc++] = newrow - k1;
                }
            }
        }
        Cp[nk] = pc;
        if (nk <= 3) {
            for (k = 0; k < nk; k++) Pblk[k] = k;
            lnz1 = nk * (nk + 1) / 2;
        } else {
            double lnz = 0;
            int r = ko_amd_order(nk, Cp, Ci, Pblk, &lnz);
            if (r < 0) {
                ok = 0;
                break;
            }
            lnz1 = (double)((int)lnz) + nk;
        }
        S->Lnz[block] = lnz1;
        for (k = 0; k < nk; k++) {
            S->P[k + k1] = Pbtf[Pblk[k] + k1];
            S->Q[k + k1] = Qbtf[Pblk[k] + k1];
        }
    }
    S->nblocks = nblocks;
    S->maxblock = maxblock;
    S->nzoff = nzoff;
    free(Pbtf);
    free(Qbtf);
    free(Work);
    free(Pinv);
    free(Cp);
    free(Ci);
    free(Pblk);
    if (!ok) {
        *status = KO_INVALID;
        ko_free_symbolic(S);
        return 0;
    }
    return S;
}

void ko_free_symbolic(ko_symbolic *S) {
    if (!S) return;
    free(S->P);
    free(S->Q);
    free(S->R);
    free(S->Lnz);
    free(S);
}

void ko_free_numeric(ko_numeric *N) {
    if (!N) return;
    free(N->Pnum);
    free(N->Pinv);
    free(N->Lp);
    free(N->Llen);
    free(N->Up);
    free(N->Ulen);
    free(N->Li);
    free(N->Ui);
    free(N->Lx);
    free(N->Ux);
    free(N->Udiag);
    free(N->Rs);
    free(N->Offp);
    free(N->Offi);
    free(N->Offx);
    free(N->Xwork);
    free(N);
}

/* ------------------------- real arithmetic ------------------------------ */
#define Entry double
#define FN(name) kr_##name
#define CLEAR(a) ((a) = 0.0)
#define IS_ZERO(a) ((a) == 0.0)
#define MULT_SUB(c, a, b) ((c) -= (a) * (b))
#define DIVE(c, a, b) ((c) = (a) / (b))
#define ABSE(s, a) ((s) = fabs(a))
#define SCALE_DIV_ASSIGN(a, c, s) ((a) = (c) / (s))
#include "klu_kernel.inc"
#undef Entry
#undef FN
#undef CLEAR
#undef IS_ZERO
#undef MULT_SUB
#undef DIVE
#undef ABSE
#undef SCALE_DIV_ASSIGN

/* ------------------------ complex arithmetic ---------------------------- */
typedef struct {
    double r, i;
} kz_entry;

/* SuiteSparse_config_hypot */
static double kz_hypot(double x, double y) {
    double s, r;
    x = fabs(x);
    y = fabs(y);
    if (x >= y) {
        if (x + y == x) {
            s = x;
        } else {
            r = y / x;
            s = x * sqrt(1.0 + r * r);
        }
    } else {
        if (y + x == y) {
            s = y;
        } else {
            r = x / y;
            s = y * sqrt(1.0 + r * r);
        }
    }
    return s;
}

/* SuiteSparse_config_divcomplex (Smith's method) */
static kz_entry kz_div(kz_entry a, kz_entry b) {
    kz_entry c;
    double tr = fabs(b.r), ti = fabs(b.i), r, den;
    if (tr >= ti) {
        r = b.i / b.r;
        den = b.r + r * b.i;
        c.r = (a.r + a.i * r) / den;
        c.i = (a.i - a.r * r) / den;
    } else {
        r = b.r / b.i;
        den = r * b.r + b.i;
        c.r = (a.r * r + a.i) / den;
        c.i = (a.i * r - a.r) / den;
    }
    return c;
}

#define Entry kz_entry
#define FN(name) kz_##name
#define CLEAR(a) ((a).r = 0.0, (a).i = 0.0)
#define IS_ZERO(a) ((a).r == 0.0 && (a).i == 0.0)
#define MULT_SUB(c, a, b)                               \
    do {                                                \
        double _cr = (a).r * (b).r - (a).i * (b).i;     \
        double _ci = (a).i * (b).r + (a).r * (b).i;     \
        (c).r -= _cr;                                   \
        (c).i -= _ci;                                   \
    } while (0)
#define DIVE(c, a, b) ((c) = kz_div((a), (b)))
#define ABSE(s, a) ((s) = kz_hypot((a).r, (a).i))
#define SCALE_DIV_ASSIGN(a, c, s) ((a).r = (c).r / (s), (a).i = (c).i / (s))
#include "klu_kernel.inc"
#undef Entry
#undef FN

/* ------------------------------ dispatch -------------------------------- */
ko_numeric *ko_factor(const int *Ap, const int *Ai, const double *Ax,
                      const ko_symbolic *S, int is_complex, int *status) {
    return is_complex ? kz_factor(Ap, Ai, Ax, S, status)
                      : kr_factor(Ap, Ai, Ax, S, status);
}
int ko_refactor(const int *Ap, const int *Ai, const double *Ax,
                const ko_symbolic *S, ko_numeric *N, int *status) {
    return N->is_complex ? kz_refactor(Ap, Ai, Ax, S, N, status)
                         : kr_refactor(Ap, Ai, Ax, S, N, status);
}
int ko_solve(const ko_symbolic *S, ko_numeric *N, int d, int nrhs, double *B) {
    return N->is_complex ? kz_solve(S, N, d, nrhs, B)
                         : kr_solve(S, N, d, nrhs, B);
}
int ko_tsolve(const ko_symbolic *S, ko_numeric *N, int d, int nrhs, double *B) {
    return N->is_complex ? kz_tsolve(S, N, d, nrhs, B)
                         : kr_tsolve(S, N, d, nrhs, B);
}
