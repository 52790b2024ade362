/*
 * btf_oracle.c — TEST INFRASTRUCTURE ONLY (see klu_oracle.h).
 *
 * Restatement of SuiteSparse BTF (btf_maxtrans.c, btf_strongcomp.c,
 * btf_order.c; v7.5.0 as pinned by /root/reference/justfile:53-56), which
 * klu_analyze calls from /root/reference/klujax.cpp:296 and :1366.
 *   - maximum transversal: Duff's MC21-style column-by-column DFS augmenting
 *     paths with a "cheap assignment" pass, rows visited in stored order;
 *   - strongly connected components: non-recursive Tarjan on the graph of A*Q.
 * maxwork = 0 (klu_defaults), so the work limit is never active.
 */
#include "klu_oracle.h"

#define EMPTY (-1)
#define FLIP(j) (-(j) - 2)
#define ISFLIPPED(j) ((j) < -1)
#define UNFLIP(j) (((j) < -1) ? FLIP(j) : (j))
#define MINI(a, b) (((a) < (b)) ? (a) : (b))

/* One augmenting-path search starting from column k. */
static int augment(int k, const int *Ap, const int *Ai, int *Match, int *Cheap,
                   int *Flag, int *Istack, int *Jstack, int *Pstack) {
    int found = 0, head = 0, i = EMPTY, j, p, pend, pstart, j2;
    Jstack[0] = k;
    while (head >= 0) {
        j = Jstack[head];
        pend = Ap[j + 1];
        if (Flag[j] != k) {
            /* first visit of column j in this search: try a cheap assignment */
            Flag[j] = k;
            for (p = Cheap[j]; p < pend && !found; p++) {
                i = Ai[p];
                found = (Match[i] == EMPTY);
            }
            Cheap[j] = p;
            if (found) {
                Istack[head] = i;
                break;
            }
            Pstack[head] = Ap[j];
        }
        pstart = Pstack[head];
        for (p = pstart; p < pend; p++) {
            i = Ai[p];
            j2 = Match[i];
            if (Flag[j2] != k) {
                Pstack[head] = p + 1;
                Istack[head] = i;
                Jstack[++head] = j2;
                break;
            }
        }
        if (p == pend) head--;
    }
    if (found) {
        for (p = head; p >= 0; p--) {
            j = Jstack[p];
            i = Istack[p];
            Match[i] = j;
        }
    }
    return found;
}

int ko_btf_maxtrans(int nrow, int ncol, const int *Ap, const int *Ai,
                    int *Match, int *Work) {
    int *Cheap = Work, *Flag = Work + ncol, *Istack = Work + 2 * ncol,
        *Jstack = Work + 3 * ncol, *Pstack = Work + 4 * ncol;
    int i, j, k, nmatch = 0;
    for (j = 0; j < ncol; j++) {
        Cheap[j] = Ap[j];
        Flag[j] = EMPTY;
    }
    for (i = 0; i < nrow; i++) Match[i] = EMPTY;
    for (k = 0; k < ncol; k++)
        if (augment(k, Ap, Ai, Match, Cheap, Flag, Istack, Jstack, Pstack))
            nmatch++;
    return nmatch;
}

#define UNVISITED (-2)
#define UNASSIGNED (-1)

static void scc_dfs(int j, const int *Ap, const int *Ai, const int *Q,
                    int *Time, int *Flag, int *Low, int *nblocks,
                    int *timestamp, int *Cstack, int *Jstack, int *Pstack) {
    int chead = 0, jhead = 0, i, p, pend, jj, parent;
    Jstack[0] = j;
    while (jhead >= 0) {
        j = Jstack[jhead];
        jj = (Q == 0) ? j : UNFLIP(Q[j]);
        pend = Ap[jj + 1];
        if (Flag[j] == UNVISITED) {
            Cstack[++chead] = j;
            (*timestamp)++;
            Time[j] = *timestamp;
            Low[j] = *timestamp;
            Flag[j] = UNASSIGNED;
            Pstack[jhead] = Ap[jj];
        }
        for (p = Pstack[jhead]; p < pend; p++) {
            i = Ai[p];
            if (Flag[i] == UNVISITED) {
                Pstack[jhead] = p + 1;
                Jstack[++jhead] = i;
                break;
            } else if (Flag[i] == UNASSIGNED) {
                Low[j] = MINI(Low[j], Time[i]);
            }
        }
        if (p == pend) {
            jhead--;
            if (Low[j] == Time[j]) {
                for (;;) {
                    i = Cstack[chead--];
                    Flag[i] = *nblocks;
                    if (i == j) break;
                }
                (*nblocks)++;
            }
            if (jhead >= 0) {
                parent = Jstack[jhead];
                Low[parent] = MINI(Low[parent], Low[j]);
            }
        }
    }
}

/* Work is 4n: Time, Flag, (Jstack|Low share per BTF: Low, Pstack), ...      */
#include <stdlib.h>
int ko_btf_strongcomp(int n, const int *Ap, const int *Ai, int *Q, int *P,
                      int *R, int *Work) {
    int *Time = Work, *Flag = Work + n, *Low = P, *Cstack = R;
    int *Jstack = Work + 2 * n, *Pstack = Work + 3 * n;
    int j, k, b, nblocks = 0, timestamp = 0;
    /* Cstack needs n+1 slots: R has n+1. Low uses P as scratch (as BTF does). */
    for (j = 0; j < n; j++) {
        Flag[j] = UNVISITED;
        Low[j] = EMPTY;
        Time[j] = EMPTY;
    }
    for (j = 0; j < n; j++)
        if (Flag[j] == UNVISITED)
            scc_dfs(j, Ap, Ai, Q, Time, Flag, Low, &nblocks, &timestamp,
                    Cstack, Jstack, Pstack);
    /* Flag[j] = block of node j; build R and P */
    for (b = 0; b < nblocks; b++) R[b] = 0;
    for (j = 0; j < n; j++) R[Flag[j]]++;
    Time[0] = 0;
    for (b = 1; b < nblocks; b++) Time[b] = Time[b - 1] + R[b - 1];
    for (b = 0; b < nblocks; b++) R[b] = Time[b];
    R[nblocks] = n;
    for (j = 0; j < n; j++) P[Time[Flag[j]]++] = j;
    if (Q != 0) {
        /* Q = Q*P' : column permutation that pairs with the symmetric P */
        for (k = 0; k < n; k++) Time[k] = Q[P[k]];
        for (k = 0; k < n; k++) Q[k] = Time[k];
    }
    return nblocks;
}

int ko_btf_order(int n, const int *Ap, const int *Ai, int *P, int *Q, int *R,
                 int *nmatch, int *Work) {
    int i, j, nbadcol;
    *nmatch = ko_btf_maxtrans(n, n, Ap, Ai, Q, Work);
    if (*nmatch < n) {
        /* complete the permutation for a structurally singular matrix:
         * unmatched rows get the unmatched columns, flagged by FLIP */
        int *Flag = Work + n;
        for (j = 0; j < n; j++) Flag[j] = 0;
        for (i = 0; i < n; i++) {
            j = Q[i];
            if (j != EMPTY) Flag[j] = 1;
        }
        nbadcol = 0;
        for (j = n - 1; j >= 0; j--)
            if (!Flag[j]) Work[nbadcol++] = j;
        for (i = 0; i < n; i++) {
            if (Q[i] == EMPTY && nbadcol > 0) {
                j = Work[--nbadcol];
                Q[i] = FLIP(j);
            }
        }
    }
    return ko_btf_strongcomp(n, Ap, Ai, Q, P, R, Work);
}
