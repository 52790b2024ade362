/*
 * amd_oracle.c — TEST INFRASTRUCTURE ONLY (see klu_oracle.h).
 *
 * Restatement of SuiteSparse AMD's amd_order with default Control
 * (dense = 10, aggressive absorption on), as klu_analyze calls it per BTF
 * block of size > 3 (KLU's analyze_worker; reached from
 * /root/reference/klujax.cpp:296 and :1366 via klu_analyze).
 * Pipeline restated: amd_valid -> (amd_preprocess if jumbled) -> amd_aat ->
 * amd_1 (build A+A') -> amd_2 (approximate minimum degree with quotient
 * graph, element absorption, supervariable hashing, mass elimination) ->
 * amd_postorder.  Tie-breaking follows the published algorithm: degree lists
 * are LIFO, hash buckets thread through Next/Last, children are postordered
 * with the biggest child last.
 */
#include <math.h>
#include <stdlib.h>
#include <limits.h>
#include "klu_oracle.h"

#define EMPTY (-1)
#define FLIP(i) (-(i) - 2)
#define MAXI(a, b) (((a) > (b)) ? (a) : (b))
#define MINI(a, b) (((a) < (b)) ? (a) : (b))

static int clear_flag(int wflg, int wbig, int *W, int n) {
    int x;
    if (wflg < 2 || wflg >= wbig) {
        for (x = 0; x < n; x++)
            if (W[x] != 0) W[x] = 1;
        wflg = 2;
    }
    return wflg;
}

static int post_tree(int root, int k, int *Child, const int *Sibling,
                     int *Order, int *Stack) {
    int f, head = 0, h, i;
    Stack[0] = root;
    while (head >= 0) {
        i = Stack[head];
        if (Child[i] != EMPTY) {
            for (f = Child[i]; f != EMPTY; f = Sibling[f]) head++;
            h = head;
            for (f = Child[i]; f != EMPTY; f = Sibling[f]) Stack[h--] = f;
            Child[i] = EMPTY;
        } else {
            head--;
            Order[i] = k++;
        }
    }
    return k;
}

static void postorder(int nn, const int *Parent, const int *Nv,
                      const int *Fsize, int *Order, int *Child, int *Sibling,
                      int *Stack) {
    int i, j, k, parent, frsize, f, fprev, maxfrsize, bigfprev, bigf, fnext;
    for (j = 0; j < nn; j++) {
        Child[j] = EMPTY;
        Sibling[j] = EMPTY;
    }
    for (j = nn - 1; j >= 0; j--) {
        if (Nv[j] > 0) {
            parent = Parent[j];
            if (parent != EMPTY) {
                Sibling[j] = Child[parent];
                Child[parent] = j;
            }
        }
    }
    /* biggest child goes last in each child list */
    for (i = 0; i < nn; i++) {
        if (Nv[i] > 0 && Child[i] != EMPTY) {
            fprev = EMPTY;
            maxfrsize = EMPTY;
            bigfprev = EMPTY;
            bigf = EMPTY;
            for (f = Child[i]; f != EMPTY; f = Sibling[f]) {
                frsize = Fsize[f];
                if (frsize >= maxfrsize) {
                    maxfrsize = frsize;
                    bigfprev = fprev;
                    bigf = f;
                }
                fprev = f;
            }
            fnext = Sibling[bigf];
            if (fnext != EMPTY) {
                if (bigfprev == EMPTY)
                    Child[i] = fnext;
                else
                    Sibling[bigfprev] = fnext;
                Sibling[bigf] = EMPTY;
                Sibling[fprev] = bigf;
            }
        }
    }
    for (i = 0; i < nn; i++) Order[i] = EMPTY;
    k = 0;
    for (i = 0; i < nn; i++)
        if (Parent[i] == EMPTY && Nv[i] > 0)
            k = post_tree(i, k, Child, Sibling, Order, Stack);
}

/* The approximate-minimum-degree core. Next/Last double as output Pinv/P.   */
static void amd2(int n, int *Pe, int *Iw, int *Len, int iwlen, int pfree,
                 int *Nv, int *Next, int *Last, int *Head, int *Elen,
                 int *Degree, int *W, double *lnz_out) {
    int deg, degme, dext, lemax, e, elenme, eln, i, ilast, inext, j, jlast,
        jnext, k, knt1, knt2, knt3, lenj, ln, me, mindeg, nel, nleft, nvi, nvj,
        nvpiv, slenme, wbig, we, wflg, wnvi, ok, ndense, dense, p, p1, p2, p3,
        p4, pdst, pend, pj, pme, pme1, pme2, pn, psrc;
    unsigned int hash;
    double f, r, lnz = 0, lnzme;

    me = EMPTY;
    mindeg = 0;
    nel = 0;
    lemax = 0;
    dense = (int)(10.0 * sqrt((double)n));
    dense = MAXI(16, dense);
    dense = MINI(n, dense);

    for (i = 0; i < n; i++) {
        Last[i] = EMPTY;
        Head[i] = EMPTY;
        Next[i] = EMPTY;
        Nv[i] = 1;
        W[i] = 1;
        Elen[i] = 0;
        Degree[i] = Len[i];
    }
    wbig = INT_MAX - n;
    wflg = clear_flag(0, wbig, W, n);

    ndense = 0;
    for (i = 0; i < n; i++) {
        deg = Degree[i];
        if (deg == 0) {
            Elen[i] = FLIP(1);
            nel++;
            Pe[i] = EMPTY;
            W[i] = 0;
        } else if (deg > dense) {
            ndense++;
            Nv[i] = 0;
            Elen[i] = EMPTY;
            nel++;
            Pe[i] = EMPTY;
        } else {
            inext = Head[deg];
            if (inext != EMPTY) Last[inext] = i;
            Next[i] = inext;
            Head[deg] = i;
        }
    }

    while (nel < n) {
        /* ---- pick the pivot of minimum approximate degree ---- */
        for (deg = mindeg; deg < n; deg++) {
            me = Head[deg];
            if (me != EMPTY) break;
        }
        mindeg = deg;
        inext = Next[me];
        if (inext != EMPTY) Last[inext] = EMPTY;
        Head[deg] = inext;

        elenme = Elen[me];
        nvpiv = Nv[me];
        nel += nvpiv;

        /* ---- construct the new element ---- */
        Nv[me] = -nvpiv;
        degme = 0;
        if (elenme == 0) {
            pme1 = Pe[me];
            pme2 = pme1 - 1;
            for (p = pme1; p <= pme1 + Len[me] - 1; p++) {
                i = Iw[p];
                nvi = Nv[i];
                if (nvi > 0) {
                    degme += nvi;
                    Nv[i] = -nvi;
                    Iw[++pme2] = i;
                    ilast = Last[i];
                    inext = Next[i];
                    if (inext != EMPTY) Last[inext] = ilast;
                    if (ilast != EMPTY)
                        Next[ilast] = inext;
                    else
                        Head[Degree[i]] = inext;
                }
            }
        } else {
            p = Pe[me];
            pme1 = pfree;
            slenme = Len[me] - elenme;
            for (knt1 = 1; knt1 <= elenme + 1; knt1++) {
                if (knt1 > elenme) {
                    e = me;
                    pj = p;
                    ln = slenme;
                } else {
                    e = Iw[p++];
                    pj = Pe[e];
                    ln = Len[e];
                }
                for (knt2 = 1; knt2 <= ln; knt2++) {
                    i = Iw[pj++];
                    nvi = Nv[i];
                    if (nvi > 0) {
                        if (pfree >= iwlen) {
                            /* garbage-collect Iw */
                            Pe[me] = p;
                            Len[me] -= knt1;
                            if (Len[me] == 0) Pe[me] = EMPTY;
                            Pe[e] = pj;
                            Len[e] = ln - knt2;
                            if (Len[e] == 0) Pe[e] = EMPTY;
                            for (j = 0; j < n; j++) {
                                pn = Pe[j];
                                if (pn >= 0) {
                                    Pe[j] = Iw[pn];
                                    Iw[pn] = FLIP(j);
                                }
                            }
                            psrc = 0;
                            pdst = 0;
                            pend = pme1 - 1;
                            while (psrc <= pend) {
                                j = FLIP(Iw[psrc++]);
                                if (j >= 0) {
                                    Iw[pdst] = Pe[j];
                                    Pe[j] = pdst++;
                                    lenj = Len[j];
                                    for (knt3 = 0; knt3 <= lenj - 2; knt3++)
                                        Iw[pdst++] = Iw[psrc++];
                                }
                            }
                            p1 = pdst;
                            for (psrc = pme1; psrc <= pfree - 1; psrc++)
                                Iw[pdst++] = Iw[psrc];
                            pme1 = p1;
                            pfree = pdst;
                            pj = Pe[e];
                            p = Pe[me];
                        }
                        degme += nvi;
                        Nv[i] = -nvi;
                        Iw[pfree++] = i;
                        ilast = Last[i];
                        inext = Next[i];
                        if (inext != EMPTY) Last[inext] = ilast;
                        if (ilast != EMPTY)
                            Next[ilast] = inext;
                        else
                            Head[Degree[i]] = inext;
                    }
                }
                if (e != me) {
                    Pe[e] = FLIP(me);
                    W[e] = 0;
                }
            }
            pme2 = pfree - 1;
        }
        Degree[me] = degme;
        Pe[me] = pme1;
        Len[me] = pme2 - pme1 + 1;
        Elen[me] = FLIP(nvpiv + degme);

        wflg = clear_flag(wflg, wbig, W, n);

        /* ---- scan 1: W[e]-wflg = |Le \ Lme| for every element touched ---- */
        for (pme = pme1; pme <= pme2; pme++) {
            i = Iw[pme];
            eln = Elen[i];
            if (eln > 0) {
                nvi = -Nv[i];
                wnvi = wflg - nvi;
                for (p = Pe[i]; p <= Pe[i] + eln - 1; p++) {
                    e = Iw[p];
                    we = W[e];
                    if (we >= wflg)
                        we -= nvi;
                    else if (we != 0)
                        we = Degree[e] + wnvi;
                    W[e] = we;
                }
            }
        }

        /* ---- scan 2: degree update, aggressive absorption, hashing ---- */
        for (pme = pme1; pme <= pme2; pme++) {
            i = Iw[pme];
            p1 = Pe[i];
            p2 = p1 + Elen[i] - 1;
            pn = p1;
            hash = 0;
            deg = 0;
            for (p = p1; p <= p2; p++) {
                e = Iw[p];
                we = W[e];
                if (we != 0) {
                    dext = we - wflg;
                    if (dext > 0) {
                        deg += dext;
                        Iw[pn++] = e;
                        hash += (unsigned int)e;
                    } else {
                        Pe[e] = FLIP(me);
                        W[e] = 0;
                    }
                }
            }
            Elen[i] = pn - p1 + 1;
            p3 = pn;
            p4 = p1 + Len[i];
            for (p = p2 + 1; p < p4; p++) {
                j = Iw[p];
                nvj = Nv[j];
                if (nvj > 0) {
                    deg += nvj;
                    Iw[pn++] = j;
                    hash += (unsigned int)j;
                }
            }
            if (Elen[i] == 1 && p3 == pn) {
                /* mass elimination */
                Pe[i] = FLIP(me);
                nvi = -Nv[i];
                degme -= nvi;
                nvpiv += nvi;
                nel += nvi;
                Nv[i] = 0;
                Elen[i] = EMPTY;
            } else {
                Degree[i] = MINI(Degree[i], deg);
                Iw[pn] = Iw[p3];
                Iw[p3] = Iw[p1];
                Iw[p1] = me;
                Len[i] = pn - p1 + 1;
                hash = hash % (unsigned int)n;
                j = Head[hash];
                if (j <= EMPTY) {
                    Next[i] = FLIP(j);
                    Head[hash] = FLIP(i);
                } else {
                    Next[i] = Last[j];
                    Last[j] = i;
                }
                Last[i] = (int)hash;
            }
        }
        Degree[me] = degme;

        lemax = MAXI(lemax, degme);
        wflg += lemax;
        wflg = clear_flag(wflg, wbig, W, n);

        /* ---- supervariable detection ---- */
        for (pme = pme1; pme <= pme2; pme++) {
            i = Iw[pme];
            if (Nv[i] < 0) {
                hash = (unsigned int)Last[i];
                j = Head[hash];
                if (j == EMPTY) {
                    i = EMPTY;
                } else if (j < EMPTY) {
                    i = FLIP(j);
                    Head[hash] = EMPTY;
                } else {
                    i = Last[j];
                    Last[j] = EMPTY;
                }
                while (i != EMPTY && Next[i] != EMPTY) {
                    ln = Len[i];
                    eln = Elen[i];
                    for (p = Pe[i] + 1; p <= Pe[i] + ln - 1; p++)
                        W[Iw[p]] = wflg;
                    jlast = i;
                    j = Next[i];
                    while (j != EMPTY) {
                        ok = (Len[j] == ln) && (Elen[j] == eln);
                        for (p = Pe[j] + 1; ok && p <= Pe[j] + ln - 1; p++)
                            if (W[Iw[p]] != wflg) ok = 0;
                        if (ok) {
                            Pe[j] = FLIP(i);
                            Nv[i] += Nv[j];
                            Nv[j] = 0;
                            Elen[j] = EMPTY;
                            j = Next[j];
                            Next[jlast] = j;
                        } else {
                            jlast = j;
                            j = Next[j];
                        }
                    }
                    wflg++;
                    i = Next[i];
                }
            }
        }

        /* ---- restore degree lists, drop non-principal vars from Lme ---- */
        p = pme1;
        nleft = n - nel;
        for (pme = pme1; pme <= pme2; pme++) {
            i = Iw[pme];
            nvi = -Nv[i];
            if (nvi > 0) {
                Nv[i] = nvi;
                deg = Degree[i] + degme - nvi;
                deg = MINI(deg, nleft - nvi);
                inext = Head[deg];
                if (inext != EMPTY) Last[inext] = i;
                Next[i] = inext;
                Last[i] = EMPTY;
                Head[deg] = i;
                mindeg = MINI(mindeg, deg);
                Degree[i] = deg;
                Iw[p++] = i;
            }
        }

        /* ---- finalize the new element ---- */
        Nv[me] = nvpiv;
        Len[me] = p - pme1;
        if (Len[me] == 0) {
            Pe[me] = EMPTY;
            W[me] = 0;
        }
        if (elenme != 0) pfree = p;

        f = nvpiv;
        r = degme + ndense;
        lnzme = f * r + (f - 1) * f / 2;
        lnz += lnzme;
    }

    f = ndense;
    lnzme = (f - 1) * f / 2;
    lnz += lnzme;
    if (lnz_out) *lnz_out = lnz;

    /* ---- postordering ---- */
    for (i = 0; i < n; i++) Pe[i] = FLIP(Pe[i]);
    for (i = 0; i < n; i++) Elen[i] = FLIP(Elen[i]);
    for (i = 0; i < n; i++) {
        if (Nv[i] == 0) {
            j = Pe[i];
            if (j == EMPTY) continue; /* dense variable: no parent */
            while (Nv[j] == 0) j = Pe[j];
            e = j;
            j = i;
            while (Nv[j] == 0) {
                jnext = Pe[j];
                Pe[j] = e;
                j = jnext;
            }
        }
    }
    postorder(n, Pe, Nv, Elen, W, Head, Next, Last);

    for (k = 0; k < n; k++) {
        Head[k] = EMPTY;
        Next[k] = EMPTY;
    }
    for (e = 0; e < n; e++) {
        k = W[e];
        if (k != EMPTY) Head[k] = e;
    }
    nel = 0;
    for (k = 0; k < n; k++) {
        e = Head[k];
        if (e == EMPTY) break;
        Next[e] = nel;
        nel += Nv[e];
    }
    for (i = 0; i < n; i++) {
        if (Nv[i] == 0) {
            e = Pe[i];
            if (e != EMPTY) {
                Next[i] = Next[e];
                Next[e]++;
            } else {
                Next[i] = nel++;
            }
        }
    }
    for (i = 0; i < n; i++) {
        k = Next[i];
        Last[k] = i;
    }
}

/* amd_aat: degrees of A+A' (no diagonal); Tp is size-n scratch */
static long aat_len(int n, const int *Ap, const int *Ai, int *Len, int *Tp) {
    int p1, p2, p, i, j, pj, pj2, k;
    long nzaat = 0;
    for (k = 0; k < n; k++) Len[k] = 0;
    for (k = 0; k < n; k++) {
        p1 = Ap[k];
        p2 = Ap[k + 1];
        for (p = p1; p < p2;) {
            j = Ai[p];
            if (j < k) {
                Len[j]++;
                Len[k]++;
                p++;
            } else if (j == k) {
                p++;
                break;
            } else {
                break;
            }
            pj2 = Ap[j + 1];
            for (pj = Tp[j]; pj < pj2;) {
                i = Ai[pj];
                if (i < k) {
                    Len[i]++;
                    Len[j]++;
                    pj++;
                } else if (i == k) {
                    pj++;
                    break;
                } else {
                    break;
                }
            }
            Tp[j] = pj;
        }
        Tp[k] = p;
    }
    for (j = 0; j < n; j++) {
        for (pj = Tp[j]; pj < Ap[j + 1]; pj++) {
            i = Ai[pj];
            Len[i]++;
            Len[j]++;
        }
    }
    for (k = 0; k < n; k++) nzaat += Len[k];
    return nzaat;
}

int ko_amd_order(int n, const int *Ap, const int *Ai, int *P, double *lnz_out) {
    int i, j, k, p, p1, p2, pj, pj2, pfree, ilast, jumbled = 0, status = 0;
    int *Rp = 0, *Ri = 0, *Len, *Pinv, *S, *Pe, *Nv, *Head, *Elen, *Degree, *W,
        *Iw, *Sp, *Tp;
    const int *Cp, *Ci;
    long nzaat, slen;
    int iwlen;

    if (lnz_out) *lnz_out = 0;
    if (n < 0) return -2;
    if (n == 0) return 0;
    /* amd_valid: sorted columns without duplicates, else "jumbled" */
    for (j = 0; j < n; j++) {
        p1 = Ap[j];
        p2 = Ap[j + 1];
        if (p1 > p2) return -2;
        ilast = EMPTY;
        for (p = p1; p < p2; p++) {
            i = Ai[p];
            if (i < 0 || i >= n) return -2;
            if (i <= ilast) jumbled = 1;
            ilast = i;
        }
    }
    Len = (int *)malloc(sizeof(int) * (size_t)n);
    Pinv = (int *)malloc(sizeof(int) * (size_t)n);
    if (jumbled) {
        /* amd_preprocess: R = A' with sorted, duplicate-free columns */
        int *Wk = Len, *Flag = Pinv;
        int nz = Ap[n];
        status = 1;
        Rp = (int *)malloc(sizeof(int) * (size_t)(n + 1));
        Ri = (int *)malloc(sizeof(int) * (size_t)(nz > 0 ? nz : 1));
        for (i = 0; i < n; i++) {
            Wk[i] = 0;
            Flag[i] = EMPTY;
        }
        for (j = 0; j < n; j++) {
            for (p = Ap[j]; p < Ap[j + 1]; p++) {
                i = Ai[p];
                if (Flag[i] != j) {
                    Wk[i]++;
                    Flag[i] = j;
                }
            }
        }
        Rp[0] = 0;
        for (i = 0; i < n; i++) Rp[i + 1] = Rp[i] + Wk[i];
        for (i = 0; i < n; i++) {
            Wk[i] = Rp[i];
            Flag[i] = EMPTY;
        }
        for (j = 0; j < n; j++) {
            for (p = Ap[j]; p < Ap[j + 1]; p++) {
                i = Ai[p];
                if (Flag[i] != j) {
                    Ri[Wk[i]++] = j;
                    Flag[i] = j;
                }
            }
        }
        Cp = Rp;
        Ci = Ri;
    } else {
        Cp = Ap;
        Ci = Ai;
    }
    nzaat = aat_len(n, Cp, Ci, Len, P);
    slen = nzaat + nzaat / 5 + 7L * n;
    S = (int *)malloc(sizeof(int) * (size_t)slen);
    iwlen = (int)(slen - 6L * n);
    Pe = S;
    Nv = Pe + n;
    Head = Nv + n;
    Elen = Head + n;
    Degree = Elen + n;
    W = Degree + n;
    Iw = W + n;

    /* amd_1: build the pattern of A+A' in Iw */
    Sp = Nv;
    Tp = W;
    pfree = 0;
    for (j = 0; j < n; j++) {
        Pe[j] = pfree;
        Sp[j] = pfree;
        pfree += Len[j];
    }
    for (k = 0; k < n; k++) {
        p1 = Cp[k];
        p2 = Cp[k + 1];
        for (p = p1; p < p2;) {
            j = Ci[p];
            if (j < k) {
                Iw[Sp[j]++] = k;
                Iw[Sp[k]++] = j;
                p++;
            } else if (j == k) {
                p++;
                break;
            } else {
                break;
            }
            pj2 = Cp[j + 1];
            for (pj = Tp[j]; pj < pj2;) {
                i = Ci[pj];
                if (i < k) {
                    Iw[Sp[i]++] = j;
                    Iw[Sp[j]++] = i;
                    pj++;
                } else if (i == k) {
                    pj++;
                    break;
                } else {
                    break;
                }
            }
            Tp[j] = pj;
        }
        Tp[k] = p;
    }
    for (j = 0; j < n; j++) {
        for (pj = Tp[j]; pj < Cp[j + 1]; pj++) {
            i = Ci[pj];
            Iw[Sp[i]++] = j;
            Iw[Sp[j]++] = i;
        }
    }
    amd2(n, Pe, Iw, Len, iwlen, pfree, Nv, Pinv, P, Head, Elen, Degree, W,
         lnz_out);
    free(S);
    free(Len);
    free(Pinv);
    free(Rp);
    free(Ri);
    return status;
}
