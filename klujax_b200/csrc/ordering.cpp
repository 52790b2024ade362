// ordering.cpp — host symbolic analysis for the B200 path (product code).
//
// Produces the same P / Q / R that KLU's klu_analyze gives klujax
// (/root/reference/klujax.cpp:1362-1366: coo_to_csc_analyze + klu_defaults +
// klu_analyze => BTF then AMD per diagonal block).  The permutations decide the
// static elimination schedule, so the traversal orders of the published
// algorithms (Duff MC21 + cheap assignment, Tarjan SCC, Amestoy/Davis/Duff AMD)
// are kept: rows are visited in stored column order, degree lists are LIFO,
// supervariables are found by hashing, the assembly tree is postordered with
// the largest child last.
#include <algorithm>
#include <climits>
#include <cmath>
#include <numeric>

#include "symbolic.hpp"

namespace kb {

namespace {
constexpr int NONE = -1;
inline int flip(int i) { return -i - 2; }
inline int unflip(int i) { return i < NONE ? flip(i) : i; }
}  // namespace

// ---------------------------------------------------------------------------
// COO -> CSC (stable by arrival order inside a column), klujax.cpp:86-125
// ---------------------------------------------------------------------------
CscPattern coo_to_csc(int n, int nz, const int* Ai, const int* Aj) {
    CscPattern A;
    A.n = n;
    A.colptr.assign(n + 1, 0);
    A.rowidx.resize(nz);
    A.coo_of.resize(nz);
    for (int e = 0; e < nz; ++e) A.colptr[Aj[e] + 1]++;
    for (int j = 0; j < n; ++j) A.colptr[j + 1] += A.colptr[j];
    std::vector<int> fill(A.colptr.begin(), A.colptr.end() - 1);
    for (int e = 0; e < nz; ++e) {
        int dst = fill[Aj[e]]++;
        A.rowidx[dst] = Ai[e];
        A.coo_of[dst] = e;
    }
    std::vector<int> seen(n, NONE);
    for (int j = 0; j < n && !A.has_duplicates; ++j)
        for (int p = A.colptr[j]; p < A.colptr[j + 1]; ++p) {
            if (seen[A.rowidx[p]] == j) {
                A.has_duplicates = true;
                break;
            }
            seen[A.rowidx[p]] = j;
        }
    return A;
}

// ---------------------------------------------------------------------------
// Maximum transversal: one DFS for an augmenting path per column.
// match[row] = column.
// ---------------------------------------------------------------------------
int maximum_transversal(int n, const std::vector<int>& colptr,
                        const std::vector<int>& rowidx, std::vector<int>& match) {
    match.assign(n, NONE);
    std::vector<int> cheap(colptr.begin(), colptr.begin() + n), visited(n, NONE);
    std::vector<int> rowstk(n), colstk(n), posstk(n);
    int nmatch = 0;
    for (int k = 0; k < n; ++k) {
        int top = 0, hit_row = NONE;
        bool found = false;
        colstk[0] = k;
        while (top >= 0) {
            const int j = colstk[top];
            const int end = colptr[j + 1];
            if (visited[j] != k) {
                visited[j] = k;
                int p = cheap[j];
                while (p < end && !found) {
                    hit_row = rowidx[p++];
                    found = (match[hit_row] == NONE);
                }
                cheap[j] = p;
                if (found) {
                    rowstk[top] = hit_row;
                    break;
                }
                posstk[top] = colptr[j];
            }
            int p = posstk[top];
            for (; p < end; ++p) {
                const int i = rowidx[p];
                const int j2 = match[i];
                if (visited[j2] != k) {
                    posstk[top] = p + 1;
                    rowstk[top] = i;
                    colstk[++top] = j2;
                    break;
                }
            }
            if (p == end) --top;
        }
        if (found) {
            for (int s = top; s >= 0; --s) match[rowstk[s]] = colstk[s];
            ++nmatch;
        }
    }
    return nmatch;
}

// ---------------------------------------------------------------------------
// Strongly connected components of the graph of A*Q (iterative Tarjan).
// On return P is the symmetric block permutation, Q is composed with it and
// R holds the block boundaries (block upper triangular form).
// ---------------------------------------------------------------------------
int strong_components(int n, const std::vector<int>& colptr,
                      const std::vector<int>& rowidx, std::vector<int>& Q,
                      std::vector<int>& P, std::vector<int>& R) {
    constexpr int UNVISITED = -2, OPEN = -1;
    std::vector<int> when(n, NONE), low(n, NONE), comp(n, UNVISITED);
    std::vector<int> open_stack(n + 1), path(n), pos(n);
    int clock = 0, nblocks = 0;
    for (int root = 0; root < n; ++root) {
        if (comp[root] != UNVISITED) continue;
        int otop = 0, ptop = 0;
        path[0] = root;
        while (ptop >= 0) {
            const int j = path[ptop];
            const int col = unflip(Q[j]);
            const int end = colptr[col + 1];
            if (comp[j] == UNVISITED) {
                open_stack[++otop] = j;
                when[j] = low[j] = ++clock;
                comp[j] = OPEN;
                pos[ptop] = colptr[col];
            }
            int p = pos[ptop];
            for (; p < end; ++p) {
                const int i = rowidx[p];
                if (comp[i] == UNVISITED) {
                    pos[ptop] = p + 1;
                    path[++ptop] = i;
                    break;
                }
                if (comp[i] == OPEN) low[j] = std::min(low[j], when[i]);
            }
            if (p == end) {
                --ptop;
                if (low[j] == when[j]) {
                    int i;
                    do {
                        i = open_stack[otop--];
                        comp[i] = nblocks;
                    } while (i != j);
                    ++nblocks;
                }
                if (ptop >= 0) low[path[ptop]] = std::min(low[path[ptop]], low[j]);
            }
        }
    }
    R.assign(n + 1, 0);
    std::vector<int> start(nblocks + 1, 0);
    for (int j = 0; j < n; ++j) start[comp[j] + 1]++;
    for (int b = 0; b < nblocks; ++b) start[b + 1] += start[b];
    for (int b = 0; b <= nblocks; ++b) R[b] = start[b];
    P.assign(n, 0);
    for (int j = 0; j < n; ++j) P[start[comp[j]]++] = j;
    std::vector<int> q2(n);
    for (int k = 0; k < n; ++k) q2[k] = Q[P[k]];
    Q.swap(q2);
    return nblocks;
}

// ---------------------------------------------------------------------------
// AMD
// ---------------------------------------------------------------------------
namespace {

struct AmdState {
    int n;
    std::vector<int> pe, nv, head, elen, degree, w, len, next, last;
    std::vector<int> iw;
    int iwlen = 0, pfree = 0;
};

int reset_marks(int wflg, int wbig, std::vector<int>& w) {
    if (wflg < 2 || wflg >= wbig) {
        for (int& x : w)
            if (x != 0) x = 1;
        wflg = 2;
    }
    return wflg;
}

// postorder of the assembly tree; fills order[e] = position (NONE for non-elements)
void assembly_postorder(int n, const std::vector<int>& parent, const std::vector<int>& nv,
                        const std::vector<int>& fsize, std::vector<int>& order) {
    std::vector<int> child(n, NONE), sibling(n, NONE), stack(n);
    for (int j = n - 1; j >= 0; --j)
        if (nv[j] > 0 && parent[j] != NONE) {
            sibling[j] = child[parent[j]];
            child[parent[j]] = j;
        }
    for (int i = 0; i < n; ++i) {
        if (nv[i] <= 0 || child[i] == NONE) continue;
        int prev = NONE, best = NONE, best_prev = NONE, best_size = NONE;
        for (int f = child[i]; f != NONE; f = sibling[f]) {
            if (fsize[f] >= best_size) {
                best_size = fsize[f];
                best_prev = prev;
                best = f;
            }
            prev = f;
        }
        const int after = sibling[best];
        if (after != NONE) {
            if (best_prev == NONE)
                child[i] = after;
            else
                sibling[best_prev] = after;
            sibling[best] = NONE;
            sibling[prev] = best;
        }
    }
    order.assign(n, NONE);
    int k = 0;
    for (int root = 0; root < n; ++root) {
        if (parent[root] != NONE || nv[root] <= 0) continue;
        int top = 0;
        stack[0] = root;
        while (top >= 0) {
            const int i = stack[top];
            if (child[i] != NONE) {
                for (int f = child[i]; f != NONE; f = sibling[f]) ++top;
                int h = top;
                for (int f = child[i]; f != NONE; f = sibling[f]) stack[h--] = f;
                child[i] = NONE;
            } else {
                --top;
                order[i] = k++;
            }
        }
    }
}

void amd_core(AmdState& s, std::vector<int>& perm, double* lnz_out) {
    const int n = s.n;
    auto &pe = s.pe, &nv = s.nv, &head = s.head, &elen = s.elen, &degree = s.degree, &w = s.w,
         &len = s.len, &next = s.next, &last = s.last, &iw = s.iw;
    int pfree = s.pfree;
    const int iwlen = s.iwlen;
    int dense = static_cast<int>(10.0 * std::sqrt(static_cast<double>(n)));
    dense = std::min(n, std::max(16, dense));
    double lnz = 0;

    for (int i = 0; i < n; ++i) {
        last[i] = head[i] = next[i] = NONE;
        nv[i] = 1;
        w[i] = 1;
        elen[i] = 0;
        degree[i] = len[i];
    }
    const int wbig = INT_MAX - n;
    int wflg = reset_marks(0, wbig, w);
    int nel = 0, ndense = 0, mindeg = 0, lemax = 0, me = NONE;

    auto unlink = [&](int i) {
        const int il = last[i], in = next[i];
        if (in != NONE) last[in] = il;
        if (il != NONE)
            next[il] = in;
        else
            head[degree[i]] = in;
    };

    for (int i = 0; i < n; ++i) {
        const int d = degree[i];
        if (d == 0) {
            elen[i] = flip(1);
            ++nel;
            pe[i] = NONE;
            w[i] = 0;
        } else if (d > dense) {
            ++ndense;
            nv[i] = 0;
            elen[i] = NONE;
            ++nel;
            pe[i] = NONE;
        } else {
            const int in = head[d];
            if (in != NONE) last[in] = i;
            next[i] = in;
            head[d] = i;
        }
    }

    while (nel < n) {
        int deg = mindeg;
        for (; deg < n; ++deg)
            if ((me = head[deg]) != NONE) break;
        mindeg = deg;
        {
            const int in = next[me];
            if (in != NONE) last[in] = NONE;
            head[deg] = in;
        }
        const int elenme = elen[me];
        int nvpiv = nv[me];
        nel += nvpiv;
        nv[me] = -nvpiv;
        int degme = 0, pme1, pme2;

        if (elenme == 0) {
            // element built in place over me's variable list
            pme1 = pe[me];
            pme2 = pme1 - 1;
            for (int p = pme1; p <= pme1 + len[me] - 1; ++p) {
                const int i = iw[p], nvi = nv[i];
                if (nvi > 0) {
                    degme += nvi;
                    nv[i] = -nvi;
                    iw[++pme2] = i;
                    unlink(i);
                }
            }
        } else {
            // element built in free space from me's elements + variables
            int p = pe[me];
            pme1 = pfree;
            const int slenme = len[me] - elenme;
            for (int knt1 = 1; knt1 <= elenme + 1; ++knt1) {
                int e, pj, ln;
                if (knt1 > elenme) {
                    e = me;
                    pj = p;
                    ln = slenme;
                } else {
                    e = iw[p++];
                    pj = pe[e];
                    ln = len[e];
                }
                for (int knt2 = 1; knt2 <= ln; ++knt2) {
                    const int i = iw[pj++], nvi = nv[i];
                    if (nvi <= 0) continue;
                    if (pfree >= iwlen) {
                        // compact the quotient-graph storage
                        pe[me] = p;
                        len[me] -= knt1;
                        if (len[me] == 0) pe[me] = NONE;
                        pe[e] = pj;
                        len[e] = ln - knt2;
                        if (len[e] == 0) pe[e] = NONE;
                        for (int j = 0; j < n; ++j) {
                            const int pn = pe[j];
                            if (pn >= 0) {
                                pe[j] = iw[pn];
                                iw[pn] = flip(j);
                            }
                        }
                        int src = 0, dst = 0;
                        const int stop = pme1 - 1;
                        while (src <= stop) {
                            const int j = flip(iw[src++]);
                            if (j >= 0) {
                                iw[dst] = pe[j];
                                pe[j] = dst++;
                                for (int c = 0; c <= len[j] - 2; ++c) iw[dst++] = iw[src++];
                            }
                        }
                        const int moved = dst;
                        for (src = pme1; src <= pfree - 1; ++src) iw[dst++] = iw[src];
                        pme1 = moved;
                        pfree = dst;
                        pj = pe[e];
                        p = pe[me];
                    }
                    degme += nvi;
                    nv[i] = -nvi;
                    iw[pfree++] = i;
                    unlink(i);
                }
                if (e != me) {
                    pe[e] = flip(me);
                    w[e] = 0;
                }
            }
            pme2 = pfree - 1;
        }
        degree[me] = degme;
        pe[me] = pme1;
        len[me] = pme2 - pme1 + 1;
        elen[me] = flip(nvpiv + degme);
        wflg = reset_marks(wflg, wbig, w);

        // |Le \ Lme| for every element adjacent to a variable of Lme
        for (int pme = pme1; pme <= pme2; ++pme) {
            const int i = iw[pme], eln = elen[i];
            if (eln <= 0) continue;
            const int nvi = -nv[i], wnvi = wflg - nvi;
            for (int p = pe[i]; p <= pe[i] + eln - 1; ++p) {
                const int e = iw[p];
                int we = w[e];
                if (we >= wflg)
                    we -= nvi;
                else if (we != 0)
                    we = degree[e] + wnvi;
                w[e] = we;
            }
        }

        // degree update, aggressive absorption, hash of the new adjacency
        for (int pme = pme1; pme <= pme2; ++pme) {
            const int i = iw[pme];
            const int p1 = pe[i], p2 = p1 + elen[i] - 1;
            int pn = p1, d = 0;
            unsigned int hash = 0;
            for (int p = p1; p <= p2; ++p) {
                const int e = iw[p], we = w[e];
                if (we == 0) continue;
                const int dext = we - wflg;
                if (dext > 0) {
                    d += dext;
                    iw[pn++] = e;
                    hash += static_cast<unsigned int>(e);
                } else {
                    pe[e] = flip(me);
                    w[e] = 0;
                }
            }
            elen[i] = pn - p1 + 1;
            const int p3 = pn, p4 = p1 + len[i];
            for (int p = p2 + 1; p < p4; ++p) {
                const int j = iw[p], nvj = nv[j];
                if (nvj > 0) {
                    d += nvj;
                    iw[pn++] = j;
                    hash += static_cast<unsigned int>(j);
                }
            }
            if (elen[i] == 1 && p3 == pn) {
                // mass elimination: i is indistinguishable from the pivot
                pe[i] = flip(me);
                const int nvi = -nv[i];
                degme -= nvi;
                nvpiv += nvi;
                nel += nvi;
                nv[i] = 0;
                elen[i] = NONE;
            } else {
                degree[i] = std::min(degree[i], d);
                iw[pn] = iw[p3];
                iw[p3] = iw[p1];
                iw[p1] = me;
                len[i] = pn - p1 + 1;
                hash %= static_cast<unsigned int>(n);
                const int j = head[hash];
                if (j <= NONE) {
                    next[i] = flip(j);
                    head[hash] = flip(i);
                } else {
                    next[i] = last[j];
                    last[j] = i;
                }
                last[i] = static_cast<int>(hash);
            }
        }
        degree[me] = degme;
        lemax = std::max(lemax, degme);
        wflg += lemax;
        wflg = reset_marks(wflg, wbig, w);

        // supervariable detection inside each hash bucket
        for (int pme = pme1; pme <= pme2; ++pme) {
            int i = iw[pme];
            if (nv[i] >= 0) continue;
            const int hash = last[i];
            const int j0 = head[hash];
            if (j0 == NONE) {
                i = NONE;
            } else if (j0 < NONE) {
                i = flip(j0);
                head[hash] = NONE;
            } else {
                i = last[j0];
                last[j0] = NONE;
            }
            while (i != NONE && next[i] != NONE) {
                const int ln = len[i], eln = elen[i];
                for (int p = pe[i] + 1; p <= pe[i] + ln - 1; ++p) w[iw[p]] = wflg;
                int jlast = i, j = next[i];
                while (j != NONE) {
                    bool same = (len[j] == ln) && (elen[j] == eln);
                    for (int p = pe[j] + 1; same && p <= pe[j] + ln - 1; ++p)
                        if (w[iw[p]] != wflg) same = false;
                    if (same) {
                        pe[j] = flip(i);
                        nv[i] += nv[j];
                        nv[j] = 0;
                        elen[j] = NONE;
                        j = next[j];
                        next[jlast] = j;
                    } else {
                        jlast = j;
                        j = next[j];
                    }
                }
                ++wflg;
                i = next[i];
            }
        }

        // put surviving principal variables back into the degree lists
        int p = pme1;
        const int nleft = n - nel;
        for (int pme = pme1; pme <= pme2; ++pme) {
            const int i = iw[pme], nvi = -nv[i];
            if (nvi <= 0) continue;
            nv[i] = nvi;
            int d = std::min(degree[i] + degme - nvi, nleft - nvi);
            const int in = head[d];
            if (in != NONE) last[in] = i;
            next[i] = in;
            last[i] = NONE;
            head[d] = i;
            mindeg = std::min(mindeg, d);
            degree[i] = d;
            iw[p++] = i;
        }
        nv[me] = nvpiv;
        len[me] = p - pme1;
        if (len[me] == 0) {
            pe[me] = NONE;
            w[me] = 0;
        }
        if (elenme != 0) pfree = p;
        const double f = nvpiv, r = degme + ndense;
        lnz += f * r + (f - 1) * f / 2;
    }
    {
        const double f = ndense;
        lnz += (f - 1) * f / 2;
    }
    if (lnz_out) *lnz_out = lnz;

    // assembly tree -> postorder -> permutation
    for (int i = 0; i < n; ++i) pe[i] = flip(pe[i]);
    for (int i = 0; i < n; ++i) elen[i] = flip(elen[i]);
    for (int i = 0; i < n; ++i) {
        if (nv[i] != 0) continue;
        int j = pe[i];
        if (j == NONE) continue;
        while (nv[j] == 0) j = pe[j];
        const int e = j;
        j = i;
        while (nv[j] == 0) {
            const int jn = pe[j];
            pe[j] = e;
            j = jn;
        }
    }
    std::vector<int> order;
    assembly_postorder(n, pe, nv, elen, order);
    std::vector<int> elem_at(n, NONE), place(n, NONE);
    for (int e = 0; e < n; ++e)
        if (order[e] != NONE) elem_at[order[e]] = e;
    int cnt = 0;
    for (int k = 0; k < n; ++k) {
        const int e = elem_at[k];
        if (e == NONE) break;
        place[e] = cnt;
        cnt += nv[e];
    }
    for (int i = 0; i < n; ++i) {
        if (nv[i] != 0) continue;
        const int e = pe[i];
        if (e != NONE) {
            place[i] = place[e];
            place[e]++;
        } else {
            place[i] = cnt++;
        }
    }
    perm.assign(n, 0);
    for (int i = 0; i < n; ++i) perm[place[i]] = i;
}

}  // namespace

bool amd_order(int n, const std::vector<int>& Ap, const std::vector<int>& Ai,
               std::vector<int>& perm, double* lnz) {
    if (lnz) *lnz = 0;
    perm.clear();
    if (n == 0) return true;
    // sorted, duplicate-free columns are used as they are; anything else goes
    // through a transpose that sorts and de-duplicates (AMD's preprocess step)
    bool tidy = true;
    for (int j = 0; j < n && tidy; ++j) {
        int prev = NONE;
        for (int p = Ap[j]; p < Ap[j + 1]; ++p) {
            if (Ai[p] < 0 || Ai[p] >= n) return false;
            if (Ai[p] <= prev) {
                tidy = false;
                break;
            }
            prev = Ai[p];
        }
    }
    std::vector<int> Tp_store, Ti_store;
    const std::vector<int>*Cp = &Ap, *Ci = &Ai;
    if (!tidy) {
        std::vector<int> cnt(n, 0), mark(n, NONE);
        for (int j = 0; j < n; ++j)
            for (int p = Ap[j]; p < Ap[j + 1]; ++p) {
                const int i = Ai[p];
                if (i < 0 || i >= n) return false;
                if (mark[i] != j) {
                    cnt[i]++;
                    mark[i] = j;
                }
            }
        Tp_store.assign(n + 1, 0);
        for (int i = 0; i < n; ++i) Tp_store[i + 1] = Tp_store[i] + cnt[i];
        Ti_store.resize(std::max(1, Tp_store[n]));
        for (int i = 0; i < n; ++i) {
            cnt[i] = Tp_store[i];
            mark[i] = NONE;
        }
        for (int j = 0; j < n; ++j)
            for (int p = Ap[j]; p < Ap[j + 1]; ++p) {
                const int i = Ai[p];
                if (mark[i] != j) {
                    Ti_store[cnt[i]++] = j;
                    mark[i] = j;
                }
            }
        Cp = &Tp_store;
        Ci = &Ti_store;
    }
    const std::vector<int>&cp = *Cp, &ci = *Ci;

    // One sweep builds A+A' (without diagonal).  It runs twice: first to count
    // the list lengths, then to fill the lists; `emit(a, b)` sees every
    // unordered off-diagonal pair once per orientation found.
    AmdState s;
    s.n = n;
    s.len.assign(n, 0);
    auto sweep = [&](auto&& emit) {
        std::vector<int> resume(n, 0);
        for (int k = 0; k < n; ++k) {
            int p = cp[k];
            const int p2 = cp[k + 1];
            while (p < p2) {
                const int j = ci[p];
                if (j < k) {
                    emit(j, k);
                    ++p;
                } else {
                    if (j == k) ++p;
                    break;
                }
                int pj = resume[j];
                const int pj2 = cp[j + 1];
                while (pj < pj2) {
                    const int i = ci[pj];
                    if (i < k) {
                        emit(i, j);
                        ++pj;
                    } else {
                        if (i == k) ++pj;
                        break;
                    }
                }
                resume[j] = pj;
            }
            resume[k] = p;
        }
        for (int j = 0; j < n; ++j)
            for (int pj = resume[j]; pj < cp[j + 1]; ++pj) emit(ci[pj], j);
    };
    sweep([&](int a, int b) {
        s.len[a]++;
        s.len[b]++;
    });
    long nzaat = 0;
    for (int k = 0; k < n; ++k) nzaat += s.len[k];
    s.iwlen = static_cast<int>(nzaat + nzaat / 5 + n);
    s.iw.assign(static_cast<size_t>(s.iwlen) + 1, 0);
    s.pe.assign(n, 0);
    std::vector<int> fillpos(n);
    int pf = 0;
    for (int j = 0; j < n; ++j) {
        s.pe[j] = fillpos[j] = pf;
        pf += s.len[j];
    }
    s.pfree = pf;
    sweep([&](int a, int b) {
        s.iw[fillpos[a]++] = b;
        s.iw[fillpos[b]++] = a;
    });
    s.nv.assign(n, 0);
    s.head.assign(n, 0);
    s.elen.assign(n, 0);
    s.degree.assign(n, 0);
    s.w.assign(n, 0);
    s.next.assign(n, 0);
    s.last.assign(n, 0);
    amd_core(s, perm, lnz);
    return true;
}

// ---------------------------------------------------------------------------
// klu_analyze equivalent: BTF, then AMD on every diagonal block larger than 3
// ---------------------------------------------------------------------------
bool analyze_pattern(const CscPattern& A, Ordering& out) {
    const int n = A.n;
    out = Ordering();
    out.n = n;
    if (n <= 0) return false;
    std::vector<int> match;
    out.structural_rank = maximum_transversal(n, A.colptr, A.rowidx, match);
    std::vector<int> Qbtf = match;  // row i is matched with column Qbtf[i]
    if (out.structural_rank < n) {
        std::vector<char> used(n, 0);
        for (int i = 0; i < n; ++i)
            if (Qbtf[i] != NONE) used[Qbtf[i]] = 1;
        std::vector<int> spare;
        for (int j = n - 1; j >= 0; --j)
            if (!used[j]) spare.push_back(j);
        for (int i = 0; i < n; ++i)
            if (Qbtf[i] == NONE && !spare.empty()) {
                Qbtf[i] = flip(spare.back());
                spare.pop_back();
            }
    }
    std::vector<int> Pbtf;
    out.nblocks = strong_components(n, A.colptr, A.rowidx, Qbtf, Pbtf, out.R);
    for (int& q : Qbtf) q = unflip(q);
    std::vector<int> pinv(n);
    for (int k = 0; k < n; ++k) pinv[Pbtf[k]] = k;
    out.P.assign(n, 0);
    out.Q.assign(n, 0);
    out.lnz_est.assign(out.nblocks, 0.0);
    out.maxblock = 1;
    std::vector<int> Cp, Ci, blk;
    for (int b = 0; b < out.nblocks; ++b) {
        const int k1 = out.R[b], k2 = out.R[b + 1], nk = k2 - k1;
        out.maxblock = std::max(out.maxblock, nk);
        Cp.assign(nk + 1, 0);
        Ci.clear();
        for (int k = k1; k < k2; ++k) {
            Cp[k - k1] = static_cast<int>(Ci.size());
            const int col = Qbtf[k];
            for (int p = A.colptr[col]; p < A.colptr[col + 1]; ++p) {
                const int r = pinv[A.rowidx[p]];
                if (r < k1)
                    out.nzoff++;
                else
                    Ci.push_back(r - k1);
            }
        }
        Cp[nk] = static_cast<int>(Ci.size());
        if (nk <= 3) {
            blk.resize(nk);
            std::iota(blk.begin(), blk.end(), 0);
            out.lnz_est[b] = nk * (nk + 1) / 2;
        } else {
            double lnz = 0;
            if (!amd_order(nk, Cp, Ci, blk, &lnz)) return false;
            out.lnz_est[b] = static_cast<double>(static_cast<int>(lnz)) + nk;
        }
        for (int k = 0; k < nk; ++k) {
            out.P[k1 + k] = Pbtf[k1 + blk[k]];
            out.Q[k1 + k] = Qbtf[k1 + blk[k]];
        }
    }
    return true;
}

}  // namespace kb
