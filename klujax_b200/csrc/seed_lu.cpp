// seed_lu.cpp — one pivoting factorization on the host, to fix the pivots.
//
// klujax re-pivots every system of a batch (klu_factor per system,
// /root/reference/klujax.cpp:300-322, :431-451).  The B200 path pivots ONCE, on
// a seed system, with KLU's rule — left-looking Gilbert-Peierls per diagonal
// block, max-abs row scaling, threshold partial pivoting with diagonal
// preference (tol = 1e-3) — and freezes the resulting row order and L/U
// patterns for the whole batch.  The numeric kernels later certify, per
// system, that KLU would have picked the same pivots (status/growth outputs).
// Only the *structure* leaves this file; seed values are discarded.
#include <algorithm>
#include <cmath>
#include <complex>

#include "symbolic.hpp"

namespace kb {

namespace {

inline double mag(double v) { return std::fabs(v); }
// overflow-safe modulus, the scaled form KLU uses for complex entries
inline double mag(const std::complex<double>& v) {
    double x = std::fabs(v.real()), y = std::fabs(v.imag());
    if (x < y) std::swap(x, y);
    if (x + y == x) return x;
    const double r = y / x;
    return x * std::sqrt(1.0 + r * r);
}
inline double over(double a, double s) { return a / s; }
inline std::complex<double> over(const std::complex<double>& a, double s) {
    return {a.real() / s, a.imag() / s};
}
inline double quot(double a, double b) { return a / b; }
// Smith's complex division
inline std::complex<double> quot(const std::complex<double>& a, const std::complex<double>& b) {
    const double br = b.real(), bi = b.imag(), ar = a.real(), ai = a.imag();
    if (std::fabs(br) >= std::fabs(bi)) {
        const double r = bi / br, den = br + r * bi;
        return {(ar + ai * r) / den, (ai - ar * r) / den};
    }
    const double r = br / bi, den = r * br + bi;
    return {(ar * r + ai) / den, (ai * r - ar) / den};
}
inline bool is_zero(double v) { return v == 0.0; }
inline bool is_zero(const std::complex<double>& v) { return v.real() == 0.0 && v.imag() == 0.0; }
inline void mult_sub(double& c, double a, double b) { c -= a * b; }
inline void mult_sub(std::complex<double>& c, const std::complex<double>& a,
                     const std::complex<double>& b) {
    c = {c.real() - (a.real() * b.real() - a.imag() * b.imag()),
         c.imag() - (a.imag() * b.real() + a.real() * b.imag())};
}
inline int flip(int i) { return -i - 2; }

}  // namespace

template <typename T>
bool seed_factor(const CscPattern& A, const Ordering& ord, const T* Ax, PivotPlan& plan) {
    const int n = A.n;
    const double tol = 1e-3;
    plan = PivotPlan();
    plan.n = n;
    plan.Pnum.assign(n, -1);
    plan.Lp.assign(n + 1, 0);
    plan.Up.assign(n + 1, 0);
    plan.Offp.assign(n + 1, 0);
    plan.diag_cand.assign(n, -2);
    if (A.has_duplicates) {
        plan.status = KLU_INVALID;  // klu_factor's KLU_scale rejects duplicates
        return false;
    }

    // max-abs row scale factors in original row order (empty rows -> 1)
    std::vector<double> rs(n, 0.0);
    for (int j = 0; j < n; ++j)
        for (int p = A.colptr[j]; p < A.colptr[j + 1]; ++p)
            rs[A.rowidx[p]] = std::max(rs[A.rowidx[p]], mag(Ax[p]));
    for (double& r : rs)
        if (r == 0.0) r = 1.0;

    std::vector<int> psinv(n);  // symbolic inverse row permutation
    for (int k = 0; k < n; ++k) psinv[ord.P[k]] = k;

    const int mb = ord.maxblock;
    std::vector<T> X(mb, T(0));
    std::vector<int> pinv(mb), prow(mb), mark(mb), stack(mb), resume(mb);
    std::vector<std::vector<int>> Lrow(mb);
    std::vector<std::vector<T>> Lval(mb);
    std::vector<int> ucol;  // pivot columns reached, topological order
    std::vector<int> off_oldrow;

    for (int b = 0; b < ord.nblocks; ++b) {
        const int k1 = ord.R[b], k2 = ord.R[b + 1], nk = k2 - k1;
        if (nk == 1) {
            const int col = ord.Q[k1];
            T s(0);
            for (int p = A.colptr[col]; p < A.colptr[col + 1]; ++p) {
                const int orow = A.rowidx[p];
                if (psinv[orow] < k1) {
                    off_oldrow.push_back(orow);
                    plan.Offsrc.push_back(p);
                } else {
                    s = over(Ax[p], rs[orow]);
                }
            }
            plan.Offp[k1 + 1] = static_cast<int>(off_oldrow.size());
            plan.Lp[k1 + 1] = static_cast<int>(plan.Li.size());
            plan.Up[k1 + 1] = static_cast<int>(plan.Ui.size());
            plan.Pnum[k1] = ord.P[k1];
            if (is_zero(s)) {
                plan.status = KLU_SINGULAR;
                return false;
            }
            continue;
        }
        for (int k = 0; k < nk; ++k) {
            prow[k] = k;
            pinv[k] = flip(k);
            mark[k] = -1;
            X[k] = T(0);
            Lrow[k].clear();
            Lval[k].clear();
        }
        for (int k = 0; k < nk; ++k) {
            const int col = ord.Q[k1 + k];
            std::vector<int>& cand = Lrow[k];  // non-pivotal rows reached = pattern of L(:,k)
            // ---- symbolic reach of A(:,k) through the graph of L
            int top = nk;
            for (int p = A.colptr[col]; p < A.colptr[col + 1]; ++p) {
                const int i0 = psinv[A.rowidx[p]] - k1;
                if (i0 < 0 || mark[i0] == k) continue;
                if (pinv[i0] < 0) {
                    mark[i0] = k;
                    cand.push_back(i0);
                    continue;
                }
                int head = 0;
                stack[0] = i0;
                while (head >= 0) {
                    const int j = stack[head];
                    const int jn = pinv[j];
                    if (mark[j] != k) {
                        mark[j] = k;
                        resume[head] = static_cast<int>(Lrow[jn].size());
                    }
                    int pos = --resume[head];
                    for (; pos >= 0; --pos) {
                        const int i = Lrow[jn][pos];
                        if (mark[i] == k) continue;
                        if (pinv[i] >= 0) {
                            resume[head] = pos;
                            stack[++head] = i;
                            break;
                        }
                        mark[i] = k;
                        cand.push_back(i);
                    }
                    if (pos == -1) {
                        --head;
                        stack[--top] = j;
                    }
                }
            }
            // ---- scatter the scaled column, split off the BTF off-diagonal part
            for (int p = A.colptr[col]; p < A.colptr[col + 1]; ++p) {
                const int orow = A.rowidx[p];
                const int i = psinv[orow] - k1;
                if (i < 0) {
                    off_oldrow.push_back(orow);
                    plan.Offsrc.push_back(p);
                } else {
                    X[i] = over(Ax[p], rs[orow]);
                }
            }
            plan.Offp[k1 + k + 1] = static_cast<int>(off_oldrow.size());
            // ---- sparse triangular solve over the topological order
            for (int s = top; s < nk; ++s) {
                const int j = stack[s], jn = pinv[j];
                const T xj = X[j];
                const auto& rows = Lrow[jn];
                const auto& vals = Lval[jn];
                for (size_t q = 0; q < rows.size(); ++q) mult_sub(X[rows[q]], vals[q], xj);
            }
            // ---- pivot: largest candidate, but keep the diagonal if it is
            //      within tol of the largest
            if (cand.empty()) {
                plan.status = KLU_SINGULAR;
                return false;
            }
            const int diagrow = prow[k];
            const int lastrow = cand.back();
            cand.pop_back();
            std::vector<T>& lv = Lval[k];
            lv.resize(cand.size());
            int pdiag = -1, ppiv = -1;
            double best = -1.0;
            for (size_t q = 0; q < cand.size(); ++q) {
                const int i = cand[q];
                lv[q] = X[i];
                X[i] = T(0);
                const double a = mag(lv[q]);
                if (i == diagrow) pdiag = static_cast<int>(q);
                if (a > best) {
                    best = a;
                    ppiv = static_cast<int>(q);
                }
            }
            double alast = mag(X[lastrow]);
            if (alast > best) {
                best = alast;
                ppiv = -1;
            }
            if (lastrow == diagrow) {
                if (alast >= tol * best) ppiv = -1;
            } else if (pdiag >= 0) {
                if (mag(lv[pdiag]) >= tol * best) ppiv = pdiag;
            }
            int pivrow;
            T pivot;
            if (ppiv >= 0) {
                pivrow = cand[ppiv];
                pivot = lv[ppiv];
                cand[ppiv] = lastrow;
                lv[ppiv] = X[lastrow];
            } else {
                pivrow = lastrow;
                pivot = X[lastrow];
            }
            X[lastrow] = T(0);
            if (is_zero(pivot)) {
                plan.status = KLU_SINGULAR;
                return false;
            }
            for (T& v : lv) v = quot(v, pivot);
            // ---- U(:,k): the pivotal rows reached, in topological order
            for (int s = top; s < nk; ++s) {
                const int j = stack[s];
                plan.Ui.push_back(pinv[j]);
                X[j] = T(0);
            }
            plan.Up[k1 + k + 1] = static_cast<int>(plan.Ui.size());
            // ---- record the pivot, keeping track of which row is "the diagonal"
            if (pivrow != diagrow) {
                plan.noffdiag++;
                plan.diag_cand[k1 + k] = pinv[diagrow] < 0 ? diagrow : -1;  // block-local row for now
                if (pinv[diagrow] < 0) {
                    const int kbar = flip(pinv[pivrow]);
                    prow[kbar] = diagrow;
                    pinv[diagrow] = flip(kbar);
                }
            }
            prow[k] = pivrow;
            pinv[pivrow] = k;
        }
        for (int k = 0; k < nk; ++k)
            if (plan.diag_cand[k1 + k] >= 0) plan.diag_cand[k1 + k] = k1 + pinv[plan.diag_cand[k1 + k]];
        // rows of L in pivot order; Pnum through the symbolic permutation
        for (int k = 0; k < nk; ++k) {
            for (int i : Lrow[k]) plan.Li.push_back(pinv[i]);
            plan.Lp[k1 + k + 1] = static_cast<int>(plan.Li.size());
            plan.Pnum[k1 + k] = ord.P[k1 + prow[k]];
        }
    }
    plan.Pinv.assign(n, 0);
    for (int k = 0; k < n; ++k) plan.Pinv[plan.Pnum[k]] = k;
    plan.Offi.resize(off_oldrow.size());
    for (size_t p = 0; p < off_oldrow.size(); ++p) plan.Offi[p] = plan.Pinv[off_oldrow[p]];
    return true;
}

template bool seed_factor<double>(const CscPattern&, const Ordering&, const double*, PivotPlan&);
template bool seed_factor<std::complex<double>>(const CscPattern&, const Ordering&,
                                                const std::complex<double>*, PivotPlan&);

}  // namespace kb
