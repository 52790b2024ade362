// tools/dep_check.cpp - host check of the dependency-driven programs (plan.hpp: ColProgram,
// PanelProgram): interprets them on one system exactly the way the kernels do (kernels_dep.cuh),
// asserts that every column a column waits for comes earlier in the order (the property the
// factor kernel's deadlock freedom rests on), and checks the backward error of x.
//
//   cd klujax_b200/csrc && g++ -O2 -std=c++17 -I. ../../tools/dep_check.cpp ordering.cpp seed_lu.cpp plan.cpp -o /tmp/dep_check
//   /tmp/dep_check pattern.txt        pattern: first line `n nz`, then `i j re im` per entry
#include <cmath>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "plan.hpp"
#include "symbolic.hpp"
using namespace kb;
using cd = std::complex<double>;

static std::vector<cd> dense_solve(int n, std::vector<cd> A, std::vector<cd> b, bool transpose) {
    if (transpose)
        for (int i = 0; i < n; ++i)
            for (int j = i + 1; j < n; ++j) std::swap(A[(size_t)i * n + j], A[(size_t)j * n + i]);
    for (int k = 0; k < n; ++k) {
        int p = k;
        for (int i = k + 1; i < n; ++i)
            if (std::abs(A[(size_t)i * n + k]) > std::abs(A[(size_t)p * n + k])) p = i;
        if (p != k) {
            for (int j = 0; j < n; ++j) std::swap(A[(size_t)k * n + j], A[(size_t)p * n + j]);
            std::swap(b[k], b[p]);
        }
        for (int i = k + 1; i < n; ++i) {
            const cd l = A[(size_t)i * n + k] / A[(size_t)k * n + k];
            if (l == cd(0)) continue;
            for (int j = k; j < n; ++j) A[(size_t)i * n + j] -= l * A[(size_t)k * n + j];
            b[i] -= l * b[k];
        }
    }
    for (int k = n - 1; k >= 0; --k) {
        for (int j = k + 1; j < n; ++j) b[k] -= A[(size_t)k * n + j] * b[j];
        b[k] /= A[(size_t)k * n + k];
    }
    return b;
}

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    FILE* f = std::fopen(argv[1], "r");
    int n, nz;
    if (!f || std::fscanf(f, "%d %d", &n, &nz) != 2) return 2;
    std::vector<int> Ai(nz), Aj(nz);
    std::vector<cd> Ax(nz);
    for (int e = 0; e < nz; ++e) {
        double re, im;
        if (std::fscanf(f, "%d %d %lf %lf", &Ai[e], &Aj[e], &re, &im) != 4) return 2;
        Ax[e] = {re, im};
    }
    CscPattern A = coo_to_csc(n, nz, Ai.data(), Aj.data());
    Ordering ord;
    analyze_pattern(A, ord);
    std::vector<cd> cv(nz);
    for (int p = 0; p < nz; ++p) cv[p] = Ax[A.coo_of[p]];
    PivotPlan piv;
    if (!seed_factor<cd>(A, ord, cv.data(), piv)) {
        std::printf("seed factorization failed\n");
        return 1;
    }
    SlotLayout lay;
    build_layout(A, ord, piv, lay);
    // scatter + max-abs row scaling, like the kernels
    std::vector<cd> V(lay.n_val, cd(0));
    for (int s = 0; s < lay.n_val; ++s)
        if (lay.slot_csc[s] >= 0) V[s] = cv[lay.slot_csc[s]];
    std::vector<double> Rs(n, 1.0);
    for (int i = 0; i < n; ++i) {
        double mx = 0;
        for (int p = lay.rs_ptr[i]; p < lay.rs_ptr[i + 1]; ++p) mx = std::max(mx, std::abs(V[lay.rs_slot[p]]));
        if (mx == 0) mx = 1;
        Rs[i] = mx;
        for (int p = lay.rs_ptr[i]; p < lay.rs_ptr[i + 1]; ++p) V[lay.rs_slot[p]] /= mx;
    }
    ColProgram cp;
    if (!build_col_program(piv, lay, cp)) {
        std::printf("column program not built\n");
        return 1;
    }
    int bad_order = 0;
    {
        std::vector<char> done(n, 0);
        std::vector<cd> col;
        for (int k = 0; k < n; ++k) {
            const int cb = cp.colinfo[4 * k], nu = cp.colinfo[4 * k + 1], nc = cp.colinfo[4 * k + 2], ub = cp.colinfo[4 * k + 3];
            col.assign(V.begin() + cb, V.begin() + cb + nc);
            for (int u = 0; u < nu; ++u) {
                const int j = cp.urec[4 * (ub + u)], lbeg = cp.urec[4 * (ub + u) + 1], lcnt = cp.urec[4 * (ub + u) + 2], dp = cp.urec[4 * (ub + u) + 3];
                if (!done[j]) ++bad_order;
                const cd uv = col[u];
                for (int q = 0; q < lcnt; ++q) col[cp.dest[dp + q]] -= V[lbeg + q] * uv;
            }
            const cd inv = cd(1) / col[nu];
            for (int p = nu + 1; p < nc; ++p) col[p] *= inv;
            col[nu] = inv;
            for (int p = 0; p < nc; ++p) V[cb + p] = col[p];
            done[k] = 1;
        }
    }
    std::printf("column program: n %d maxcol %d depth %d mac %lld order violations %d\n", n, cp.maxcol, cp.depth, (long long)cp.n_mac, bad_order);
    // dense copy of A and a right-hand side
    std::vector<cd> D;
    if (n <= 3000) {
        D.assign((size_t)n * n, cd(0));
        for (int e = 0; e < nz; ++e) D[(size_t)Ai[e] * n + Aj[e]] += Ax[e];
    }
    std::vector<cd> b(n);
    for (int i = 0; i < n; ++i) b[i] = cd(std::sin(0.37 * i) + 1.0, std::cos(0.11 * i));
    int rc = bad_order ? 1 : 0;
    // panel programs (plan.hpp: PanelProgram), interpreted the way kernels_dep.cuh: panel_solve does
    for (int tr = 0; tr < 2; ++tr) {
        PanelProgram pp;
        if (!build_panel_program(ord, piv, lay, tr != 0, pp)) {
            std::printf("panel program not built\n");
            return 1;
        }
        std::vector<cd> X(n + pp.ntemps, cd(0));
        for (int k = 0; k < n; ++k) {
            const int src = tr ? ord.Q[k] : piv.Pnum[k];
            X[k] = tr ? b[src] : b[src] / Rs[src];
        }
        auto coef = [&](int slot) { return slot == -2 ? cd(-1) : V[slot]; };
        long long steps = 0;
        for (int p = 0; p < pp.npanels; ++p) {
            const int E = pp.hdr[4 * p], I = pp.hdr[4 * p + 1], eo = pp.hdr[4 * p + 2], io = pp.hdr[4 * p + 3];
            cd acc[32], dv[32];
            int tcnt[32];
            unsigned any = 0;
            for (int ln = 0; ln < 32; ++ln) {
                const int* li = &pp.lane[4 * ((size_t)p * 32 + ln)];
                acc[ln] = (li[3] & 1) ? cd(0) : X[li[0]];
                dv[ln] = li[1] >= 0 ? V[li[1]] : cd(1);
                for (int e = 0; e < E; ++e) {
                    const int idx = pp.ext_idx[((size_t)eo + e) * 32 + ln];
                    if (idx >= 0) acc[ln] -= coef(pp.ext_slot[((size_t)eo + e) * 32 + ln]) * X[idx];
                }
                tcnt[ln] = 0;
                any |= (unsigned)li[2];
            }
            for (int sft = 0; sft < 32; ++sft) {
                if (!((any >> sft) & 1)) continue;
                ++steps;
                const int* ls = &pp.lane[4 * ((size_t)p * 32 + sft)];
                const cd xs = (ls[3] & 4) ? acc[sft] * dv[sft] : acc[sft];
                for (int ln = sft + 1; ln < 32; ++ln) {
                    const int* li = &pp.lane[4 * ((size_t)p * 32 + ln)];
                    if (((unsigned)li[2] >> sft) & 1) {
                        if (tcnt[ln] >= I) return 3;
                        acc[ln] -= coef(pp.in_slot[((size_t)io + tcnt[ln]) * 32 + ln]) * xs;
                        tcnt[ln]++;
                    }
                }
            }
            for (int ln = 0; ln < 32; ++ln) {
                const int* li = &pp.lane[4 * ((size_t)p * 32 + ln)];
                if (!(li[3] & 2)) X[li[0]] = (li[3] & 4) ? acc[ln] * dv[ln] : acc[ln];
            }
        }
        std::vector<cd> x(n);
        for (int k = 0; k < n; ++k) {
            const int dst = tr ? piv.Pnum[k] : ord.Q[k];
            x[dst] = tr ? X[k] / Rs[dst] : X[k];
        }
        std::vector<cd> res(b);
        std::vector<double> scale(n, 0.0);
        for (int i = 0; i < n; ++i) scale[i] = std::abs(b[i]);
        for (int e = 0; e < nz; ++e) {
            if (!tr) res[Ai[e]] -= Ax[e] * x[Aj[e]], scale[Ai[e]] += std::abs(Ax[e]) * std::abs(x[Aj[e]]);
            else res[Aj[e]] -= Ax[e] * x[Ai[e]], scale[Aj[e]] += std::abs(Ax[e]) * std::abs(x[Ai[e]]);
        }
        double err = 0;
        for (int i = 0; i < n; ++i) err = std::max(err, std::abs(res[i]) / scale[i]);
        size_t ext_lines = pp.ext_slot.size() / 32, in_lines = pp.in_slot.size() / 32;
        std::printf("panel program (transpose %d): %d panels, %d temporaries, max external depth %d, max in-panel depth %d, %zu external lines, %zu in-panel lines, %lld shuffle steps, backward error %.3e\n",
                    tr, pp.npanels, pp.ntemps, pp.maxE, pp.maxI, ext_lines, in_lines, steps, err);
        if (!(err < 1e-9)) rc = 1;  // a wrong program is off by O(1); badly scaled test matrices reach 1e-11
    }

    // group programs (plan.hpp: GroupProgram), interpreted level by level the way lv_solve_grouped does;
    // external reads once from a snapshot taken at the start of the level (what a concurrent group sees
    // if it runs first) and once in place in sorted order (if it runs last): both must give x
    for (int tr = 0; tr < 2; ++tr)
        for (int snap = 0; snap < 2; ++snap) {
            GroupProgram gp;
            if (!build_group_program(ord, piv, lay, tr != 0, nullptr, gp)) {
                std::printf("group program not built\n");
                return 1;
            }
            std::vector<cd> X(n), S;
            for (int k = 0; k < n; ++k) {
                const int src = tr ? ord.Q[k] : piv.Pnum[k];
                X[k] = tr ? b[src] : b[src] / Rs[src];
            }
            for (int l = 0; l < gp.nlevels; ++l) {
                if (snap) S = X;
                const std::vector<cd>& R = snap ? S : X;
                for (int g = gp.level_ptr[l]; g < gp.level_ptr[l + 1]; ++g) {
                    const int nk = gp.ghdr[12 * g], ko = gp.ghdr[12 * g + 1], nt = gp.ghdr[12 * g + 2] & 255, mb = (gp.ghdr[12 * g + 2] >> 8) & 255;
                    const bool partial = (gp.ghdr[12 * g + 2] >> 30) & 1;
                    const int* aux = &gp.gaux[(size_t)kGroupAux * gp.ghdr[12 * g + 3]];
                    for (int i = 0; i < nt; ++i)
                        if (gp.ghdr[12 * g + 4 + i] != aux[i]) return 4;
                    cd acc[kGroupRows];
                    for (int i = 0; i < nt; ++i) acc[i] = R[aux[i]];
                    for (int k = 0; k < nk; ++k) {
                        const int* kr = &gp.krec[12 * ((size_t)ko + k)];
                        const cd xv = R[kr[0]];
                        for (int i = 0; i < nt; ++i)
                            if (kr[1 + i] >= 0) acc[i] -= V[kr[1 + i]] * xv;
                    }
                    for (int i = 0; i < nt; ++i) {
                        if (partial) {
                            X[aux[i]] = acc[i];
                            continue;
                        }
                        const cd val = ((mb >> i) & 1) ? acc[i] * V[aux[8 + i]] : acc[i];
                        X[aux[i]] = val;
                        for (int j = i + 1; j < nt; ++j) {
                            const int sl = aux[16 + j * (j - 1) / 2 + i];
                            if (sl >= 0) acc[j] -= V[sl] * val;
                        }
                    }
                }
            }
            std::vector<cd> x(n);
            for (int k = 0; k < n; ++k) {
                const int dst = tr ? piv.Pnum[k] : ord.Q[k];
                x[dst] = tr ? X[k] / Rs[dst] : X[k];
            }
            std::vector<cd> res(b);
            std::vector<double> scale(n, 0.0);
            for (int i = 0; i < n; ++i) scale[i] = std::abs(b[i]);
            for (int e = 0; e < nz; ++e) {
                if (!tr) res[Ai[e]] -= Ax[e] * x[Aj[e]], scale[Ai[e]] += std::abs(Ax[e]) * std::abs(x[Aj[e]]);
                else res[Aj[e]] -= Ax[e] * x[Ai[e]], scale[Aj[e]] += std::abs(Ax[e]) * std::abs(x[Ai[e]]);
            }
            double err = 0;
            for (int i = 0; i < n; ++i) err = std::max(err, std::abs(res[i]) / scale[i]);
            if (snap == 0)
                std::printf("group program (transpose %d): %d groups, %d levels, %lld multiply-adds in %lld k-steps (%.2f rows per read of X)\n", tr,
                            gp.ngroups, gp.nlevels, (long long)gp.n_mac, (long long)gp.n_ksteps, (double)gp.n_mac / std::max<long long>(1, gp.n_ksteps));
            std::printf("  %s: backward error %.3e\n", snap ? "level snapshot" : "in place", err);
            if (!(err < 1e-9)) rc = 1;
        }
    std::printf(rc ? "FAILED\n" : "OK\n");
    return rc;
}
