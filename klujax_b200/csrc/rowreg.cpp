// rowreg.cpp — planner of the register-resident regime (rowreg.hpp) and its host interpreter.
//
// The planner decides, once per symbolic handle:
//   * which lane owns which row (rows that one pivot updates together want different lanes, so
//     that one predicated instruction serves them all; rows of one elimination level too);
//   * which value register holds which entry (entries of one COLUMN in rows that are updated
//     together want the SAME register name: the update is then one instruction);
//   * the order of elimination (level sets of the pivot dependency graph), which rows stay in
//     registers until the back substitution ("pinned") and when the others are parked in
//     shared memory.
// Follows /root/reference/klujax.cpp:431-451 in effect (klu_factor + klu_solve per system) with
// the static pivot sequence of the seed factorization (seed_lu.cpp).
#include "rowreg.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <numeric>
#include <unordered_map>

namespace kb {
namespace rr {
namespace {

constexpr int kMaxRegsHard = 192;

struct Ent {
    int row = 0, col = 0, csc = -1;
    int reg = -1;
    int slot = -1;  // Ubuf slot once parked
    int lane = 0;   // lane that holds the entry (the row's lane unless the entry was moved)
    bool alive = false, frozen = false;
};

struct Builder {
    // problem
    int n = 0, nz = 0, r = 1;
    bool transpose = false;
    std::vector<std::vector<std::pair<int, int>>> rows;  // owner row -> (col, csc)
    std::vector<int> blk;                                // pivot -> block id
    std::vector<int> R;                                  // block boundaries
    std::vector<int> rhs_src, rhs_scale, out_dst, out_scale;
    Options opt;
    // output
    Plan* pl = nullptr;
    // register file state
    std::vector<uint32_t> occ;  // per register: lanes in use
    int Pmax = 0;
    // entries
    std::vector<Ent> ents;
    std::unordered_map<long long, int> emap;
    std::vector<int> lane_of, g_of;  // per pivot row
    std::vector<int> stack_depth;    // (unused)
    int block_planes = 0;            // (unused)
    int time_base = 0;               // global time of the start of the current block
    // shared-memory slot pool: planes of 32 values; a slot is busy over [from, to] (global time):
    // entries of A from the staging to their load, parked values from the dump to the block's end
    std::vector<std::vector<std::vector<std::pair<int, int>>>> pool;
    std::vector<int> csc_tabs;       // tables whose entries are CSC positions (translated to slots at the end)
    std::vector<uint16_t> place;     // CSC position -> slot
    // finds a plane where (lane, [from,to]) is free for every member; marks it
    int pool_place(const std::vector<int>& lanes, const std::vector<std::pair<int, int>>& iv) {
        for (size_t pi = 0;; ++pi) {
            if (pi == pool.size()) pool.emplace_back(kLanes);
            bool ok = true;
            for (size_t q = 0; q < lanes.size() && ok; ++q)
                for (auto [b, d] : pool[pi][lanes[q]])
                    if (iv[q].first <= d && b <= iv[q].second) {
                        ok = false;
                        break;
                    }
            if (!ok) continue;
            for (size_t q = 0; q < lanes.size(); ++q) pool[pi][lanes[q]].push_back(iv[q]);
            return static_cast<int>(pi);
        }
    }
    bool failed = false;
    std::string why;

    long long key(int row, int col) const { return static_cast<long long>(row) * (n + 8) + col; }
    int find(int row, int col) const {
        auto it = emap.find(key(row, col));
        return it == emap.end() ? -1 : it->second;
    }
    int add_ent(int row, int col, int csc) {
        Ent e;
        e.row = row;
        e.col = col;
        e.csc = csc;
        ents.push_back(e);
        emap[key(row, col)] = static_cast<int>(ents.size()) - 1;
        return static_cast<int>(ents.size()) - 1;
    }
    int add_tab(const uint16_t* v) {
        const int id = static_cast<int>(pl->tabs.size() / kLanes);
        pl->tabs.insert(pl->tabs.end(), v, v + kLanes);
        return id;
    }
    void emit(const Ins& i) { pl->ins.push_back(i); }
    void fail(const std::string& w) {
        if (!failed) why = w;
        failed = true;
    }

    // lowest register free in all lanes of `lanes`, starting at `from`
    int first_fit(uint32_t lanes, int from = 0) {
        for (int p = from; p < kMaxRegsHard; ++p)
            if (!(occ[p] & lanes)) return p;
        fail("out of value registers");
        return 0;
    }
    void take(int e, int p) {
        const uint32_t bit = 1u << lane_of[ents[e].row];
        if (occ[p] & bit) fail("internal: register slot taken twice");
        occ[p] |= bit;
        ents[e].reg = p;
        ents[e].alive = true;
        Pmax = std::max(Pmax, p + 1);
    }
    void release(int e) {
        occ[ents[e].reg] &= ~(1u << lane_of[ents[e].row]);
        ents[e].alive = false;
    }

    // ---- one BTF block ----------------------------------------------------------------
    void plan_block(int k0, int k1);
};

void Builder::plan_block(int k0, int k1) {
    const int nb = k1 - k0;
    auto inblk = [&](int c) { return c >= k0 && c < k1; };
    // --- structure ---
    std::vector<std::vector<int>> Lrows(nb), Ucols(nb), Ucolrows(nb);
    std::vector<std::vector<std::pair<int, int>>> offs(nb);  // (col, csc)
    for (int i = k0; i < k1; ++i)
        for (auto [c, csc] : rows[i]) {
            if (!inblk(c)) {
                offs[i - k0].push_back({c, csc});
                continue;
            }
            if (c < i) Lrows[c - k0].push_back(i);
            if (c > i) {
                Ucols[i - k0].push_back(c);
                Ucolrows[c - k0].push_back(i);
            }
        }
    for (auto& v : Lrows) std::sort(v.begin(), v.end());
    for (auto& v : Ucols) std::sort(v.begin(), v.end());
    // --- levels of the pivot dependency graph (elimination and back substitution) ---
    std::vector<int> lev(nb, 0);
    for (int k = k0; k < k1; ++k) {
        // row k depends on pivot j < k if L(k,j) != 0 (it is updated by j) or U(j,k) != 0
        int l = 0;
        for (auto [c, csc] : rows[k])
            if (inblk(c) && c < k) l = std::max(l, lev[c - k0] + 1);
        for (int j : Ucolrows[k - k0]) l = std::max(l, lev[j - k0] + 1);
        lev[k - k0] = l;
    }
    std::vector<int> sched(nb);
    std::iota(sched.begin(), sched.end(), k0);
    // Order of elimination.  The pivot order of the analysis is a post-order of the elimination
    // tree (AMD): walking it keeps few columns alive at a time, and the number of columns alive is
    // what the register file has to hold (every live column is one register name).  Walking the
    // level sets instead (all branches of the tree at once) gives wider epochs but needs about
    // twice the registers (cfg 5: 73 instead of 4x).
    if (opt.level_order)
        std::stable_sort(sched.begin(), sched.end(), [&](int a, int b) { return lev[a - k0] < lev[b - k0]; });
    std::vector<int> spos(nb);
    for (int t = 0; t < nb; ++t) spos[sched[t] - k0] = t;
    const int npin = std::min(nb, opt.pinned_rows);
    std::vector<char> pinned(nb, 0);
    for (int t = nb - npin; t < nb; ++t) pinned[sched[t] - k0] = 1;
    pl->st.pinned += npin;

    // --- lanes: greedy, last rows of the schedule first ---
    {
        std::vector<double> load(kLanes, 0.0);
        std::vector<std::vector<int>> lane_rows(kLanes);
        // pivots that update row i (i in Lrows[k])
        std::vector<std::vector<int>> updaters(nb);
        for (int k = 0; k < nb; ++k)
            for (int i : Lrows[k]) updaters[i - k0].push_back(k);
        std::vector<int> assigned(nb, -1);
        for (int t = nb - 1; t >= 0; --t) {
            const int i = sched[t];
            double best = 1e300;
            int bl = 0;
            std::vector<double> pen(kLanes, 0.0);
            for (int k : updaters[i - k0]) {
                const double w = static_cast<double>(Ucols[k].size()) + r + 1.0;
                uint32_t used = 0;
                for (int i2 : Lrows[k])
                    if (assigned[i2 - k0] >= 0) used |= 1u << assigned[i2 - k0];
                for (int l = 0; l < kLanes; ++l)
                    if (used >> l & 1u) pen[l] += w;
            }
            // rows that THIS row updates as a pivot: its own Lrows do not conflict with it.
            // same level: one row per lane and epoch
            for (int i2 = k0; i2 < k1; ++i2)
                if (i2 != i && assigned[i2 - k0] >= 0 && lev[i2 - k0] == lev[i - k0]) pen[assigned[i2 - k0]] += 8.0;
            // two pinned rows in one lane cost registers during the dense tail
            if (pinned[i - k0])
                for (int i2 = k0; i2 < k1; ++i2)
                    if (i2 != i && assigned[i2 - k0] >= 0 && pinned[i2 - k0]) pen[assigned[i2 - k0]] += 1000.0;
            const double w_self = static_cast<double>(rows[i].size());
            for (int l = 0; l < kLanes; ++l) {
                const double c = pen[l] + 0.05 * load[l] + 2.0 * static_cast<double>(lane_rows[l].size());
                if (c < best) {
                    best = c;
                    bl = l;
                }
            }
            assigned[i - k0] = bl;
            load[bl] += w_self;
            lane_rows[bl].push_back(i);
        }
        for (int i = k0; i < k1; ++i) lane_of[i] = assigned[i - k0];
        // rank of a row among its lane's rows (schedule order): its right-hand-side register
        for (int l = 0; l < kLanes; ++l) {
            auto& v = lane_rows[l];
            std::sort(v.begin(), v.end(), [&](int a, int b) { return spos[a - k0] < spos[b - k0]; });
            for (size_t q = 0; q < v.size(); ++q) g_of[v[q]] = static_cast<int>(q);
        }
    }
    int G = 0;
    for (int i = k0; i < k1; ++i) G = std::max(G, g_of[i] + 1);

    // --- registers -------------------------------------------------------------------------
    // Right-hand sides: dedicated registers yr(g, c).  Everything else: every entry (i,j) has a
    // life span [birth, death] in schedule time; entries of one column in rows that some pivot
    // updates together form a class and want ONE register name; classes are packed into
    // registers by first fit over (lane, time).
    auto yr = [&](int g, int c) { return g * r + c; };
    const int first_reg = G * r;
    Pmax = std::max(Pmax, first_reg);
    const size_t ent0 = ents.size();
    for (int i = k0; i < k1; ++i) {
        for (auto [c, csc] : rows[i])
            if (inblk(c)) ents[add_ent(i, c, csc)].lane = lane_of[i];
        for (int c = 0; c < r; ++c) {
            const int e = add_ent(i, n + c, -2);
            ents[e].reg = yr(g_of[i], c);
            ents[e].lane = lane_of[i];
            ents[e].alive = true;
        }
    }
    auto is_a = [&](int e) { return ents[e].csc >= 0; };
    // epochs: pivots of one level, one row per lane
    struct Epoch {
        std::vector<int> piv;
    };
    std::vector<Epoch> epochs;
    std::vector<int> epoch_of(nb, 0);
    {
        int t = 0;
        std::vector<char> in_ep(nb, 0);
        while (t < nb) {
            Epoch ep;
            uint32_t lanes = 0;
            while (t < nb) {
                const int k = sched[t];
                const uint32_t bit = 1u << lane_of[k];
                if (lanes & bit) break;  // one row per lane and epoch
                bool dep = false;        // rows of one epoch do not depend on each other
                for (auto [c, csc] : rows[k])
                    if (inblk(c) && c != k && in_ep[c - k0]) dep = true;
                for (int j : Ucolrows[k - k0])
                    if (in_ep[j - k0]) dep = true;
                for (int i : Lrows[k - k0])
                    if (in_ep[i - k0]) dep = true;
                if (dep) break;
                lanes |= bit;
                in_ep[k - k0] = 1;
                epoch_of[k - k0] = static_cast<int>(epochs.size());
                ep.piv.push_back(k);
                ++t;
            }
            for (int k : ep.piv) in_ep[k - k0] = 0;
            epochs.push_back(ep);
        }
    }
    pl->st.epochs += static_cast<int>(epochs.size());
    constexpr int kInf = 1 << 30;
    auto T = [&](int k) { return 2 * spos[k - k0] + 2; };  // time of pivot k
    std::vector<int> epoch_end(epochs.size(), 0);
    for (size_t q = 0; q < epochs.size(); ++q) epoch_end[q] = T(epochs[q].piv.back()) + 1;
    std::vector<int> birth(ents.size(), -1), death(ents.size(), kInf);
    for (size_t e = ent0; e < ents.size(); ++e) {
        const Ent& en = ents[e];
        if (en.col >= n) continue;
        (void)0;  // entries of A are born when the epoch that first touches them starts (below)
        if (en.col < en.row)
            death[e] = T(en.col);
        else
            death[e] = pinned[en.row - k0] ? kInf : epoch_end[epoch_of[en.row - k0]];
    }
    // symbolic elimination: birth of the fill-in; classes
    std::vector<int> parent(ents.size());
    std::iota(parent.begin(), parent.end(), 0);
    auto root = [&](int x) {
        while (parent[x] != x) x = parent[x] = parent[parent[x]];
        return x;
    };
    std::vector<int> load_ep(ents.size(), -1);  // entries of A: epoch whose start loads them
    auto LT = [&](int ep) { return T(epochs[ep].piv[0]) - 1; };
    auto touch = [&](int e, int k) {
        if (birth[e] >= 0) return;
        if (is_a(e)) {
            load_ep[e] = opt.lazy_load ? epoch_of[k - k0] : 0;
            birth[e] = LT(load_ep[e]);
        } else {
            birth[e] = T(k);
        }
    };
    for (int t = 0; t < nb; ++t) {
        const int k = sched[t];
        const auto& I = Lrows[k - k0];
        touch(find(k, k), k);
        for (int j : Ucols[k - k0]) touch(find(k, j), k);
        if (I.empty()) continue;
        std::vector<int> cols;
        cols.push_back(k);
        for (int j : Ucols[k - k0]) cols.push_back(j);
        for (int j : cols) {
            int first = -1;
            for (int i : I) {
                const int e = find(i, j);
                if (e < 0) {
                    fail("pattern is not closed under elimination");
                    return;
                }
                touch(e, k);
                if (first < 0)
                    first = e;
                else
                    parent[root(e)] = root(first);
            }
        }
    }
    for (size_t e = ent0; e < ents.size(); ++e)
        if (ents[e].col < n && birth[e] < 0) {
            // structural entry nothing ever touches (kept for safety): born when it is first read
            const Ent& en = ents[e];
            touch(static_cast<int>(e), en.col < en.row ? en.col : en.row);
        }
    // --- move entries out of overloaded lanes ------------------------------------------------
    // A row is atomic only as far as its multipliers, its diagonal and its right-hand side go;
    // entries of U in rows that will be parked may live in any lane (they get the multiplier by
    // one shuffle per pivot).  Lanes that hold two long rows at the same time hand entries over
    // to lanes with room, until the peak number of live values per lane stops shrinking.
    // Rows that are parked but reach far into the columns of the pinned rows (the last rows
    // before the dense tail) would each need a register of their own per column, because their
    // lane also holds a pinned row.  Their entries in pinned columns are striped over the lane
    // groups instead: row q of the set keeps its entry of column j in lane q + W * (j mod F), so
    // that F columns share one register and every lane needs the multiplier of one fixed row.
    if (opt.stripe && nb > kLanes) {
        std::vector<int> pcol(nb, -1);
        int np_ = 0;
        for (int t = 0; t < nb; ++t)
            if (pinned[sched[t] - k0]) pcol[sched[t] - k0] = np_++;
        std::vector<int> S;
        for (int t = 0; t < nb; ++t) {
            const int i = sched[t];
            if (pinned[i - k0]) continue;
            int cnt = 0;
            for (int j : Ucols[i - k0]) cnt += pinned[j - k0] ? 1 : 0;
            if (cnt >= opt.stripe_min) S.push_back(i);
        }
        if (!S.empty() && S.size() <= 16) {
            const int F = S.size() <= 8 ? 4 : 2, Wd = kLanes / F;
            for (size_t q = 0; q < S.size(); ++q)
                for (int j : Ucols[S[q] - k0])
                    if (pinned[j - k0]) {
                        ents[find(S[q], j)].lane = static_cast<int>(q) + Wd * (pcol[j - k0] % F);
                        pl->st.moved++;
                    }
        }
    }
    if (opt.rebalance && nb > kLanes) {
        const int TT = 2 * nb + 6;
        auto clampT = [&](int t) { return std::min(t, TT - 1); };
        std::vector<std::vector<int>> load(kLanes, std::vector<int>(TT, 0));
        auto add_iv = [&](int l, int b, int d, int w) {
            for (int t = b; t <= clampT(d); ++t) load[l][t] += w;
        };
        for (size_t e = ent0; e < ents.size(); ++e)
            if (ents[e].col < n) add_iv(ents[e].lane, birth[e], death[e], 1);
        auto lane_peak = [&](int l) {
            int m = 0;
            for (int t = 0; t < TT; ++t) m = std::max(m, load[l][t]);
            return m;
        };
        for (int iter = 0; iter < 4000; ++iter) {
            int wl = 0, wp = -1;
            for (int l = 0; l < kLanes; ++l) {
                const int pk = lane_peak(l);
                if (pk > wp) wp = pk, wl = l;
            }
            int wt = 0;
            for (int t = 0; t < TT; ++t)
                if (load[wl][t] == wp) {
                    wt = t;
                    break;
                }
            // a movable entry that is alive at the peak: U part of a row that gets parked
            int best_e = -1, best_l = -1, best_gain = 0;
            for (size_t e = ent0; e < ents.size(); ++e) {
                const Ent& en = ents[e];
                if (en.col >= n || en.lane != wl || en.col <= en.row || pinned[en.row - k0]) continue;
                if (birth[e] > wt || death[e] < wt) continue;
                for (int l = 0; l < kLanes; ++l) {
                    if (l == wl) continue;
                    int m = 0;
                    for (int t = birth[e]; t <= clampT(death[e]); ++t) m = std::max(m, load[l][t]);
                    const int gain = wp - (m + 1);
                    if (gain > best_gain) best_gain = gain, best_e = static_cast<int>(e), best_l = l;
                }
            }
            if (best_e < 0 || best_gain < 2) break;
            add_iv(wl, birth[best_e], death[best_e], -1);
            add_iv(best_l, birth[best_e], death[best_e], +1);
            ents[best_e].lane = best_l;
            pl->st.moved++;
        }
    }
    if (std::getenv("KLUJAX_CUDA_RR_DEBUG")) {
        // lower bound of the register need: peak number of live entries in one lane
        int worst = 0, wl = 0, wt = 0;
        for (int l = 0; l < kLanes; ++l) {
            std::vector<std::pair<int, int>> ev;
            for (size_t e = ent0; e < ents.size(); ++e)
                if (ents[e].col < n && ents[e].lane == l) {
                    ev.push_back({birth[e], +1});
                    ev.push_back({std::min(death[e], kInf - 1) + 1, -1});
                }
            std::sort(ev.begin(), ev.end());
            int cur = 0;
            for (auto [t, d] : ev) {
                cur += d;
                if (cur > worst) worst = cur, wl = l, wt = t;
            }
        }
        std::fprintf(stderr, "block [%d,%d): peak live entries in one lane = %d (lane %d, time %d) + %d right-hand-side registers\n", k0,
                     k1, worst, wl, wt, first_reg);
    }
    // pack the classes
    {
        std::map<int, std::vector<int>> cls;
        for (size_t e = ent0; e < ents.size(); ++e)
            if (ents[e].col < n) cls[root(static_cast<int>(e))].push_back(static_cast<int>(e));
        std::vector<std::vector<int>> order;
        for (auto& [rt, v] : cls) order.push_back(v);
        auto cbirth = [&](const std::vector<int>& v) {
            int b = kInf;
            for (int e : v) b = std::min(b, birth[e]);
            return b;
        };
        auto cweight = [&](const std::vector<int>& v) {  // lane-time the class occupies
            long long w = 0;
            for (int e : v) w += std::min(death[e], 2 * nb + 6) - birth[e] + 1;
            return w;
        };
        const int pack_order = std::getenv("KLUJAX_CUDA_RR_PACK") ? std::atoi(std::getenv("KLUJAX_CUDA_RR_PACK")) : 1;
        std::stable_sort(order.begin(), order.end(), [&](const std::vector<int>& a, const std::vector<int>& b) {
            if (pack_order == 1) {  // heavy classes first: they define the register names, the rest fills gaps
                const long long wa = cweight(a), wb = cweight(b);
                if (wa != wb) return wa > wb;
            }
            const int ba = cbirth(a), bb = cbirth(b);
            if (ba != bb) return ba < bb;
            return a.size() > b.size();
        });
        // busy[p][lane] = list of [birth, death]
        std::vector<std::vector<std::vector<std::pair<int, int>>>> busy(
            kMaxRegsHard, std::vector<std::vector<std::pair<int, int>>>(kLanes));
        auto is_free = [&](int p, int lane, int b, int d) {
            for (auto [bb, dd] : busy[p][lane])
                if (b <= dd && bb <= d) return false;
            return true;
        };
        for (size_t q = 0; q < order.size() && !failed; ++q) {
            std::vector<int> v = order[q];
            // long-lived members first: the rows that stay to the end keep one common register
            std::stable_sort(v.begin(), v.end(), [&](int a, int b) {
                if (death[a] != death[b]) return death[a] > death[b];
                return birth[a] < birth[b];
            });
            while (!v.empty() && !failed) {
                // members that can share one register: no two overlapping in one lane
                std::vector<int> grp, left;
                for (int e : v) {
                    bool clash = false;
                    for (int f : grp)
                        if (ents[f].lane == ents[e].lane && birth[e] <= death[f] && birth[f] <= death[e]) clash = true;
                    (clash ? left : grp).push_back(e);
                }
                int p = first_reg;
                for (; p < kMaxRegsHard; ++p) {
                    bool ok = true;
                    for (int e : grp)
                        if (!is_free(p, ents[e].lane, birth[e], death[e])) {
                            ok = false;
                            break;
                        }
                    if (ok) break;
                }
                if (p >= kMaxRegsHard) {
                    fail("out of value registers");
                    break;
                }
                for (int e : grp) {
                    busy[p][ents[e].lane].push_back({birth[e], death[e]});
                    ents[e].reg = p;
                }
                Pmax = std::max(Pmax, p + 1);
                if (std::getenv("KLUJAX_CUDA_RR_DEBUG") && p >= first_reg + 40) {
                    std::fprintf(stderr, "  class -> reg %d:", p);
                    for (int e : grp) std::fprintf(stderr, " (%d,%d l%d %d..%d)", ents[e].row, ents[e].col, ents[e].lane, birth[e], death[e] == kInf ? -1 : death[e]);
                    std::fprintf(stderr, "\n");
                }
                v.swap(left);
            }
        }
    }
    if (failed) return;
    if (std::getenv("KLUJAX_CUDA_RR_DEBUG") && nb > kLanes) {
        // registers in use over time, and what they hold at the busiest moment
        const int TT = 2 * nb + 6;
        int bt = 0, bc = 0;
        for (int t = 0; t < TT; ++t) {
            std::vector<char> used(kMaxRegsHard, 0);
            for (size_t e = ent0; e < ents.size(); ++e)
                if (ents[e].col < n && birth[e] <= t && death[e] >= t) used[ents[e].reg] = 1;
            int c = 0;
            for (char u : used) c += u;
            if (c > bc) bc = c, bt = t;
        }
        std::fprintf(stderr, "registers in use at the busiest time %d: %d (+%d rhs)\n", bt, bc, first_reg);
        for (int p = first_reg; p < Pmax; ++p) {
            std::map<int, int> cols;
            int cnt = 0;
            for (size_t e = ent0; e < ents.size(); ++e)
                if (ents[e].col < n && ents[e].reg == p && birth[e] <= bt && death[e] >= bt) cols[ents[e].col]++, cnt++;
            std::fprintf(stderr, "  reg %2d: %2d lanes:", p, cnt);
            for (auto [c, k] : cols) std::fprintf(stderr, " col%d x%d", c, k);
            std::fprintf(stderr, "\n");
        }
    }
    // (entries of A become alive when their load instruction is emitted)
    // `take` of an entry that is born during the elimination: the register is already chosen
    auto born_now = [&](int e) { ents[e].alive = true; };

    // --- shared-memory slots ------------------------------------------------------------------
    // Entries of A wait in the pool from the staging until the epoch that first touches them
    // (one load instruction per (epoch, register): its entries share a plane, address =
    // immediate + lane); parked entries (k,j) sit at [plane][lane of row k] from their dump to the
    // end of the block, all parked entries of one column in one plane (one load instruction per
    // column in the back substitution).  Slots are reused over time.
    const int block_end = time_base + 2 * nb + 6;
    std::vector<std::map<std::pair<int, int>, uint32_t>> loads(epochs.size());  // epoch -> (reg, plane) -> lanes
    {
        std::map<std::pair<int, int>, std::vector<int>> grp;  // (epoch, reg) -> entries
        for (size_t e = ent0; e < ents.size(); ++e)
            if (ents[e].col < n && is_a(static_cast<int>(e))) grp[{load_ep[e], ents[e].reg}].push_back(static_cast<int>(e));
        for (auto& [key2, v] : grp) {
            std::vector<int> lanes;
            std::vector<std::pair<int, int>> iv;
            for (int e : v) {
                lanes.push_back(ents[e].lane);
                iv.push_back({0, time_base + birth[e]});
            }
            const int pi = pool_place(lanes, iv);
            uint32_t m = 0;
            for (int e : v) {
                place[ents[e].csc] = static_cast<uint16_t>(pi * kLanes + ents[e].lane);
                m |= 1u << ents[e].lane;
            }
            loads[key2.first][{key2.second, pi}] |= m;
        }
        // off-diagonal entries: read once at the start of the block, from the row's lane
        for (int i = k0; i < k1; ++i)
            for (auto [c, csc] : offs[i - k0]) {
                const int pi = pool_place({lane_of[i]}, {{0, time_base}});
                place[csc] = static_cast<uint16_t>(pi * kLanes + lane_of[i]);
            }
        auto park = [&](const std::vector<int>& es, int from) {
            std::vector<int> rest = es;
            while (!rest.empty()) {
                uint32_t m = 0;
                std::vector<int> g2, left, lanes;
                std::vector<std::pair<int, int>> iv;
                for (int e : rest) {
                    const uint32_t bit = 1u << lane_of[ents[e].row];
                    if (m & bit) {
                        left.push_back(e);
                        continue;
                    }
                    m |= bit;
                    g2.push_back(e);
                    lanes.push_back(lane_of[ents[e].row]);
                    iv.push_back({time_base + (from >= 0 ? from : epoch_end[epoch_of[ents[e].row - k0]]), block_end});
                }
                const int pi = pool_place(lanes, iv);
                for (int e : g2) ents[e].slot = pi * kLanes + lane_of[ents[e].row];
                rest.swap(left);
            }
        };
        for (int j = k1 - 1; j >= k0; --j) {
            std::vector<int> es;
            for (int k : Ucolrows[j - k0])
                if (!pinned[k - k0]) es.push_back(find(k, j));
            if (!es.empty()) park(es, -1);
        }
        for (const Epoch& ep : epochs) {
            std::vector<int> es;
            for (int k : ep.piv)
                if (!pinned[k - k0]) es.push_back(find(k, k));
            if (!es.empty()) park(es, -1);
        }
    }

    // --- loads: right-hand sides, off-diagonal blocks, entries of A ---
    for (int g = 0; g < G; ++g)
        for (int c = 0; c < r; ++c) {
            uint16_t tb[kLanes], ts[kLanes];
            std::fill(tb, tb + kLanes, kNone);
            std::fill(ts, ts + kLanes, kNone);
            for (int i = k0; i < k1; ++i)
                if (g_of[i] == g) {
                    tb[lane_of[i]] = static_cast<uint16_t>(rhs_src[i] * r + c);
                    if (rhs_scale[i] >= 0) ts[lane_of[i]] = static_cast<uint16_t>(rhs_scale[i]);
                }
            Ins in;
            in.op = OP_LOADB;
            in.a = static_cast<int16_t>(yr(g, c));
            in.tab = add_tab(tb);
            in.tab2 = add_tab(ts);
            emit(in);
        }
    {
        // off-diagonal entries of a row: rhs -= A(i,j) * x(j), j in a block solved before
        size_t maxoff = 0;
        for (auto& v : offs) maxoff = std::max(maxoff, v.size());
        for (size_t q = 0; q < maxoff; ++q)
            for (int g = 0; g < G; ++g) {
                bool any = false;
                uint16_t ta[kLanes];
                std::fill(ta, ta + kLanes, kNone);
                std::vector<uint16_t> tx(static_cast<size_t>(kLanes) * r, kNone);
                for (int i = k0; i < k1; ++i)
                    if (g_of[i] == g && offs[i - k0].size() > q) {
                        any = true;
                        ta[lane_of[i]] = static_cast<uint16_t>(offs[i - k0][q].second);
                        for (int c = 0; c < r; ++c)
                            tx[static_cast<size_t>(c) * kLanes + lane_of[i]] = static_cast<uint16_t>(offs[i - k0][q].first * r + c);
                    }
                if (!any) continue;
                const int tabA = add_tab(ta);
                csc_tabs.push_back(tabA);
                for (int c = 0; c < r; ++c) {
                    Ins in;
                    in.op = OP_OFFMAC;
                    in.a = static_cast<int16_t>(yr(g, c));
                    in.tab = tabA;
                    in.tab2 = add_tab(tx.data() + static_cast<size_t>(c) * kLanes);
                    emit(in);
                }
                pl->need_xbuf = true;
            }
    }
    // --- elimination, epoch by epoch ---
    auto zero_one = [&](int e) {  // structural entry nothing wrote before its first read
        born_now(e);
        Ins z;
        z.op = OP_ZERO;
        z.a = static_cast<int16_t>(ents[e].reg);
        z.mask = 1u << ents[e].lane;
        emit(z);
    };
    for (size_t epi = 0; epi < epochs.size(); ++epi) {
        const Epoch& ep = epochs[epi];
        for (auto& [rp, m] : loads[epi]) {
            Ins in;
            in.op = OP_LOADV;
            in.a = static_cast<int16_t>(rp.first);
            in.t = rp.second;
            in.mask = m;
            emit(in);
            pl->st.loadv++;
        }
        for (size_t e = ent0; e < ents.size(); ++e)
            if (ents[e].col < n && load_ep[e] == static_cast<int>(epi)) ents[e].alive = true;
        for (int k : ep.piv) {
            const int ed = find(k, k);
            if (ed < 0) {
                fail("missing diagonal entry");
                return;
            }
            if (!ents[ed].alive) zero_one(ed);
            const int tinv = pl->n_inv++;
            Ins in;
            in.op = OP_PIVOT;
            in.a = static_cast<int16_t>(ents[ed].reg);
            in.b = static_cast<int16_t>(lane_of[k]);
            in.t = tinv;
            emit(in);
            in.op = OP_STOREINV;
            emit(in);
            pl->st.pivots++;
            const auto& I = Lrows[k - k0];
            if (!I.empty()) {
                std::map<int, uint32_t> lgroups;
                for (int i : I) {
                    const int e = find(i, k);
                    if (!ents[e].alive) zero_one(e);
                    lgroups[ents[e].reg] |= 1u << lane_of[i];
                }
                for (auto [p, m] : lgroups) {
                    Ins lm;
                    lm.op = OP_LMUL;
                    lm.a = static_cast<int16_t>(p);
                    lm.mask = m;
                    lm.t = tinv;
                    emit(lm);
                    pl->st.lmul++;
                }
                std::vector<int> cols = Ucols[k - k0];
                for (int c = 0; c < r; ++c) cols.push_back(n + c);
                // entries that were moved away from their row's lane get the multiplier by shuffle
                struct Spread {
                    int reg;
                    int src[kLanes];
                    int row[kLanes];
                    int t;
                };
                std::vector<Spread> spreads;
                std::map<std::pair<int, int>, int> spread_of;  // (lane, row) -> spread
                for (int j : cols)
                    for (int i : I) {
                        const int e = find(i, j);
                        if (e < 0) {
                            fail("pattern is not closed under elimination");
                            return;
                        }
                        const int l2 = ents[e].lane;
                        if (l2 == lane_of[i] || spread_of.count({l2, i})) continue;
                        const int pl_ = ents[find(i, k)].reg;
                        int g = -1;
                        for (size_t q = 0; q < spreads.size(); ++q)
                            if (spreads[q].reg == pl_ && spreads[q].row[l2] < 0) {
                                g = static_cast<int>(q);
                                break;
                            }
                        if (g < 0) {
                            Spread sp;
                            sp.reg = pl_;
                            for (int l = 0; l < kLanes; ++l) sp.src[l] = l, sp.row[l] = -1;
                            sp.t = pl->n_l++;
                            spreads.push_back(sp);
                            g = static_cast<int>(spreads.size()) - 1;
                        }
                        spreads[g].src[l2] = lane_of[i];
                        spreads[g].row[l2] = i;
                        spread_of[{l2, i}] = g;
                    }
                for (const Spread& sp : spreads) {
                    uint16_t tb[kLanes];
                    for (int l = 0; l < kLanes; ++l) tb[l] = static_cast<uint16_t>(sp.src[l]);
                    Ins si;
                    si.op = OP_LSPREAD;
                    si.a = static_cast<int16_t>(sp.reg);
                    si.t = sp.t;
                    si.tab = add_tab(tb);
                    emit(si);
                    pl->st.spread++;
                }
                for (int j : cols) {
                    const int es = find(k, j);
                    if (es < 0) {
                        fail("missing pivot-row entry");
                        return;
                    }
                    if (!ents[es].alive) zero_one(es);
                    const int tu = pl->n_u++;
                    Ins bc;
                    bc.op = OP_BCAST;
                    bc.a = static_cast<int16_t>(ents[es].reg);
                    bc.b = static_cast<int16_t>(ents[es].lane);
                    bc.t = tu;
                    emit(bc);
                    pl->st.bcast++;
                    // fill-in born here starts from zero
                    std::map<int, uint32_t> zgroups;
                    for (int i : I) {
                        const int e = find(i, j);
                        if (!ents[e].alive) {
                            born_now(e);
                            zgroups[ents[e].reg] |= 1u << ents[e].lane;
                        }
                    }
                    for (auto [p, m] : zgroups) {
                        Ins z;
                        z.op = OP_ZERO;
                        z.a = static_cast<int16_t>(p);
                        z.mask = m;
                        emit(z);
                        pl->st.zero++;
                    }
                    // (destination register, multiplier register or - spread temporary - 1) -> lanes
                    std::map<std::pair<int, int>, uint32_t> mgroups;
                    for (int i : I) {
                        const int e = find(i, j), el = find(i, k);
                        const int l2 = ents[e].lane;
                        if (l2 == lane_of[i])
                            mgroups[{ents[e].reg, ents[el].reg}] |= 1u << l2;
                        else
                            mgroups[{ents[e].reg, -1 - spreads[spread_of[{l2, i}]].t}] |= 1u << l2;
                    }
                    for (auto& [pr, m] : mgroups) {
                        Ins mc;
                        mc.op = OP_MAC;
                        mc.a = static_cast<int16_t>(pr.first);
                        if (pr.second >= 0)
                            mc.b = static_cast<int16_t>(pr.second);
                        else
                            mc.b = -1, mc.t2 = -1 - pr.second;
                        mc.mask = m;
                        mc.t = tu;
                        emit(mc);
                        pl->st.mac_ins++;
                        pl->st.useful_mac += __builtin_popcount(m);
                    }
                }
                for (int i : I) ents[find(i, k)].alive = false;
            }
            if (std::getenv("KLUJAX_CUDA_RR_DEBUG"))
                std::fprintf(stderr, "pivot %3d lane %2d lev %2d | Lrows %2zu Ucols %2zu | mac instr so far %d (ideal %d) P so far %d\n", k,
                             lane_of[k], lev[k - k0], I.size(), Ucols[k - k0].size(), pl->st.mac_ins, pl->st.bcast, Pmax);
            // the row is final: U(k,:), the reciprocal pivot
            ents[ed].frozen = true;
            for (int j : Ucols[k - k0]) {
                const int e = find(k, j);
                if (!ents[e].alive) zero_one(e);
                ents[e].frozen = true;
            }
        }
        // park the rows of this epoch that are not pinned
        std::map<std::pair<int, int>, uint32_t> direct;  // (register, plane) -> lanes, entry in its row's lane
        std::map<int, std::vector<int>> moved;           // register -> entries that live in another lane
        for (int k : ep.piv) {
            if (pinned[k - k0]) continue;
            std::vector<int> es;
            es.push_back(find(k, k));
            for (int j : Ucols[k - k0]) es.push_back(find(k, j));
            for (int e : es) {
                if (ents[e].lane == lane_of[k])
                    direct[{ents[e].reg, ents[e].slot / kLanes}] |= 1u << lane_of[k];
                else
                    moved[ents[e].reg].push_back(e);
                ents[e].alive = false;
            }
        }
        for (auto& [rp, m] : direct) {
            Ins d;
            d.op = OP_DUMP;
            d.a = static_cast<int16_t>(rp.first);
            d.t = rp.second;
            d.mask = m;
            emit(d);
            pl->st.dump++;
        }
        for (auto& [p, v] : moved) {
            uint16_t tb[kLanes];
            std::fill(tb, tb + kLanes, kNone);
            uint32_t m = 0;
            for (int e : v) {
                tb[ents[e].lane] = static_cast<uint16_t>(ents[e].slot);
                m |= 1u << ents[e].lane;
            }
            Ins d;
            d.op = OP_DUMPT;
            d.a = static_cast<int16_t>(p);
            d.mask = m;
            d.tab = add_tab(tb);
            emit(d);
            pl->st.dump++;
        }
    }
    if (failed) return;

    // --- back substitution, epochs in reverse ---
    {
        Ins s;
        s.op = OP_SYNC;
        emit(s);
    }
    for (int ei = static_cast<int>(epochs.size()) - 1; ei >= 0; --ei) {
        const Epoch& ep = epochs[ei];
        // x of the rows of this epoch
        std::vector<std::vector<int>> xt(nb, std::vector<int>(r, -1));  // X temporary per (row, c)
        // (g, diag register, or -1 - plane for a parked row) -> rows
        std::map<std::pair<int, int>, std::vector<int>> groups;
        for (int k : ep.piv)
            groups[{g_of[k], pinned[k - k0] ? ents[find(k, k)].reg : -1 - ents[find(k, k)].slot / kLanes}].push_back(k);
        for (auto& [gk, v] : groups) {
            uint32_t m = 0;
            for (int k : v) m |= 1u << lane_of[k];
            for (int c = 0; c < r; ++c) {
                const int tx = pl->n_x++;
                Ins x;
                x.a = static_cast<int16_t>(yr(gk.first, c));
                x.t = tx;
                x.mask = m;
                if (gk.second >= 0) {
                    x.op = OP_XMUL;
                    x.b = static_cast<int16_t>(gk.second);
                } else {
                    x.op = OP_XMUL_S;
                    x.t2 = -1 - gk.second;
                }
                emit(x);
                pl->st.xmul++;
                Ins sx;
                sx.op = OP_SETX;
                sx.a = x.a;
                sx.t = tx;
                sx.mask = m;
                emit(sx);
                for (int k : v) xt[k - k0][c] = tx;
            }
        }
        // apply x(j) to the rows above: y(k) -= U(k,j) x(j)
        for (int j : ep.piv) {
            const auto& K = Ucolrows[j - k0];
            if (K.empty()) continue;
            for (int c = 0; c < r; ++c) {
                const int tu = pl->n_u++;
                Ins bc;
                bc.op = OP_BCASTX;
                bc.t = xt[j - k0][c];
                bc.t2 = tu;
                bc.b = static_cast<int16_t>(lane_of[j]);
                emit(bc);
                pl->st.bcast++;
                std::map<std::pair<int, int>, uint32_t> rg;  // pinned rows: (y reg, U reg)
                std::map<int, std::vector<int>> sg;          // parked rows: y reg -> entries
                for (int k : K) {
                    const int e = find(k, j);
                    if (pinned[k - k0])
                        rg[{yr(g_of[k], c), ents[e].reg}] |= 1u << lane_of[k];
                    else
                        sg[yr(g_of[k], c)].push_back(e);
                }
                for (auto& [pr, m] : rg) {
                    Ins mc;
                    mc.op = OP_MAC;
                    mc.a = static_cast<int16_t>(pr.first);
                    mc.b = static_cast<int16_t>(pr.second);
                    mc.mask = m;
                    mc.t = tu;
                    emit(mc);
                    pl->st.mac_ins++;
                    pl->st.useful_mac += __builtin_popcount(m);
                }
                for (auto& [py, v] : sg) {
                    std::map<int, uint32_t> byplane;
                    for (int e : v) byplane[ents[e].slot / kLanes] |= 1u << lane_of[ents[e].row];
                    for (auto [pi, m] : byplane) {
                        Ins mc;
                        mc.op = OP_MAC_S;
                        mc.a = static_cast<int16_t>(py);
                        mc.mask = m;
                        mc.t = tu;
                        mc.t2 = pi;
                        emit(mc);
                        pl->st.mac_s++;
                        pl->st.useful_mac += __builtin_popcount(m);
                    }
                }
            }
        }
    }
    // --- results of this block ---
    for (int g = 0; g < G; ++g)
        for (int c = 0; c < r; ++c) {
            uint16_t tb[kLanes], ts[kLanes], tx[kLanes];
            std::fill(tb, tb + kLanes, kNone);
            std::fill(ts, ts + kLanes, kNone);
            std::fill(tx, tx + kLanes, kNone);
            for (int i = k0; i < k1; ++i)
                if (g_of[i] == g) {
                    tb[lane_of[i]] = static_cast<uint16_t>(out_dst[i] * r + c);
                    if (out_scale[i] >= 0) ts[lane_of[i]] = static_cast<uint16_t>(out_scale[i]);
                    tx[lane_of[i]] = static_cast<uint16_t>(i * r + c);
                }
            Ins in;
            in.op = OP_STOREX;
            in.a = static_cast<int16_t>(yr(g, c));
            in.tab = add_tab(tb);
            in.tab2 = add_tab(ts);
            emit(in);
            Ins xb;
            xb.op = OP_XBUF;  // dropped at the end if no block needs it
            xb.a = in.a;
            xb.tab = add_tab(tx);
            emit(xb);
        }
    {
        Ins s;
        s.op = OP_SYNC;
        emit(s);
    }
    time_base = block_end + 2;
}

}  // namespace

bool build_plan(const CscPattern& A, const Ordering& ord, const PivotPlan& piv, const SlotLayout& lay,
                bool is_complex, const Options& opt, Plan& out) {
    out = Plan();
    Builder B;
    B.pl = &out;
    B.opt = opt;
    const int n = B.n = lay.n;
    B.nz = static_cast<int>(A.rowidx.size());
    B.r = std::max(1, opt.r);
    B.transpose = opt.transpose;
    out.n = n;
    out.nz = B.nz;
    out.r = B.r;
    out.is_complex = is_complex;
    out.transpose = opt.transpose;
    if (n <= 0 || n * B.r >= kNone || B.nz >= kNone || n > 4096) {
        out.why = "matrix too large for 16-bit tables";
        return false;
    }
    B.rows.assign(n, {});
    for (int s = 0; s < lay.n_val; ++s) {
        int rw = lay.slot_row[s], cl = lay.slot_col[s];
        if (opt.transpose) std::swap(rw, cl);
        B.rows[rw].push_back({cl, lay.slot_csc[s]});
    }
    B.R.assign(ord.R.begin(), ord.R.begin() + ord.nblocks + 1);
    B.blk.assign(n, 0);
    for (int b = 0; b < ord.nblocks; ++b)
        for (int k = B.R[b]; k < B.R[b + 1]; ++k) B.blk[k] = b;
    B.rhs_src.resize(n);
    B.rhs_scale.resize(n);
    B.out_dst.resize(n);
    B.out_scale.resize(n);
    for (int k = 0; k < n; ++k) {
        if (!opt.transpose) {  // klu_solve: X[k] = B[Pnum[k]] / Rs ; x[Q[k]] = X[k]
            B.rhs_src[k] = piv.Pnum[k];
            B.rhs_scale[k] = piv.Pnum[k];
            B.out_dst[k] = ord.Q[k];
            B.out_scale[k] = -1;
        } else {  // klu_tsolve: X[k] = B[Q[k]] ; x[Pnum[k]] = X[k] / Rs
            B.rhs_src[k] = ord.Q[k];
            B.rhs_scale[k] = -1;
            B.out_dst[k] = piv.Pnum[k];
            B.out_scale[k] = piv.Pnum[k];
        }
    }
    B.occ.assign(kMaxRegsHard, 0u);
    B.lane_of.assign(n, 0);
    B.g_of.assign(n, 0);
    B.stack_depth.assign(kLanes, 0);
    B.place.assign(B.nz, kNone);

    // row scale factors: original row i is handled by lane i % 32, lane-row q = i / 32
    {
        std::vector<std::vector<int>> arow(n);
        for (int j = 0; j < n; ++j)
            for (int p = A.colptr[j]; p < A.colptr[j + 1]; ++p) arow[A.rowidx[p]].push_back(p);
        const int Q = (n + kLanes - 1) / kLanes;
        for (int q = 0; q < Q; ++q) {
            size_t ml = 0;
            for (int i = q * kLanes; i < std::min(n, (q + 1) * kLanes); ++i) ml = std::max(ml, arow[i].size());
            if (ml == 0) ml = 1;
            int first = -1;
            for (size_t e = 0; e < ml; ++e) {
                uint16_t tb[kLanes];
                std::fill(tb, tb + kLanes, kNone);
                for (int i = q * kLanes; i < std::min(n, (q + 1) * kLanes); ++i)
                    if (arow[i].size() > e) tb[i - q * kLanes] = static_cast<uint16_t>(arow[i][e]);
                const int id = B.add_tab(tb);
                B.csc_tabs.push_back(id);
                if (first < 0) first = id;
            }
            Ins in;
            in.op = OP_ROWSCALE;
            in.a = static_cast<int16_t>(first);
            in.tab = first;
            in.b = static_cast<int16_t>(ml);
            in.c = static_cast<int16_t>(q);
            if (first > 32767 || ml > 32767) {
                out.why = "row too long";
                return false;
            }
            out.ins.push_back(in);
        }
        Ins s;
        s.op = OP_SYNC;
        out.ins.push_back(s);
    }
    // blocks: last first (klu_solve), first first for the transposed system (klu_tsolve)
    for (int t = 0; t < ord.nblocks && !B.failed; ++t) {
        const int b = opt.transpose ? t : ord.nblocks - 1 - t;
        B.plan_block(B.R[b], B.R[b + 1]);
    }
    if (B.failed) {
        out.why = B.why;
        return false;
    }
    if (!out.need_xbuf) {
        std::vector<Ins> keep;
        for (const Ins& i : out.ins)
            if (i.op != OP_XBUF) keep.push_back(i);
        out.ins.swap(keep);
    }
    // tables of CSC positions -> slots of the pool
    for (int id : B.csc_tabs)
        for (int l = 0; l < kLanes; ++l) {
            uint16_t& v = out.tabs[static_cast<size_t>(id) * kLanes + l];
            if (v != kNone) v = B.place[v];
        }
    out.place = B.place;
    for (uint16_t v : out.place)
        if (v == kNone) {
            out.why = "internal: entry of A without a slot";
            return false;
        }
    out.P = B.Pmax;
    out.st.P = B.Pmax;
    out.planes = static_cast<int>(B.pool.size());
    out.st.planes = out.planes;
    if (out.P > opt.max_regs) {
        out.why = "needs " + std::to_string(out.P) + " value registers";
        return false;
    }
    out.ok = true;
    return true;
}

// ---------------------------------------------------------------------------------------
// host interpreter (tests): executes the instruction list on 32 emulated lanes
// ---------------------------------------------------------------------------------------
namespace {
inline double mag2(double a) { return a * a; }
inline double mag2(const std::complex<double>& a) { return a.real() * a.real() + a.imag() * a.imag(); }
inline double amax(double a) { return std::fabs(a); }
inline double amax(const std::complex<double>& a) { return std::max(std::fabs(a.real()), std::fabs(a.imag())); }
inline bool finite_(double a) { return std::isfinite(a); }
inline bool finite_(const std::complex<double>& a) { return std::isfinite(a.real()) && std::isfinite(a.imag()); }
inline double scale_by(double a, double s) { return a * s; }
inline std::complex<double> scale_by(const std::complex<double>& a, double s) { return {a.real() * s, a.imag() * s}; }
// complex multiply the way the kernel does it (four fused multiply-adds, no library call)
inline double cmul(double a, double b) { return a * b; }
inline std::complex<double> cmul(const std::complex<double>& a, const std::complex<double>& b) {
    return {a.real() * b.real() - a.imag() * b.imag(), a.real() * b.imag() + a.imag() * b.real()};
}
inline double recip_(double a) { return 1.0 / a; }
inline std::complex<double> recip_(const std::complex<double>& a) {
    const double s = std::max(std::fabs(a.real()), std::fabs(a.imag()));
    if (s == 0.0 || !std::isfinite(s)) return {1.0 / 0.0, 0.0};
    const double xr = a.real() / s, xi = a.imag() / s, m = xr * xr + xi * xi;
    return {xr / (m * s), -xi / (m * s)};
}
}  // namespace

template <typename T>
int emulate(const Plan& p, const T* Ax_csc, const T* b, T* x, double* growth) {
    const int n = p.n, r = p.r;
    std::vector<T> Ubuf(static_cast<size_t>(std::max(1, p.planes)) * kLanes, T(0));
    for (int q = 0; q < p.nz; ++q) Ubuf[p.place[q]] = Ax_csc[q];
    std::vector<T>& Abuf = Ubuf;  // one pool: entries of A until they are loaded, parked values later
    std::vector<T> xbuf(static_cast<size_t>(n) * r, T(0));
    std::vector<double> Rs(n, 1.0);
    std::vector<std::vector<T>> Rg(p.P, std::vector<T>(kLanes, T(0)));
    std::vector<std::vector<T>> U(std::max(1, p.n_u), std::vector<T>(kLanes, T(0)));
    std::vector<std::vector<T>> X(std::max(1, p.n_x), std::vector<T>(kLanes, T(0)));
    std::vector<T> INV(std::max(1, p.n_inv), T(0));
    std::vector<std::vector<T>> Lt(std::max(1, p.n_l), std::vector<T>(kLanes, T(0)));
    bool bad = false;
    double g2 = 0.0;
    auto tab = [&](int id, int l) { return p.tabs[static_cast<size_t>(id) * kLanes + l]; };
    for (const Ins& in : p.ins) {
        switch (in.op) {
            case OP_ROWSCALE:
                for (int l = 0; l < kLanes; ++l) {
                    const int i = in.c * kLanes + l;
                    if (i >= n) continue;
                    double c = 0.0;
                    for (int e = 0; e < in.b; ++e) {
                        const uint16_t s = tab(in.tab + e, l);
                        if (s != kNone) c = std::max(c, amax(Abuf[s]));
                    }
                    double rs = 1.0;
                    if (c > 0.0 && std::isfinite(c)) {
                        const double ic = 1.0 / c;
                        double m = 0.0;
                        for (int e = 0; e < in.b; ++e) {
                            const uint16_t s = tab(in.tab + e, l);
                            if (s != kNone) m = std::max(m, mag2(scale_by(Abuf[s], ic)));
                        }
                        rs = c * std::sqrt(m);
                    }
                    Rs[i] = rs;
                    const double ir = 1.0 / rs;
                    for (int e = 0; e < in.b; ++e) {
                        const uint16_t s = tab(in.tab + e, l);
                        if (s != kNone) Abuf[s] = scale_by(Abuf[s], ir);
                    }
                }
                break;
            case OP_LOADV:
                for (int l = 0; l < kLanes; ++l)
                    if (in.mask >> l & 1u) Rg[in.a][l] = Abuf[static_cast<size_t>(in.t) * kLanes + l];
                break;
            case OP_LOADB:
                for (int l = 0; l < kLanes; ++l) {
                    const uint16_t s = tab(in.tab, l), q = tab(in.tab2, l);
                    T v = s == kNone ? T(0) : b[s];
                    if (q != kNone) v = v / Rs[q];
                    Rg[in.a][l] = v;
                }
                break;
            case OP_OFFMAC:
                for (int l = 0; l < kLanes; ++l) {
                    const uint16_t s = tab(in.tab, l), q = tab(in.tab2, l);
                    if (s != kNone && q != kNone) Rg[in.a][l] -= cmul(Abuf[s], xbuf[q]);
                }
                break;
            case OP_PIVOT: {
                const T d = Rg[in.a][in.b];
                if (d == T(0) || !finite_(d)) bad = true;
                INV[in.t] = recip_(d);
                break;
            }
            case OP_STOREINV:
                Rg[in.a][in.b] = INV[in.t];
                break;
            case OP_LMUL:
                for (int l = 0; l < kLanes; ++l)
                    if (in.mask >> l & 1u) {
                        Rg[in.a][l] = cmul(Rg[in.a][l], INV[in.t]);
                        g2 = std::max(g2, mag2(Rg[in.a][l]));
                    }
                break;
            case OP_BCAST:
                for (int l = 0; l < kLanes; ++l) U[in.t][l] = Rg[in.a][in.b];
                break;
            case OP_ZERO:
                for (int l = 0; l < kLanes; ++l)
                    if (in.mask >> l & 1u) Rg[in.a][l] = T(0);
                break;
            case OP_LSPREAD:
                for (int l = 0; l < kLanes; ++l) Lt[in.t][l] = Rg[in.a][tab(in.tab, l)];
                break;
            case OP_MAC:
                for (int l = 0; l < kLanes; ++l)
                    if (in.mask >> l & 1u) Rg[in.a][l] -= cmul(in.b >= 0 ? Rg[in.b][l] : Lt[in.t2][l], U[in.t][l]);
                break;
            case OP_DUMP:
                for (int l = 0; l < kLanes; ++l)
                    if (in.mask >> l & 1u) Ubuf[static_cast<size_t>(in.t) * kLanes + l] = Rg[in.a][l];
                break;
            case OP_DUMPT:
                for (int l = 0; l < kLanes; ++l)
                    if (in.mask >> l & 1u) Ubuf[tab(in.tab, l)] = Rg[in.a][l];
                break;
            case OP_XMUL:
                for (int l = 0; l < kLanes; ++l) X[in.t][l] = cmul(Rg[in.a][l], Rg[in.b][l]);
                break;
            case OP_XMUL_S:
                for (int l = 0; l < kLanes; ++l)
                    if (in.mask >> l & 1u) X[in.t][l] = cmul(Rg[in.a][l], Ubuf[static_cast<size_t>(in.t2) * kLanes + l]);
                break;
            case OP_SETX:
                for (int l = 0; l < kLanes; ++l)
                    if (in.mask >> l & 1u) Rg[in.a][l] = X[in.t][l];
                break;
            case OP_BCASTX:
                for (int l = 0; l < kLanes; ++l) U[in.t2][l] = X[in.t][in.b];
                break;
            case OP_MAC_S:
                for (int l = 0; l < kLanes; ++l)
                    if (in.mask >> l & 1u) Rg[in.a][l] -= cmul(Ubuf[static_cast<size_t>(in.t2) * kLanes + l], U[in.t][l]);
                break;
            case OP_STOREX:
                for (int l = 0; l < kLanes; ++l) {
                    const uint16_t s = tab(in.tab, l), q = tab(in.tab2, l);
                    if (s == kNone) continue;
                    T v = Rg[in.a][l];
                    if (q != kNone) v = v / Rs[q];
                    x[s] = v;
                }
                break;
            case OP_XBUF:
                for (int l = 0; l < kLanes; ++l) {
                    const uint16_t s = tab(in.tab, l);
                    if (s != kNone) xbuf[s] = Rg[in.a][l];
                }
                break;
            case OP_SYNC:
            default:
                break;
        }
    }
    if (growth) *growth = std::sqrt(g2);
    return bad ? 1 : 0;
}

template int emulate<double>(const Plan&, const double*, const double*, double*, double*);
template int emulate<std::complex<double>>(const Plan&, const std::complex<double>*, const std::complex<double>*,
                                           std::complex<double>*, double*);

}  // namespace rr
}  // namespace kb
