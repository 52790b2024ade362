// plan.cpp — builds the static schedule (see plan.hpp).
#include "plan.hpp"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <numeric>

namespace kb {

namespace {

std::vector<int> block_of_column(const Ordering& ord) {
    std::vector<int> k1(ord.n);
    for (int b = 0; b < ord.nblocks; ++b)
        for (int k = ord.R[b]; k < ord.R[b + 1]; ++k) k1[k] = ord.R[b];
    return k1;
}

}  // namespace

void build_layout(const CscPattern& A, const Ordering& ord, const PivotPlan& piv, SlotLayout& lay) {
    const int n = A.n;
    lay = SlotLayout();
    lay.n = n;
    const std::vector<int> k1of = block_of_column(ord);
    lay.col_begin.assign(n + 1, 0);
    lay.diag_slot.assign(n, 0);
    // slot order inside a column: U rows ascending, pivot, L rows ascending
    for (int k = 0; k < n; ++k) {
        const int k1 = k1of[k];
        std::vector<int> ur(piv.Ui.begin() + piv.Up[k], piv.Ui.begin() + piv.Up[k + 1]);
        std::vector<int> lr(piv.Li.begin() + piv.Lp[k], piv.Li.begin() + piv.Lp[k + 1]);
        std::sort(ur.begin(), ur.end());
        std::sort(lr.begin(), lr.end());
        lay.col_begin[k] = static_cast<int>(lay.slot_row.size());
        for (int j : ur) {
            lay.slot_row.push_back(k1 + j);
            lay.slot_col.push_back(k);
        }
        lay.diag_slot[k] = static_cast<int>(lay.slot_row.size());
        lay.slot_row.push_back(k);
        lay.slot_col.push_back(k);
        for (int i : lr) {
            lay.slot_row.push_back(k1 + i);
            lay.slot_col.push_back(k);
        }
        lay.nnz_u += static_cast<int64_t>(ur.size());
        lay.nnz_l += static_cast<int64_t>(lr.size());
    }
    lay.n_lu = static_cast<int>(lay.slot_row.size());
    lay.col_begin[n] = lay.n_lu;
    lay.n_off = static_cast<int>(piv.Offi.size());
    lay.n_val = lay.n_lu + lay.n_off;
    lay.slot_csc.assign(lay.n_val, -1);
    for (int k = 0; k < n; ++k)
        for (int p = piv.Offp[k]; p < piv.Offp[k + 1]; ++p) {
            lay.slot_row.push_back(piv.Offi[p]);
            lay.slot_col.push_back(k);
            lay.slot_csc[lay.n_lu + p] = piv.Offsrc[p];
        }
    // where each entry of A lands
    std::vector<int> pos(n, -1);
    std::vector<int> csc_slot(A.rowidx.size(), -1);
    for (int p = 0; p < lay.n_off; ++p) csc_slot[piv.Offsrc[p]] = lay.n_lu + p;
    for (int k = 0; k < n; ++k) {
        const int k1 = k1of[k];
        for (int s = lay.col_begin[k]; s < lay.col_begin[k + 1]; ++s) pos[lay.slot_row[s]] = s;
        const int col = ord.Q[k];
        for (int p = A.colptr[col]; p < A.colptr[col + 1]; ++p) {
            const int prow = piv.Pinv[A.rowidx[p]];
            if (prow < k1) continue;  // off-diagonal block entry, already placed
            const int s = pos[prow];
            lay.slot_csc[s] = p;
            csc_slot[p] = s;
        }
        for (int s = lay.col_begin[k]; s < lay.col_begin[k + 1]; ++s) pos[lay.slot_row[s]] = -1;
    }
    // (col,row)-sorted copy of A's pattern for the per-call COO -> slot map
    lay.sorted_colptr = A.colptr;
    lay.sorted_row.resize(A.rowidx.size());
    lay.sorted_slot.resize(A.rowidx.size());
    std::vector<int> ordv;
    for (int j = 0; j < n; ++j) {
        const int b = A.colptr[j], e = A.colptr[j + 1];
        ordv.resize(e - b);
        std::iota(ordv.begin(), ordv.end(), b);
        std::sort(ordv.begin(), ordv.end(),
                  [&](int x, int y) { return A.rowidx[x] < A.rowidx[y]; });
        for (int t = 0; t < e - b; ++t) {
            lay.sorted_row[b + t] = A.rowidx[ordv[t]];
            lay.sorted_slot[b + t] = csc_slot[ordv[t]];
        }
    }
    // row lists for the max-abs row scale factors
    lay.rs_ptr.assign(n + 1, 0);
    for (size_t p = 0; p < A.rowidx.size(); ++p) lay.rs_ptr[A.rowidx[p] + 1]++;
    for (int i = 0; i < n; ++i) lay.rs_ptr[i + 1] += lay.rs_ptr[i];
    lay.rs_slot.resize(A.rowidx.size());
    std::vector<int> fill(lay.rs_ptr.begin(), lay.rs_ptr.end() - 1);
    for (size_t p = 0; p < A.rowidx.size(); ++p) lay.rs_slot[fill[A.rowidx[p]]++] = csc_slot[p];
}

namespace {

// Assigns levels while tasks are appended in a valid sequential order.
//
// split = false: one task per destination, at the level where its last operand is final
//                (pure "dot" form).
// split = true : the multiply-adds of a destination are applied as soon as their operands
//                are final: one partial task per (destination, level), the last one carrying
//                the finishing operation.  Same critical path, but the work of a long dot
//                product is spread over the levels before it and chains (dense trailing
//                blocks, back substitution) become one short update per level instead of
//                one long reduction per level.
struct Leveler {
    std::vector<int> val_level, read_level;
    TaskProgram& prog;
    bool split;
    bool single = false;
    int min_group = 1, max_group = 0;
    std::vector<std::pair<int, int>> order;  // (ready level, pair index)
    Leveler(int nidx, TaskProgram& p, bool split_) : val_level(nidx, 0), read_level(nidx, 0), prog(p), split(split_) {}

    void emit(uint8_t kind, int out, int div, int lvl, const std::vector<std::pair<int, int>>& pairs,
              int ob, int oe) {
        Task t;
        t.kind = kind;
        t.out = out;
        t.div = div;
        t.level = lvl;
        t.pbeg = static_cast<int>(prog.pa.size());
        t.plen = oe - ob;
        for (int q = ob; q < oe; ++q) {
            const auto& pr = pairs[order[q].second];
            prog.pa.push_back(pr.first);
            prog.pb.push_back(pr.second);
            read_level[pr.first] = std::max(read_level[pr.first], lvl);
            read_level[pr.second] = std::max(read_level[pr.second], lvl);
        }
        if (div >= 0) read_level[div] = std::max(read_level[div], lvl);
        prog.n_mac += t.plen;
        prog.tasks.push_back(t);
    }

    void add(uint8_t kind, int out, int div, const std::vector<std::pair<int, int>>& pairs) {
        // nothing may touch `out` before its previous version is final and has been read
        const int base = std::max(val_level[out], read_level[out]) + 1;
        const int np = static_cast<int>(pairs.size());
        order.resize(np);
        int last = base;
        for (int q = 0; q < np; ++q) {
            int lp = std::max(val_level[pairs[q].first], val_level[pairs[q].second]) + 1;
            lp = std::max(lp, base);
            order[q] = {lp, q};
            last = std::max(last, lp);
        }
        if (div >= 0) last = std::max(last, val_level[div] + 1);
        if (!split) {
            for (int q = 0; q < np; ++q) order[q].first = last;
        } else {
            std::stable_sort(order.begin(), order.end(),
                             [](const std::pair<int, int>& x, const std::pair<int, int>& y) { return x.first < y.first; });
            if (single) {  // at most one multiply-add per destination and level
                for (int q = 1; q < np; ++q) order[q].first = std::max(order[q].first, order[q - 1].first + 1);
                if (np) last = std::max(last, order[np - 1].first);
            }
        }
        int ob = 0;
        if (max_group > 0 && split) {
            // bounded groups (multi-warp shared-memory kernel): a partial update holds at most
            // max_group multiply-adds; it is cut where the next pair only becomes ready at a
            // later level (once min_group pairs are in), and a burst that is ready all at once
            // is chained over consecutive levels.  The last group carries the finish.
            int prev_level = 0;
            while (ob < np) {
                int oe = ob + 1;
                while (oe < np && oe - ob < max_group &&
                       (order[oe].first == order[oe - 1].first || oe - ob < min_group))
                    ++oe;
                const int lvl = std::max(order[oe - 1].first, prev_level + 1);
                if (oe == np) {
                    last = std::max(last, lvl);
                    break;
                }
                emit(TK_SUB, out, -1, lvl, pairs, ob, oe);
                prev_level = lvl;
                ob = oe;
            }
        } else
        while (ob < np) {
            // a partial update takes every pair that is ready at one level, and keeps taking
            // later ones until it holds at least min_group pairs
            int oe = ob;
            while (oe < np && (order[oe].first == order[ob].first || oe - ob < min_group)) ++oe;
            while (oe < np && order[oe].first == order[oe - 1].first) ++oe;
            if (oe == np) break;  // the last group carries the finishing operation
            emit(TK_SUB, out, -1, order[oe - 1].first, pairs, ob, oe);
            ob = oe;
        }
        // last group (possibly empty) + finish; it may have to wait for the pivot
        if (np == 0 && kind == TK_SUB) return;  // V[out] -= nothing
        emit(kind, out, div, last, pairs, ob, np);
        val_level[out] = last;
        read_level[out] = 0;
    }
};

}  // namespace

namespace {

// Step balancing for the multi-warp kernel.  A level of c tasks costs ceil(c / lanes) steps
// whatever the fill of its last step.  A partial update of the refactorization
// (V[out] -= V[a]*V[b] with a, b finished entries of L and U, which never change again) may run
// at any later level before the task that finishes its destination, as long as no other task
// of that destination sits there.  So the tasks in a level's last, partly filled step are
// deferred into the free lanes of a later level's last step whenever that removes a step.
void rebalance_levels(TaskProgram& prog, int n_val, int lanes) {
    const int nl = prog.nlevels;
    if (nl < 2) return;
    std::vector<int> final_level(n_val, 0);
    std::vector<std::vector<int>> by_level(nl + 1);
    for (size_t t = 0; t < prog.tasks.size(); ++t) {
        const Task& k = prog.tasks[t];
        by_level[k.level].push_back(static_cast<int>(t));
        if (k.out < n_val) final_level[k.out] = std::max(final_level[k.out], k.level);
    }
    // (destination, level) pairs in use: one writer per destination and level
    std::vector<std::vector<int>> dest_levels(n_val);
    for (const Task& k : prog.tasks)
        if (k.out < n_val) dest_levels[k.out].push_back(k.level);
    auto busy = [&](int out, int lvl) {
        for (int l : dest_levels[out])
            if (l == lvl) return true;
        return false;
    };
    auto movable = [&](const Task& k) {
        if (k.kind != TK_SUB || k.out >= n_val || k.plen == 0) return false;
        for (int q = 0; q < k.plen; ++q)
            if (prog.pa[k.pbeg + q] >= n_val || prog.pb[k.pbeg + q] >= n_val) return false;
        return k.level < final_level[k.out];
    };
    const int horizon = 16;
    // moves exactly m tasks out of level l into free lanes of the last steps of later levels
    // (possibly several of them); all or nothing
    auto try_shrink = [&](int l, int m) {
        std::vector<std::pair<int, int>> plan;  // (index in by_level[l], target level)
        std::vector<char> taken(by_level[l].size(), 0);
        int placed = 0;
        for (int l2 = l + 1; l2 <= std::min(nl, l + horizon) && placed < m; ++l2) {
            const int used2 = static_cast<int>(by_level[l2].size()) % lanes;
            if (used2 == 0) continue;
            int room = lanes - used2;
            for (int idx = static_cast<int>(by_level[l].size()) - 1; idx >= 0 && room > 0 && placed < m; --idx) {
                if (taken[idx]) continue;
                const Task& k = prog.tasks[by_level[l][idx]];
                if (!movable(k) || l2 >= final_level[k.out] || busy(k.out, l2)) continue;
                taken[idx] = 1;
                plan.emplace_back(idx, l2);
                --room;
                ++placed;
            }
        }
        if (placed < m) return false;
        std::sort(plan.begin(), plan.end(), [](const std::pair<int, int>& x, const std::pair<int, int>& y) { return x.first > y.first; });
        for (const auto& mv : plan) {  // descending indices: erase is safe
            const int t = by_level[l][mv.first];
            Task& k = prog.tasks[t];
            for (int& dl : dest_levels[k.out])
                if (dl == l) dl = mv.second;
            k.level = mv.second;
            by_level[mv.second].push_back(t);
            by_level[l].erase(by_level[l].begin() + mv.first);
        }
        return true;
    };
    for (int l = 1; l <= nl; ++l) {  // trim partly filled EXTRA steps
        const int c = static_cast<int>(by_level[l].size());
        if (c >= lanes && c % lanes) try_shrink(l, c % lanes);
    }
    for (int pass = 0; pass < 4; ++pass) {  // then whole steps, while free lanes remain downstream
        bool any = false;
        for (int l = 1; l <= nl; ++l) {
            const int c = static_cast<int>(by_level[l].size());
            if (c >= 2 * lanes && c % lanes == 0) any |= try_shrink(l, lanes);
        }
        if (!any) break;
    }
    std::stable_sort(prog.tasks.begin(), prog.tasks.end(), [](const Task& a, const Task& b) {
        if (a.level != b.level) return a.level < b.level;
        return a.plen > b.plen;
    });
    prog.level_ptr.assign(nl + 1, 0);
    for (const Task& t : prog.tasks) prog.level_ptr[t.level]++;
    for (int l = 0; l < nl; ++l) prog.level_ptr[l + 1] += prog.level_ptr[l];
}

// Linear-scan slot allocation over the level sets (see TaskProgram::remap).
void compact_slots(const SlotLayout& lay, TaskProgram& prog, int bank_cols) {
    const int nv = lay.n_val, nl = prog.nlevels;
    auto in_tail = [&](int k) { return prog.tail_t0 >= 0 && k >= prog.tail_t0 && k < prog.tail_t0 + kTail; };
    const int NEVER = 1 << 30;
    std::vector<int> birth(nv, NEVER), death(nv, 0), first_write(nv, -1);
    for (int s = 0; s < nv; ++s)
        if (lay.slot_csc[s] >= 0) birth[s] = 0;  // scattered (and row-scaled) before level 1
    for (size_t ti = 0; ti < prog.tasks.size(); ++ti) {
        const Task& t = prog.tasks[ti];
        auto use = [&](int s, bool write) {
            if (s >= nv) return;
            if (birth[s] == NEVER) {
                // fill-in: born at its first write; a read before any write means the value
                // must be the zero of the load phase
                birth[s] = write ? t.level : 0;
                if (write) first_write[s] = static_cast<int>(ti);
            }
            death[s] = std::max(death[s], t.level);
        };
        for (int q = 0; q < t.plen; ++q) {
            use(prog.pa[t.pbeg + q], false);
            use(prog.pb[t.pbeg + q], false);
        }
        if (t.div >= 0) use(t.div, false);
        use(t.out, true);
    }
    for (int s = 0; s < nv; ++s)
        if (birth[s] == NEVER) birth[s] = 0;  // never touched by a task: only the load phase sees it
    // entries of the dense tail are read (and their U part rewritten) by the register stage
    for (int s = 0; s < lay.n_lu; ++s)
        if (in_tail(lay.slot_row[s]) && in_tail(lay.slot_col[s])) death[s] = std::max(death[s], prog.tail_level);
    // slots by birth level; slots that die at level d are free again from level d + 1
    std::vector<std::vector<int>> born(nl + 2), dies(nl + 2);
    for (int s = 0; s < nv; ++s) {
        born[std::min(birth[s], nl + 1)].push_back(s);
        dies[std::min(death[s], nl + 1)].push_back(s);
    }
    prog.remap.assign(nv, -1);
    // A slot keeps its shared-memory bank column (slot mod bank_cols): the layout numbers the
    // entries of a column consecutively, which is what keeps the lanes of a dense level in
    // different banks, and the packer's lane order was chosen for those columns.  Free ids are
    // kept per bank column (sorted descending: pop_back gives the smallest).
    const int G = std::max(1, bank_cols);
    std::vector<std::vector<int>> free_ids(G);
    int n_phys = 0;
    auto take = [&](int col) {
        std::vector<int>& f = free_ids[col];
        if (!f.empty()) {
            const int id = f.back();
            f.pop_back();
            return id;
        }
        while (n_phys % G != col) {  // ids skipped on the way become free ids of their columns
            std::vector<int>& h = free_ids[n_phys % G];
            h.insert(h.begin(), n_phys);  // larger than every id in the list: keeps it descending
            ++n_phys;
        }
        return n_phys++;
    };
    for (int l = 0; l <= nl + 1; ++l) {
        if (l >= 1) {
            for (int s : dies[l - 1]) free_ids[prog.remap[s] % G].push_back(prog.remap[s]);
            for (auto& f : free_ids) std::sort(f.begin(), f.end(), std::greater<int>());
        }
        for (int s : born[l]) {
            prog.remap[s] = take(s % G);
            if (l >= 1 && first_write[s] >= 0) prog.tasks[first_write[s]].fresh = 1;
        }
    }
    auto rn = [&](int s) { return s < nv ? prog.remap[s] : s - nv + n_phys; };
    for (Task& t : prog.tasks) {
        t.out = rn(t.out);
        if (t.div >= 0) t.div = rn(t.div);
    }
    for (int& s : prog.pa) s = rn(s);
    for (int& s : prog.pb) s = rn(s);
    prog.n_val = n_phys;
}

}  // namespace

int pick_tail(const Ordering& ord, const PivotPlan& piv, const SlotLayout& lay) {
    // the largest block; its last kTail pivots
    int best = -1, bsz = 0;
    for (int b = 0; b < ord.nblocks; ++b)
        if (ord.R[b + 1] - ord.R[b] > bsz) bsz = ord.R[b + 1] - ord.R[b], best = b;
    if (best < 0 || bsz < kTail + 8 || lay.n < 128) return -1;
    const int t0 = ord.R[best + 1] - kTail;
    for (int k = t0; k < t0 + kTail; ++k)
        if (!piv.diag_cand.empty() && piv.diag_cand[k] != -2) return -1;  // the stage certifies diagonal pivots only
    long long inside = 0;
    for (int s = 0; s < lay.n_lu; ++s)
        if (lay.slot_row[s] >= t0 && lay.slot_row[s] < t0 + kTail && lay.slot_col[s] >= t0 && lay.slot_col[s] < t0 + kTail) ++inside;
    if (inside * 10 < 6LL * kTail * kTail) return -1;  // not dense enough to pay for lock-step elimination
    for (int k = t0; k < t0 + kTail; ++k)  // every diagonal entry of the tail has a slot (always true)
        if (lay.diag_slot[k] < 0) return -1;
    return t0;
}

void build_program(const Ordering& ord, const PivotPlan& piv, const SlotLayout& lay, int mode,
                   int rc, TaskProgram& prog, int split, int compact, int tail_t0) {
    const int n = lay.n;
    prog = TaskProgram();
    if (mode != PM_FACTOR_SOLVE) tail_t0 = -1;
    prog.tail_t0 = tail_t0;
    auto in_tail = [&](int k) { return tail_t0 >= 0 && k >= tail_t0 && k < tail_t0 + kTail; };
    prog.mode = mode;
    prog.rc = rc;
    prog.n_val = lay.n_val;
    const bool do_factor = (mode == PM_FACTOR || mode == PM_FACTOR_SOLVE || mode == PM_FACTOR_TSOLVE);
    const bool do_solve = (mode == PM_FACTOR_SOLVE || mode == PM_SOLVE);
    const bool do_tsolve = (mode == PM_FACTOR_TSOLVE || mode == PM_TSOLVE);
    const int nx = (do_solve || do_tsolve) ? n * rc : 0;
    // split: 0 = dot form everywhere, 1 = as-soon-as-possible updates in the triangular
    // solves only, 2 = everywhere, 3 = everywhere in groups of at least 8 multiply-adds
    Leveler lv(lay.n_val + nx, prog, split >= 2);
    if (split == 3) lv.min_group = 8;
    if (split == 5) lv.min_group = 16;
    if (split == 6) lv.min_group = 32;
    if (split == 4) lv.single = true;  // single multiply-add tasks
    if (split >= 100) {                // 100 + 10*min + max: bounded groups (multi-warp shared-memory kernel)
        lv.min_group = (split - 100) / 10;
        lv.max_group = (split - 100) % 10;
    }
    const std::vector<int> k1of = block_of_column(ord);
    auto xs = [&](int k, int c) { return lay.n_val + k * rc + c; };
    std::vector<std::pair<int, int>> pairs;

    if (do_factor) {
        // Crout form of klu_refactor: every entry of column k is a dot product
        // of finished entries of L (rows) and of U(:,k).
        std::vector<int> pos(n, -1);
        std::vector<std::vector<std::pair<int, int>>> dest;
        for (int k = 0; k < n; ++k) {
            const int cb = lay.col_begin[k], ce = lay.col_begin[k + 1], dg = lay.diag_slot[k];
            if (ce - cb == 1) {
                pairs.clear();
                lv.add(TK_RECIP, dg, -1, pairs);
                continue;
            }
            for (int s = cb; s < ce; ++s) pos[lay.slot_row[s]] = s;
            dest.assign(ce - cb, {});
            for (int su = cb; su < dg; ++su) {
                const int j = lay.slot_row[su];  // U(j,k)
                if (in_tail(k) && in_tail(j)) continue;  // updates by tail pivots: register stage
                for (int sl = lay.diag_slot[j] + 1; sl < lay.col_begin[j + 1]; ++sl) {
                    const int d = pos[lay.slot_row[sl]];  // (i,k) with L(i,j) != 0
                    dest[d - cb].emplace_back(sl, su);
                }
            }
            for (int s = cb; s < ce; ++s) {
                const auto& pr = dest[s - cb];
                if (in_tail(k) && in_tail(lay.slot_row[s])) {
                    // entry of the dense tail: only the Schur-complement part; pivot, L scaling and
                    // the certificate happen in the register stage
                    if (!pr.empty()) lv.add(TK_SUB, s, -1, pr);
                    continue;
                }
                if (s < dg) {
                    if (!pr.empty()) lv.add(TK_SUB, s, -1, pr);
                } else if (s == dg) {
                    lv.add(TK_RECIP, s, -1, pr);
                } else {
                    uint8_t kind = TK_MUL_TRACK;
                    if (!piv.diag_cand.empty() && piv.diag_cand[k] != -2)  // off-diagonal static pivot (plan.hpp)
                        kind = lay.slot_row[s] == piv.diag_cand[k] ? TK_MUL_TRACK_DIAG : TK_MUL_TRACK_OFF;
                    lv.add(kind, s, dg, pr);
                }
            }
            for (int s = cb; s < ce; ++s) pos[lay.slot_row[s]] = -1;
        }
    }

    lv.split = split >= 1;
    if (do_solve) {
        // row-wise views of L, U and the off-diagonal blocks
        std::vector<std::vector<std::pair<int, int>>> lrow(n), urow(n), offrow(n);
        for (int k = 0; k < n; ++k) {
            for (int s = lay.col_begin[k]; s < lay.diag_slot[k]; ++s)
                urow[lay.slot_row[s]].emplace_back(s, k);
            for (int s = lay.diag_slot[k] + 1; s < lay.col_begin[k + 1]; ++s)
                lrow[lay.slot_row[s]].emplace_back(s, k);
        }
        for (int i = 0; i < n; ++i) std::reverse(urow[i].begin(), urow[i].end());
        for (int b = ord.nblocks - 1; b >= 0; --b)
            for (int k = ord.R[b]; k < ord.R[b + 1]; ++k)
                for (int p = piv.Offp[k]; p < piv.Offp[k + 1]; ++p)
                    offrow[piv.Offi[p]].emplace_back(lay.n_lu + p, k);
        // klu_solve: blocks last -> first; L then U inside a block
        for (int b = ord.nblocks - 1; b >= 0; --b) {
            const int k1 = ord.R[b], k2 = ord.R[b + 1];
            if (k2 - k1 == 1) {
                for (int c = 0; c < rc; ++c) {
                    pairs.clear();
                    for (auto& e : offrow[k1]) pairs.emplace_back(e.first, xs(e.second, c));
                    lv.add(TK_MUL, xs(k1, c), lay.diag_slot[k1], pairs);
                }
                continue;
            }
            for (int i = k1; i < k2; ++i) {
                if (offrow[i].empty() && lrow[i].empty()) continue;
                for (int c = 0; c < rc; ++c) {
                    pairs.clear();
                    for (auto& e : offrow[i]) pairs.emplace_back(e.first, xs(e.second, c));
                    for (auto& e : lrow[i])
                        if (!(in_tail(i) && in_tail(e.second))) pairs.emplace_back(e.first, xs(e.second, c));
                    lv.add(TK_SUB, xs(i, c), -1, pairs);
                }
            }
            if (tail_t0 >= k1 && tail_t0 < k2) {
                // everything the register stage reads is final below this level; what it writes
                // (x of the tail rows) is final at it
                int top = 0;
                for (const Task& t : prog.tasks) top = std::max(top, t.level);
                prog.tail_level = top + 1;
                for (int i = tail_t0; i < tail_t0 + kTail; ++i)
                    for (int c = 0; c < rc; ++c) lv.val_level[xs(i, c)] = prog.tail_level;
            }
            for (int i = k2 - 1; i >= k1; --i)
                for (int c = 0; c < rc; ++c) {
                    if (in_tail(i)) continue;  // back substitution of the tail: register stage
                    pairs.clear();
                    for (auto& e : urow[i]) pairs.emplace_back(e.first, xs(e.second, c));
                    lv.add(TK_MUL, xs(i, c), lay.diag_slot[i], pairs);
                }
        }
    }

    if (do_tsolve) {
        // klu_tsolve: blocks first -> last; U' then L' inside a block
        for (int b = 0; b < ord.nblocks; ++b) {
            const int k1 = ord.R[b], k2 = ord.R[b + 1];
            for (int k = k1; k < k2; ++k)
                for (int c = 0; c < rc; ++c) {
                    pairs.clear();
                    for (int p = piv.Offp[k]; p < piv.Offp[k + 1]; ++p)
                        pairs.emplace_back(lay.n_lu + p, xs(piv.Offi[p], c));
                    for (int s = lay.col_begin[k]; s < lay.diag_slot[k]; ++s)
                        pairs.emplace_back(s, xs(lay.slot_row[s], c));
                    lv.add(TK_MUL, xs(k, c), lay.diag_slot[k], pairs);
                }
            if (k2 - k1 == 1) continue;
            for (int k = k2 - 1; k >= k1; --k) {
                if (lay.diag_slot[k] + 1 == lay.col_begin[k + 1]) continue;
                for (int c = 0; c < rc; ++c) {
                    pairs.clear();
                    for (int s = lay.diag_slot[k] + 1; s < lay.col_begin[k + 1]; ++s)
                        pairs.emplace_back(s, xs(lay.slot_row[s], c));
                    lv.add(TK_SUB, xs(k, c), -1, pairs);
                }
            }
        }
    }

    // stable sort by level (longest dots first inside a level)
    int maxlvl = 0;
    for (const Task& t : prog.tasks) maxlvl = std::max(maxlvl, t.level);
    if (tail_t0 >= 0) maxlvl = std::max(maxlvl, prog.tail_level);
    prog.nlevels = maxlvl;
    std::stable_sort(prog.tasks.begin(), prog.tasks.end(), [](const Task& a, const Task& b) {
        if (a.level != b.level) return a.level < b.level;
        return a.plen > b.plen;
    });
    // levels are 1-based: level_ptr[l] .. level_ptr[l+1] are the tasks of level l+1
    prog.level_ptr.assign(maxlvl + 1, 0);
    for (const Task& t : prog.tasks) prog.level_ptr[t.level]++;
    for (int l = 0; l < maxlvl; ++l) prog.level_ptr[l + 1] += prog.level_ptr[l];
    if (split >= 100 && do_factor &&
        !(std::getenv("KLUJAX_CUDA_REBALANCE") && std::atoi(std::getenv("KLUJAX_CUDA_REBALANCE")) == 0))
        rebalance_levels(prog, lay.n_val, 128);
    if (compact && do_factor && (do_solve || do_tsolve)) compact_slots(lay, prog, compact);
    if (tail_t0 >= 0) {
        prog.tail_tab.assign(static_cast<size_t>(kTail) * kTail, 0xFFFF);
        for (int s = 0; s < lay.n_lu; ++s)
            if (in_tail(lay.slot_row[s]) && in_tail(lay.slot_col[s])) {
                const int ps = prog.remap.empty() ? s : prog.remap[s];
                prog.tail_tab[static_cast<size_t>(lay.slot_col[s] - tail_t0) * kTail + (lay.slot_row[s] - tail_t0)] =
                    static_cast<uint16_t>(ps);
            }
    }
}

bool pack_program(const SlotLayout& lay, const TaskProgram& prog, PackedProgram& out) {
    out = PackedProgram();
    constexpr int LW = 32, LINES = kChunkWords / LW;
    const int nx = (prog.mode == PM_FACTOR) ? 0 : lay.n * prog.rc;
    out.zero_slot = prog.n_val + nx;  // then ONE (= 1), then TRASH (write-only)
    out.n_slots = out.zero_slot + 3;
    if (out.n_slots > 16383) return false;
    const uint32_t ZERO = out.zero_slot, ONE = ZERO + 1, TRASH = ZERO + 2;
    auto opword = [](uint32_t a, uint32_t b) { return (a << 18) | (b << 4); };
    // bits 2-3: certificate class of a tracked entry (plan.hpp)
    auto hdword = [](uint32_t o, uint32_t d, int track, bool pivot) {
        return (o << 18) | (d << 4) | (track ? 1u : 0u) | (pivot ? 2u : 0u) | (track > 1 ? static_cast<uint32_t>(track - 1) << 2 : 0u);
    };
    const uint32_t idle_op = opword(ZERO, ZERO), idle_hd = hdword(TRASH, ONE, 0, false);
    // a batch = header line + 2*trips operand lines
    const int MAXTRIPS = (LINES - 1) / 2;
    int line = LINES;  // forces a first chunk
    size_t chunk_base = 0;
    auto open_chunk = [&]() {
        chunk_base = out.stream.size();
        out.stream.resize(chunk_base + kChunkWords, idle_op);
        out.chunk_ptr.push_back(static_cast<uint16_t>(out.nbatch));
        line = 0;
        out.nchunks++;
    };
    std::vector<long> f;
    std::vector<int> choice, grp;
    // measured on B200 (in-kernel cycle counters): a batch costs ~700 cycles before its first
    // operand line, ~75 per pair of operand lines, ~40 per reduction step
    long cost_fixed = 18, cost_trip = 2, cost_lg = 1;
    if (const char* e = std::getenv("KLUJAX_CUDA_PACK_COST")) std::sscanf(e, "%ld,%ld,%ld", &cost_fixed, &cost_trip, &cost_lg);
    for (int l = 0; l < prog.nlevels; ++l) {
        const int tb = prog.level_ptr[l], te = prog.level_ptr[l + 1];
        {
            grp.clear();
            for (int t = tb; t < te; ++t) grp.push_back(t);
            const int T = static_cast<int>(grp.size());
            if (!T) continue;
            // Tasks are sorted by dot length (longest first).  Dynamic program over that
            // order: a batch takes the next 32/g tasks and gives each g lanes; its cost is
            // a fixed part, one unit per pair of operand lines and one per reduction step.
            f.assign(T + 1, 0);
            choice.assign(T, 0);
            for (int i = T - 1; i >= 0; --i) {
                long best = -1;
                int best_lg = -1;
                for (int lg = 0; lg <= 5; ++lg) {
                    const int g = 1 << lg, cnt = std::min(LW >> lg, T - i);
                    const int run = (prog.tasks[grp[i]].plen + g - 1) / g;
                    const int trips = std::max(1, (run + 1) / 2);
                    if (trips > MAXTRIPS) continue;
                    const long c = cost_fixed + cost_trip * trips + cost_lg * lg + f[i + cnt];
                    if (best < 0 || c < best) {
                        best = c;
                        best_lg = lg;
                    }
                }
                if (best_lg < 0) return false;  // a dot product too long for one chunk
                f[i] = best;
                choice[i] = best_lg;
            }
            for (int i = 0; i < T;) {
                const int lg = choice[i], g = 1 << lg, cnt = std::min(LW >> lg, T - i);
                const int run = (prog.tasks[grp[i]].plen + g - 1) / g;
                const int trips = std::max(1, (run + 1) / 2);
                if (line + 1 + 2 * trips > LINES) open_chunk();
                uint32_t cw = static_cast<uint32_t>(trips) | (static_cast<uint32_t>(lg) << 8);
                const size_t base = chunk_base + static_cast<size_t>(line) * LW;
                for (int lane = 0; lane < LW; ++lane) out.stream[base + lane] = idle_hd;
                for (int k = 0; k < cnt; ++k) {
                    const Task& tk = prog.tasks[grp[i + k]];
                    const int lead = k * g;
                    const uint32_t div = tk.div >= 0 ? static_cast<uint32_t>(tk.div) : ONE;
                    if (tk.kind == TK_RECIP) cw |= kCwPivot;
                    if (tk.div >= 0) cw |= kCwMul;
                    if (tk.kind >= TK_MUL_TRACK) cw |= kCwTrack;
                    out.stream[base + lead] = hdword(tk.out, div, tk.kind >= TK_MUL_TRACK ? tk.kind - TK_MUL_TRACK + 1 : 0, tk.kind == TK_RECIP);
                    for (int q = 0; q < tk.plen; ++q) {
                        const int lane = lead + (q % g), row = q / g;
                        out.stream[base + static_cast<size_t>(LW) * (1 + row) + lane] =
                            opword(prog.pa[tk.pbeg + q], prog.pb[tk.pbeg + q]);
                    }
                }
                out.ctrl.push_back(cw);
                line += 1 + 2 * trips;
                out.nbatch++;
                out.lines_used += 1 + 2 * trips;
                i += cnt;
            }
        }
    }
    out.chunk_ptr.push_back(static_cast<uint16_t>(out.nbatch));
    return out.nbatch <= kMaxBatches && out.nchunks <= kMaxChunks;
}

}  // namespace kb

namespace kb {

// Packs a single-multiply-add program (build_program with split = 111) for the multi-warp
// shared-memory kernel (kernels_wide.cuh).  A level is cut into 128-lane steps; a step is, per
// lane, one header word (out<<18 | pivot-or-ONE<<4 | fresh<<3 | kind) and one operand word
// (a<<18 | b<<4; ZERO, ZERO for a lane without a pair), stored side by side.
// ctrl[step] = 1 | level-ends<<4; steps never straddle a chunk.
namespace {

// Lane order of the tasks of one level for the multi-warp kernel.  A 16-byte (8-byte)
// shared-memory access of a warp is served in groups of 8 (16) consecutive lanes; a group costs
// one wavefront per DISTINCT address that shares a 16-byte (8-byte) bank column with another
// one.  Tasks of a level are independent, so they are dealt greedily into lane groups such that
// inside a group the destinations, the first operands and the second operands each fall into
// different bank columns (equal addresses are a broadcast and cost nothing).
void order_for_banks(const TaskProgram& prog, std::vector<int>& order, int group) {
    const size_t n = order.size();
    if (n <= 1) return;
    auto slot_of = [&](int t, int cls) -> int {
        const Task& tk = prog.tasks[t];
        if (cls == 0) return tk.out;
        if (tk.plen == 0) return -1;
        return cls == 1 ? prog.pa[tk.pbeg] : prog.pb[tk.pbeg];
    };
    std::vector<char> used(n, 0);
    std::vector<int> result;
    result.reserve(n);
    std::vector<int> members;
    size_t first_free = 0;
    while (result.size() < n) {
        while (used[first_free]) ++first_free;
        members.clear();
        members.push_back(static_cast<int>(first_free));
        used[first_free] = 1;
        while (static_cast<int>(members.size()) < group && result.size() + members.size() < n) {
            int best = -1, best_cost = 1 << 30, seen = 0;
            for (size_t c = first_free + 1; c < n && best_cost > 0 && seen < 512; ++c) {  // bounded scan
                if (used[c]) continue;
                ++seen;
                int cost = 0;
                for (int cls = 0; cls < 3; ++cls) {
                    const int sc = slot_of(order[c], cls);
                    if (sc < 0) continue;
                    bool same = false, clash = false;
                    for (int m : members) {
                        const int sm = slot_of(order[m], cls);
                        if (sm < 0) continue;
                        if (sm == sc) same = true;
                        else if ((sm % group) == (sc % group)) clash = true;
                    }
                    if (!same && clash) ++cost;
                }
                if (cost < best_cost) {
                    best_cost = cost;
                    best = static_cast<int>(c);
                }
            }
            if (best < 0) break;
            used[best] = 1;
            members.push_back(best);
        }
        for (int m : members) result.push_back(order[m]);
    }
    order.swap(result);
}

}  // namespace

bool pack_program_wide(const SlotLayout& lay, const TaskProgram& prog, PackedProgram& out, int elem_bytes) {
    out = PackedProgram();
    constexpr int LANES = 128;
    const bool bank_order = !(std::getenv("KLUJAX_CUDA_BANK_ORDER") && std::atoi(std::getenv("KLUJAX_CUDA_BANK_ORDER")) == 0);
    const int nx = (prog.mode == PM_FACTOR) ? 0 : lay.n * prog.rc;
    out.zero_slot = prog.n_val + nx;  // then ONE (= 1), then TRASH
    out.n_slots = out.zero_slot + 3;
    if (out.n_slots > 16383) return false;
    const uint32_t ZERO = out.zero_slot, ONE = out.zero_slot + 1;
    const uint32_t idle_op = (ZERO << 18) | (ZERO << 4);  // an operand line without a pair: 0 * 0
    size_t chunk_base = 0;
    int used = kChunkWords;  // forces a first chunk
    auto open_chunk = [&]() {
        chunk_base = out.stream.size();
        out.stream.resize(chunk_base + kChunkWords, 0u);
        out.chunk_ptr.push_back(static_cast<uint16_t>(out.ctrl.size()));
        used = 0;
        out.nchunks++;
    };
    std::vector<int> order;
    bool tail_pending = prog.tail_t0 >= 0;
    for (int l = 0; l < prog.nlevels; ++l) {
        const int tb = prog.level_ptr[l], te = prog.level_ptr[l + 1];
        if (tail_pending && l + 1 >= prog.tail_level && te == tb && l + 1 == prog.nlevels) {
            // nothing left after the register stage: an idle step carries its flag
            if (used > 0) open_chunk();
            uint32_t* base = out.stream.data() + chunk_base + used;
            for (int k = 0; k < LANES; ++k) base[2 * k] = 0u, base[2 * k + 1] = idle_op;
            out.ctrl.push_back(1u | 16u | 32u);
            used += LANES * 2;
            tail_pending = false;
            continue;
        }
        order.resize(te - tb);
        for (int t = tb; t < te; ++t) order[t - tb] = t;  // already sorted by length, longest first
        if (bank_order) order_for_banks(prog, order, 128 / elem_bytes);
        for (size_t i = 0; i < order.size(); i += LANES) {
            const size_t cnt = std::min<size_t>(LANES, order.size() - i);
            // the kernel's step format is fixed: header word + ONE operand word per lane (split 111)
            for (size_t k = 0; k < cnt; ++k)
                if (prog.tasks[order[i + k]].plen > 1) return false;
            const int P = 1;
            const int words = LANES * (1 + P);
            // the register stage is entered at a chunk boundary (kernels_wide.cuh)
            const bool tail_here = tail_pending && l + 1 >= prog.tail_level;
            if (used + words > kChunkWords || (tail_here && used > 0)) open_chunk();
            uint32_t* base = out.stream.data() + chunk_base + used;
            for (size_t k = 0; k < cnt; ++k) {
                const Task& tk = prog.tasks[order[i + k]];
                const uint32_t d = tk.div >= 0 ? static_cast<uint32_t>(tk.div) : ONE;
                const uint32_t kind = tk.kind == TK_RECIP ? 4u : tk.kind == TK_MUL_TRACK ? 3u : tk.kind == TK_MUL_TRACK_OFF ? 5u : tk.kind == TK_MUL_TRACK_DIAG ? 6u : tk.kind == TK_MUL ? 2u : 1u;
                // lane k: (header word, operand word) side by side
                base[2 * k] = (static_cast<uint32_t>(tk.out) << 18) | (d << 4) | (tk.fresh ? 8u : 0u) | kind;
                base[2 * k + 1] = tk.plen ? (static_cast<uint32_t>(prog.pa[tk.pbeg]) << 18) |
                                                (static_cast<uint32_t>(prog.pb[tk.pbeg]) << 4)
                                          : idle_op;
            }
            const bool level_end = (i + LANES >= order.size());
            uint32_t cwv = static_cast<uint32_t>(P) | (level_end ? 16u : 0u);
            if (tail_pending && l + 1 >= prog.tail_level) {  // the register stage runs before this step
                cwv |= 32u;
                tail_pending = false;
            }
            out.ctrl.push_back(cwv);
            used += words;
            out.lines_used += 1 + P;
        }
    }
    if (tail_pending) {  // no step at or after the stage's level: an idle step carries the flag
        if (used > 0) open_chunk();
        uint32_t* base = out.stream.data() + chunk_base + used;
        for (int k = 0; k < LANES; ++k) base[2 * k] = 0u, base[2 * k + 1] = idle_op;
        out.ctrl.push_back(1u | 16u | 32u);
        used += LANES * 2;
    }
    out.chunk_ptr.push_back(static_cast<uint16_t>(out.ctrl.size()));
    out.nbatch = static_cast<int>(out.ctrl.size());
    return out.nbatch > 0 && out.nbatch <= kMaxBatches && out.nchunks <= kMaxChunks;
}

}  // namespace kb

// ---------------------------------------------------------------------------------------------
// Dependency-driven regime (plan.hpp: ColProgram, RowProgram)
// ---------------------------------------------------------------------------------------------
namespace kb {

bool build_col_program(const PivotPlan& piv, const SlotLayout& lay, ColProgram& out) {
    const int n = lay.n;
    out = ColProgram();
    out.n = n;
    out.colinfo.resize(4 * static_cast<size_t>(n));
    out.ccand.assign(n, -2);
    size_t nu = 0;
    for (int k = 0; k < n; ++k) nu += static_cast<size_t>(lay.diag_slot[k] - lay.col_begin[k]);
    out.urec.resize(4 * nu);
    std::vector<int> pos(n, -1), depth(n, 0);
    size_t u = 0;
    for (int k = 0; k < n; ++k) {
        const int cb = lay.col_begin[k], dg = lay.diag_slot[k], ce = lay.col_begin[k + 1];
        const int nc = ce - cb;
        out.maxcol = std::max(out.maxcol, nc);
        out.colinfo[4 * k + 0] = cb;
        out.colinfo[4 * k + 1] = dg - cb;
        out.colinfo[4 * k + 2] = nc;
        out.colinfo[4 * k + 3] = static_cast<int>(u);
        for (int s = cb; s < ce; ++s) pos[lay.slot_row[s]] = s - cb;
        int d = 0;
        for (int su = cb; su < dg; ++su, ++u) {  // ascending pivot rows: a topological order
            const int j = lay.slot_row[su];
            const int lbeg = lay.diag_slot[j] + 1, lcnt = lay.col_begin[j + 1] - lbeg;
            if (out.dest.size() + static_cast<size_t>(lcnt) >= (1ull << 31)) return false;
            out.urec[4 * u + 0] = j;
            out.urec[4 * u + 1] = lbeg;
            out.urec[4 * u + 2] = lcnt;
            out.urec[4 * u + 3] = static_cast<int>(out.dest.size());
            for (int sl = lbeg; sl < lbeg + lcnt; ++sl) {
                const int p = pos[lay.slot_row[sl]];
                if (p < 0) return false;  // the pattern of column k always contains L(:,j) below row j
                out.dest.push_back(p);
            }
            out.n_mac += lcnt;
            d = std::max(d, depth[j]);
        }
        depth[k] = d + 1;
        out.depth = std::max(out.depth, depth[k]);
        if (!piv.diag_cand.empty() && piv.diag_cand[k] != -2) {
            const int c = piv.diag_cand[k];
            out.ccand[k] = (c >= 0 && pos[c] > dg - cb) ? pos[c] : -1;
        }
        for (int s = cb; s < ce; ++s) pos[lay.slot_row[s]] = -1;
    }
    if (out.dest.empty()) out.dest.push_back(0);
    if (out.urec.empty()) out.urec.assign(4, 0);
    return true;
}

bool build_panel_program(const Ordering& ord, const PivotPlan& piv, const SlotLayout& lay, bool transpose,
                         PanelProgram& out) {
    const int n = lay.n;
    out = PanelProgram();
    out.n = n;
    struct Opnd {
        int slot, row;  // slot -2: the constant -1
        int wseq;       // sequence number of the task that produced the operand (-1: the loaded b)
    };
    struct PTask {
        int row, div;
        bool mul, noinit;
        std::vector<Opnd> ops;
    };
    std::vector<PTask> seq;
    std::vector<int> last_writer(n, -1);  // emission-order index of the latest task that wrote row k
    // Temporaries (partial sums of split dot products) come from a small ring: a partial is consumed
    // by the final task of its row, which follows its partials immediately, and panels run one after
    // the other - a slot is dead once the panel of that final task is done.  Two rows' worth of
    // partials plus two panels of slack never collide.
    long long maxops = 0;
    if (!transpose) {
        std::vector<long long> cnt(n, 0);
        for (int k = 0; k < n; ++k) {
            for (int s = lay.col_begin[k]; s < lay.diag_slot[k]; ++s) cnt[lay.slot_row[s]]++;   // U row (second pass)
            for (int p = piv.Offp[k]; p < piv.Offp[k + 1]; ++p) cnt[piv.Offi[p]]++;
        }
        std::vector<long long> cl(n, 0);
        for (int k = 0; k < n; ++k)
            for (int s = lay.diag_slot[k] + 1; s < lay.col_begin[k + 1]; ++s) cl[lay.slot_row[s]]++;
        for (int k = 0; k < n; ++k) maxops = std::max(maxops, cnt[k] + cl[k]);
    } else {
        for (int k = 0; k < n; ++k)
            maxops = std::max<long long>(maxops, (lay.col_begin[k + 1] - lay.col_begin[k]) + (piv.Offp[k + 1] - piv.Offp[k]));
    }
    const int ring = static_cast<int>(((2 * (maxops / kPanelChunk + 1) + 64 + 31) / 32) * 32);
    int ntemps = 0, temp_next = 0;
    auto emit = [&](int row, int div, bool mul, std::vector<Opnd>& ops) {
        for (Opnd& o : ops) o.wseq = last_writer[o.row];
        // earliest-produced operands first: they are the ones that can be summed ahead of time
        std::stable_sort(ops.begin(), ops.end(), [](const Opnd& a, const Opnd& b) { return a.wseq < b.wseq; });
        size_t done = 0;
        std::vector<int> temps;
        while (ops.size() - done > static_cast<size_t>(kPanelChunk)) {
            PTask pt;
            pt.row = n + temp_next;
            pt.div = -1;
            pt.mul = false;
            pt.noinit = true;
            pt.ops.reserve(kPanelChunk);
            for (int q = 0; q < kPanelChunk; ++q) pt.ops.push_back(ops[done + q]);
            done += kPanelChunk;
            temps.push_back(n + temp_next);
            temp_next = (temp_next + 1) % ring;
            ntemps = std::min(ring, ntemps + 1);
            seq.push_back(std::move(pt));
        }
        PTask t;
        t.row = row;
        t.div = div;
        t.mul = mul;
        t.noinit = false;
        for (size_t q = done; q < ops.size(); ++q) t.ops.push_back(ops[q]);
        for (int tr : temps) t.ops.push_back({-2, tr, 0});  // acc -= (-1) * partial
        out.n_mac += static_cast<int64_t>(ops.size());
        last_writer[row] = static_cast<int>(seq.size());
        seq.push_back(std::move(t));
    };
    std::vector<Opnd> ops;
    if (!transpose) {
        std::vector<std::vector<std::pair<int, int>>> lrow(n), urow(n), offrow(n);
        for (int k = 0; k < n; ++k) {
            for (int s = lay.col_begin[k]; s < lay.diag_slot[k]; ++s) urow[lay.slot_row[s]].emplace_back(s, k);
            for (int s = lay.diag_slot[k] + 1; s < lay.col_begin[k + 1]; ++s) lrow[lay.slot_row[s]].emplace_back(s, k);
            for (int p = piv.Offp[k]; p < piv.Offp[k + 1]; ++p) offrow[piv.Offi[p]].emplace_back(lay.n_lu + p, k);
        }
        // klu_solve: blocks last -> first; L then U inside a block
        for (int b = ord.nblocks - 1; b >= 0; --b) {
            const int k1 = ord.R[b], k2 = ord.R[b + 1];
            if (k2 - k1 == 1) {
                ops.clear();
                for (auto& e : offrow[k1]) ops.push_back({e.first, e.second, 0});
                emit(k1, lay.diag_slot[k1], true, ops);
                continue;
            }
            for (int i = k1; i < k2; ++i) {
                if (offrow[i].empty() && lrow[i].empty()) continue;
                ops.clear();
                for (auto& e : offrow[i]) ops.push_back({e.first, e.second, 0});
                for (auto& e : lrow[i]) ops.push_back({e.first, e.second, 0});
                emit(i, -1, false, ops);
            }
            for (int i = k2 - 1; i >= k1; --i) {
                ops.clear();
                for (auto& e : urow[i]) ops.push_back({e.first, e.second, 0});
                emit(i, lay.diag_slot[i], true, ops);
            }
        }
    } else {
        // klu_tsolve: blocks first -> last; U' then L' inside a block
        for (int b = 0; b < ord.nblocks; ++b) {
            const int k1 = ord.R[b], k2 = ord.R[b + 1];
            for (int k = k1; k < k2; ++k) {
                ops.clear();
                for (int p = piv.Offp[k]; p < piv.Offp[k + 1]; ++p) ops.push_back({lay.n_lu + p, piv.Offi[p], 0});
                for (int s = lay.col_begin[k]; s < lay.diag_slot[k]; ++s) ops.push_back({s, lay.slot_row[s], 0});
                emit(k, lay.diag_slot[k], true, ops);
            }
            if (k2 - k1 == 1) continue;
            for (int k = k2 - 1; k >= k1; --k) {
                if (lay.diag_slot[k] + 1 == lay.col_begin[k + 1]) continue;
                ops.clear();
                for (int s = lay.diag_slot[k] + 1; s < lay.col_begin[k + 1]; ++s) ops.push_back({s, lay.slot_row[s], 0});
                emit(k, -1, false, ops);
            }
        }
    }
    out.ntemps = ntemps;  // slots of the ring in use
    const int ntask = static_cast<int>(seq.size());
    out.npanels = (ntask + 31) / 32;
    if (out.npanels == 0) return false;
    out.hdr.assign(4 * static_cast<size_t>(out.npanels), 0);
    out.lane.assign(4 * 32 * static_cast<size_t>(out.npanels), 0);
    // classification pass: who wrote each operand last, in final numbering
    std::vector<int> writer(n + ntemps, -1);
    struct LaneOps {
        std::vector<std::pair<int, int>> ext;  // (slot, row)
        std::vector<std::pair<int, int>> in;   // (producing lane, slot)
    };
    std::vector<LaneOps> lo(32);
    for (int p = 0; p < out.npanels; ++p) {
        const int g0 = p * 32, g1 = std::min(ntask, g0 + 32);
        int E = 0, I = 0;
        for (int g = g0; g < g1; ++g) {
            const PTask& t = seq[g];
            LaneOps& l = lo[g - g0];
            l.ext.clear();
            l.in.clear();
            bool noinit = t.noinit;
            unsigned mask = 0;
            auto classify = [&](int slot, int row) {
                const int w = writer[row];
                if (w >= g0) {
                    l.in.emplace_back(w - g0, slot);
                    mask |= 1u << (w - g0);
                } else {
                    l.ext.emplace_back(slot, row);
                }
            };
            if (!noinit && writer[t.row] >= g0) {  // the row's own previous value was produced in this panel
                classify(-2, t.row);
                noinit = true;
            }
            for (const Opnd& o : t.ops) classify(o.slot, o.row);
            // one operand per producing lane: two entries of a row never share a column
            std::sort(l.in.begin(), l.in.end());
            for (size_t q = 1; q < l.in.size(); ++q)
                if (l.in[q].first == l.in[q - 1].first) return false;
            E = std::max<int>(E, static_cast<int>(l.ext.size()));
            I = std::max<int>(I, static_cast<int>(l.in.size()));
            int* li = &out.lane[4 * static_cast<size_t>(g)];
            li[0] = t.row;
            li[1] = t.div;
            li[2] = static_cast<int>(mask);
            li[3] = (noinit ? 1 : 0) | (t.mul ? 4 : 0);
            if (writer[t.row] >= g0) out.lane[4 * static_cast<size_t>(writer[t.row]) + 3] |= 2;  // superseded inside the panel
            writer[t.row] = g;
        }
        for (int g = g1; g < g0 + 32; ++g) {  // idle lanes of the last panel
            int* li = &out.lane[4 * static_cast<size_t>(g)];
            li[0] = 0;
            li[1] = -1;
            li[2] = 0;
            li[3] = 1 | 2;
            lo[g - g0].ext.clear();
            lo[g - g0].in.clear();
        }
        out.hdr[4 * p + 0] = E;
        out.hdr[4 * p + 1] = I;
        out.hdr[4 * p + 2] = static_cast<int>(out.ext_slot.size() / 32);
        out.hdr[4 * p + 3] = static_cast<int>(out.in_slot.size() / 32);
        out.maxE = std::max(out.maxE, E);
        out.maxI = std::max(out.maxI, I);
        for (int e = 0; e < E; ++e)
            for (int ln = 0; ln < 32; ++ln) {
                const bool has = e < static_cast<int>(lo[ln].ext.size());
                out.ext_slot.push_back(has ? lo[ln].ext[e].first : -1);
                out.ext_idx.push_back(has ? lo[ln].ext[e].second : -1);
            }
        for (int t = 0; t < I; ++t)
            for (int ln = 0; ln < 32; ++ln)
                out.in_slot.push_back(t < static_cast<int>(lo[ln].in.size()) ? lo[ln].in[t].second : -1);
    }
    if (out.ext_slot.empty()) {
        out.ext_slot.assign(32, -1);
        out.ext_idx.assign(32, -1);
    }
    if (out.in_slot.empty()) out.in_slot.assign(32, -1);
    return true;
}

}  // namespace kb

namespace kb {

bool build_group_program(const Ordering& ord, const PivotPlan& piv, const SlotLayout& lay, bool transpose,
                         const int* xrow, GroupProgram& out) {
    const int n = lay.n;
    out = GroupProgram();
    out.n = n;
    auto name = [&](int k) { return xrow ? xrow[k] : k; };
    struct RTask {
        int row, div;
        bool mul;
        std::vector<std::pair<int, int>> ops;  // (value slot, operand row), pivot-order rows
    };
    std::vector<RTask> seq;
    auto emit = [&](int row, int div, bool mul, std::vector<std::pair<int, int>>& ops) {
        RTask t;
        t.row = row;
        t.div = div;
        t.mul = mul;
        t.ops = ops;
        out.n_mac += static_cast<int64_t>(ops.size());
        seq.push_back(std::move(t));
    };
    std::vector<std::pair<int, int>> ops;
    if (!transpose) {
        std::vector<std::vector<std::pair<int, int>>> lrow(n), urow(n), offrow(n);
        for (int k = 0; k < n; ++k) {
            for (int s = lay.col_begin[k]; s < lay.diag_slot[k]; ++s) urow[lay.slot_row[s]].emplace_back(s, k);
            for (int s = lay.diag_slot[k] + 1; s < lay.col_begin[k + 1]; ++s) lrow[lay.slot_row[s]].emplace_back(s, k);
            for (int p = piv.Offp[k]; p < piv.Offp[k + 1]; ++p) offrow[piv.Offi[p]].emplace_back(lay.n_lu + p, k);
        }
        for (int b = ord.nblocks - 1; b >= 0; --b) {  // klu_solve: blocks last -> first; L then U inside a block
            const int k1 = ord.R[b], k2 = ord.R[b + 1];
            if (k2 - k1 == 1) {
                ops = offrow[k1];
                emit(k1, lay.diag_slot[k1], true, ops);
                continue;
            }
            for (int i = k1; i < k2; ++i) {
                if (offrow[i].empty() && lrow[i].empty()) continue;
                ops = offrow[i];
                ops.insert(ops.end(), lrow[i].begin(), lrow[i].end());
                emit(i, -1, false, ops);
            }
            for (int i = k2 - 1; i >= k1; --i) {
                ops = urow[i];
                emit(i, lay.diag_slot[i], true, ops);
            }
        }
    } else {
        for (int b = 0; b < ord.nblocks; ++b) {  // klu_tsolve: blocks first -> last; U' then L' inside a block
            const int k1 = ord.R[b], k2 = ord.R[b + 1];
            for (int k = k1; k < k2; ++k) {
                ops.clear();
                for (int p = piv.Offp[k]; p < piv.Offp[k + 1]; ++p) ops.emplace_back(lay.n_lu + p, piv.Offi[p]);
                for (int s = lay.col_begin[k]; s < lay.diag_slot[k]; ++s) ops.emplace_back(s, lay.slot_row[s]);
                emit(k, lay.diag_slot[k], true, ops);
            }
            if (k2 - k1 == 1) continue;
            for (int k = k2 - 1; k >= k1; --k) {
                if (lay.diag_slot[k] + 1 == lay.col_begin[k + 1]) continue;
                ops.clear();
                for (int s = lay.diag_slot[k] + 1; s < lay.col_begin[k + 1]; ++s) ops.emplace_back(s, lay.slot_row[s]);
                emit(k, -1, false, ops);
            }
        }
    }
    // ---- grouping: consecutive tasks that depend on a member of the group or share operands with it
    struct Grp {
        std::vector<int> tasks;
        int level = 0;
        std::vector<int> aux, krec;
        int nk = 0;
    };
    std::vector<Grp> groups;
    {
        std::vector<int> stamp(n, -1);  // operand rows of the current group (stamp = group id)
        std::vector<int> member(n, -1);  // rows written by the current group
        int gid = -1;
        for (int t = 0; t < static_cast<int>(seq.size()); ++t) {
            const RTask& rt = seq[t];
            bool join = false;
            // a group neither writes a row twice nor writes a row one of its members reads from outside
            // (the partial tasks update the group's rows in place while its k-steps are still being read)
            if (gid >= 0 && static_cast<int>(groups[gid].tasks.size()) < kGroupRows && member[rt.row] != gid && stamp[rt.row] != gid) {
                int shared = 0, dep = 0;
                for (const auto& o : rt.ops) {
                    if (stamp[o.second] == gid) ++shared;
                    if (member[o.second] == gid) ++dep;
                }
                join = dep > 0 || (!rt.ops.empty() && 2 * shared >= static_cast<int>(rt.ops.size()));
            }
            if (!join) {
                groups.emplace_back();
                gid = static_cast<int>(groups.size()) - 1;
            }
            groups[gid].tasks.push_back(t);
            member[rt.row] = gid;
            for (const auto& o : rt.ops) stamp[o.second] = gid;
        }
    }
    // ---- records + levels (dependencies and anti-dependencies on the rows of X, which is updated in place).
    // The k-steps of a group are sorted by the level at which their operand row becomes final and cut
    // into chunks of kGroupChunk: every chunk but the last is a PARTIAL task (X[rows] -= sum, in place)
    // that runs as early as its operands allow, on whatever warp is free - without this the thousands
    // of k-steps of a top-separator group would all sit on the critical path of one warp.
    constexpr int kGroupChunk = 64;
    struct Entry {
        int level, nk, koff, z, aux;
    };
    std::vector<Entry> entries;
    std::vector<int> lw(n, 0), lr(n, 0), pos_in_group(n, -1), kidx(n, -1);
    std::vector<int> gkrec, order_k;
    for (size_t g = 0; g < groups.size(); ++g) {
        Grp& G = groups[g];
        const int nt = static_cast<int>(G.tasks.size());
        G.aux.assign(kGroupAux, -1);
        for (int i = 0; i < nt; ++i) pos_in_group[seq[G.tasks[i]].row] = i;
        std::vector<int> krows;
        gkrec.clear();
        int base = 0, mulbits = 0;
        for (int i = 0; i < nt; ++i) {
            const RTask& rt = seq[G.tasks[i]];
            G.aux[i] = name(rt.row);
            G.aux[8 + i] = rt.mul ? rt.div : -1;
            if (rt.mul) mulbits |= 1 << i;
            base = std::max(base, std::max(lw[rt.row], lr[rt.row]));
            for (const auto& o : rt.ops) {
                const int pj = pos_in_group[o.second];
                if (pj >= 0 && pj < i) {  // produced by an earlier row of this group
                    G.aux[16 + i * (i - 1) / 2 + pj] = o.first;
                } else {
                    if (pj >= 0) return false;  // would read a row the group is updating in place (excluded by the grouping)
                    if (kidx[o.second] < 0) {
                        kidx[o.second] = static_cast<int>(krows.size());
                        krows.push_back(o.second);
                        gkrec.resize(gkrec.size() + 12, -1);
                        gkrec[12 * kidx[o.second]] = name(o.second);
                    }
                    gkrec[12 * kidx[o.second] + 1 + i] = o.first;
                }
            }
        }
        const int nk = static_cast<int>(krows.size());
        order_k.resize(nk);
        for (int q = 0; q < nk; ++q) order_k[q] = q;
        std::stable_sort(order_k.begin(), order_k.end(), [&](int x, int y) { return lw[krows[x]] < lw[krows[y]]; });
        const int aux_index = static_cast<int>(out.gaux.size() / kGroupAux);
        out.gaux.insert(out.gaux.end(), G.aux.begin(), G.aux.end());
        int prev = base, q0 = 0;
        const int nchunks = std::max(1, (nk + kGroupChunk - 1) / kGroupChunk);
        for (int c = 0; c < nchunks; ++c) {
            const int q1 = std::min(nk, q0 + kGroupChunk);
            const bool last = c == nchunks - 1;
            int lvl = prev;
            for (int q = q0; q < q1; ++q) lvl = std::max(lvl, lw[krows[order_k[q]]]);
            lvl += 1;
            Entry e;
            e.level = lvl;
            e.nk = q1 - q0;
            e.koff = static_cast<int>(out.krec.size() / 12);
            e.z = nt | (last ? (mulbits << 8) : (1 << 30));
            e.aux = aux_index;
            for (int q = q0; q < q1; ++q) {
                int* src = &gkrec[12 * order_k[q]];
                int mask = 0;
                for (int i = 0; i < kGroupRows; ++i)
                    if (src[1 + i] >= 0) mask |= 1 << i;
                src[9] = mask;  // rows of the group that have an entry at this k-step
                out.krec.insert(out.krec.end(), src, src + 12);
                lr[krows[order_k[q]]] = std::max(lr[krows[order_k[q]]], lvl);
            }
            entries.push_back(e);
            out.n_ksteps += e.nk;
            prev = lvl;
            q0 = q1;
        }
        for (int r2 : krows) kidx[r2] = -1;
        for (int i = 0; i < nt; ++i) {
            const int row = seq[G.tasks[i]].row;
            lw[row] = prev;
            lr[row] = 0;
            pos_in_group[row] = -1;
        }
        out.nlevels = std::max(out.nlevels, prev);
    }
    // sort by level, emit
    std::vector<int> order(entries.size());
    for (size_t g = 0; g < entries.size(); ++g) order[g] = static_cast<int>(g);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return entries[a].level < entries[b].level; });
    out.ngroups = static_cast<int>(entries.size());
    out.level_ptr.assign(out.nlevels + 1, 0);
    for (const Entry& e : entries) out.level_ptr[e.level]++;
    for (int l = 0; l < out.nlevels; ++l) out.level_ptr[l + 1] += out.level_ptr[l];
    out.ghdr.reserve(12 * entries.size());
    for (int g : order) {
        const Entry& e = entries[g];
        out.ghdr.push_back(e.nk);
        out.ghdr.push_back(e.koff);
        out.ghdr.push_back(e.z);
        out.ghdr.push_back(e.aux);
        for (int i = 0; i < kGroupRows; ++i) out.ghdr.push_back(out.gaux[static_cast<size_t>(kGroupAux) * e.aux + i]);  // row names
    }
    if (out.krec.empty()) out.krec.assign(12, -1);
    return out.ngroups > 0;
}

}  // namespace kb
