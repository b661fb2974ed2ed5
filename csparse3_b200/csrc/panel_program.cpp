// panel_program.cpp -- host compiler of the "panel" refactor program executed by lu_panel.cu.
//
// No reference counterpart.  The arithmetic the program encodes is the frozen-pattern
// left-looking cs_lu column update of oracle/csp3_oracle.c (orc_csc_lu_refactor): for every entry of every column
// the same operations in the same order, so the factors are bit-identical in exact mode.
//
// Formulation.  Columns are eliminated in PANELS of one or two adjacent columns (power-flow Jacobians carry the
// theta / V unknowns of a bus as adjacent columns with nearly the same pattern) whose accumulators acc0 / acc1 live
// in shared memory, and the source columns of a panel are applied in TASKS of one or two adjacent source columns
// with nested patterns (a fundamental supernode pair).  A task keeps its multipliers U(j,k) in registers and
// walks the rows of L(:,j): a 2 x 2 task costs 2 loads of L, 2 accumulator loads and 2 stores for 4 multiply-
// subtracts where the scalar formulation (wide_program.cpp) needs 4 shared-memory accesses per operation.
//
// Order.  Bit-exactness needs, per accumulator entry, the update sequence of cs_lu (sources in the stored
// topological order of U(:,k)), and final multipliers.  Per column the tasks follow the stored order of U(:,k);
// two columns share a task only where both orders agree, and two sources are fused only when they are adjacent in
// that order, so every entry sees exactly the sequence of the oracle.
//
// Program = sequence of STEPS of `groups` 8-byte records, one record per lane group (row group); lanes of a group
// are the systems of the bundle.  Encoding: see panel_program.hpp.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "panel_program.hpp"
#include "symbolic.hpp"

namespace csp3 {

namespace {

struct Task {
    i32 j = 0;          // first source column
    int ws = 1;         // source columns (1 or 2)
    int mask = 0;       // bit 0: acc0 (first column of the panel), bit 1: acc1
};

struct Emitter {
    std::vector<uint64_t> words;
    int G;
    i64 steps = 0;
    i64 count[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    explicit Emitter(int g) : G(g) {}
    // one step: records.size() <= G, missing ones become invalid records of the same opcode
    void step(int op, const std::vector<uint64_t> &recs)
    {
        for (int g = 0; g < G; ++g) words.push_back(g < (int)recs.size() ? recs[(size_t)g] : panel_word(op, 0, 0, 0, 0));
        ++steps;
        ++count[op];
    }
    void uniform(int op, uint64_t rec)
    {
        for (int g = 0; g < G; ++g) words.push_back(rec);
        ++steps;
        ++count[op];
    }
    int last_op() const { return words.empty() ? -1 : panel_op(words.back()); }
};

}  // namespace

bool compile_panel_refactor(i64 n64, const i32 *Ap, const i32 *Ai, const std::vector<i32> &q, const Factor &F, i32 width,
                            i32 groups, PanelProgram &P, const char **why)
{
    P = PanelProgram();
    const i32 n = (i32)n64;
    const std::vector<i32> &Lp = F.Lp, &Li = F.Li, &Up = F.Up, &Ui = F.Ui, &pinv = F.pinv;
    if (n == 0) { *why = "empty matrix"; return false; }
    if ((i64)Li.size() >= (1ll << 20) || (i64)Ui.size() >= (1ll << 32) || (i64)n >= (1ll << 31)) {
        *why = "panel program: factor too large for the 20-bit L entry index";
        return false;
    }
    const bool no_pair = getenv("CSP3_PANEL_NOPAIR") != nullptr;      // ws = 1 only (debug)
    const bool no_panel = getenv("CSP3_PANEL_NOPANEL") != nullptr;    // wt = 1 only (debug)

    // position of row i inside L(:,j): sorted (row, pos) per column for set comparisons and lookups
    std::vector<std::vector<std::pair<i32, i32>>> lrow((size_t)n);
    for (i32 j = 0; j < n; ++j) {
        auto &v = lrow[(size_t)j];
        for (i32 p = Lp[j] + 1; p < Lp[j + 1]; ++p) v.push_back({Li[p], p});
        std::sort(v.begin(), v.end());
        for (size_t t = 1; t < v.size(); ++t)
            if (v[t].first == v[t - 1].first) { *why = "panel program: duplicate row in a column of L"; return false; }
    }
    auto lpos = [&](i32 j, i32 row) -> i32 {
        const auto &v = lrow[(size_t)j];
        auto it = std::lower_bound(v.begin(), v.end(), std::make_pair(row, (i32)-1));
        return (it != v.end() && it->first == row) ? it->second : -1;
    };
    // fundamental supernode pair: rows(L(:,j)) \ {j+1} == rows(L(:,j+1)) and j+1 in rows(L(:,j))
    std::vector<char> sn_pair((size_t)n, 0);
    for (i32 j = 0; j + 1 < n && !no_pair; ++j) {
        const auto &a = lrow[(size_t)j], &b = lrow[(size_t)j + 1];
        if (a.size() != b.size() + 1 || a.empty() || a[0].first != j + 1) continue;
        bool same = true;
        for (size_t t = 0; t < b.size() && same; ++t) same = a[t + 1].first == b[t].first;
        sn_pair[(size_t)j] = same;
    }

    // ---- pass 1: panels and their task lists ------------------------------------------------------------
    struct Panel {
        i32 k = 0;
        int wt = 1;
        bool internal = false;             // U(k, k+1) is in the pattern: column k updates column k+1 while it is finalised
        std::vector<Task> tasks, post;     // post: sources of column k+1 that follow k in its stored order
        std::vector<i32> rows;             // sorted union of the rows of the panel's columns (slot = rank)
    };
    std::vector<Panel> panels;
    std::vector<i32> mark((size_t)n, -1), posA((size_t)n, -1);
    auto pair_up = [&](const std::vector<std::pair<i32, int>> &seq, std::vector<Task> &out) {
        for (size_t t = 0; t < seq.size(); ++t) {
            Task T; T.j = seq[t].first; T.mask = seq[t].second; T.ws = 1;
            if (t + 1 < seq.size() && seq[t + 1].first == T.j + 1 && seq[t + 1].second == T.mask && sn_pair[(size_t)T.j]) {
                T.ws = 2; ++t;
            }
            out.push_back(T);
        }
    };
    for (i32 k = 0; k < n;) {
        Panel pn; pn.k = k;
        const i32 a0 = Up[k], a1 = Up[k + 1] - 1;                      // sources of column k: Ui[a0, a1)
        bool two = false;
        std::vector<std::pair<i32, int>> seq;
        if (k + 1 < n && !no_panel) {
            const i32 b0 = Up[k + 1], b1 = Up[k + 2] - 1;
            i32 bk = -1;                                               // position of k in U(:,k+1)
            for (i32 p = b0; p < b1; ++p) if (Ui[p] == k) bk = p;
            const i32 pre1 = bk >= 0 ? bk : b1;
            // common sources: in U(:,k) and before k in U(:,k+1); their relative order must agree
            for (i32 p = a0; p < a1; ++p) { mark[(size_t)Ui[p]] = k; posA[(size_t)Ui[p]] = p; }
            bool consistent = true;
            i32 last = -1, ncommon = 0;
            for (i32 p = b0; p < pre1; ++p) {
                const i32 j = Ui[p];
                if (mark[(size_t)j] == k) { if (posA[(size_t)j] < last) consistent = false; last = posA[(size_t)j]; ++ncommon; }
            }
            if (consistent && (ncommon > 0 || bk >= 0)) {
                two = true;
                // merge the two orders: runs of one-sided sources are kept intact (they may hold supernode pairs)
                std::vector<char> common((size_t)(a1 - a0), 0);
                for (i32 p = b0; p < pre1; ++p) if (mark[(size_t)Ui[p]] == k) common[(size_t)(posA[(size_t)Ui[p]] - a0)] = 1;
                i32 ia = a0, ib = b0;
                while (ia < a1 || ib < pre1) {
                    if (ia < a1 && !common[(size_t)(ia - a0)]) { seq.push_back({Ui[ia], 1}); ++ia; continue; }
                    if (ib < pre1 && mark[(size_t)Ui[ib]] != k) { seq.push_back({Ui[ib], 2}); ++ib; continue; }
                    // both at a common source (the same one, by consistency)
                    seq.push_back({Ui[ia], 3}); ++ia; ++ib;
                }
                pair_up(seq, pn.tasks);
                if (bk >= 0) {
                    pn.internal = true;
                    std::vector<std::pair<i32, int>> ps;
                    for (i32 p = bk + 1; p < b1; ++p) ps.push_back({Ui[p], 2});
                    pair_up(ps, pn.post);
                }
            }
        }
        if (!two) {
            for (i32 p = a0; p < a1; ++p) seq.push_back({Ui[p], 1});
            pair_up(seq, pn.tasks);
        }
        pn.wt = two ? 2 : 1;
        // rows of the panel: U rows + diagonal + L rows of each column
        for (int c = 0; c < pn.wt; ++c) {
            const i32 kc = k + c;
            for (i32 p = Up[kc]; p < Up[kc + 1]; ++p) pn.rows.push_back(Ui[p]);
            for (i32 p = Lp[kc] + 1; p < Lp[kc + 1]; ++p) pn.rows.push_back(Li[p]);
        }
        std::sort(pn.rows.begin(), pn.rows.end());
        pn.rows.erase(std::unique(pn.rows.begin(), pn.rows.end()), pn.rows.end());
        P.nslots = std::max<i32>(P.nslots, (i32)pn.rows.size());
        k += pn.wt;
        panels.push_back(std::move(pn));
    }
    if (P.nslots >= 4096) { *why = "panel program: column too long for the 12-bit slot index"; return false; }
    P.nslots = (P.nslots + 1) & ~1;
    const i32 NS = P.nslots;

    // ---- pass 2: emit ---------------------------------------------------------------------------------------
    Emitter E(groups);
    std::vector<i32> slot_of((size_t)n, -1);
    std::vector<i32> qinv_dummy;
    i64 upd_rows = 0, upd_slots = 0, ops = 0;
    i64 tasks_by[3][4] = {{0}};
    auto emit_tasks = [&](const std::vector<Task> &tasks) -> bool {
        for (const Task &T : tasks) {
            const i32 j = T.j;
            const int m0 = T.mask & 1, m1 = (T.mask >> 1) & 1;
            ++tasks_by[T.ws][T.mask];
            if (slot_of[(size_t)j] < 0) { *why = "panel program: source row without a slot"; return false; }
            int fl = (T.ws == 2 ? kPanelWS2 : 0) | (m0 ? kPanelM0 : 0) | (m1 ? kPanelM1 : 0) | kPanelValid;
            if (T.ws == 2) {
                const i32 tri = lpos(j, j + 1);
                if (tri < 0 || slot_of[(size_t)j + 1] < 0) { *why = "panel program: broken supernode pair"; return false; }
                E.uniform(kPanelLoadU, panel_word(kPanelLoadU, fl, (uint32_t)tri, (uint32_t)slot_of[(size_t)j + 1], (uint32_t)slot_of[(size_t)j]));
                ops += m0 + m1;
            } else {
                E.uniform(kPanelLoadU, panel_word(kPanelLoadU, fl, 0, 0, (uint32_t)slot_of[(size_t)j]));
            }
            // rows: storage order of the last source column
            const i32 jl = j + T.ws - 1;
            std::vector<uint64_t> recs;
            for (i32 p = Lp[jl] + 1; p < Lp[jl + 1]; ++p) {
                const i32 row = Li[p];
                const i32 sl = slot_of[(size_t)row];
                if (sl < 0) { *why = "panel program: target row without a slot"; return false; }
                uint32_t l0 = (uint32_t)p, l1 = 0;
                if (T.ws == 2) { l0 = (uint32_t)lpos(j, row); l1 = (uint32_t)p; }
                recs.push_back(panel_word(kPanelUpd, fl, l0, l1, (uint32_t)sl));
                ops += (m0 + m1) * T.ws;
                ++upd_rows;
                if ((int)recs.size() == groups) { E.step(kPanelUpd, recs); recs.clear(); upd_slots += groups; }
            }
            if (!recs.empty()) { E.step(kPanelUpd, recs); upd_slots += groups; }
        }
        return true;
    };
    auto emit_final = [&](i32 kc, int acc, bool fused) {
        // pivot first (the diagonal is still in the accumulator), then U entries, then L entries
        const i32 base = acc * NS;
        E.uniform(kPanelPiv, panel_word_wide(kPanelPiv, kPanelValid | (fused ? kPanelFused : 0), (uint64_t)(kc + 1), (uint32_t)(base + slot_of[(size_t)kc])));
        std::vector<uint64_t> recs;
        for (i32 p = Up[kc]; p < Up[kc + 1]; ++p) {
            recs.push_back(panel_word_wide(kPanelFinU, kPanelValid, (uint64_t)p, (uint32_t)(base + slot_of[(size_t)Ui[p]])));
            if ((int)recs.size() == groups) { E.step(kPanelFinU, recs); recs.clear(); }
        }
        if (!recs.empty()) { E.step(kPanelFinU, recs); recs.clear(); }
        for (i32 p = Lp[kc] + 1; p < Lp[kc + 1]; ++p) {
            recs.push_back(panel_word_wide(kPanelFinL, kPanelValid | (fused ? kPanelFused : 0), (uint64_t)p, (uint32_t)(base + slot_of[(size_t)Li[p]])));
            if ((int)recs.size() == groups) { E.step(kPanelFinL, recs); recs.clear(); }
            if (fused) ++ops;
        }
        if (!recs.empty()) E.step(kPanelFinL, recs);
    };
    for (const Panel &pn : panels) {
        for (size_t t = 0; t < pn.rows.size(); ++t) slot_of[(size_t)pn.rows[t]] = (i32)t;
        // (an UPD step never directly follows a FINL step -- tasks start with LOADU -- so the L operands it loads one
        // step ahead are never the ones a FINL step is writing)
        // scatter A(:, q[k + c]) into acc_c; an entry written twice in a column (duplicates: last wins) goes to a later step
        std::vector<uint64_t> recs;
        std::vector<i32> used;
        for (int c = 0; c < pn.wt; ++c) {
            const i32 col = q.empty() ? pn.k + c : q[(size_t)(pn.k + c)];
            for (i32 p = Ap[col]; p < Ap[col + 1]; ++p) {
                const i32 row = pinv[(size_t)Ai[p]];
                const i32 sl = slot_of[(size_t)row];
                if (sl < 0) { *why = "panel program: entry of A outside the factor pattern"; return false; }
                const i32 dst = c * NS + sl;
                if (std::find(used.begin(), used.end(), dst) != used.end()) { E.step(kPanelScatter, recs); recs.clear(); used.clear(); }
                recs.push_back(panel_word_wide(kPanelScatter, kPanelValid, (uint64_t)p, (uint32_t)dst));
                used.push_back(dst);
                if ((int)recs.size() == groups) { E.step(kPanelScatter, recs); recs.clear(); used.clear(); }
            }
        }
        if (!recs.empty()) E.step(kPanelScatter, recs);
        if (!emit_tasks(pn.tasks)) return false;
        emit_final(pn.k, 0, pn.internal);
        if (pn.wt == 2) {
            if (!emit_tasks(pn.post)) return false;
            emit_final(pn.k + 1, 1, false);
        }
        for (i32 r : pn.rows) slot_of[(size_t)r] = -1;
    }
    for (int t = 0; t < 3; ++t) E.uniform(kPanelEnd, panel_word(kPanelEnd, 0, 0, 0, 0));

    P.ok = true;
    P.width = width; P.groups = groups;
    P.steps = (i32)E.steps;
    P.ops = ops;
    P.smem_bytes = (size_t)2 * NS * width * 8;
    P.prog.bytes.resize(E.words.size() * 8);
    std::memcpy(P.prog.bytes.data(), E.words.data(), P.prog.bytes.size());
    P.prog.stage = groups * 8;
    for (int i = 0; i < 8; ++i) P.step_count[i] = E.count[i];
    P.upd_rows = upd_rows; P.upd_row_slots = upd_slots; P.npanels = (i32)panels.size();
    if (getenv("CSP3_DEBUG")) {
        i64 two = 0;
        for (const Panel &pn : panels) two += pn.wt == 2;
        fprintf(stderr,
                "csp3: panel program: %d panels (%lld of 2), %d slots, %lld steps (scatter %lld loadu %lld upd %lld piv %lld finu %lld finl %lld nop %lld), "
                "upd rows %lld in %lld row slots, ops %lld, tasks ws1 [m1 %lld m2 %lld m3 %lld] ws2 [m1 %lld m2 %lld m3 %lld], %zu bytes\n",
                P.npanels, (long long)two, NS, (long long)E.steps, (long long)E.count[kPanelScatter], (long long)E.count[kPanelLoadU],
                (long long)E.count[kPanelUpd], (long long)E.count[kPanelPiv], (long long)E.count[kPanelFinU], (long long)E.count[kPanelFinL],
                (long long)E.count[kPanelNop], (long long)upd_rows, (long long)upd_slots, (long long)ops,
                (long long)tasks_by[1][1], (long long)tasks_by[1][2], (long long)tasks_by[1][3],
                (long long)tasks_by[2][1], (long long)tasks_by[2][2], (long long)tasks_by[2][3], P.prog.bytes.size());
    }
    return true;
}

}  // namespace csp3
