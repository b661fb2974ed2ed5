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

// a word before encoding
struct Word {
    int op = kPanelNone, flags = 0;
    uint32_t a = 0, b = 0, c = 0;
    uint64_t ab = 0;
    bool wide = false;
    uint64_t encode() const { return wide ? panel_word_wide(op, flags, ab, c) : panel_word(op, flags, a, b, c); }
};
// an operand that has to be fetched from global memory into a landing entry before step `consumer`
struct Req {
    i64 consumer;       // logical step
    int word;           // word of that step
    int field;          // 0: a, 1: b, 2: ab
    int kindA;          // 1: A value (index into Ax), 0: L value (entry of the bundle's L array)
    i64 src;
    i64 ready;          // first logical step at which the data may be read from global memory
};
struct Step {
    std::vector<Word> w;        // <= kPanelStepWords
};

}  // namespace

bool compile_panel_refactor(i64 n64, const i32 *Ap, const i32 *Ai, const std::vector<i32> &q, const Factor &F, i32 width,
                            size_t smem_budget, PanelProgram &P, const char **why)
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

    // ---- pass 2: logical steps -----------------------------------------------------------------------------
    constexpr int SW = kPanelStepWords, D = kPanelLookahead;
    const size_t EB = (size_t)width * 8;
    const size_t prog_ring = (size_t)kPanelProgStages * kPanelStageSteps * SW * 8;
    auto envi = [](const char *k, int dflt) { const char *v = getenv(k); return v ? atoi(v) : dflt; };
    i32 LN = envi("CSP3_PANEL_LANDING", 64);                       // landing entries
    if ((size_t)(2 * NS + LN + 32) * EB + prog_ring > smem_budget) { *why = "panel program: shared-memory budget too small"; return false; }
    i32 R = (i32)((smem_budget - prog_ring) / EB) - 2 * NS - LN;   // ring entries: what is left
    if (getenv("CSP3_PANEL_RING")) R = std::min(R, std::max(32, envi("CSP3_PANEL_RING", R)));
    i32 maxl = 1;
    for (i32 j = 0; j < n; ++j) maxl = std::max(maxl, Lp[j + 1] - Lp[j] - 1);
    if (R < maxl) { *why = "panel program: L column longer than the ring"; return false; }
    R = std::min<i32>(R, 4096);
    if (2 * NS >= 8192 || R + LN >= 8192) { *why = "panel program: index field overflow"; return false; }

    std::vector<Step> steps;
    std::vector<Req> reqs;
    std::vector<i32> slot_of((size_t)n, -1);
    std::vector<i32> ring_pos((size_t)n, -1);        // ring entry of L(:,j)'s first off-diagonal value
    std::vector<i64> ring_dead((size_t)n, -1);       // logical step of the FINL word that overwrites (part of) it
    std::vector<i64> fin_done((size_t)n, -1);        // first logical step at which L(:,j) may be read from global memory
    std::vector<i32> ring_owner((size_t)R, -1);
    i32 ring_ptr = 0;
    i64 upd_rows = 0, ops = 0, ring_rows = 0;
    i64 tasks_by[3][4] = {{0}};
    auto new_step = [&]() -> Step & { steps.emplace_back(); return steps.back(); };
    auto cur_step = [&]() -> i64 { return (i64)steps.size() - 1; };

    // operand L(p) of column j for a word of the current step: ring entry when the column is still there, else a
    // landing entry (request recorded, index patched in pass 3)
    auto l_operand = [&](i32 j, i32 p, int word, int field, bool &from_ring) -> uint32_t {
        const i64 here = cur_step();
        if (ring_pos[(size_t)j] >= 0 && (ring_dead[(size_t)j] < 0 || ring_dead[(size_t)j] > here)) {
            from_ring = true;
            return (uint32_t)(ring_pos[(size_t)j] + (p - Lp[j] - 1));
        }
        from_ring = false;
        reqs.push_back({here, word, field, 0, p, fin_done[(size_t)j]});
        return 0;
    };
    auto emit_tasks = [&](const std::vector<Task> &tasks) -> bool {
        for (const Task &T : tasks) {
            const i32 j = T.j;
            const int m0 = T.mask & 1, m1 = (T.mask >> 1) & 1;
            ++tasks_by[T.ws][T.mask];
            if (slot_of[(size_t)j] < 0) { *why = "panel program: source row without a slot"; return false; }
            const int fl = (T.ws == 2 ? kPanelWS2 : 0) | (m0 ? kPanelM0 : 0) | (m1 ? kPanelM1 : 0);
            Step *st = &new_step();
            Word hd; hd.op = kPanelHdrU; hd.flags = fl; hd.c = (uint32_t)slot_of[(size_t)j];
            st->w.push_back(hd);
            if (T.ws == 2) {
                const i32 tri = lpos(j, j + 1);
                if (tri < 0 || slot_of[(size_t)j + 1] < 0) { *why = "panel program: broken supernode pair"; return false; }
                bool fr;
                st->w[0].a = l_operand(j, tri, 0, 0, fr);
                st->w[0].b = (uint32_t)slot_of[(size_t)j + 1];
                ops += m0 + m1;
            }
            const i32 jl = j + T.ws - 1;            // rows: storage order of the last source column
            for (i32 p = Lp[jl] + 1; p < Lp[jl + 1]; ++p) {
                const i32 row = Li[p];
                const i32 sl = slot_of[(size_t)row];
                if (sl < 0) { *why = "panel program: target row without a slot"; return false; }
                if ((int)st->w.size() == SW) st = &new_step();
                Word u; u.op = kPanelUpd; u.flags = fl; u.c = (uint32_t)sl;
                const int wi = (int)st->w.size();
                st->w.push_back(u);
                bool fr0 = false, fr1 = false;
                if (T.ws == 2) { st->w[(size_t)wi].a = l_operand(j, lpos(j, row), wi, 0, fr0); st->w[(size_t)wi].b = l_operand(j + 1, p, wi, 1, fr1); }
                else st->w[(size_t)wi].a = l_operand(j, p, wi, 0, fr0);
                ops += (m0 + m1) * T.ws;
                ++upd_rows;
                ring_rows += fr0 ? 1 : 0;
            }
        }
        return true;
    };
    auto emit_final = [&](i32 kc, int acc, bool fused) {
        // header: pivot (the diagonal is still in the accumulator); U entries; L entries (also kept in the ring)
        const i32 base = acc * NS;
        Step *st = &new_step();
        Word hd; hd.op = kPanelHdrP; hd.flags = fused ? kPanelFused : 0; hd.wide = true; hd.ab = (uint64_t)(kc + 1);
        hd.c = (uint32_t)(base + slot_of[(size_t)kc]);
        st->w.push_back(hd);
        for (i32 p = Up[kc]; p < Up[kc + 1]; ++p) {
            if ((int)st->w.size() == SW) st = &new_step();
            Word u; u.op = kPanelFinU; u.wide = true; u.ab = (uint64_t)p; u.c = (uint32_t)(base + slot_of[(size_t)Ui[p]]);
            st->w.push_back(u);
        }
        const i32 len = Lp[kc + 1] - Lp[kc] - 1;
        i32 rpos = -1;
        if (len > 0) {
            if (ring_ptr + len > R) ring_ptr = 0;
            rpos = ring_ptr;
            ring_ptr += len;
            // L entries never share a step with U entries of the same column: the division needs the header's pivot,
            // which is fine, but keeping them apart keeps the kernel's word loop uniform
            st = &new_step();
        }
        for (i32 p = Lp[kc] + 1; p < Lp[kc + 1]; ++p) {
            if ((int)st->w.size() == SW) st = &new_step();
            const i32 re = rpos + (p - Lp[kc] - 1);
            const i32 old = ring_owner[(size_t)re];
            if (old >= 0 && old != kc && (ring_dead[(size_t)old] < 0)) ring_dead[(size_t)old] = cur_step();
            ring_owner[(size_t)re] = kc;
            Word u; u.op = kPanelFinL; u.wide = true; u.flags = (fused ? kPanelFused : 0) | kPanelX;
            u.ab = (uint64_t)p | ((uint64_t)re << 24); u.c = (uint32_t)(base + slot_of[(size_t)Li[p]]);
            st->w.push_back(u);
            if (fused) ++ops;
        }
        ring_pos[(size_t)kc] = rpos;
        ring_dead[(size_t)kc] = -1;
        fin_done[(size_t)kc] = cur_step() + 1;
    };
    for (const Panel &pn : panels) {
        for (size_t t = 0; t < pn.rows.size(); ++t) slot_of[(size_t)pn.rows[t]] = (i32)t;
        // scatter A(:, q[k + c]) into acc_c; an entry written twice in a column (duplicates: last wins) goes to a later step
        Step *st = &new_step();
        std::vector<i32> used;
        for (int c = 0; c < pn.wt; ++c) {
            const i32 col = q.empty() ? pn.k + c : q[(size_t)(pn.k + c)];
            for (i32 p = Ap[col]; p < Ap[col + 1]; ++p) {
                const i32 row = pinv[(size_t)Ai[p]];
                const i32 sl = slot_of[(size_t)row];
                if (sl < 0) { *why = "panel program: entry of A outside the factor pattern"; return false; }
                const i32 dst = c * NS + sl;
                if ((int)st->w.size() == SW || std::find(used.begin(), used.end(), dst) != used.end()) { st = &new_step(); used.clear(); }
                Word u; u.op = kPanelScatter; u.wide = true; u.c = (uint32_t)dst;
                reqs.push_back({cur_step(), (int)st->w.size(), 2, 1, p, -1});
                st->w.push_back(u);
                used.push_back(dst);
            }
        }
        if (st->w.empty()) steps.pop_back();
        if (!emit_tasks(pn.tasks)) return false;
        emit_final(pn.k, 0, pn.internal);
        if (pn.wt == 2) {
            if (!emit_tasks(pn.post)) return false;
            emit_final(pn.k + 1, 1, false);
        }
        for (i32 r : pn.rows) slot_of[(size_t)r] = -1;
    }

    // ---- pass 3: schedule the fetches ---------------------------------------------------------------------------
    // Requests are in consumer order.  A FETCH word is placed into a free word of a logical step at least D steps before
    // the consumer (pure fetch steps are inserted when the deadline comes without a free word); landing entries are
    // handed out round-robin and an entry is reused only after its consumer has executed.  A request whose data becomes
    // final less than D steps before its consumer is served synchronously by the consumer (flag X / Y).
    const i64 NL = (i64)steps.size();
    const int EARLY = envi("CSP3_PANEL_EARLY", 6);
    std::vector<std::vector<Word>> extra((size_t)NL + 1);          // pure fetch steps inserted BEFORE logical step i
    i64 issued = 0, sync_loads = 0, fetch_words = 0;
    std::vector<char> is_sync(reqs.size(), 0);
    std::vector<i32> landing_of(reqs.size(), -1);
    for (size_t r = 0; r < reqs.size(); ++r)                        // requests that can never be fetched in time
        if (reqs[r].consumer - std::max<i64>(reqs[r].ready, 0) < D) { is_sync[r] = 1; ++sync_loads; }
    std::vector<i64> issued_consumer;                               // consumer step of every issued fetch, in issue order
    size_t cptr = 0;                                                // fetches whose consumer step is over
    size_t rq = 0;
    auto next_async = [&]() { while (rq < reqs.size() && is_sync[rq]) ++rq; };
    next_async();
    for (i64 i = 0; i < NL && rq < reqs.size(); ++i) {
        while (cptr < issued_consumer.size() && issued_consumer[cptr] < i) ++cptr;
        auto in_use = [&]() { return (i64)(issued_consumer.size() - cptr); };
        auto place = [&](std::vector<Word> &dst) {
            const Req &r = reqs[rq];
            Word f; f.op = r.kindA ? kPanelFetchA : kPanelFetchL; f.wide = true; f.ab = (uint64_t)r.src;
            landing_of[rq] = (i32)(issued % LN);
            f.c = (uint32_t)(R + landing_of[rq]);
            dst.push_back(f);
            issued_consumer.push_back(r.consumer);
            ++issued; ++fetch_words; ++rq;
            next_async();
        };
        auto can_issue = [&]() {
            if (rq >= reqs.size()) return false;
            const Req &r = reqs[rq];
            if (r.ready > i) return false;                                         // data not final yet
            if (r.consumer - i > D + EARLY) return false;                          // too early
            return in_use() < LN;                                                  // the landing entry it would get is free
        };
        while (can_issue() && (int)steps[(size_t)i].w.size() < SW) place(steps[(size_t)i].w);
        // deadline: everything whose consumer is step i + D must be issued before step i
        while (rq < reqs.size() && reqs[rq].consumer - i <= D) {
            if (reqs[rq].consumer - i < D || reqs[rq].ready > i || in_use() >= LN) {
                is_sync[rq] = 1; ++sync_loads;                                     // cannot be fetched in time after all
                ++rq; next_async();
                continue;
            }
            place(extra[(size_t)i]);
        }
    }
    // patch the consumers
    for (size_t r = 0; r < reqs.size(); ++r) {
        Word &w = steps[(size_t)reqs[r].consumer].w[(size_t)reqs[r].word];
        if (is_sync[r]) {
            if (reqs[r].field == 0) { w.a = (uint32_t)reqs[r].src; w.flags |= kPanelX; }
            else if (reqs[r].field == 1) { w.b = (uint32_t)reqs[r].src; w.flags |= kPanelY; }
            else { w.ab = (uint64_t)reqs[r].src; w.flags |= kPanelX; }
        } else {
            const uint32_t e = (uint32_t)(R + landing_of[r]);
            if (reqs[r].field == 0) w.a = e;
            else if (reqs[r].field == 1) w.b = e;
            else w.ab = e;
        }
    }
    // ---- encode -----------------------------------------------------------------------------------------------------
    std::vector<uint64_t> words;
    i64 real_steps = 0, fetch_steps = 0;
    auto out_step = [&](const Word *w, size_t cnt) {
        for (size_t t = 0; t < (size_t)SW; ++t) words.push_back(t < cnt ? w[t].encode() : panel_word(kPanelNone, 0, 0, 0, 0));
        ++real_steps;
    };
    for (i64 i = 0; i < NL; ++i) {
        const auto &ex = extra[(size_t)i];
        for (size_t o = 0; o < ex.size(); o += SW) { out_step(ex.data() + o, std::min<size_t>(SW, ex.size() - o)); ++fetch_steps; }
        // a header must stay word 0: fetch words were appended behind the step's own words
        out_step(steps[(size_t)i].w.data(), steps[(size_t)i].w.size());
    }
    Word endw; endw.op = kPanelEnd;
    for (int t = 0; t < 2; ++t) out_step(&endw, 1);
    while (real_steps % kPanelStageSteps) out_step(&endw, 1);

    P.ok = true;
    P.width = width; P.groups = kPanelGroups;
    P.nslots = NS; P.ring = R; P.landing = LN;
    P.steps = (i32)real_steps;
    P.ops = ops;
    P.smem_bytes = (size_t)(2 * NS + R + LN) * EB + prog_ring;
    P.prog.bytes.resize(words.size() * 8);
    std::memcpy(P.prog.bytes.data(), words.data(), P.prog.bytes.size());
    P.prog.stage = kPanelStageSteps * SW * 8;
    P.upd_rows = upd_rows; P.npanels = (i32)panels.size(); P.sync_loads = sync_loads; P.fetch_steps = fetch_steps; P.ring_rows = ring_rows;
    if (getenv("CSP3_DEBUG")) {
        i64 two = 0;
        for (const Panel &pn : panels) two += pn.wt == 2;
        fprintf(stderr,
                "csp3: panel program: %d panels (%lld of 2), %d slots, ring %d, landing %d, %lld steps (%lld pure fetch steps), upd rows %lld "
                "(first operand from the ring: %lld), fetch words %lld, synchronous loads %lld, ops %lld, tasks ws1 [m1 %lld m2 %lld m3 %lld] "
                "ws2 [m1 %lld m2 %lld m3 %lld], %zu bytes, smem %zu\n",
                P.npanels, (long long)two, NS, R, LN, (long long)real_steps, (long long)fetch_steps, (long long)upd_rows, (long long)ring_rows,
                (long long)fetch_words, (long long)sync_loads, (long long)ops,
                (long long)tasks_by[1][1], (long long)tasks_by[1][2], (long long)tasks_by[1][3],
                (long long)tasks_by[2][1], (long long)tasks_by[2][2], (long long)tasks_by[2][3], P.prog.bytes.size(), P.smem_bytes);
    }
    return true;
}

}  // namespace csp3
