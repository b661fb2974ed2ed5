// wide_program.cpp -- host compiler of the "wide" (lane = system) refactor program executed by lu_wide.cu.
// Record formats: program.hpp.  No reference counterpart; the arithmetic the program
// encodes is the frozen-pattern left-looking cs_lu column update of oracle/csp3_oracle.c, operation for
// operation and, for every entry of every column, in the same order (which is all that bit-exactness needs: the
// list scheduler below reorders operations on different entries freely).
// Knobs (environment, read once): CSP3_WIDE_SCHED=0 in-order chunk packer, CSP3_WIDE_PAIRS source columns a column
// offers to the scheduler at a time (6), CSP3_WIDE_RUN longest landing run in entries, CSP3_WIDE_GA A entries per
// group (<= kWideGroupA), CSP3_WIDE_ACC accumulator slots.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "program.hpp"
#include "symbolic.hpp"

namespace csp3 {

namespace {

struct Bytes {
    std::vector<uint8_t> &out;
    explicit Bytes(std::vector<uint8_t> &o) : out(o) {}
    void i32v(i32 v) { const uint8_t *b = (const uint8_t *)&v; out.insert(out.end(), b, b + 4); }
    void u16v(i64 v) { const uint16_t w = (uint16_t)v; const uint8_t *b = (const uint8_t *)&w; out.insert(out.end(), b, b + 2); }
    void pad(size_t a) { while (out.size() % a) out.push_back(0); }
};

struct Fetch { i32 len = 0, dst = 0, src = 0; };

struct Op { i32 pair, t, base; };   // update t of pair `pair` (index into Schedule::pairs); slot base of its column

struct FinEnt { i32 gout, slot, cache; };       // cache: lsrc entry or -1

struct Rec {
    int kind = 0;               // 0 group header, 1 chunk, 2 finalisation
    i32 group = -1;             // group index (-1: preamble)
    std::vector<Op> ops;        // chunk records
    std::vector<FinEnt> fins;   // finalisation records
    i32 new_far = -1;           // pair whose landing run is first used by this chunk (-1: none)
    bool immediate = false;
    Fetch fetch;
};

struct Group {
    std::vector<i32> cols, base;
    i32 nslots = 0, a_cnt = 0;
    i32 first_rec = 0, chunk_cnt = 0, fin_cnt = 0;
};

inline size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

bool compile_wide_refactor(const Schedule &S, const Factor &F, i32 width, i32 groups, size_t smem_budget,
                           i32 ring_override, i32 stage_override, WideProgram &W, const char **why)
{
    W = WideProgram();
    const i32 n = (i32)S.cols.size();
    const std::vector<i32> &Lp = F.Lp, &Up = F.Up, &Ui = F.Ui;
    if (width != 4 && width != 8 && width != 16 && width != 32) { *why = "wide bundle width must be 4, 8, 16 or 32"; return false; }
    if (n == 0) { *why = "empty matrix"; return false; }
    if (F.Li.size() >= (1u << 28) || F.Ui.size() >= (1u << 28)) { *why = "wide refactor: factor too large"; return false; }
    const i32 cap = 2 * groups;                                  // operations per chunk / finalisation record
    const i32 cover = kWideARegs * groups;                       // A values the kernel keeps in registers
    i32 max_llen = 1, max_acnt = 0;
    for (const PairDesc &pd : S.pairs) max_llen = std::max(max_llen, pd.llen);
    for (const ColDesc &cd : S.cols) max_acnt = std::max(max_acnt, cd.a_cnt);
    static const int ga_env = getenv("CSP3_WIDE_GA") ? atoi(getenv("CSP3_WIDE_GA")) : 0;
    const i32 group_a_limit = ga_env > 0 ? std::min<i32>(ga_env, kWideGroupA) : kWideGroupA;
    const i32 group_a = std::max<i32>(kWideGroupA, max_acnt);
    // ---- record sizes -> program stage size ----------------------------------------------------------------
    const size_t chunk_rec = (size_t)kWideChunkHeader + 8 * (size_t)cap;
    const size_t max_grp_rec = round_up(round_up((size_t)kWideColHeader + 16 * (size_t)kWideGroupCols + 2 * (size_t)group_a, 4) +
                                        4 * (size_t)group_a, 16);
    const size_t max_rec = std::max(chunk_rec, max_grp_rec);
    size_t stage = 512;
    while (stage < max_rec) stage *= 2;
    // ---- shared-memory geometry ---------------------------------------------------------------------------
    // value area: acc_slots accumulator slots | 2 * kWideGroupCols pivot / reciprocal entries | L cache | landing area
    const size_t entry = (size_t)width * 8;
    const i32 table = 2 * kWideGroupCols;
    i32 stage_entries = stage_override > 0 ? stage_override : (i32)round_up((size_t)(kWideLookahead + 1) * max_llen, 8);
    stage_entries = std::max(stage_entries, (i32)round_up((size_t)max_llen, 8));
    // Accumulator: the longest column must fit; up to 45 % more slots (larger groups: fewer group prologues, better
    // filled chunks) are taken from the L cache as long as the cache keeps max(32, longest L column) entries --
    // measured on config 3: 86 -> 124 slots, 828 -> 702 groups, -1.3 % time.
    static const int acc_env = getenv("CSP3_WIDE_ACC") ? atoi(getenv("CSP3_WIDE_ACC")) : 0;
    const i32 acc_need = (S.max_col_len + 1 + 1) & ~1;
    i32 acc_slots = acc_need;
    if (acc_env > 0) acc_slots = std::max(acc_need, acc_env & ~1);
    else if (smem_budget > (size_t)kWideProgStages * stage) {
        const i64 avail = (i64)((smem_budget - (size_t)kWideProgStages * stage) / entry) - table;
        const i64 room = avail - stage_entries - std::max<i64>(32, max_llen);
        acc_slots = (i32)std::max<i64>(acc_need, std::min<i64>((i64)acc_need * 29 / 20, room) & ~(i64)1);
    }

    // ---- phase 0: list schedule of the columns into groups -------------------------------------------------------
    // sources of column k: the off-diagonal rows of U(:,k)
    std::vector<char> done((size_t)n, 0), in_group((size_t)n, 0);
    std::vector<Group> G;
    {
        i32 lo = 0;                                              // oldest unfinished column
        const i32 window = 96;
        while (lo < n) {
            Group g;
            i32 a_total = 0;
            for (i32 k = lo; k < std::min(n, lo + window) && (i32)g.cols.size() < kWideGroupCols; ++k) {
                if (done[k]) continue;
                const ColDesc &cd = S.cols[k];
                const i32 len = cd.ucnt + cd.lcnt - 1;
                bool ready = true;
                for (i32 p = Up[k]; p < Up[k + 1] - 1 && ready; ++p) ready = done[Ui[p]] && !in_group[Ui[p]];
                if (!ready) continue;
                if (!g.cols.empty() && (g.nslots + len > acc_slots || a_total + cd.a_cnt > group_a_limit)) continue;
                g.cols.push_back(k); g.base.push_back(g.nslots);
                g.nslots += len; a_total += cd.a_cnt;
                in_group[k] = 1;
                if (len > acc_slots / 2) break;                  // a long column runs alone
            }
            if (g.cols.empty()) { *why = "wide refactor: internal error (no ready column)"; return false; }
            for (i32 k : g.cols) { done[k] = 1; in_group[k] = 0; }
            g.a_cnt = a_total;
            G.push_back(std::move(g));
            while (lo < n && done[lo]) ++lo;
        }
    }
    const i32 ngroups = (i32)G.size();

    auto geometry = [&](size_t stg, i32 &ring_entries) -> bool {
        const size_t prog_ring = (size_t)kWideProgStages * stg;
        const size_t fixed = ((size_t)acc_slots + table) * entry + prog_ring + (size_t)stage_entries * entry;
        if (fixed + 8 * entry > smem_budget) return false;
        ring_entries = (i32)((smem_budget - fixed) / entry);
        if (ring_override > 0) ring_entries = std::min(ring_entries, ring_override);
        if (((size_t)acc_slots + table + ring_entries + stage_entries) * entry > 0xfff0) {
            const size_t room = (0xfff0 / entry);
            if (room <= (size_t)acc_slots + table + (size_t)stage_entries + 8) return false;
            ring_entries = (i32)(room - (size_t)acc_slots - table - (size_t)stage_entries);
        }
        return true;
    };

    // Two chunk packers.  The list scheduler (mode 1) also places the landing runs itself; when it cannot make
    // progress (landing area full of runs whose pairs wait for each other) the program is compiled again with the
    // in-order packer (mode 0), whose runs are consumed one after the other.
    static const int sched_env = getenv("CSP3_WIDE_SCHED") ? atoi(getenv("CSP3_WIDE_SCHED")) : 1;
    static const int pair_window = getenv("CSP3_WIDE_PAIRS") ? std::max(1, atoi(getenv("CSP3_WIDE_PAIRS"))) : 6;
    static const int run_cap_env = getenv("CSP3_WIDE_RUN") ? atoi(getenv("CSP3_WIDE_RUN")) : 0;
    const size_t stage0 = stage;
    for (int mode = sched_env ? 1 : 0; mode >= 0; --mode) {
    const bool listsched = mode == 1;
    const i32 run_cap = run_cap_env > 0 ? run_cap_env : std::max(max_llen, stage_entries / 5);
    bool sched_failed = false;
    stage = stage0;
    for (int attempt = 0; attempt < 4 && !sched_failed; ++attempt, stage *= 2) {
        W = WideProgram();
        i32 ring_entries = 0;
        if (!geometry(stage, ring_entries)) { *why = "wide refactor working set exceeds the shared-memory budget"; return false; }
        const size_t prog_ring = (size_t)kWideProgStages * stage;
        const size_t lsrc0 = (size_t)acc_slots + table;          // first lsrc entry, in entries of the value area
        W.width = width; W.groups = groups; W.acc_slots = acc_slots; W.ring_entries = ring_entries; W.stage_entries = stage_entries;
        W.prog.stage = (i32)stage;
        W.smem_bytes = (lsrc0 + (size_t)ring_entries + stage_entries) * entry + prog_ring;

        // ---- phase 1: L cache (ring) simulation group by group; chunk packing; finalisation records -----------------
        std::vector<i32> ring_pos((size_t)n, -1), owner((size_t)std::max(ring_entries, 1), -1);
        std::vector<i32> pair_src((size_t)S.pairs.size(), -1), pair_col((size_t)S.pairs.size(), -1);
        std::vector<char> pair_far((size_t)S.pairs.size(), 0);
        i32 cur = 0;
        auto ring_valid = [&](i32 j) {
            const i32 pos = ring_pos[j], len = Lp[j + 1] - Lp[j] - 1;
            if (pos < 0) return false;
            for (i32 t = 0; t < len; ++t) if (owner[pos + t] != j) return false;
            return true;
        };
        std::vector<Rec> recs;
        recs.reserve(S.pairs.size() + 2 * (size_t)n + 1);
        { Rec r; r.kind = 0; r.group = -1; recs.push_back(r); }
        std::vector<i32> stamp_t((size_t)acc_slots + 1, -1);     // accumulator slot -> chunk id that writes it
        i32 chunk_id = 0;
        // record after which column j's L entries are in global memory: the first record AFTER its group
        std::vector<i32> final_rec((size_t)n, 0);
        // list scheduler: landing entry -> first record that may overwrite it (INT32_MAX while a run is in use)
        std::vector<i32> free_at((size_t)stage_entries, 0);
        i32 land_cur = 0;
        auto land_alloc = [&](i32 len, i32 x) -> i32 {           // run of `len` entries writable from record x on
            if (len > stage_entries) return -1;
            for (int pass = 0; pass < 2; ++pass) {
                const i32 lo = pass == 0 ? land_cur : 0, hi = pass == 0 ? stage_entries : std::min(stage_entries, land_cur + len - 1);
                i32 run = 0;
                for (i32 e = lo; e < hi; ++e) {
                    run = free_at[(size_t)e] <= x ? run + 1 : 0;
                    if (run >= len) { land_cur = e + 1; return e + 1 - len; }
                }
            }
            return -1;
        };
        for (i32 gi = 0; gi < ngroups; ++gi) {
            Group &g = G[gi];
            g.first_rec = (i32)recs.size();
            g.chunk_cnt = g.fin_cnt = 0;
            { Rec r; r.kind = 0; r.group = gi; recs.push_back(r); }
            // per column: cursor over its pairs / entries
            struct Cur { i32 pi, pend, t; bool first_of_pair; };
            std::vector<Cur> cs;
            for (size_t c = 0; c < g.cols.size(); ++c) {
                const ColDesc &cd = S.cols[g.cols[c]];
                cs.push_back({cd.pair_ptr, cd.pair_ptr + cd.pair_cnt, 0, true});
                for (i32 pi = cd.pair_ptr; pi < cd.pair_ptr + cd.pair_cnt; ++pi) {
                    const PairDesc &pd = S.pairs[pi];
                    if (pd.llen == 0) continue;
                    const i32 j = (i32)(std::upper_bound(Lp.begin(), Lp.end(), pd.lstart - 1) - Lp.begin()) - 1;
                    pair_col[pi] = j;
                    if (ring_valid(j)) { pair_src[pi] = ring_pos[j]; W.near_fma += pd.llen; }
                    else { pair_far[pi] = 1; W.far_fma += pd.llen; }
                }
            }
            if (listsched) {
            // chunks: list schedule of the group's update operations.  What must hold (and all that must hold) for the
            // results to stay bit-identical: the operations on one accumulator slot run in their natural order, and a
            // pair starts only when its multiplier is final, i.e. when nothing is left that targets the multiplier's
            // slot (every such operation precedes the pair in natural order, because the pairs of a column are in
            // topological order).  Ready operations are taken by height (longest dependent chain first); a column
            // offers the operations of its first `pair_window` unfinished pairs only, which bounds the landing runs.
            struct GOp { i32 pair, t, base, tgt, mult, height, next_same, col; };
            std::vector<GOp> gops;
            std::vector<i32> col_first(g.cols.size()), col_end(g.cols.size());
            for (size_t c = 0; c < g.cols.size(); ++c) {
                const ColDesc &cd = S.cols[g.cols[c]];
                col_first[c] = cd.pair_ptr; col_end[c] = cd.pair_ptr + cd.pair_cnt;
                for (i32 pi = cd.pair_ptr; pi < cd.pair_ptr + cd.pair_cnt; ++pi) {
                    const PairDesc &pd = S.pairs[pi];
                    for (i32 t = 0; t < pd.llen; ++t)
                        gops.push_back({pi, t, g.base[c], g.base[c] + S.upd_map[(size_t)pd.mapstart + t], g.base[c] + pd.moff, 0, -1, (i32)c});
                }
            }
            const i32 nops = (i32)gops.size();
            std::vector<i32> head((size_t)g.nslots, -1), pending((size_t)g.nslots, 0);
            {
                std::vector<i32> hq((size_t)g.nslots, 0), hm((size_t)g.nslots, 0);
                for (i32 o = nops - 1; o >= 0; --o) {
                    GOp &u = gops[(size_t)o];
                    u.height = 1 + std::max(hq[u.tgt], hm[u.tgt]);
                    hq[u.tgt] = u.height;
                    hm[u.mult] = std::max(hm[u.mult], u.height);
                    u.next_same = head[u.tgt];
                    head[u.tgt] = o;
                    ++pending[u.tgt];
                }
            }
            i32 pair_lo = INT32_MAX, pair_hi = -1;
            for (size_t c = 0; c < g.cols.size(); ++c) { pair_lo = std::min(pair_lo, col_first[c]); pair_hi = std::max(pair_hi, col_end[c]); }
            std::vector<char> started;                            // far pairs whose landing run is placed
            std::vector<i32> pair_left;                           // operations a pair still has to go
            if (pair_hi > pair_lo) { started.assign((size_t)(pair_hi - pair_lo), 0); pair_left.assign((size_t)(pair_hi - pair_lo), 0); }
            for (const GOp &u : gops) ++pair_left[(size_t)(u.pair - pair_lo)];
            struct Run1 { i32 start, len, left; };                // landing run: entries and operations still to read it
            std::vector<Run1> runs;
            std::vector<i32> run_of(pair_left.size(), -1), pair_owner(pair_left.size(), 0);
            for (size_t c = 0; c < g.cols.size(); ++c) for (i32 pi = col_first[c]; pi < col_end[c]; ++pi) pair_owner[(size_t)(pi - pair_lo)] = (i32)c;
            i32 left = nops;
            std::vector<i32> ready;
            struct Unit { i32 a, b, prio; };
            std::vector<Unit> units, taken;
            while (left > 0 && !sched_failed) {
                Rec ch; ch.kind = 1; ch.group = gi;
                const i32 r = (i32)recs.size();                   // index of this chunk record
                // pair window of every column
                std::vector<i32> limit(g.cols.size());
                for (size_t c = 0; c < g.cols.size(); ++c) {
                    while (col_first[c] < col_end[c] && pair_left[(size_t)(col_first[c] - pair_lo)] == 0) ++col_first[c];
                    i32 seen = 0, pi = col_first[c];
                    for (; pi < col_end[c] && seen < pair_window; ++pi) if (pair_left[(size_t)(pi - pair_lo)] > 0) ++seen;
                    limit[c] = pi;
                }
                ready.clear();
                for (i32 sl = 0; sl < g.nslots; ++sl) {
                    const i32 o = head[(size_t)sl];
                    if (o < 0) continue;
                    const GOp &q = gops[(size_t)o];
                    if (pending[(size_t)q.mult] == 0 && q.pair < limit[(size_t)q.col]) ready.push_back(o);
                }
                // lane-group units: up to two ready operations of one pair (they share the multiplier load)
                std::sort(ready.begin(), ready.end(), [&](i32 x, i32 y) {
                    const GOp &a1 = gops[(size_t)x], &b1 = gops[(size_t)y];
                    if (a1.pair != b1.pair) return a1.pair < b1.pair;
                    if (a1.height != b1.height) return a1.height > b1.height;
                    return x < y;
                });
                units.clear();
                for (size_t x = 0; x < ready.size();) {
                    const bool two = x + 1 < ready.size() && gops[(size_t)ready[x + 1]].pair == gops[(size_t)ready[x]].pair;
                    units.push_back({ready[x], two ? ready[x + 1] : -1, gops[(size_t)ready[x]].height});
                    x += two ? 2 : 1;
                }
                std::stable_sort(units.begin(), units.end(), [&](const Unit &x, const Unit &y) {
                    if (x.prio != y.prio) return x.prio > y.prio;
                    return (x.b >= 0) > (y.b >= 0);
                });
                taken.clear();
                // a new landing run: fetched by one of the records r - lookahead - 3 .. r - lookahead when the source
                // column is final by then and a fetch slot and landing entries are free; an immediate fetch (the
                // kernel waits for it) only when the chunk would otherwise be empty
                auto place = [&](i32 pi, bool immediate) -> bool {
                    // the run: this pair and, when the following pairs of the column take the next L columns (a
                    // supernode-like chain), those too -- their values are adjacent in the workspace, one fetch
                    const i32 c = pair_owner[(size_t)(pi - pair_lo)];
                    i32 last = pi;
                    const i32 src0 = S.pairs[pi].lstart;
                    i32 len = S.pairs[pi].llen, need = final_rec[pair_col[pi]];
                    if (!immediate) {
                        for (i32 pn = pi + 1; pn < col_end[(size_t)c]; ++pn) {
                            const PairDesc &pn_d = S.pairs[pn];
                            if (pn_d.llen == 0 || !pair_far[pn] || started[(size_t)(pn - pair_lo)] || pair_col[pn] != pair_col[last] + 1) break;
                            const i32 nlen = pn_d.lstart + pn_d.llen - src0;
                            if (nlen > run_cap) break;
                            len = nlen; last = pn;
                            need = std::max(need, final_rec[pair_col[pn]]);
                        }
                    }
                    i32 s0 = -1, x = -1;
                    if (!immediate) {
                        for (x = r - kWideLookahead; x >= std::max(0, r - kWideLookahead - 3); --x) {
                            if (recs[(size_t)x].fetch.len != 0 || need > x || need == 0) continue;   // slot taken / source not final
                            s0 = land_alloc(len, x);
                            if (s0 >= 0) break;
                        }
                        if (s0 < 0) return false;
                        recs[(size_t)x].fetch.len = len; recs[(size_t)x].fetch.dst = ring_entries + s0; recs[(size_t)x].fetch.src = src0;
                    } else {
                        if (ch.fetch.len != 0) return false;
                        s0 = land_alloc(len, r);
                        if (s0 < 0) return false;
                        ch.fetch.len = len; ch.fetch.dst = ring_entries + s0; ch.fetch.src = src0;
                        ch.immediate = true;
                        ++W.immediate_fetches;
                    }
                    for (i32 e = 0; e < len; ++e) free_at[(size_t)(s0 + e)] = INT32_MAX;
                    runs.push_back({s0, len, 0});
                    for (i32 pn = pi; pn <= last; ++pn) {
                        if (S.pairs[pn].llen == 0) continue;
                        pair_src[pn] = ring_entries + s0 + (S.pairs[pn].lstart - src0);
                        started[(size_t)(pn - pair_lo)] = 1;
                        run_of[(size_t)(pn - pair_lo)] = (i32)runs.size() - 1;
                        runs.back().left += pair_left[(size_t)(pn - pair_lo)];
                    }
                    return true;
                };
                for (int force = 0; force < 2 && taken.empty(); ++force) {
                    for (const Unit &u : units) {
                        if ((i32)taken.size() >= groups) break;
                        const i32 pi = gops[(size_t)u.a].pair;
                        if (pair_far[pi] && !started[(size_t)(pi - pair_lo)]) {
                            if (!place(pi, force != 0)) continue;
                        }
                        taken.push_back(u);
                    }
                }
                if (taken.empty()) { sched_failed = true; break; }
                // the emitter wants the operations of a pair next to each other
                std::stable_sort(taken.begin(), taken.end(), [&](const Unit &x, const Unit &y) { return gops[(size_t)x.a].pair < gops[(size_t)y.a].pair; });
                for (const Unit &u : taken) {
                    for (i32 o : {u.a, u.b}) {
                        if (o < 0) continue;
                        const GOp &q = gops[(size_t)o];
                        ch.ops.push_back({q.pair, q.t, q.base});
                        head[(size_t)q.tgt] = q.next_same;
                        --pending[(size_t)q.tgt];
                        --left;
                        --pair_left[(size_t)(q.pair - pair_lo)];
                        if (pair_far[q.pair]) {
                            Run1 &ru = runs[(size_t)run_of[(size_t)(q.pair - pair_lo)]];
                            if (--ru.left == 0)                                   // the run may be overwritten from the next record on
                                for (i32 e = 0; e < ru.len; ++e) free_at[(size_t)(ru.start + e)] = r + 1;
                        }
                    }
                }
                recs.push_back(ch); ++chunk_id; ++g.chunk_cnt;
            }
            if (sched_failed) break;
            } else {
            // chunks: round robin over the columns, every column strictly in its own order
            size_t remaining = g.cols.size();
            std::vector<char> fin_col(g.cols.size(), 0);
            while (remaining > 0) {
                Rec ch; ch.kind = 1; ch.group = gi;
                bool any = false;
                // the two entries of a lane group share one multiplier load: they must belong to the same pair, so a
                // chunk holds at most `groups` lane groups' worth of (pair, up to two operations)
                i32 lane_groups = 0, run_pair = -1, run_cnt = 0;
                for (size_t c = 0; c < g.cols.size(); ++c) {
                    Cur &u = cs[c];
                    while (!fin_col[c]) {
                        while (u.pi < u.pend && (S.pairs[u.pi].llen == 0 || u.t >= S.pairs[u.pi].llen)) { ++u.pi; u.t = 0; u.first_of_pair = true; }
                        if (u.pi >= u.pend) { fin_col[c] = 1; --remaining; break; }
                        const PairDesc &pd = S.pairs[u.pi];
                        const i32 tgt = g.base[c] + S.upd_map[(size_t)pd.mapstart + u.t], mult = g.base[c] + pd.moff;
                        const bool new_group = (u.pi != run_pair) || (run_cnt % 2 == 0);       // needs another lane group
                        if ((new_group && lane_groups >= groups) || stamp_t[tgt] == chunk_id || stamp_t[mult] == chunk_id) break;
                        // at most one pair per chunk may bring in a new landing run (one fetch slot per record)
                        if (u.first_of_pair && pair_far[u.pi] && ch.new_far >= 0) break;
                        if (u.first_of_pair && pair_far[u.pi]) ch.new_far = u.pi;
                        u.first_of_pair = false;
                        if (u.pi != run_pair) { run_pair = u.pi; run_cnt = 0; }
                        if (run_cnt % 2 == 0) ++lane_groups;
                        ++run_cnt;
                        ch.ops.push_back({u.pi, u.t, g.base[c]});
                        stamp_t[tgt] = chunk_id;
                        ++u.t;
                        any = true;
                    }
                }
                if (any) { recs.push_back(ch); ++chunk_id; ++g.chunk_cnt; }
            }
            }
            // finalisation: cache the strict L part of every column; fixed-size finalisation records, U entries (copy)
            // and L entries (divide) in separate records so that the kernel's two paths never diverge inside a record
            std::vector<FinEnt> fu, fl;
            for (size_t c = 0; c < g.cols.size(); ++c) {
                const i32 k = g.cols[c];
                const ColDesc &cd = S.cols[k];
                const i32 len = Lp[k + 1] - Lp[k] - 1;
                i32 pos = -1;
                if (len > 0 && len <= ring_entries) {
                    if (cur + len > ring_entries) cur = 0;
                    pos = cur;
                    for (i32 t = 0; t < len; ++t) owner[pos + t] = k;
                    cur += len;
                    ring_pos[k] = pos;
                }
                for (i32 t = 0; t < cd.ucnt; ++t) fu.push_back({cd.up + t, g.base[c] + t, -1});
                for (i32 t = 0; t < cd.lcnt - 1; ++t)
                    fl.push_back({(i32)(0x80000000u | ((uint32_t)c << 28) | (uint32_t)(cd.lp + 1 + t)), g.base[c] + cd.ucnt + t, pos >= 0 ? pos + t : -1});
            }
            for (int kind = 0; kind < 2; ++kind) {
                const std::vector<FinEnt> &fe = kind ? fl : fu;
                for (size_t o = 0; o < fe.size(); o += (size_t)cap) {
                    Rec fr; fr.kind = 2; fr.group = gi;
                    fr.fins.assign(fe.begin() + (long)o, fe.begin() + (long)std::min(fe.size(), o + (size_t)cap));
                    recs.push_back(fr);
                    ++g.fin_cnt;
                }
            }
            for (i32 k : g.cols) final_rec[k] = g.first_rec + 1 + g.chunk_cnt + g.fin_cnt;   // == recs.size()
        }
        if (sched_failed) break;
        const i32 nrec = (i32)recs.size();
        std::vector<i32> last_use((size_t)S.pairs.size(), -1);
        for (i32 r = 0; r < nrec; ++r)
            for (const Op &o : recs[r].ops) if (pair_far[o.pair]) last_use[o.pair] = r;

        // ---- phase 2 (in-order packer only): landing-area allocation in issue order -----------------------------------
        if (!listsched) {
        struct Run { i32 pair, start, len; };
        std::vector<Run> live;
        i32 scur = 0;
        auto overlaps = [&](i32 s, i32 len) {
            for (const Run &u : live) if (s < u.start + u.len && u.start < s + len) return true;
            return false;
        };
        auto alloc = [&](i32 len, i32 pair) -> i32 {
            const i32 positions = stage_entries - len + 1;
            if (positions <= 0) return -1;
            const i32 first = scur < positions ? scur : 0;
            for (i32 probe = 0; probe < positions; ++probe) {
                const i32 s = (first + probe) % positions;
                if (!overlaps(s, len)) { live.push_back({pair, s, len}); scur = s + len; return s; }
            }
            return -1;
        };
        auto drop = [&](i32 pair) {
            for (size_t u = 0; u < live.size(); ++u) if (live[u].pair == pair) { live.erase(live.begin() + (long)u); return; }
        };
        std::vector<char> served((size_t)S.pairs.size(), 0);
        bool failed = false;
        for (i32 x = 0; x < nrec && !failed; ++x) {
            const i32 r = x + kWideLookahead;
            if (r < nrec && recs[r].new_far >= 0) {
                const i32 pi = recs[r].new_far, j = pair_col[pi];
                const PairDesc &pd = S.pairs[pi];
                const bool final_by_now = final_rec[j] <= x;         // column j was finalised before record x starts
                const i32 s = final_by_now ? alloc(pd.llen, pi) : -1;
                if (s >= 0) {
                    recs[x].fetch.len = pd.llen; recs[x].fetch.dst = ring_entries + s; recs[x].fetch.src = pd.lstart;
                    pair_src[pi] = ring_entries + s;
                    served[pi] = 1;
                }
            }
            if (recs[x].new_far >= 0 && !served[recs[x].new_far]) {
                const i32 pi = recs[x].new_far;
                const PairDesc &pd = S.pairs[pi];
                if (recs[x].fetch.len != 0) {
                    const i32 p2 = recs[x + kWideLookahead].new_far;
                    drop(p2);
                    recs[x].fetch = Fetch();
                    served[p2] = 0;
                }
                i32 s = alloc(pd.llen, pi);
                while (s < 0) {
                    i32 victim = -1, vrec = -1;
                    for (i32 y = x + 1; y < std::min(nrec, x + kWideLookahead + 1); ++y)
                        if (recs[y].new_far >= 0 && served[recs[y].new_far]) { victim = recs[y].new_far; vrec = y; }
                    if (victim < 0) { failed = true; break; }
                    recs[vrec - kWideLookahead].fetch = Fetch();
                    served[victim] = 0;
                    drop(victim);
                    s = alloc(pd.llen, pi);
                }
                if (failed) break;
                recs[x].fetch.len = pd.llen; recs[x].fetch.dst = ring_entries + s; recs[x].fetch.src = pd.lstart;
                pair_src[pi] = ring_entries + s;
                served[pi] = 1;
                recs[x].immediate = true;
                ++W.immediate_fetches;
            }
            for (size_t u = 0; u < live.size();) {
                if (last_use[live[u].pair] == x) live.erase(live.begin() + (long)u); else ++u;
            }
        }
        if (failed) { *why = "wide refactor: landing area too small"; return false; }
        }

        // ---- phase 3: emit ---------------------------------------------------------------------------------------------
        std::vector<std::vector<uint8_t>> blobs((size_t)nrec);
        for (i32 r = 0; r < nrec; ++r) {
            const Rec &R = recs[r];
            std::vector<uint8_t> &b = blobs[r];
            Bytes B(b);
            const i64 fdst16 = (i64)((lsrc0 + R.fetch.dst) * entry / 16), funits = (i64)((size_t)R.fetch.len * entry / 16);
            const i32 fsrc16 = (i32)((size_t)R.fetch.src * entry / 16);
            if (R.kind == 0) {
                const bool pre = R.group < 0;
                const Group *g = pre ? nullptr : &G[R.group];
                const Group *nx = (R.group + 1 < ngroups) ? &G[R.group + 1] : nullptr;
                const Group *pf = (R.group + kWidePfGroups < ngroups && R.group + kWidePfGroups >= 0) ? &G[R.group + kWidePfGroups] : nullptr;
                // A lists of a group: its columns one after the other
                auto a_list = [&](const Group &gg, std::vector<std::pair<i32, i32>> &out) {       // (slot, src)
                    for (size_t c = 0; c < gg.cols.size(); ++c) {
                        const ColDesc &cd = S.cols[gg.cols[c]];
                        for (i32 t = 0; t < cd.a_cnt; ++t) out.emplace_back(gg.base[c] + S.a_off[cd.a_ptr + t], S.a_src[cd.a_ptr + t]);
                    }
                };
                std::vector<std::pair<i32, i32>> own, next;
                if (g) a_list(*g, own);
                if (nx) a_list(*nx, next);
                const i32 an_cnt = std::min((i32)next.size(), cover);
                B.i32v(0); B.i32v(0);
                B.u16v(pre ? 0 : (i64)g->cols.size()); B.u16v(pre ? 0 : g->nslots);
                B.u16v((i64)own.size()); B.u16v(pre ? 0 : g->chunk_cnt);
                B.u16v(pre ? 0 : g->fin_cnt); B.u16v(an_cnt);
                B.u16v(fdst16); B.u16v(funits); B.i32v(fsrc16);
                B.u16v(pf ? (i64)pf->cols.size() : 0); B.u16v(0); B.u16v(0); B.u16v(0);
                B.i32v(0); B.i32v(0); B.i32v(0);
                for (i32 c = 0; c < kWideGroupCols; ++c) {
                    if (g && c < (i32)g->cols.size()) {
                        const ColDesc &cd = S.cols[g->cols[c]];
                        B.i32v(g->cols[c] + 1); B.u16v((i64)((size_t)(g->base[c] + cd.ucnt - 1) * entry)); B.u16v(0);
                    } else { B.i32v(0); B.u16v(0); B.u16v(0); }
                }
                for (i32 c = 0; c < kWideGroupCols; ++c) {
                    i32 pf_src = -1, pf_cnt = 0;
                    if (pf && c < (i32)pf->cols.size()) {
                        const ColDesc &cd = S.cols[pf->cols[c]];
                        i32 lo = INT32_MAX, hi = -1;
                        for (i32 t = 0; t < cd.a_cnt; ++t) { lo = std::min(lo, S.a_src[cd.a_ptr + t]); hi = std::max(hi, S.a_src[cd.a_ptr + t]); }
                        if (hi >= 0) { pf_src = lo; pf_cnt = hi - lo + 1; }
                    }
                    B.i32v(pf_src); B.i32v(pf_cnt);
                }
                for (auto &a : own) B.u16v((i64)((size_t)a.first * entry));
                B.pad(4);
                for (i32 t = 0; t < an_cnt; ++t) B.i32v(next[(size_t)t].second);
                for (size_t t = (size_t)cover; t < own.size(); ++t) B.i32v(own[t].second);         // own overflow
                B.pad(16);
            } else if (R.kind == 1) {
                B.i32v(fsrc16); B.u16v(fdst16); B.u16v(funits);
                B.u16v(R.immediate ? 1 : 0); B.u16v(0); B.i32v(0);
                // Slot assignment.  Lane group e executes entries e and e + groups with ONE multiplier load, so both come
                // from the same pair.  Entries 2q and 2q+1 of either half are served by the same shared-memory wavefront
                // (two 64-byte entries per 128-byte wavefront for 8-system bundles): they collide when they are different
                // entries of the same parity.
                std::vector<i32> slot_op((size_t)cap, -1);
                auto op_off = [&](i32 x, int which) -> i64 {                   // byte offset of an operand: 0 source, 1 multiplier, 2 target
                    const Op &o = R.ops[(size_t)x];
                    const PairDesc &pd = S.pairs[o.pair];
                    if (which == 0) return (i64)((lsrc0 + pair_src[o.pair] + o.t) * entry);
                    if (which == 1) return (i64)((size_t)(o.base + pd.moff) * entry);
                    return (i64)((size_t)(o.base + S.upd_map[(size_t)pd.mapstart + o.t]) * entry);
                };
                for (const Op &o : R.ops)
                    if (pair_src[o.pair] < 0) { *why = "wide refactor: internal error (unresolved source)"; return false; }
                if (entry == 64 && groups == 8) {
                    // Units of two operations of one pair, then the perfect matching of the 8 units into wavefront
                    // partners (and which operation of a unit goes to which half) with the fewest collisions over the
                    // source loads, the multiplier load and the accumulator load + store: 105 matchings, exhaustive.
                    struct Unit2 { i32 a, b; };
                    std::vector<Unit2> un;
                    for (size_t o = 0; o < R.ops.size();) {
                        const bool two = o + 1 < R.ops.size() && R.ops[o + 1].pair == R.ops[o].pair;
                        un.push_back({(i32)o, two ? (i32)o + 1 : -1});
                        o += two ? 2 : 1;
                    }
                    if ((i32)un.size() > groups) { *why = "wide refactor: internal error (lane groups)"; return false; }
                    while ((i32)un.size() < groups) un.push_back({-1, -1});
                    auto clash = [&](i32 x, i32 y, int which) -> int {
                        if (x < 0 || y < 0) return 0;
                        const i64 p = op_off(x, which), q = op_off(y, which);
                        return (p != q && ((p >> 6) & 1) == ((q >> 6) & 1)) ? 1 : 0;
                    };
                    int cost[8][8], swp[8][8];
                    for (int u = 0; u < 8; ++u)
                        for (int v = u + 1; v < 8; ++v) {
                            int best = 1 << 20, bs = 0;
                            for (int sw = 0; sw < 2; ++sw) {
                                if (sw && un[(size_t)v].b < 0) continue;       // entry a of a lane group must be its valid one
                                const i32 va = sw ? un[(size_t)v].b : un[(size_t)v].a, vb = sw ? un[(size_t)v].a : un[(size_t)v].b;
                                const int c = clash(un[(size_t)u].a, va, 1) + clash(un[(size_t)u].a, va, 0) + 2 * clash(un[(size_t)u].a, va, 2) +
                                              clash(un[(size_t)u].b, vb, 0) + 2 * clash(un[(size_t)u].b, vb, 2);
                                if (c < best) { best = c; bs = sw; }
                            }
                            cost[u][v] = best; swp[u][v] = bs;
                        }
                    int best_total = 1 << 20, best_mate[8], mate[8];
                    for (int i = 0; i < 8; ++i) mate[i] = best_mate[i] = -1;
                    struct Rec2 {
                        static void go(int total, int (&cost)[8][8], int (&mate)[8], int &best_total, int (&best_mate)[8]) {
                            int i = 0;
                            while (i < 8 && mate[i] >= 0) ++i;
                            if (i == 8) { if (total < best_total) { best_total = total; for (int k = 0; k < 8; ++k) best_mate[k] = mate[k]; } return; }
                            if (total >= best_total) return;
                            for (int j = i + 1; j < 8; ++j) {
                                if (mate[j] >= 0) continue;
                                mate[i] = j; mate[j] = i;
                                go(total + cost[i][j], cost, mate, best_total, best_mate);
                                mate[i] = mate[j] = -1;
                            }
                        }
                    };
                    Rec2::go(0, cost, mate, best_total, best_mate);
                    i32 lg = 0;
                    for (int u = 0; u < 8; ++u) {
                        const int v = best_mate[u];
                        if (v < u) continue;
                        const bool sw = swp[u][v] != 0;
                        slot_op[(size_t)lg] = un[(size_t)u].a; slot_op[(size_t)(lg + groups)] = un[(size_t)u].b;
                        slot_op[(size_t)(lg + 1)] = sw ? un[(size_t)v].b : un[(size_t)v].a;
                        slot_op[(size_t)(lg + 1 + groups)] = sw ? un[(size_t)v].a : un[(size_t)v].b;
                        // an empty lane group next to a used one: keep its valid entry (if any) in half a
                        if (slot_op[(size_t)lg] < 0 && slot_op[(size_t)(lg + groups)] >= 0) std::swap(slot_op[(size_t)lg], slot_op[(size_t)(lg + groups)]);
                        W.bank_clashes += cost[u][v];
                        lg += 2;
                    }
                } else {
                    // other bundle widths: the operations of a pair are handed out by target parity (lane group 2q gets
                    // even targets, 2q+1 odd ones) whenever possible
                    const size_t half = entry >= 128 ? 0 : 128 / entry;
                    i32 lg = 0;
                    size_t o = 0;
                    while (o < R.ops.size()) {
                        size_t o2 = o;
                        while (o2 < R.ops.size() && R.ops[o2].pair == R.ops[o].pair) ++o2;
                        std::vector<i32> ev, od;                               // operations of this pair by target parity
                        for (size_t x = o; x < o2; ++x) {
                            const i32 tg = R.ops[x].base + S.upd_map[(size_t)S.pairs[R.ops[x].pair].mapstart + R.ops[x].t];
                            ((half && (tg % (i32)half) >= (i32)half / 2) ? od : ev).push_back((i32)x);
                        }
                        size_t a = 0, b2 = 0;
                        while (a < ev.size() || b2 < od.size()) {
                            // even lane groups prefer even targets, odd lane groups odd targets; two operations each
                            std::vector<i32> &first = (lg % 2 == 0) ? ev : od, &second = (lg % 2 == 0) ? od : ev;
                            size_t &fi = (lg % 2 == 0) ? a : b2, &si = (lg % 2 == 0) ? b2 : a;
                            i32 take[2] = {-1, -1};
                            for (int q = 0; q < 2; ++q) {
                                if (fi < first.size()) take[q] = first[fi++];
                                else if (si < second.size()) take[q] = second[si++];
                            }
                            if (lg >= groups) { *why = "wide refactor: internal error (lane groups)"; return false; }
                            slot_op[(size_t)lg] = take[0];
                            slot_op[(size_t)(lg + groups)] = take[1];
                            ++lg;
                        }
                        o = o2;
                    }
                }
                // Unused entries are still loaded (the kernel does not predicate): they repeat the addresses of their
                // wavefront partner (a broadcast, no extra wavefront) and are marked invalid.
                for (i32 u = 0; u < cap; ++u) {
                    const i32 x = slot_op[(size_t)u] >= 0 ? slot_op[(size_t)u] : slot_op[(size_t)(u ^ 1)];
                    if (x >= 0) {
                        B.u16v(op_off(x, 0)); B.u16v(op_off(x, 1)); B.u16v(op_off(x, 2));
                        B.u16v(slot_op[(size_t)u] >= 0 ? 1 : 0);
                    } else {
                        B.u16v((i64)(lsrc0 * entry)); B.u16v(0); B.u16v(0); B.u16v(0);
                    }
                }
                W.chunk_ops += (i64)R.ops.size();
                ++W.chunks;
            } else {
                B.i32v(fsrc16); B.u16v(fdst16); B.u16v(funits);
                B.u16v((!R.fins.empty() && R.fins[0].gout < 0) ? 16 : 0); B.u16v(0); B.i32v(0);       // bit 4: L entries
                for (i32 u = 0; u < cap; ++u) {
                    if (u < (i32)R.fins.size()) {
                        const FinEnt &f = R.fins[(size_t)u];
                        B.i32v(f.gout); B.u16v((i64)((size_t)f.slot * entry));
                        B.u16v(f.cache >= 0 ? (i64)((lsrc0 + f.cache) * entry) : 0xffff);
                    } else { B.i32v(0); B.u16v(0xffff); B.u16v(0xffff); }
                }
            }
            if (b.size() > max_rec || b.size() > stage) { *why = "wide refactor: internal error (record larger than planned)"; return false; }
        }
        std::vector<uint8_t> &out = W.prog.bytes;
        size_t pos = 0, cur_stage = 0;
        std::vector<size_t> rec_stage((size_t)nrec);
        bool ok = true;
        for (i32 r = 0; r < nrec && ok; ++r) {
            const size_t st = pos / stage;
            const size_t adv = st - cur_stage;
            cur_stage = st;
            rec_stage[r] = st;
            if (adv > 2) { ok = false; break; }
            size_t pad = 0;
            bool wrap = false;
            if (r + 1 < nrec) {
                const size_t next_start = pos + blobs[r].size();
                const size_t next_end = next_start + blobs[r + 1].size();
                if (next_start / prog_ring != (next_end - 1) / prog_ring) { pad = round_up(next_start, prog_ring) - next_start; wrap = true; }
                else if (next_start % prog_ring == 0) wrap = true;
            }
            const uint8_t fl = (uint8_t)((adv << 1) | (wrap ? 8 : 0));
            blobs[r][recs[r].kind == 0 ? 34 : 8] |= fl;
            out.insert(out.end(), blobs[r].begin(), blobs[r].end());
            out.insert(out.end(), pad, 0);
            pos = out.size();
        }
        // The reader requests stage t when it enters stage t - (kWideProgStages - 2) and first touches it up to two
        // records before the first record that starts in it (headers are read ahead); the request rides in the cp.async
        // group of its record and is only waited for kWideLookahead records later.  Check that on the real layout.
        if (ok) {
            std::vector<i32> first_in((size_t)cur_stage + 2, -1);
            for (i32 r = nrec - 1; r >= 0; --r) first_in[rec_stage[r]] = r;
            for (size_t t = (size_t)kWideProgStages - 1; t <= cur_stage && ok; ++t) {
                if (first_in[t] < 0) continue;
                size_t q = t - (size_t)(kWideProgStages - 2);        // the reader requests stage t when it passes stage q
                while (q <= cur_stage && first_in[q] < 0) ++q;
                const i32 a = first_in[q];
                // first touch: one record before the first record of stage t (records are read one ahead, and that
                // record may itself extend into t), relying on the wait of the record before: + 3 records of margin
                if (first_in[t] - a < kWideLookahead + 4) ok = false;
            }
        }
        if (!ok) continue;                                       // larger stages
        while (out.size() % stage) out.push_back(0);
        out.insert(out.end(), stage, 0);                          // guard stage
        W.records = nrec; W.ngroups = ngroups;
        if (getenv("CSP3_DEBUG")) {
            i64 fins = 0;
            for (const Group &g : G) fins += g.fin_cnt;
            fprintf(stderr, "csp3: wide refactor: %d columns in %d groups, %lld chunk records (%.1f ops each), %lld finalisation records, %d records, stage %zu, %s packer, %d immediate fetches, %.2f modelled bank clashes per chunk\n",
                    n, ngroups, (long long)W.chunks, W.chunks ? (double)W.chunk_ops / (double)W.chunks : 0.0, (long long)fins, nrec, stage,
                    listsched ? "list" : "in-order", W.immediate_fetches, W.chunks ? (double)W.bank_clashes / (double)W.chunks : 0.0);
        }
        W.ok = true;
        return true;
    }
    if (!sched_failed) { *why = "wide refactor: program stream does not fit the ring"; return false; }
    }
    *why = "wide refactor: internal error (scheduler)";
    return false;
}

}  // namespace csp3
