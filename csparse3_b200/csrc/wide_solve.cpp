// wide_solve.cpp -- host compiler of the wide triangular-sweep programs executed by lu_wide.cu (record format:
// program.hpp).  The operations and their order are those of CSparse cs_lsolve / cs_usolve as restated in
// oracle/csp3_oracle.c (orc_csc_lsolve, orc_csc_usolve): column after column, every update of a column in
// storage order, so each entry of the solution sees its subtractions in exactly the sequential order.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "program.hpp"
#include "symbolic.hpp"

namespace csp3 {

namespace {

struct Upd { i32 g, mult_row, tgt_row; };
struct Fin { i32 row, div_g, out; };
struct SRec {
    std::vector<Fin> fins;
    std::vector<Upd> upds;
    std::vector<std::pair<i32, i32>> loads;      // (row, gidx)
};

inline void put32(std::vector<uint8_t> &o, i32 v) { const uint8_t *b = (const uint8_t *)&v; o.insert(o.end(), b, b + 4); }
inline void put16(std::vector<uint8_t> &o, i64 v) { const uint16_t w = (uint16_t)v; const uint8_t *b = (const uint8_t *)&w; o.insert(o.end(), b, b + 2); }

}  // namespace

// lower == true : forward sweep on L (columns ascending, no division, factor entries index Lx)
// lower == false: backward sweep on U (columns descending, division by the diagonal, factor entries index Ux)
static bool compile_sweep_window(const Factor &F, bool lower, i32 width, i32 groups, i32 fwd_window, WideSweep &W, const char **why)
{
    W = WideSweep();
    const i32 n = (i32)F.Lp.size() - 1;
    const std::vector<i32> &Gp = lower ? F.Lp : F.Up, &Gi = lower ? F.Li : F.Ui;
    const i32 E = groups / 2, cap_u = 2 * groups, LA = kSweepLookahead;      // E: load / finalisation entries per record
    const size_t entry = (size_t)width * 8;
    if (n <= 0) { *why = "empty matrix"; return false; }
    // ---- list scheduling of the columns into records ---------------------------------------------------------------
    // Bit-exactness only needs (a) every update of row i to use a final multiplier and (b) the updates of row i to
    // be applied in the sequential order (ascending column for cs_lsolve, descending for cs_usolve).  Any column
    // order that respects those two constraints gives the same bits, so the compiler is free to pick, among the
    // columns that are ready, the one that fits the current record and keeps the fewest rows alive: this packs
    // independent columns into one record and (measured on the 2,000-bus Jacobian) cuts the live rows of the
    // backward sweep from 567 to about 200.
    std::vector<std::vector<i32>> touch((size_t)n);          // columns that update row i, in required order
    auto col_begin = [&](i32 j) { return lower ? Gp[j] + 1 : Gp[j]; };
    auto col_end = [&](i32 j) { return lower ? Gp[j + 1] : Gp[j + 1] - 1; };
    for (i32 step = 0; step < n; ++step) {
        const i32 j = lower ? step : n - 1 - step;
        for (i32 p = col_begin(j); p < col_end(j); ++p) touch[Gi[p]].push_back(j);
    }
    std::vector<i32> ptr((size_t)n, 0);
    std::vector<char> done((size_t)n, 0), queued((size_t)n, 0);
    auto is_ready = [&](i32 j) {
        if (ptr[j] != (i32)touch[j].size()) return false;
        for (i32 p = col_begin(j); p < col_end(j); ++p) { const i32 i = Gi[p]; if (touch[i][ptr[i]] != j) return false; }
        return true;
    };
    // priority: the backward sweep finishes short columns first (fewest rows kept alive); the forward sweep keeps the
    // sequential order, which already is a postorder-like walk with a small front
    auto key = [&](i32 j) { return std::make_pair(lower ? 0 : col_end(j) - col_begin(j), lower ? j : n - 1 - j); };
    std::vector<std::pair<std::pair<i32, i32>, i32>> ready;   // sorted by key
    auto push_ready = [&](i32 j) {
        if (done[j] || queued[j] || !is_ready(j)) return;
        queued[j] = 1;
        const auto item = std::make_pair(key(j), j);
        ready.insert(std::upper_bound(ready.begin(), ready.end(), item), item);
    };
    for (i32 j = 0; j < n; ++j) push_ready(j);
    std::vector<SRec> recs(1);
    std::vector<i32> tgt_stamp((size_t)n, -1);           // row -> record that updates it
    auto cur = [&]() -> SRec & { return recs.back(); };
    // Right-hand sides reach their slot through the E load entries of the records at least LA records earlier:
    // a column is only scheduled when the load entries issued so far cover the rows it touches for the first time.
    const i32 pre_extra = 8;
    std::vector<char> touched((size_t)n, 0);
    i64 credit = (i64)E * (pre_extra + 1);
    auto new_rows = [&](i32 j) {
        i32 c = touched[j] ? 0 : 1;
        for (i32 p = col_begin(j); p < col_end(j); ++p) c += touched[Gi[p]] ? 0 : 1;
        return c;
    };
    auto new_record = [&]() { recs.emplace_back(); credit += E; };
    i32 ndone = 0;
    // forward sweep: only columns close to the oldest unfinished one are candidates, otherwise the scheduler runs
    // ahead into far subtrees and the front (live rows) grows several-fold
    static const bool split_columns = !(getenv("CSP3_SWEEP_SPLIT") && atoi(getenv("CSP3_SWEEP_SPLIT")) == 0);
    const i32 window = lower ? fwd_window : n;
    i32 oldest = 0;                                       // natural position of the oldest unfinished column
    auto natural = [&](i32 j) { return lower ? j : n - 1 - j; };
    auto col_at = [&](i32 pos) { return lower ? pos : n - 1 - pos; };
    while (!ready.empty()) {
        while (oldest < n && done[col_at(oldest)]) ++oldest;
        // first ready column that fits the current record entirely
        const i32 rid = (i32)recs.size() - 1;
        size_t pick = ready.size(), spill = ready.size();
        const size_t scan = std::min<size_t>(ready.size(), 96);
        for (size_t c = 0; c < scan && pick == ready.size(); ++c) {
            const i32 j = ready[c].second;
            if (natural(j) > oldest + window) continue;
            if (tgt_stamp[j] == rid || (i32)cur().fins.size() >= E) continue;
            if (new_rows(j) > credit) continue;
            // a column whose first update can still go into this record may start here and spill into the next
            // record(s): its row is finalised now, so the updates it has left are free to follow
            if (spill == ready.size() && split_columns && (i32)cur().upds.size() + 2 <= cap_u && col_begin(j) < col_end(j) &&
                tgt_stamp[Gi[col_begin(j)]] != rid)
                spill = c;
            if ((i32)cur().upds.size() + (col_end(j) - col_begin(j)) > cap_u) continue;
            bool ok = true;
            for (i32 p = col_begin(j); p < col_end(j) && ok; ++p) ok = tgt_stamp[Gi[p]] != rid;
            if (ok) pick = c;
        }
        if (pick == ready.size() && spill != ready.size()) pick = spill;
        if (pick == ready.size()) {
            if (!cur().fins.empty() || !cur().upds.empty()) { new_record(); continue; }
            pick = 0;                                     // empty record: take the best column, it may span records
            if (new_rows(ready[0].second) > credit) { new_record(); continue; }      // wait for load entries
        }
        const i32 j = ready[pick].second;
        ready.erase(ready.begin() + (long)pick);
        credit -= new_rows(j);
        touched[j] = 1;
        for (i32 p = col_begin(j); p < col_end(j); ++p) touched[Gi[p]] = 1;
        cur().fins.push_back({j, lower ? -1 : Gp[j + 1] - 1, j});
        for (i32 p = col_begin(j); p < col_end(j); ++p) {
            const i32 i = Gi[p];
            if ((i32)cur().upds.size() >= cap_u || tgt_stamp[i] == (i32)recs.size() - 1) new_record();
            cur().upds.push_back({p, j, i});
            tgt_stamp[i] = (i32)recs.size() - 1;
        }
        done[j] = 1; ++ndone;
        for (i32 p = col_begin(j); p < col_end(j); ++p) {
            const i32 i = Gi[p];
            ++ptr[i];
            if (ptr[i] < (i32)touch[i].size()) push_ready(touch[i][ptr[i]]);
            else push_ready(i);
        }
    }
    if (ndone != n) { *why = "wide sweep: internal error (scheduler did not finish)"; return false; }
    // ---- first / last use of every row; right-hand-side loads ---------------------------------------------------
    const i32 pre = LA + pre_extra;                       // empty preamble records that only issue loads / prefetches
    const i32 nrec = (i32)recs.size() + pre;
    std::vector<SRec> all((size_t)nrec);
    for (size_t r = 0; r < recs.size(); ++r) all[r + (size_t)pre] = std::move(recs[r]);
    std::vector<i32> first_use((size_t)n, -1), last_use((size_t)n, -1);
    for (i32 r = 0; r < nrec; ++r) {
        for (const Fin &f : all[r].fins) { if (first_use[f.row] < 0) first_use[f.row] = r; last_use[f.row] = r; }
        for (const Upd &u : all[r].upds) {
            if (first_use[u.tgt_row] < 0) first_use[u.tgt_row] = r;
            last_use[u.mult_row] = std::max(last_use[u.mult_row], r);
        }
    }
    std::vector<i32> order((size_t)n), load_rec((size_t)n, -1), load_cnt((size_t)nrec, 0);
    for (i32 i = 0; i < n; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](i32 a, i32 b) { return first_use[a] < first_use[b]; });
    for (i32 i : order) {
        i32 r = first_use[i] - LA;
        while (r >= 0 && load_cnt[r] >= E) --r;
        if (r < 0) { *why = "wide sweep: no room for the right-hand-side loads"; return false; }
        load_rec[i] = r; ++load_cnt[r];
        all[r].loads.emplace_back(i, i);                  // zin is indexed by row
    }
    // ---- slots: a row owns its slot from the record that loads it to its last use ---------------------------------
    std::vector<i32> slot((size_t)n, -1), by_start(order);
    std::stable_sort(by_start.begin(), by_start.end(), [&](i32 a, i32 b) { return load_rec[a] < load_rec[b]; });
    std::vector<std::pair<i32, i32>> busy;                // (free-after record, slot), min-heap
    std::vector<i32> free_slots;
    i32 nslots = 0;
    auto cmp = [](const std::pair<i32, i32> &a, const std::pair<i32, i32> &b) { return a.first > b.first; };
    for (i32 i : by_start) {
        while (!busy.empty() && busy.front().first < load_rec[i]) {
            std::pop_heap(busy.begin(), busy.end(), cmp);
            free_slots.push_back(busy.back().second);
            busy.pop_back();
        }
        i32 s;
        if (!free_slots.empty()) { s = free_slots.back(); free_slots.pop_back(); }
        else s = nslots++;
        slot[i] = s;
        busy.emplace_back(last_use[i], s);
        std::push_heap(busy.begin(), busy.end(), cmp);
    }
    if ((size_t)nslots * entry / 16 > 0xfff0) { *why = "wide sweep: too many live rows for 16-bit slot offsets"; return false; }
    // ---- order of the updates inside a record ------------------------------------------------------------------------
    // Update u is executed by lane group u % groups; entries 2q and 2q+1 of either half share a shared-memory
    // wavefront (two 64-byte slots per 128 bytes for 8-system bundles) and collide when they are different slots of
    // the same parity.  The order inside a record is free (distinct targets, no multiplier is a target), so partners
    // are chosen with opposite target parity, from the same column (same multiplier: a broadcast) when possible.
    if (entry == 64) {
        for (SRec &R : all) {
            if (R.upds.size() < 2) continue;
            std::vector<Upd> ev, od, out_u;
            std::stable_sort(R.upds.begin(), R.upds.end(), [&](const Upd &a, const Upd &b) { return a.mult_row < b.mult_row; });
            std::vector<Upd> lev, lod;                        // leftovers of the columns
            for (size_t o = 0; o < R.upds.size();) {
                size_t o2 = o;
                ev.clear(); od.clear();
                while (o2 < R.upds.size() && R.upds[o2].mult_row == R.upds[o].mult_row) { ((slot[R.upds[o2].tgt_row] & 1) ? od : ev).push_back(R.upds[o2]); ++o2; }
                const size_t m = std::min(ev.size(), od.size());
                for (size_t k = 0; k < m; ++k) { out_u.push_back(ev[k]); out_u.push_back(od[k]); }
                for (size_t k = m; k < ev.size(); ++k) lev.push_back(ev[k]);
                for (size_t k = m; k < od.size(); ++k) lod.push_back(od[k]);
                o = o2;
            }
            // leftovers: opposite target parity first (prefer partners whose multipliers do not collide either)
            while (!lev.empty() && !lod.empty()) {
                const Upd a = lev.back(); lev.pop_back();
                size_t pick = 0;
                for (size_t k = 0; k < lod.size(); ++k)
                    if ((slot[lod[k].mult_row] & 1) != (slot[a.mult_row] & 1)) { pick = k; break; }
                out_u.push_back(a); out_u.push_back(lod[pick]);
                lod.erase(lod.begin() + (long)pick);
            }
            for (const Upd &u : lev) out_u.push_back(u);
            for (const Upd &u : lod) out_u.push_back(u);
            R.upds.swap(out_u);
        }
    }
    // ---- geometry ------------------------------------------------------------------------------------------------
    const size_t rec_bytes = (size_t)wide_solve_record_bytes(groups);
    size_t stage = 512;
    while (stage < rec_bytes || (size_t)(kSweepProgStages - 3) * (stage / rec_bytes) < (size_t)LA + 2) stage *= 2;   // + 1: records are read one ahead
    const size_t prog_ring = (size_t)kSweepProgStages * stage;
    W.width = width; W.groups = groups; W.nslots = nslots; W.records = nrec;
    W.landing_entries = (LA + 1) * (lower ? cap_u : cap_u + E);          // per set: 2*groups update values (+ groups/2 divisors)
    W.prog.stage = (i32)stage;
    W.smem_bytes = ((size_t)nslots + (size_t)W.landing_entries) * entry + prog_ring;
    // ---- emit ----------------------------------------------------------------------------------------------------
    std::vector<uint8_t> &out = W.prog.bytes;
    size_t pos = 0, cur_stage = 0;
    for (i32 r = 0; r < nrec; ++r) {
        const SRec &R = all[r];
        const SRec *A = r + LA < nrec ? &all[r + LA] : nullptr;
        std::vector<uint8_t> b;
        const size_t st = pos / stage, adv = st - cur_stage;
        cur_stage = st;
        if (adv > 2) { *why = "wide sweep: internal error (stage skip)"; return false; }
        size_t pad = 0;
        bool wrap = false;
        if (r + 1 < nrec) {
            const size_t ns = pos + rec_bytes, ne = ns + rec_bytes;
            if (ns / prog_ring != (ne - 1) / prog_ring) { pad = (ns + prog_ring - 1) / prog_ring * prog_ring - ns; wrap = true; }
            else if (ns % prog_ring == 0) wrap = true;
        }
        put16(b, (i64)((adv << 1) | (wrap ? 8 : 0))); put16(b, 0); put32(b, 0); put32(b, 0); put32(b, 0);
        for (i32 e = 0; e < E; ++e) {
            if (e < (i32)R.loads.size()) { put32(b, R.loads[e].second); put16(b, (i64)((size_t)slot[R.loads[e].first] * entry / 16)); put16(b, 0); }
            else { put32(b, -1); put16(b, 0); put16(b, 0); }
        }
        for (i32 u = 0; u < cap_u; ++u) put32(b, (A && u < (i32)A->upds.size()) ? A->upds[u].g : -1);
        for (i32 e = 0; e < E; ++e) put32(b, (A && e < (i32)A->fins.size()) ? A->fins[e].div_g : -1);
        for (i32 e = 0; e < E; ++e) {
            if (e < (i32)R.fins.size()) {
                put32(b, R.fins[e].out); put16(b, (i64)((size_t)slot[R.fins[e].row] * entry / 16)); put16(b, R.fins[e].div_g >= 0 ? 1 : 0);
            } else { put32(b, -1); put16(b, 0); put16(b, 0); }
        }
        for (i32 u = 0; u < cap_u; ++u) {
            if (u < (i32)R.upds.size()) { put16(b, (i64)((size_t)slot[R.upds[u].mult_row] * entry / 16)); put16(b, (i64)((size_t)slot[R.upds[u].tgt_row] * entry / 16)); }
            else { put16(b, 0); put16(b, 0xffff); }
        }
        if (b.size() != rec_bytes) { *why = "wide sweep: internal error (record size)"; return false; }
        out.insert(out.end(), b.begin(), b.end());
        out.insert(out.end(), pad, 0);
        pos = out.size();
        W.ops += (i64)R.upds.size();
    }
    while (out.size() % stage) out.push_back(0);
    out.insert(out.end(), stage, 0);
    W.ok = true;
    return true;
}

// The forward sweep's scheduling window trades records (parallelism) against live rows (shared memory): take the
// widest window whose working set fits the budget that keeps every bundle of a large batch resident.
bool compile_wide_sweep(const Factor &F, bool lower, i32 width, i32 groups, size_t smem_budget, WideSweep &W, const char **why)
{
    if (!lower) return compile_sweep_window(F, lower, width, groups, 0, W, why);
    // the widest window whose live rows fit the shared-memory budget
    const i32 windows[8] = {64, 48, 32, 24, 16, 12, 8, 4};
    bool any = false;
    for (int t = 0; t < 8; ++t) {
        WideSweep T;
        if (!compile_sweep_window(F, lower, width, groups, windows[t], T, why)) continue;
        any = true;
        W = std::move(T);
        if (getenv("CSP3_DEBUG")) fprintf(stderr, "csp3: forward sweep window %d: %d records, %d slots, %zu bytes\n", windows[t], W.records, W.nslots, W.smem_bytes);
        if (W.smem_bytes <= smem_budget) break;
    }
    return any;
}

}  // namespace csp3
