// wide_program.cpp -- host compiler of the "wide" (lane = system) refactor program executed by lu_wide.cu.
// Record formats: program.hpp.  No reference counterpart; the arithmetic the program
// encodes is the frozen-pattern left-looking cs_lu column update of oracle/csp3_oracle.c, operation for
// operation and in the same order.
#include <algorithm>
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

struct Op { i32 pair, t; };      // update t of pair `pair` (index into Schedule::pairs)

struct Rec {
    bool is_col = false;
    i32 k = -1;                 // column (-1: preamble)
    std::vector<Op> ops;        // chunk records
    i32 new_far = -1;           // pair whose landing run is first used by this chunk (-1: none)
    bool immediate = false;
    Fetch fetch;
    i32 ring = 0xffff;          // column records
};

inline size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

bool compile_wide_refactor(const Schedule &S, const Factor &F, i32 width, i32 groups, size_t smem_budget,
                           i32 ring_override, i32 stage_override, WideProgram &W, const char **why)
{
    W = WideProgram();
    const i32 n = (i32)S.cols.size();
    const std::vector<i32> &Lp = F.Lp;
    if (width != 4 && width != 8 && width != 16 && width != 32) { *why = "wide bundle width must be 4, 8, 16 or 32"; return false; }
    if (n == 0) { *why = "empty matrix"; return false; }
    const i32 cap = 2 * groups;                                  // update operations per chunk
    i32 max_llen = 1, max_acnt = 0;
    for (const PairDesc &pd : S.pairs) max_llen = std::max(max_llen, pd.llen);
    for (const ColDesc &cd : S.cols) max_acnt = std::max(max_acnt, cd.a_cnt);
    // ---- record sizes -> program stage size ----------------------------------------------------------------
    const size_t chunk_rec = (size_t)kWideChunkHeader + 8 * (size_t)cap;
    const size_t max_col_rec = round_up(round_up((size_t)kWideColHeader + 2 * (size_t)max_acnt, 4) + 4 * (size_t)max_acnt, 16);
    const size_t max_rec = std::max(chunk_rec, max_col_rec);
    // The reader requests stage cur + kWideProgStages - 2 when it enters stage cur and relies on stage cur + 1
    // having landed: that load was requested kWideProgStages - 3 stages earlier and must be at least
    // kWideLookahead + 3 records old (its group is only waited for kWideLookahead records later, and the kernel
    // reads records up to two ahead).
    size_t stage = 512;
    while (stage < max_rec || (size_t)(kWideProgStages - 3) * (stage / max_rec) < (size_t)kWideLookahead + 3) stage *= 2;
    const size_t prog_ring = (size_t)kWideProgStages * stage;
    // ---- shared-memory geometry ---------------------------------------------------------------------------
    const size_t entry = (size_t)width * 8;
    const i32 acc_slots = (S.max_col_len + 1 + 1) & ~1;
    i32 stage_entries = stage_override > 0 ? stage_override : (i32)round_up((size_t)(kWideLookahead + 1) * max_llen, 8);
    stage_entries = std::max(stage_entries, (i32)round_up((size_t)max_llen, 8));
    const size_t fixed = (size_t)acc_slots * entry + prog_ring + (size_t)stage_entries * entry;
    if (fixed + 8 * entry > smem_budget) { *why = "wide refactor working set exceeds the shared-memory budget"; return false; }
    i32 ring_entries = (i32)((smem_budget - fixed) / entry);
    if (ring_override > 0) ring_entries = std::min(ring_entries, ring_override);
    if ((size_t)(acc_slots + ring_entries + stage_entries) * entry > 0xfff0) {
        // all value offsets are 16-bit byte offsets
        const size_t room = (0xfff0 / entry);
        if (room <= (size_t)acc_slots + (size_t)stage_entries + 8) { *why = "wide refactor: value area too large for 16-bit offsets"; return false; }
        ring_entries = (i32)(room - (size_t)acc_slots - (size_t)stage_entries);
    }
    W.width = width; W.groups = groups; W.acc_slots = acc_slots; W.ring_entries = ring_entries; W.stage_entries = stage_entries;
    W.prog.stage = (i32)stage;
    W.smem_bytes = (size_t)acc_slots * entry + (size_t)(ring_entries + stage_entries) * entry + prog_ring;

    // ---- pass A: L cache (ring) simulation, column by column -> near / far for every pair; chunk packing -------
    // Column k's strict L part is placed in the ring when the column is finalised (after its last record), so every
    // pair of column k sees the ring as the columns before k left it.
    std::vector<i32> ring_pos((size_t)n, -1), owner((size_t)std::max(ring_entries, 1), -1);
    std::vector<i32> pair_src((size_t)S.pairs.size(), -1);       // lsrc entry of the pair's source column (near pairs)
    std::vector<i32> pair_col((size_t)S.pairs.size(), -1);       // source column j
    std::vector<char> pair_far((size_t)S.pairs.size(), 0);
    i32 cur = 0;
    auto ring_valid = [&](i32 j) {
        const i32 pos = ring_pos[j], len = Lp[j + 1] - Lp[j] - 1;
        if (pos < 0) return false;
        for (i32 t = 0; t < len; ++t) if (owner[pos + t] != j) return false;
        return true;
    };
    std::vector<Rec> recs;
    recs.reserve(S.pairs.size() + (size_t)n + 1);
    std::vector<i32> col_rec((size_t)n + 1, 0);          // index of the column record of k (n: one past the end)
    { Rec r; r.is_col = true; r.k = -1; recs.push_back(r); }
    std::vector<i32> stamp_t(65536, -1);                 // accumulator slot -> chunk id that writes it
    i32 chunk_id = 0;
    for (i32 k = 0; k < n; ++k) {
        col_rec[k] = (i32)recs.size();
        { Rec r; r.is_col = true; r.k = k; recs.push_back(r); }
        const ColDesc &cd = S.cols[k];
        Rec ch;
        ch.k = k;
        bool open = false;
        auto flush = [&]() { if (open) { recs.push_back(ch); ch = Rec(); ch.k = k; open = false; ++chunk_id; } };
        for (i32 pi = cd.pair_ptr; pi < cd.pair_ptr + cd.pair_cnt; ++pi) {
            const PairDesc &pd = S.pairs[pi];
            if (pd.llen == 0) continue;
            const i32 j = (i32)(std::upper_bound(Lp.begin(), Lp.end(), pd.lstart - 1) - Lp.begin()) - 1;
            pair_col[pi] = j;
            if (ring_valid(j)) { pair_src[pi] = ring_pos[j]; W.near_fma += pd.llen; }
            else { pair_far[pi] = 1; W.far_fma += pd.llen; }
            bool first_of_pair = true;
            for (i32 t = 0; t < pd.llen; ++t) {
                const i32 tgt = S.upd_map[(size_t)pd.mapstart + t];
                bool fits = open && (i32)ch.ops.size() < cap && stamp_t[tgt] != chunk_id && stamp_t[pd.moff] != chunk_id;
                // at most one pair per chunk may bring in a new landing run (one fetch slot per record)
                if (fits && first_of_pair && pair_far[pi] && ch.new_far >= 0) fits = false;
                if (!fits) { flush(); open = true; }
                if (first_of_pair && pair_far[pi]) ch.new_far = pi;
                first_of_pair = false;
                ch.ops.push_back({pi, t});
                stamp_t[tgt] = chunk_id;
            }
        }
        flush();
        // finalisation of column k: cache its strict L part
        const i32 len = Lp[k + 1] - Lp[k] - 1;
        i32 pos = 0xffff;
        if (len > 0 && len <= ring_entries) {
            if (cur + len > ring_entries) cur = 0;
            pos = cur;
            for (i32 t = 0; t < len; ++t) owner[pos + t] = k;
            cur += len;
            ring_pos[k] = pos;
        }
        recs[col_rec[k]].ring = pos;
    }
    col_rec[n] = (i32)recs.size();
    const i32 nrec = (i32)recs.size();
    // last record that reads each far pair's landing run
    std::vector<i32> last_use((size_t)S.pairs.size(), -1);
    for (i32 r = 0; r < nrec; ++r)
        for (const Op &o : recs[r].ops) if (pair_far[o.pair]) last_use[o.pair] = r;

    // ---- pass B: landing-area allocation in issue order ------------------------------------------------------
    struct Run { i32 pair, start, len; };
    std::vector<Run> live;
    i32 scur = 0;
    auto overlaps = [&](i32 s, i32 len) {
        for (const Run &u : live) if (s < u.start + u.len && u.start < s + len) return true;
        return false;
    };
    auto alloc = [&](i32 len, i32 pair) -> i32 {
        // first fit from the cursor, wrapping once; a run is never split by the wrap
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
    for (i32 x = 0; x < nrec; ++x) {
        // (1) look-ahead fetch for the chunk kWideLookahead records from now
        const i32 r = x + kWideLookahead;
        if (r < nrec && recs[r].new_far >= 0) {
            const i32 pi = recs[r].new_far, j = pair_col[pi];
            const PairDesc &pd = S.pairs[pi];
            const bool final_by_now = col_rec[j + 1] <= x;      // column j was finalised before record x starts
            const i32 s = final_by_now ? alloc(pd.llen, pi) : -1;
            if (s >= 0) {
                recs[x].fetch.len = pd.llen; recs[x].fetch.dst = ring_entries + s; recs[x].fetch.src = pd.lstart;
                pair_src[pi] = ring_entries + s;
                served[pi] = 1;
            }
        }
        // (2) chunks whose new run could not be requested ahead of time fetch for themselves and wait
        if (recs[x].new_far >= 0 && !served[recs[x].new_far]) {
            const i32 pi = recs[x].new_far;
            const PairDesc &pd = S.pairs[pi];
            if (recs[x].fetch.len != 0) {
                // this record already carries a look-ahead fetch for a later chunk: that one becomes immediate
                const i32 p2 = recs[x + kWideLookahead].new_far;
                drop(p2);
                recs[x].fetch = Fetch();
                served[p2] = 0;
            }
            i32 s = alloc(pd.llen, pi);
            while (s < 0) {
                // make room: cancel the youngest look-ahead fetch (its chunk will fetch for itself later)
                i32 victim = -1, vrec = -1;
                for (i32 y = x + 1; y < std::min(nrec, x + kWideLookahead + 1); ++y)
                    if (recs[y].new_far >= 0 && served[recs[y].new_far]) { victim = recs[y].new_far; vrec = y; }
                if (victim < 0) { *why = "wide refactor: landing area too small"; return false; }
                recs[vrec - kWideLookahead].fetch = Fetch();
                served[victim] = 0;
                drop(victim);
                s = alloc(pd.llen, pi);
            }
            recs[x].fetch.len = pd.llen; recs[x].fetch.dst = ring_entries + s; recs[x].fetch.src = pd.lstart;
            pair_src[pi] = ring_entries + s;
            served[pi] = 1;
            recs[x].immediate = true;
            ++W.immediate_fetches;
        }
        // (3) runs whose last reader is record x are free again
        for (size_t u = 0; u < live.size();) {
            if (last_use[live[u].pair] == x) live.erase(live.begin() + (long)u); else ++u;
        }
    }

    // ---- emit ----------------------------------------------------------------------------------------------------
    std::vector<uint8_t> &out = W.prog.bytes;
    std::vector<std::vector<uint8_t>> blobs((size_t)nrec);
    for (i32 r = 0; r < nrec; ++r) {
        const Rec &R = recs[r];
        std::vector<uint8_t> &b = blobs[r];
        Bytes B(b);
        const i64 fdst16 = (i64)(((size_t)acc_slots + R.fetch.dst) * entry / 16), funits = (i64)((size_t)R.fetch.len * entry / 16);
        const i32 fsrc16 = (i32)((size_t)R.fetch.src * entry / 16);
        if (R.is_col) {
            const i32 k = R.k;
            const bool pre = k < 0;
            const ColDesc *cd = pre ? nullptr : &S.cols[k];
            const ColDesc *nx = (k + 1 < n) ? &S.cols[k + 1] : nullptr;
            const ColDesc *pf = (k + kWidePfCols < n && k + kWidePfCols >= 0) ? &S.cols[k + kWidePfCols] : nullptr;
            const i32 chunk_cnt = pre ? 0 : (k + 1 < n ? col_rec[k + 1] : nrec) - col_rec[k] - 1;
            if (chunk_cnt > 0xffff) { *why = "wide refactor: too many chunks in a column"; return false; }
            B.i32v(pre ? -1 : cd->up); B.i32v(pre ? 0 : cd->lp);
            B.u16v(pre ? 0 : cd->ucnt); B.u16v(pre ? 1 : cd->lcnt);
            B.u16v(pre ? 0 : cd->a_cnt); B.u16v(chunk_cnt);
            const i32 cover = kWideARegs * groups;          // A values the kernel keeps in registers
            const i32 an_cnt = nx ? std::min(nx->a_cnt, cover) : 0;
            B.u16v(R.ring); B.u16v(an_cnt);
            B.u16v(fdst16); B.u16v(funits); B.i32v(fsrc16);
            i32 pf_src = -1, pf_cnt = 0;
            if (pf && pf->a_cnt > 0) {
                i32 lo = INT32_MAX, hi = -1;
                for (i32 t = 0; t < pf->a_cnt; ++t) { lo = std::min(lo, S.a_src[pf->a_ptr + t]); hi = std::max(hi, S.a_src[pf->a_ptr + t]); }
                pf_src = lo; pf_cnt = std::min(hi - lo + 1, 0xffff);
            }
            B.i32v(pf_src); B.u16v(pf_cnt); B.u16v(0);
            B.i32v(0); B.i32v(0); B.i32v(0);
            if (!pre) for (i32 t = 0; t < cd->a_cnt; ++t) B.u16v((i64)(S.a_off[cd->a_ptr + t] * entry));
            B.pad(4);
            for (i32 t = 0; t < an_cnt; ++t) B.i32v(S.a_src[nx->a_ptr + t]);
            if (!pre) for (i32 t = cover; t < cd->a_cnt; ++t) B.i32v(S.a_src[cd->a_ptr + t]);      // own overflow
            B.pad(16);
        } else {
            B.i32v(fsrc16); B.u16v(fdst16); B.u16v(funits);
            B.u16v(R.immediate ? 1 : 0); B.u16v(0); B.i32v(0);
            // Entry order inside a chunk is free (the operations are independent).  Entries 2q and 2q+1 are served
            // by the same shared-memory wavefront: give them targets in different halves of a 128-byte bank row
            // (different slot parity for 64-byte entries) whenever possible, so the accumulator accesses are
            // conflict-free.  An operation whose multiplier differs from the first half's goes last.
            std::vector<Op> ord;
            {
                std::vector<Op> ev, od;
                const size_t half = entry >= 128 ? 0 : 128 / entry;          // entries per bank row (0: no pairing needed)
                for (const Op &o : R.ops) {
                    const i32 tg = S.upd_map[(size_t)S.pairs[o.pair].mapstart + o.t];
                    ((half && (tg % (i32)half) >= (i32)half / 2) ? od : ev).push_back(o);
                }
                size_t a = 0, b2 = 0;
                while (a < ev.size() || b2 < od.size()) {
                    if (a < ev.size()) ord.push_back(ev[a++]);
                    if (b2 < od.size()) ord.push_back(od[b2++]);
                }
            }
            for (i32 u = 0; u < cap; ++u) {
                if (u < (i32)ord.size()) {
                    const Op &o = ord[(size_t)u];
                    const PairDesc &pd = S.pairs[o.pair];
                    if (pair_src[o.pair] < 0) { *why = "wide refactor: internal error (unresolved source)"; return false; }
                    B.u16v((i64)(((size_t)acc_slots + pair_src[o.pair] + o.t) * entry));
                    B.u16v((i64)(pd.moff * entry));
                    B.u16v((i64)(S.upd_map[(size_t)pd.mapstart + o.t] * entry));
                    B.u16v(1);
                } else {
                    B.u16v((i64)((size_t)acc_slots * entry)); B.u16v(0); B.u16v(0); B.u16v(0);
                }
            }
            W.chunk_ops += (i64)R.ops.size();
            ++W.chunks;
        }
        if (b.size() > max_rec) { *why = "wide refactor: internal error (record larger than planned)"; return false; }
    }
    size_t pos = 0, cur_stage = 0;
    for (i32 r = 0; r < nrec; ++r) {
        // flags of record r describe its own start (stages entered) and whether the NEXT record starts at the
        // ring base
        const size_t st = pos / stage;
        const size_t adv = st - cur_stage;
        cur_stage = st;
        if (adv > 2) { *why = "wide refactor: internal error (stage skip)"; return false; }
        size_t pad = 0;
        bool wrap = false;
        if (r + 1 < nrec) {
            const size_t next_start = pos + blobs[r].size();
            const size_t next_end = next_start + blobs[r + 1].size();
            if (next_start / prog_ring != (next_end - 1) / prog_ring) { pad = round_up(next_start, prog_ring) - next_start; wrap = true; }
            else if (next_start % prog_ring == 0) wrap = true;          // lands exactly on the ring base
        }
        const uint8_t fl = (uint8_t)((adv << 1) | (wrap ? 8 : 0));
        blobs[r][recs[r].is_col ? 34 : 8] |= fl;
        out.insert(out.end(), blobs[r].begin(), blobs[r].end());
        out.insert(out.end(), pad, 0);
        pos = out.size();
    }
    while (out.size() % stage) out.push_back(0);
    out.insert(out.end(), stage, 0);                          // guard stage
    W.records = nrec;
    W.ok = true;
    return true;
}

}  // namespace csp3
