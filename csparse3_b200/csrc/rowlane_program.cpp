// rowlane_program.cpp -- host compiler of the "row-lane" refactor program executed by lu_rowlane.cu.
//
// No reference counterpart.  The arithmetic the program encodes is the frozen-pattern
// left-looking cs_lu column update of oracle/csp3_oracle.c (orc_csc_lu_refactor): for every entry of every column the
// same operations in the same order, so the factors are bit-identical.  Format: rowlane_program.hpp.
//
// The compiler's two decisions:
//  * ORDER of the columns.  Any order in which the sources of a column come first gives the same bits.  The kernel
//    requests the L operands of a whole stage (4 quads) one stage ahead, so a source column should have been finalised
//    two stages before its first use; AMD orders chains of short columns back to back (column k's last source is
//    k - 1).  The columns are list-scheduled: among the first ready columns (in the natural order, which keeps a
//    subtree's columns close to their users for L1 / L2) the one with the fewest update quads that would have to read
//    their operands late, ties to the smallest column number.
//  * PLACEMENT of the operations of a source column in the lane groups of its records.  Lane groups 2i and 2i + 1
//    share a shared-memory wavefront (8 lanes x 16 bytes); two accumulator entries of 64 bytes collide when their
//    slots have the same parity, so an even and an odd slot are paired wherever possible.
#include <algorithm>
#include <array>
#include <cstdio>
#include <cstdlib>
#include <set>

#include "rowlane_program.hpp"

namespace csp3 {

namespace {

struct Op { i32 slot, off; };
using Rec = std::array<uint32_t, kRlOps>;

// ops of ONE role (and, for UPDATE, one source column) -> records of 8 lane words
void arrange(const std::vector<Op> &ops, std::vector<Rec> &recs, i64 &conflicts)
{
    std::vector<Op> ev, od;
    for (const Op &o : ops) (o.slot & 1 ? od : ev).push_back(o);
    const size_t nrec = (ops.size() + kRlOps - 1) / kRlOps;
    std::vector<std::pair<Op, Op>> pairs;          // slot < 0: empty
    const Op none{-1, 0};
    const size_t mixed = std::min(ev.size(), od.size());
    for (size_t t = 0; t < mixed; ++t) pairs.push_back({ev[t], od[t]});
    std::vector<Op> rest(ev.begin() + mixed, ev.end());
    rest.insert(rest.end(), od.begin() + mixed, od.end());
    // leftovers of one parity: alone while pair positions are free, two by two (a 2-way conflict) otherwise
    size_t free_pairs = nrec * (kRlOps / 2) - pairs.size();
    size_t r = 0;
    while (r < rest.size()) {
        const size_t left = rest.size() - r;
        if (left <= free_pairs) { pairs.push_back({rest[r], none}); ++r; }
        else { pairs.push_back({rest[r], rest[r + 1]}); r += 2; ++conflicts; }
        --free_pairs;
    }
    while (pairs.size() < nrec * (kRlOps / 2)) pairs.push_back({none, none});
    for (size_t rec = 0; rec < nrec; ++rec) {
        Rec w;
        for (int g = 0; g < kRlOps; ++g) {
            const auto &pr = pairs[rec * (kRlOps / 2) + g / 2];
            const Op &o = (g & 1) ? pr.second : pr.first;
            const Op &other = (g & 1) ? pr.first : pr.second;
            // an empty lane group repeats its partner's slot (same shared-memory address: no extra wavefront)
            w[g] = o.slot >= 0 ? (0x80000000u | ((uint32_t)o.off << 16) | ((uint32_t)o.slot << 6))
                               : (other.slot >= 0 ? ((uint32_t)other.slot << 6) : 0u);
        }
        recs.push_back(w);
    }
}

struct Quad {
    uint32_t w[kRlQuadWords];
    Quad() { for (auto &x : w) x = 0; for (int i = 44; i < kRlQuadWords; ++i) w[i] = 0xffffffffu; }
    void kind(int k) { w[0] = (w[0] & ~7u) | (uint32_t)k; }
    void mslot(int r, i32 slot) { uint32_t &x = w[1 + r / 2]; x = (r & 1) ? ((x & 0xffffu) | ((uint32_t)slot << 22)) : ((x & 0xffff0000u) | ((uint32_t)slot << 6)); }
    void base(int r, uint32_t b) { w[4 + r] = b; }
    void rec(int r, const Rec &v) { for (int g = 0; g < kRlOps; ++g) w[12 + g * kRlQuadRecords + r] = v[g]; }
    // address words of record r: base + off * scale for the valid lane words
    void addr(int r, uint32_t b, uint32_t scale)
    {
        for (int g = 0; g < kRlOps; ++g) {
            const uint32_t v = w[12 + g * kRlQuadRecords + r];
            w[44 + g * kRlQuadRecords + r] = (v >> 31) ? b + ((v >> 16) & 0x7fffu) * scale : 0xffffffffu;
        }
    }
};

}  // namespace

bool compile_rowlane_refactor(i64 n64, const i32 *Ap, const std::vector<i32> &q, const Factor &F, const Schedule &S, i32 warps,
                              i32 stage_quads, RowlaneProgram &P, const char **why)
{
    P = RowlaneProgram();
    const i32 n = (i32)n64;
    if (n <= 0) { *why = "empty matrix"; return false; }
    if (warps < 1 || warps > kRlMaxWarps || stage_quads < 1 || stage_quads > 4) { *why = "row-lane program: bad geometry"; return false; }
    const std::vector<i32> &Lp = F.Lp, &Up = F.Up, &Ui = F.Ui;
    if ((i64)Lp[n] >= (1ll << 26) || (i64)Up[n] >= (1ll << 26) || (i64)Ap[n] >= (1ll << 29)) {
        *why = "row-lane program: factor too large for 32-bit byte offsets";
        return false;
    }
    if (S.max_col_len > kRlMaxSlots) { *why = "row-lane program: column longer than 1024 entries"; return false; }
    if (warps > 1 && n > 65000 * warps) { *why = "row-lane program: too many columns per warp for 16-bit progress counters"; return false; }
    for (i32 k = 0; k < n; ++k) {
        const i32 col = q.empty() ? k : q[k];
        if (S.cols[k].ucnt > kRlMaxOff || S.cols[k].lcnt > kRlMaxOff || Ap[col + 1] - Ap[col] > kRlMaxOff) {
            *why = "row-lane program: column longer than 32767 entries";
            return false;
        }
    }
    P.warps = warps; P.stage_quads = stage_quads;
    P.nslots = (std::max(S.max_col_len, 1) + 1) & ~1;
    const i64 NQ = stage_quads;

    // users of every column (columns k with U(j,k) != 0) and the count of unfinished sources
    std::vector<i32> uptr((size_t)n + 1, 0), users, ndeps((size_t)n, 0);
    for (i32 k = 0; k < n; ++k)
        for (i32 p = Up[k]; p < Up[k + 1] - 1; ++p) { ++uptr[(size_t)Ui[p] + 1]; ++ndeps[k]; }
    for (i32 j = 0; j < n; ++j) uptr[(size_t)j + 1] += uptr[j];
    users.resize((size_t)uptr[n]);
    {
        std::vector<i32> cur(uptr.begin(), uptr.end() - 1);
        for (i32 k = 0; k < n; ++k)
            for (i32 p = Up[k]; p < Up[k + 1] - 1; ++p) users[(size_t)cur[Ui[p]]++] = k;
    }
    std::set<i32> ready;
    for (i32 k = 0; k < n; ++k) if (ndeps[k] == 0) ready.insert(k);
    const int window = getenv("CSP3_RL_WINDOW") ? std::max(1, atoi(getenv("CSP3_RL_WINDOW"))) : 48;
    const i64 margin = getenv("CSP3_RL_MARGIN") ? atoi(getenv("CSP3_RL_MARGIN")) : 2 * NQ;      // quads of slack for a cross-warp source

    // one stream of quads per warp of the bundle; a column lives on one warp (its accumulator is that warp's)
    struct Stream {
        std::vector<Quad> quads;
        bool have_fin = false;
        Quad fin;
        i32 ncols = 0;
        std::vector<std::array<uint16_t, kRlMaxWarps>> req;       // per stage: columns the other warps must have finished
        i64 clock() const { return (i64)quads.size() + (have_fin ? 1 : 0); }
    };
    std::vector<Stream> st((size_t)warps);
    std::vector<i32> owner((size_t)n, 0), colseq((size_t)n, 0);
    std::vector<i64> stored_stage((size_t)n, 1ll << 40);         // stage (of the owner's stream) whose execution finalises L(:,k)
    std::vector<i64> finish_clock((size_t)n, 1ll << 40);          // owner's quad count when the column is finalised (time proxy)
    auto stage_of = [NQ](i64 quad) { return quad / NQ; };
    // homogeneous quads of `recs[from, to)`, one base for all records
    auto emit_role4 = [&](Stream &T, int kind, const std::vector<Rec> &recs, size_t from, size_t to, uint32_t base, bool &first_store, i32 pivot_slot, i32 k) {
        for (size_t r0 = from; r0 < to; r0 += kRlQuadRecords) {
            Quad Q;
            const size_t cnt = std::min<size_t>(kRlQuadRecords, to - r0);
            Q.kind(kind);
            Q.w[0] |= (uint32_t)cnt << 16;
            for (size_t r = 0; r < cnt; ++r) {
                Q.rec((int)r, recs[r0 + r]); Q.base((int)r, base);
                if (kind == kRlLoad4) Q.addr((int)r, base, 8);
                Q.w[0] |= (uint32_t)(kind == kRlStoreL4 ? kRlRoleL : kind == kRlStoreU4 ? kRlRoleU : kRlRoleA) << (24 + 2 * r);
            }
            if (kind != kRlLoad4 && first_store) { Q.w[0] |= kRlFlagP; Q.mslot(0, pivot_slot); Q.w[3] = (uint32_t)k + 1u; first_store = false; }
            T.quads.push_back(Q);
        }
    };
    struct URec { Rec w; bool isnew; i32 mslot; uint32_t base; i32 j; };
    std::vector<Op> ops;
    std::vector<Rec> recs;
    std::vector<URec> urecs;
    // update quads of column k that would stall if it were eliminated next on warp w: a source of the same warp finalised
    // less than two stages before (late read), or a source of another warp that is finalised about then (wait)
    auto stall_cost = [&](i32 k, i32 w) {
        const Stream &T = st[(size_t)w];
        const ColDesc &cd = S.cols[k];
        const i64 a_recs = (cd.a_cnt + kRlOps - 1) / kRlOps;
        const i64 load4 = T.have_fin ? (std::max<i64>(a_recs, 1) - 1 + kRlQuadRecords - 1) / kRlQuadRecords : (a_recs + kRlQuadRecords - 1) / kRlQuadRecords;
        const i64 pos = T.clock() + load4;
        i64 rec = 0, cost = 0, last_quad = -1;
        for (i32 t = 0; t < cd.pair_cnt; ++t) {
            const PairDesc &pd = S.pairs[(size_t)cd.pair_ptr + t];
            const i32 j = Ui[Up[k] + t];
            const i64 nrec = (pd.llen + kRlOps - 1) / kRlOps;
            for (i64 r = 0; r < nrec; ++r, ++rec) {
                const i64 qd = pos + rec / kRlQuadRecords;
                if (qd == last_quad) continue;
                const bool stall = owner[j] == w ? stored_stage[j] > stage_of(qd) - 2 : finish_clock[j] + margin > qd;
                if (stall) { ++cost; last_quad = qd; }
            }
        }
        return cost;
    };
    P.order.reserve((size_t)n);
    for (i32 done = 0; done < n; ++done) {
        if (ready.empty()) { *why = "row-lane program: dependency cycle (internal error)"; return false; }
        // the warp that is free first (fewest quads so far) takes the next column
        i32 w = 0;
        for (i32 v = 1; v < warps; ++v) if (st[(size_t)v].clock() < st[(size_t)w].clock()) w = v;
        i32 best = -1;
        i64 best_cost = 0;
        int seen = 0;
        for (auto it = ready.begin(); it != ready.end() && seen < window; ++it, ++seen) {
            const i64 c = stall_cost(*it, w);
            if (best < 0 || c < best_cost) { best = *it; best_cost = c; }
            if (c == 0) break;
        }
        const i32 k = best;
        ready.erase(k);
        P.order.push_back(k);
        Stream &T = st[(size_t)w];
        owner[k] = w; colseq[k] = T.ncols++;
        const ColDesc &cd = S.cols[k];
        const i32 col = q.empty() ? k : q[k];
        bool unused = false;
        // ---- A(:,q[k]): the first record rides in the FIN quad of the warp's previous column -----------------------
        ops.clear(); recs.clear();
        for (i32 t = 0; t < cd.a_cnt; ++t) ops.push_back({(i32)S.a_off[(size_t)cd.a_ptr + t], S.a_src[(size_t)cd.a_ptr + t] - Ap[col]});
        arrange(ops, recs, P.conflict_pairs);
        size_t from = 0;
        if (T.have_fin) {
            if (!recs.empty()) { T.fin.w[0] |= kRlHasA | ((uint32_t)kRlRoleA << 28); T.fin.rec(2, recs[0]); T.fin.base(2, (uint32_t)Ap[col] * 8u); T.fin.addr(2, (uint32_t)Ap[col] * 8u, 8); from = 1; }
            T.quads.push_back(T.fin);
            T.have_fin = false;
        }
        emit_role4(T, kRlLoad4, recs, from, recs.size(), (uint32_t)Ap[col] * 8u, unused, 0, k);
        // ---- UPDATE, one source column after the other in the stored order of U(:,k) ------------------------------
        urecs.clear();
        for (i32 t = 0; t < cd.pair_cnt; ++t) {
            const PairDesc &pd = S.pairs[(size_t)cd.pair_ptr + t];
            if (pd.llen == 0) continue;
            const i32 j = Ui[Up[k] + t];
            ops.clear(); recs.clear();
            for (i32 e = 0; e < pd.llen; ++e) ops.push_back({(i32)S.upd_map[(size_t)pd.mapstart + e], e});
            arrange(ops, recs, P.conflict_pairs);
            for (size_t r = 0; r < recs.size(); ++r) urecs.push_back({recs[r], r == 0, pd.moff, (uint32_t)pd.lstart * 64u, j});
            P.ops += pd.llen;
        }
        P.update_records += (i64)urecs.size();
        for (size_t r0 = 0; r0 < urecs.size(); r0 += kRlQuadRecords) {
            Quad Q;
            // A warp waits for the cross-warp sources of a WHOLE stage before it executes the stage: a column this warp
            // finalises earlier in the same stage could be what the other warp is waiting for.  Such a quad starts a new stage.
            bool cross = false;
            for (size_t r = r0; r < std::min(urecs.size(), r0 + kRlQuadRecords); ++r) cross = cross || owner[urecs[r].j] != w;
            if (cross) {
                const size_t first = (size_t)(stage_of((i64)T.quads.size()) * NQ);
                bool fin_before = false;
                for (size_t t = first; t < T.quads.size(); ++t) fin_before = fin_before || (T.quads[t].w[0] & 7u) == (uint32_t)kRlFin;
                if (fin_before) { while (T.quads.size() % (size_t)NQ != 0) { T.quads.push_back(Quad()); ++P.pad_quads; } }
            }
            const i64 stg = stage_of((i64)T.quads.size());
            if ((i64)T.req.size() <= stg) T.req.resize((size_t)stg + 1, std::array<uint16_t, kRlMaxWarps>{});
            bool late = false;
            for (size_t r = r0; r < std::min(urecs.size(), r0 + kRlQuadRecords); ++r) {
                const URec &u = urecs[r];
                const int ri = (int)(r - r0);
                Q.rec(ri, u.w); Q.base(ri, u.base); Q.mslot(ri, u.mslot); Q.addr(ri, u.base, 64);
                if (u.isnew) Q.w[0] |= 0x100u << ri;
                if (owner[u.j] == w) {
                    if (stored_stage[u.j] > stg) { *why = "row-lane program: source column not finalised (internal error)"; return false; }
                    late = late || stored_stage[u.j] > stg - 2;
                } else {
                    // a source of another warp: that warp must have finished colseq + 1 columns before this stage's operands are read
                    uint16_t &rq = T.req[(size_t)stg][(size_t)owner[u.j]];
                    rq = std::max<uint16_t>(rq, (uint16_t)(colseq[u.j] + 1));
                    ++P.cross_records;
                }
            }
            Q.kind(late ? kRlUpdLate : kRlUpdate);
            T.quads.push_back(Q);
            ++P.update_quads;
            if (late) ++P.late_quads;
        }
        // ---- stores: whole records of 8 in STOREL4 / STOREU4 quads, the last record of each role in the FIN quad -----
        bool first_store = true;
        const i32 pivot_slot = cd.ucnt - 1;
        std::vector<Rec> lrecs, urecs2;
        ops.clear();
        for (i32 t = 1; t < cd.lcnt; ++t) ops.push_back({cd.ucnt + t - 1, t - 1});
        arrange(ops, lrecs, P.conflict_pairs);
        ops.clear();
        for (i32 t = 0; t < cd.ucnt; ++t) ops.push_back({t, t});
        arrange(ops, urecs2, P.conflict_pairs);
        if (lrecs.size() > 1) emit_role4(T, kRlStoreL4, lrecs, 0, lrecs.size() - 1, (uint32_t)(cd.lp + 1) * 64u, first_store, pivot_slot, k);
        if (urecs2.size() > 1) emit_role4(T, kRlStoreU4, urecs2, 0, urecs2.size() - 1, (uint32_t)cd.up * 64u, first_store, pivot_slot, k);
        T.fin = Quad();
        T.fin.kind(kRlFin);
        if (!lrecs.empty()) { T.fin.w[0] |= kRlHasL | ((uint32_t)kRlRoleL << 24); T.fin.rec(0, lrecs.back()); T.fin.base(0, (uint32_t)(cd.lp + 1) * 64u); }
        if (!urecs2.empty()) { T.fin.w[0] |= kRlHasU | ((uint32_t)kRlRoleU << 26); T.fin.rec(1, urecs2.back()); T.fin.base(1, (uint32_t)cd.up * 64u); }
        if (first_store) { T.fin.w[0] |= kRlFlagP; T.fin.mslot(0, pivot_slot); T.fin.w[3] = (uint32_t)k + 1u; }
        T.have_fin = true;
        stored_stage[k] = stage_of((i64)T.quads.size());         // the FIN quad is the next quad this warp emits
        finish_clock[k] = T.clock();
        for (i32 p = uptr[k]; p < uptr[(size_t)k + 1]; ++p)
            if (--ndeps[users[(size_t)p]] == 0) ready.insert(users[(size_t)p]);
    }
    for (i32 w = 0; w < warps; ++w)
        if (warps > 1 && st[(size_t)w].ncols > 65535) { *why = "row-lane program: more than 65535 columns on one warp (16-bit progress counters)"; return false; }
    Quad endq;
    endq.kind(kRlEnd);
    P.quads = 0;
    for (i32 w = 0; w < warps; ++w) {
        Stream &T = st[(size_t)w];
        if (T.have_fin) T.quads.push_back(T.fin);
        P.stream_quads[w] = (i32)T.quads.size();
        P.quads += (i32)T.quads.size();
        const size_t padded = (T.quads.size() + (size_t)NQ) / (size_t)NQ * (size_t)NQ + 4 * (size_t)NQ;
        while (T.quads.size() < padded) T.quads.push_back(endq);
        // cross-warp requirements of a stage ride in the spare header words of its first quad (8 x 16 bits)
        for (size_t sg = 0; sg < T.req.size(); ++sg)
            for (int v = 0; v < kRlMaxWarps; ++v) {
                T.quads[sg * (size_t)NQ].w[8 + v / 2] |= (uint32_t)T.req[sg][(size_t)v] << (16 * (v & 1));
                if (T.req[sg][(size_t)v]) T.quads[sg * (size_t)NQ].w[0] |= kRlFlagCross;
            }
        P.stream_off[w] = (i64)P.words.size() / kRlQuadWords;
        const size_t base = P.words.size();
        P.words.resize(base + T.quads.size() * kRlQuadWords);
        for (size_t i = 0; i < T.quads.size(); ++i) std::copy(T.quads[i].w, T.quads[i].w + kRlQuadWords, P.words.begin() + base + i * kRlQuadWords);
    }
    P.smem_bytes = (size_t)warps * ((size_t)P.nslots * 64 + (size_t)kRlRingStages * (size_t)NQ * kRlQuadWords * 4) + 64 + 32 * kRlMaxWarps;
    P.ok = true;
    return true;
}

}  // namespace csp3
