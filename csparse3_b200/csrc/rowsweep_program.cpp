// rowsweep_program.cpp -- host compiler of the row-sweep programs (format: rowsweep_program.hpp).
//
// No reference counterpart.  The arithmetic is that of CSparse cs_lsolve / cs_usolve as restated in
// oracle/csp3_oracle.c: per row the updates of the column sweep in the column sweep's order.
#include <algorithm>

#include "rowsweep_program.hpp"

namespace csp3 {

bool compile_row_sweep(const Factor &F, const Schedule &S, bool lower, RowSweepProgram &P, const char **why)
{
    P = RowSweepProgram();
    const i32 n = (i32)F.pinv.size();
    if (n <= 0) { *why = "empty matrix"; return false; }
    const std::vector<i32> &Fp = lower ? F.Lp : F.Up;
    if ((i64)Fp[n] >= (1ll << 26) || (i64)n >= (1ll << 26)) { *why = "row sweep: factor too large for 32-bit byte offsets"; return false; }
    const LevelSet &L = lower ? S.lev_lsolve : S.lev_usolve;
    const std::vector<i32> &rp = lower ? S.lrow_ptr : S.urow_ptr, &rc = lower ? S.lrow_col : S.urow_col, &rpos = lower ? S.lrow_pos : S.urow_pos;
    if ((i64)L.order.size() != n || (i64)rp.size() != (i64)n + 1) { *why = "row sweep: schedule without level sets"; return false; }
    P.levels = (i32)L.nlev();
    std::vector<std::vector<uint32_t>> stream((size_t)P.warps);
    for (i32 lv = 0; lv < P.levels; ++lv) {
        const i32 r0 = L.lptr[lv], r1 = L.lptr[lv + 1];
        for (i32 p0 = r0, pi = 0; p0 < r1; p0 += 8, ++pi) {
            std::vector<uint32_t> &T = stream[(size_t)(pi % P.warps)];
            const i32 rows = std::min(8, r1 - p0);
            i32 maxterms = 0;
            for (i32 g = 0; g < rows; ++g) { const i32 i = L.order[p0 + g]; maxterms = std::max(maxterms, rp[i + 1] - rp[i]); }
            const i32 nch = (maxterms + 7) / 8;
            const size_t base = T.size();
            T.resize(base + kRsPanelHeaderWords + (size_t)nch * kRsChunkWords, kRsNone);
            T[base + 0] = (uint32_t)lv; T[base + 1] = (uint32_t)nch; T[base + 2] = 0; T[base + 3] = 0;
            for (i32 g = 0; g < rows; ++g) {
                const i32 i = L.order[p0 + g];
                T[base + 4 + g] = (uint32_t)i * 64u;
                if (!lower) T[base + 12 + g] = (uint32_t)(F.Up[i + 1] - 1) * 64u;
                const i32 cnt = rp[i + 1] - rp[i];
                for (i32 t = 0; t < cnt; ++t) {
                    // forward: increasing column; backward: decreasing column (the order of the column sweep)
                    const i32 e = lower ? rp[i] + t : rp[i + 1] - 1 - t;
                    const size_t w = base + kRsPanelHeaderWords + (size_t)(t / 8) * kRsChunkWords + (size_t)((t % 8) * 8 + g) * 2;
                    T[w] = (uint32_t)rpos[e] * 64u;
                    T[w + 1] = (uint32_t)rc[e] * 64u;
                }
                P.terms += cnt;
            }
            P.chunks += nch;
            ++P.panels;
        }
    }
    for (i32 w = 0; w < P.warps; ++w) {
        std::vector<uint32_t> &T = stream[(size_t)w];
        const size_t base = T.size();
        T.resize(base + kRsPanelHeaderWords, kRsNone);
        T[base + 0] = kRsEndLevel; T[base + 1] = 0; T[base + 2] = 0; T[base + 3] = 0;
        P.stream_off[w] = (i64)P.words.size();
        P.words.insert(P.words.end(), T.begin(), T.end());
    }
    P.ok = true;
    return true;
}

}  // namespace csp3
