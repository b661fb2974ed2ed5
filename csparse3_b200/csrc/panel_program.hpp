// panel_program.hpp -- format of the "panel" refactor program (compiler: panel_program.cpp, kernel: lu_panel.cu).
//
// One warp owns a bundle of S = 8 systems.  Lane = (row group g = lane / 8, system s = lane % 8): the G = 4 row
// groups take different rows of a source column, every lane works on one system.  Values live in the bundle-
// interleaved factor arrays [entry][S] of the workspace path (same layout as the wide kernels, so the sweeps of
// lu_wide.cu read what this kernel writes) and in two shared-memory accumulators acc0 / acc1 of `nslots` entries
// each (one per column of the panel; entry = S doubles; combined index = slot + acc * nslots).
//
// The program is a sequence of STEPS; a step is G consecutive 64-bit words, word g is read by row group g.  All
// words of a step carry the same opcode.
//
//   bits 60-63  opcode      bits 53-59  flags (7 bits)      bits 40-52  c (13 bits)
//   bits 20-39  b (20 bits)                                     bits  0-19  a (20 bits)
// SCATTER / PIV / FINU / FINL carry one 40-bit value `ab` in bits 0-39 instead of a and b.
//
//   END      -- end of the program (the stream is padded with three of them: records are read two steps ahead)
//   NOP      -- separates a FINL step from a step whose L operands are loaded one step ahead
//   SCATTER  ab = index into the system's Ax, c = combined accumulator index        acc[c] = Ax[ab]
//   LOADU    c = slot of row j (multipliers u0x = acc_x[c]); WS2: a = L entry of L(j+1,j), b = slot of row j+1:
//            u1x = acc_x[b] - L[a] * u0x is computed, kept, and stored back to acc_x[b].  M0 / M1: which accumulators
//            the task updates.  All words of the step are equal.
//   UPD      a = L entry of (row, j), WS2: b = L entry of (row, j+1), c = slot of the row:
//            acc_x[c] = (acc_x[c] - L[a] * u0x) [- L[b] * u1x]      for x in the task's accumulators
//   PIV      ab = column + 1 (status code), c = combined index of the diagonal; FUSED: also load U(k,k+1) = acc1[slot]
//            (all words equal).  Loads the pivot, its refined reciprocal, checks it.
//   FINU     ab = entry of the bundle's U array, c = combined index:        U[ab] = acc[c]; acc[c] = 0
//   FINL     ab = entry of the bundle's L array, c = combined index:        L[ab] = acc[c] / pivot; acc[c] = 0;
//            FUSED (first column of a two-column panel): acc1[slot] -= L[ab] * U(k,k+1)
#pragma once
#include <cstdint>
#include <vector>

#include "symbolic.hpp"

namespace csp3 {

enum : int { kPanelEnd = 0, kPanelScatter = 1, kPanelLoadU = 2, kPanelUpd = 3, kPanelPiv = 4, kPanelFinU = 5, kPanelFinL = 6, kPanelNop = 7 };
enum : int { kPanelValid = 1, kPanelWS2 = 2, kPanelM0 = 4, kPanelM1 = 8, kPanelFused = 16 };

#if defined(__CUDACC__)
#define CSP3_HD __host__ __device__
#else
#define CSP3_HD
#endif

CSP3_HD inline uint64_t panel_word(int op, int flags, uint32_t a, uint32_t b, uint32_t c)
{
    return ((uint64_t)(unsigned)op << 60) | ((uint64_t)(unsigned)(flags & 0x7f) << 53) | ((uint64_t)(c & 0x1fffu) << 40) |
           ((uint64_t)(b & 0xfffffu) << 20) | (uint64_t)(a & 0xfffffu);
}
// SCATTER / PIV / FINU / FINL carry one wide index in a:b
CSP3_HD inline uint64_t panel_word_wide(int op, int flags, uint64_t ab, uint32_t c)
{
    return ((uint64_t)(unsigned)op << 60) | ((uint64_t)(unsigned)(flags & 0x7f) << 53) | ((uint64_t)(c & 0x1fffu) << 40) |
           (ab & 0xffffffffffull);
}
CSP3_HD inline int panel_op(uint64_t w) { return (int)(w >> 60); }

struct PanelProgram {
    bool ok = false;
    i32 width = 0, groups = 0, nslots = 0, steps = 0, npanels = 0;
    size_t smem_bytes = 0;
    i64 ops = 0, upd_rows = 0, upd_row_slots = 0;
    i64 step_count[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    Program prog;
};

bool compile_panel_refactor(i64 n, const i32 *Ap, const i32 *Ai, const std::vector<i32> &q, const Factor &F, i32 width,
                            i32 groups, PanelProgram &P, const char **why);

}  // namespace csp3
