// panel_program.hpp -- format of the "panel" refactor program (compiler: panel_program.cpp, kernel: lu_panel.cu).
//
// One warp owns a bundle of S = 8 systems.  Lane = (row group g = lane / 4, h = lane % 4): every lane carries the two
// adjacent systems 2h and 2h + 1 (16-byte accesses), the G = 8 row groups take different words of a step.  Values
// live in the bundle-interleaved factor arrays [entry][S] of the workspace path (same layout as the wide kernels,
// so the sweeps of lu_wide.cu read what this kernel writes) and in shared memory (entry = S doubles = 64 bytes):
//   acc0 / acc1   two accumulators of `nslots` entries, one per column of the panel (combined index = slot + acc * nslots)
//   lsrc          [ring | landing]: `ring` entries hold the L values of the most recently finalised columns (compile-
//                 time managed FIFO, a column is never split by the wrap), `landing` entries receive what the FETCH
//                 words copy from global memory with cp.async (L values of older columns, A values) kPanelLookahead
//                 steps before the word that consumes them
//   program ring  kPanelProgStages stages of kPanelStageSteps steps, filled with cp.async
// Every step commits exactly one cp.async group, so "wait_group kPanelLookahead" at step t guarantees that everything
// requested at steps <= t - kPanelLookahead has landed (program stages ride in the same groups).
//
// The program is a sequence of STEPS of 2 * G = 16 words of 64 bits; row group g reads words 2g and 2g + 1 (one
// 16-byte access).  Every word has its own opcode; word 0 of a step may be a HEADER that all lanes execute.
//
//   bits 60-63  opcode      bits 53-59  flags (7 bits)      bits 40-52  c (13 bits)
//   bits 20-39  b (20 bits)                                     bits  0-19  a (20 bits)          ab = bits 0-39
//
//   NONE     empty word
//   END      (header) end of the program
//   HDRU     (header) start of a task.  c = slot of row j: multipliers u0x = acc_x[c]; WS2: a = lsrc entry of L(j+1,j)
//            (X: entry of the bundle's L array, read synchronously), b = slot of row j+1: u1x = acc_x[b] - L * u0x is
//            computed, kept, and stored back to acc_x[b].  M0 / M1: which accumulators the task updates.  The UPD words
//            that follow (this step and the next ones) belong to this task.
//   UPD      a = lsrc entry of L(row, j), WS2: b = lsrc entry of L(row, j+1), c = slot of the row:
//            acc_x[c] = (acc_x[c] - L[a] * u0x) [- L[b] * u1x]     for x in the task's accumulators
//            (X / Y: a / b is an entry of the bundle's L array, read synchronously -- only when the compiler could
//            neither keep the column in the ring nor fetch it in time)
//   HDRP     (header) ab = column + 1 (status code), c = combined index of the diagonal; FUSED: also load
//            U(k,k+1) = acc1[slot].  Loads the pivot, its refined reciprocal, checks it.
//   FINU     ab = entry of the bundle's U array, c = combined index:        U[ab] = acc[c]; acc[c] = 0
//   FINL     ab = entry of the bundle's L array (bits 0-23) | ring entry (bits 24-39), c = combined index:
//            L[ab] = acc[c] / pivot; acc[c] = 0; X: ring[..] = the same value;
//            FUSED (first column of a two-column panel): acc1[slot] -= L[ab] * U(k,k+1)
//   SCATTER  ab = lsrc entry holding the A value (X: index into the system's Ax, read synchronously), c = combined
//            accumulator index:        acc[c] = value
//   FETCHL   ab = entry of the bundle's L array, c = lsrc entry:     cp.async 64 bytes
//   FETCHA   ab = index into the systems' Ax, c = lsrc entry:        cp.async 8 bytes per system
#pragma once
#include <cstdint>
#include <vector>

#include "symbolic.hpp"

namespace csp3 {

enum : int { kPanelNone = 0, kPanelScatter = 1, kPanelHdrU = 2, kPanelUpd = 3, kPanelHdrP = 4, kPanelFinU = 5, kPanelFinL = 6,
             kPanelFetchL = 7, kPanelFetchA = 8, kPanelEnd = 15 };
enum : int { kPanelWS2 = 2, kPanelM0 = 4, kPanelM1 = 8, kPanelFused = 16, kPanelX = 32, kPanelY = 64 };
constexpr int kPanelLookahead = 8;      // steps between a FETCH word and the word that reads what it fetched
constexpr int kPanelGroups = 8;         // row groups
constexpr int kPanelStepWords = 16;     // words per step
constexpr int kPanelStageSteps = 4;     // steps per program stage (512 bytes: one 16-byte piece per lane)
constexpr int kPanelProgStages = 4;     // stages in the program ring

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
CSP3_HD inline uint64_t panel_word_wide(int op, int flags, uint64_t ab, uint32_t c)
{
    return ((uint64_t)(unsigned)op << 60) | ((uint64_t)(unsigned)(flags & 0x7f) << 53) | ((uint64_t)(c & 0x1fffu) << 40) |
           (ab & 0xffffffffffull);
}
CSP3_HD inline int panel_op(uint64_t w) { return (int)(w >> 60); }

struct PanelProgram {
    bool ok = false;
    i32 width = 0, groups = 0, nslots = 0, steps = 0, npanels = 0, ring = 0, landing = 0;
    size_t smem_bytes = 0;
    i64 ops = 0, upd_rows = 0, sync_loads = 0, fetch_steps = 0, ring_rows = 0;
    Program prog;
};

// smem_budget: bytes one bundle may use (accumulators + ring + landing + program ring)
bool compile_panel_refactor(i64 n, const i32 *Ap, const i32 *Ai, const std::vector<i32> &q, const Factor &F, i32 width,
                            size_t smem_budget, PanelProgram &P, const char **why);

}  // namespace csp3
