// rowsweep_program.hpp -- format of the "row sweep" programs (compiler: rowsweep_program.cpp, kernel: lu_sweep_rows_kernel
// in lu_wide.cu): the forward / backward triangular sweeps of a SMALL batch on several warps per bundle.
//
// cs_lsolve / cs_usolve (oracle/csp3_oracle.c orc_csc_lsolve / orc_csc_usolve) are column sweeps: x[i] -= F(i,j) x[j] for
// the columns j in order.  Row i therefore receives its updates in increasing j (forward) / decreasing j (backward), and
// the same sequence per row -- unfused multiply / subtract, then the division by U(i,i) -- gives the same bits when it is
// executed ROW BY ROW.  Rows of one level of the dependency graph are independent: the W warps of a CTA (one bundle of 8
// systems, lane = (g = lane / 4, h = lane % 4) as in lu_rowlane.cu) take the rows of a level eight at a time, a CTA
// barrier separates the levels.  Values live in the bundle-interleaved factor arrays and in the z scratch vector
// [n][8] of the workspace (in place).
//
// One stream of PANELS per warp (u32 words), in level order:
//   w[0] level, w[1] chunks, w[2..3] spare                 a panel with level 0x7fffffff ends the stream
//   w[4 + g]   byte offset of the row of lane group g in z (row * 64), 0xffffffff: none
//   w[12 + g]  backward: byte offset of U(i,i) in the bundle's U array
//   chunks x 128 words: chunk c, step t (0..7), lane group g: w[20 + c * 128 + (t * 8 + g) * 2] = byte offset of F(i,j)
//   in the factor array (0xffffffff: no term), the next word = byte offset of x[j] in z
#pragma once
#include <cstdint>
#include <vector>

#include "symbolic.hpp"

namespace csp3 {

constexpr int kRsWarps = 8;                 // warps per bundle
constexpr int kRsPanelHeaderWords = 20;
constexpr int kRsChunkWords = 128;          // 8 steps x 8 lane groups x 2 words
constexpr uint32_t kRsNone = 0xffffffffu;
constexpr uint32_t kRsEndLevel = 0x7fffffffu;

struct RowSweepProgram {
    bool ok = false;
    i32 warps = kRsWarps, levels = 0, panels = 0;
    i64 terms = 0, chunks = 0;
    std::vector<uint32_t> words;
    i64 stream_off[kRsWarps] = {};          // first word of every warp's stream
};

// lower: forward sweep over L (unit diagonal, no division); otherwise backward sweep over U
bool compile_row_sweep(const Factor &F, const Schedule &S, bool lower, RowSweepProgram &P, const char **why);

}  // namespace csp3
