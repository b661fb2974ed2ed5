// rowlane_program.hpp -- format of the "row-lane" refactor program (compiler: rowlane_program.cpp, kernel: lu_rowlane.cu).
//
// One warp owns a bundle of S = 8 systems.  Lane = (g = lane / 4, h = lane % 4): every lane carries the two adjacent
// systems 2h, 2h + 1 (16-byte accesses) and the eight lane groups g take the eight OPERATIONS of a record.  Values live
// in the bundle-interleaved factor arrays [entry][S] of the workspace path (same layout as lu_wide.cu, whose sweeps
// read what this kernel writes).  Only the accumulator of the column being eliminated lives in shared memory (`nslots`
// entries of 64 bytes); L operands and A values go from global memory (L1 / L2 / HBM) straight to registers.
//
// Program = one STREAM per warp of the bundle (1, 2, 4 or 8 warps eliminate different columns at the same time, each
// in its own accumulator; a column's sources may belong to other warps: the first quad of a stage carries, per other
// warp, how many columns that warp must have finished before the stage's operands may be read).  A stream = STAGES of
// `stage_quads` QUADS of 4 RECORDS of 8 lane words.  A quad is homogeneous (one kind), so the kernel
// dispatches once per quad and runs straight-line, predicated code for its records.  All global loads of a warp share
// one scoreboard slot, so the operands of a WHOLE stage (12 records) are requested together, one stage ahead; the
// program itself is copied with cp.async into a shared-memory ring, three stages ahead.
//
//   UPDATE   record r: acc[slot] -= L[base[r] / 64 + off] * m; NEW[r]: m = acc[mslot[r]] first (= U(j,k), final by
//            then); one source column j per record, records without NEW continue the previous record's column
//   UPDLATE  the same, but a source column is finalised less than two stages before the quad: operands are read
//            when the quad executes (the compiler orders the columns so that this is rare)
//   FIN      column boundary, roles by position: record 0 = L(:,k) entries: L[base[0] / 64 + off] = acc[slot] / pivot;
//            record 1 = U(:,k) entries: U[base[1] / 64 + off] = acc[slot]; both clear the slot; record 2 = A values
//            of the NEXT column of the elimination order: acc[slot] = Ax[base[2] / 8 + off]
//   STOREL4 / STOREU4 / LOAD4   `count` records of one role, for columns with more than 8 entries of it
//   P flag on the first store quad of a column: pivot = acc[mslot[0]], status code h0.w = k + 1 when it is zero or
//   not finite, shared reciprocal for the divisions of the column.
//   END      end of the program (padded with END quads to whole stages plus four stages).
//
// Per accumulator entry the operations are those of cs_lu in the order of cs_lu (oracle/csp3_oracle.c
// orc_csc_lu_refactor): sources in the stored order of U(:,k); columns may be eliminated in any order in which
// sources come first.  Multiply and subtract are not fused, the division is IEEE: factors are bit-identical.
//
// Quad = 76 words:
//   h0.x    bits 0-2 kind; bits 8-11 UPDATE: NEW[r], FIN: bit 8 has L, bit 9 has U, bit 10 has A; bit 12 P;
//           bit 13 (first quad of a stage) the stage reads columns of other warps (h2);
//           bits 16-18 count (STOREL4 / STOREU4 / LOAD4); bits 24-31 FIN / STOREL4 / STOREU4 / LOAD4: role of record r
//           in bits 24 + 2 r (0 none, 1 L store, 2 U store, 3 A load), what the kernel's shared loop executes
//   h0.y    mslot[0] * 64 | mslot[1] * 64 << 16        h0.z   mslot[2] * 64 | mslot[3] * 64 << 16
//   h0.w    P: k + 1
//   h1      base[0..3]: byte offsets (L / U array of the bundle, or a system's Ax)
//   h2      first quad of a stage: 8 x 16 bits, columns warp v of the bundle must have finished (0: no requirement)
//   lane words [g][record] (a lane group reads its four words with one 16-byte access): bit 31 valid, bits 16-30 off
//           (entries relative to base), bits 6-15 slot * 64 (an empty lane group repeats its partner's slot)
//   address words [g][record]: byte offset of the operand the record reads from global memory (base + off, scaled:
//           L array of the bundle for UPDATE / UPDLATE, a system's Ax for LOAD4 / FIN record 2), 0xffffffff: none
#pragma once
#include <cstdint>
#include <vector>

#include "symbolic.hpp"

namespace csp3 {

enum : int { kRlNop = 0, kRlLoad4 = 1, kRlUpdate = 2, kRlStoreU4 = 3, kRlStoreL4 = 4, kRlUpdLate = 5, kRlFin = 6, kRlEnd = 7 };
enum : unsigned { kRlRoleL = 1, kRlRoleU = 2, kRlRoleA = 3 };
enum : unsigned { kRlFlagCross = 1u << 13 };      // first quad of a stage: the stage has cross-warp requirements (h2 != 0)
enum : unsigned { kRlFlagP = 1u << 12, kRlHasL = 1u << 8, kRlHasU = 1u << 9, kRlHasA = 1u << 10 };
constexpr int kRlOps = 8;              // operations per record (lane groups)
constexpr int kRlMaxSlots = 1024;      // accumulator slots (slot * 64 is a 16-bit byte offset)
constexpr int kRlMaxOff = 32767;       // relative entry offset of a lane word
constexpr int kRlQuadRecords = 4;
constexpr int kRlQuadWords = 12 + 2 * kRlQuadRecords * kRlOps;  // 76
constexpr int kRlRingStages = 4;
constexpr int kRlMaxWarps = 8;         // warps of a bundle (each eliminates its own columns in its own accumulator)

struct RowlaneProgram {
    bool ok = false;
    i32 warps = 1;                     // warps per bundle = streams of the program
    i32 stage_quads = 3;               // quads per stage (operands of a stage are requested together, one stage ahead)
    i32 nslots = 0;                    // accumulator entries (per warp)
    i32 quads = 0;                     // all streams, without the END padding
    i32 stream_quads[kRlMaxWarps] = {};   // quads of every stream (without padding)
    i64 stream_off[kRlMaxWarps] = {};     // first quad of every stream in `words`
    i64 cross_records = 0;             // update records whose source column belongs to another warp
    i64 pad_quads = 0;                 // empty quads (a quad with cross-warp sources never shares a stage with an earlier FIN)
    size_t smem_bytes = 0;
    std::vector<uint32_t> words;       // 76 words per quad (quads + padding)
    i64 ops = 0, update_records = 0, update_quads = 0, late_quads = 0, conflict_pairs = 0;
    std::vector<i32> order;            // elimination order of the columns
};

bool compile_rowlane_refactor(i64 n, const i32 *Ap, const std::vector<i32> &q, const Factor &F, const Schedule &S, i32 warps,
                              i32 stage_quads, RowlaneProgram &P, const char **why);

}  // namespace csp3
