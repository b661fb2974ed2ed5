// program.hpp -- byte-level formats of the compiled elimination / substitution programs.
//
// The host compiles the static schedule of a pattern into three byte streams (refactor, forward sweep,
// backward sweep).  All bundles execute the same stream, so it stays L2-resident; every warp pulls it through
// a small shared-memory ring with cp.async and reads the records with plain LDS off a running offset.
// Every record and every section inside a record is 8-byte aligned.
//
// Refactor program.  Column record (one per column k, in elimination order):
//   +0   i32  up        Up[k]
//   +4   i32  lp        Lp[k]
//   +8   u16  ucnt      entries of U(:,k) incl. diagonal      +10  u16 lcnt   entries of L(:,k) incl. unit diagonal
//   +12  u16  a_cnt     entries of A(:,q[k]) to scatter        +14  u16 pair_cnt
//   +16  u16  pf_cnt    A entries of column k + kPfCols to prefetch into L1
//   +18  u16  mpf_cnt   far-back source columns of column k + kPfMissCols to prefetch into L1     +20..23 reserved
//   +24  a_cnt x i32 src (index into the system's Ax), then a_cnt x u16 accumulator slot; pad to 8
//        pf_cnt x i32 src; pad to 8
//        mpf_cnt x { i32 lstart, u16 llen, u16 0 }
// followed by pair_cnt pair records:
//   +0 i32 lstart (Lp[j]+1)   +4 u16 moff (slot of U(j,k))   +6 u16 llen
//   +8 llen x u16 accumulator slot of row Li[lstart+t]; pad to 8
//
// Sweep programs (forward: columns ascending, backward: columns descending), one record per column j:
//   +0   i32  start     first off-diagonal value of the column in Lx (Lp[j]+1) / Ux (Up[j])
//   +4   i32  rhs       forward: original row of b that initialises y[j] (prow[j]); backward: j
//   +8   i16  slot      slot holding y[j] when the column is reached, -1: never touched (use rhs)
//   +10  u16  len       off-diagonal entries        +12 u16 nalloc      +14 u16 pf_cnt
//   +16  i32  out       forward: j; backward: q[j] (where x goes)       +20 reserved
//   +24  len x u16 slot of the updated row; pad to 8
//        nalloc x i32 rhs index (as `rhs`), then nalloc x u16 slot; pad to 8
//        pf_cnt x i32 rhs index of the allocations of the column kPfCols steps ahead (L1 prefetch); pad to 8
//
// Backward sweep, row-oriented (what the solve kernel executes for U): one record per row i, rows descending.
// x_i = (y_i - sum_{j>i} U(i,j) x_j) / U(i,i) with the subtractions in DESCENDING j, which is exactly the order in
// which cs_usolve's column sweep updates row i.  Only the x_j still needed by a later row are kept on chip.
//   +0   i32  diagpos   position of U(i,i) in Ux
//   +4   i32  xpos      q[i]: where y_i was parked by the forward sweep and where x_i goes
//   +8   i16  slot_out  slot that receives x_i (-1: no later row needs it)      +10 u16 len
//   +12  u16  pf_cnt    +14 reserved
//   +16  i32  pf_x      xpos of the row kPfCols steps ahead (-1: none)           +20 reserved
//   +24  len x { i32 pos (in Ux), u16 slot of x_j, u16 0 }   entries in descending j
//        pf_cnt x i32   Ux positions (entries and diagonal) of the row kPfCols steps ahead; pad to 8
#pragma once
#include <cstdint>

namespace csp3 {

constexpr int kPfCols = 8;        // prefetch distance, in columns, of the L1 prefetch directives
constexpr int kPfMissCols = 2;    // same for far-back L columns (sources older than kCompileWindow entries)
constexpr int kCompileWindow = 256;   // recent-L ring size (entries) assumed when the directives are compiled
constexpr int kRfHeader = 24;
constexpr int kSvHeader = 24;

}  // namespace csp3
