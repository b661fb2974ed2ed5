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

// ---------------------------------------------------------------------------------------------------------
// "Wide" programs (lu_wide.cu): lane = system.  One warp owns a bundle of S systems (S = 16: lanes are
// 2 entry groups x 16 systems; S = 32: one entry group), every quantity below is uniform over the systems of
// a bundle.  Values live in bundle-interleaved arrays [entry][S], so one entry of all systems of a bundle is
// one contiguous 8*S-byte run.
//
// Shared memory of a refactor warp:  acc[acc_slots][S] | lsrc[ring_entries + stage_entries][S] | program ring
//   lsrc[0, ring_entries)        compile-time managed cache of recently finalised L columns (sequential
//                                allocation, a column is never split by the wrap)
//   lsrc[ring_entries, +stage)   landing area of far (not cached) source columns, fetched from the bundle's L
//                                array with cp.async kWideLookahead (or up to 3 more) records before the pair that
//                                consumes them
// Every record issues at most one fetch and commits exactly one cp.async group, so "wait_group kWideLookahead"
// at record r guarantees everything requested at records <= r - kWideLookahead has landed (this also covers the
// program ring itself, whose stage loads ride in the groups of the records that trigger them).
//
// Records are 16-byte aligned and never straddle the end of the program ring; the reader keeps a running pointer
// into the ring and never masks: flag bits tell it when the next record starts at the ring base (wrap) and how
// many program stages the record enters (each entry requests one more stage).  All shared-memory positions are
// BYTE offsets from the start of the value area (acc first, lsrc right behind it), all global positions of the
// fetches are 16-byte units, so decoding is adds only.  EB = 8*S is the size of one bundle entry.
//
// Group record (48-byte header).  A group is a set of up to kWideGroupCols mutually independent columns (none is
// a source of another, all their sources were finalised by earlier groups) that are eliminated together: their
// accumulators occupy consecutive slot ranges, their updates share chunk records, one pass finalises them.  The
// elimination order inside the program is a list schedule of the columns (any order in which sources come first
// gives the same bits: a column's own arithmetic is untouched); long columns form groups of one.  The first record
// of a program is a preamble (ncols == 0) that only loads the A values of the first group.
//   +0   8 bytes reserved
//   +8   u16 ncols        +10 u16 nslots         accumulator slots of the group (cleared by the finalisation)
//   +12  u16 a_cnt        +14 u16 chunk_cnt
//   +16  u16 fin_cnt      finalisation records that follow the chunk records
//   +18  u16 an_cnt       A entries of the NEXT group that are loaded into registers while this group is
//                         eliminated: min(a_cnt of the next group, kWideARegs * E), E = lane groups of the kernel
//   +20  u16 fetch_dst16  +22 u16 fetch_units (16-byte units, 0: none)      +24 i32 fetch_src16
//   +28  u16 npf          A runs to pull into L2 (those of the group kWidePfGroups ahead)
//   +30  u16 reserved     +32 u16 reserved      +34 u8 flags (bits 1-2: stages entered, bit 3: wrap)
//   +48  kWideGroupCols x { i32 col1 (column + 1, 0: none), u16 pivot_off, u16 0 }
//   +..  kWideGroupCols x { i32 pf_src (first A entry of a run, -1: none), i32 pf_cnt }
//   +..  a_cnt x u16 accumulator byte offset; pad to 4; an_cnt x i32 index into the system's Ax (next group);
//        then the indices of this group's own entries beyond the register-held ones; pad to 16
// Finalisation record (16-byte header + C entries of 8 bytes, C = 2 * lane groups; same size as a chunk record):
//   +0   i32 fetch_src16  +4 u16 fetch_dst16     +6 u16 fetch_units (0: none)      (a look-ahead fetch, as in chunks)
//   +8   u16 flags (bits 1-2 stages entered, bit 3 wrap, bit 4: the record holds L entries, otherwise U entries;
//        a record never mixes the two)
//   +16  C x { i32 gout, u16 slot_off, u16 cache_off }
//        gout: bit 31 = L entry (value = slot / pivot of its column), bits 28-30 = index of the column inside the
//        group, bits 0-27 = position in the bundle's L or U array; slot_off == 0xffff: no entry;
//        cache_off != 0xffff: the L entry is also kept in the L cache (byte offset in the value area)
// Chunk record (16-byte header + C entries of 8 bytes, C = 2 * lane groups): up to C update operations
//   acc[tgt] -= lsrc[src] * acc[mult]
// that are mutually independent (no two share a target, no multiplier is a target of the chunk).  Across chunks the
// operations on one accumulator slot keep the order of the column's update sequence (pair after pair in the stored
// topological order of U(:,k)) and a pair starts only after the last operation that targets its multiplier, so
// executing "all loads, then all stores" per chunk reproduces the sequential result bit for bit; inside these two
// rules the compiler's list scheduler is free (wide_program.cpp).
// Lane group e of the kernel executes entries e and e + C/2 with ONE multiplier load: when both are valid they have
// the same mult_off (they belong to the same source column) and entry e is the valid one when only one is.  An
// unused entry repeats the offsets of its wavefront partner (entry e ^ 1) with valid == 0: the kernel loads it
// anyway, as a broadcast.
// A fetch covers the strict L parts of one or several ADJACENT columns of the bundle's L array (the unit-diagonal
// positions between them are read and never used).
//   +0   i32 fetch_src16  +4 u16 fetch_dst16     +6 u16 fetch_units (0: none)
//   +8   u16 flags: bit 0 = the fetch of THIS record must land before it is used (immediate),
//                   bits 1-2 = stages entered, bit 3 = wrap                   +10..15 reserved
//   +16  C x { u16 src_off, u16 mult_off, u16 tgt_off, u16 valid }   (byte offsets in the value area)

namespace csp3 {
constexpr int kWideLookahead = 4;     // records between a fetch and its use == pending cp.async groups allowed
constexpr int kWideARegs = 6;         // A values (per system) a lane holds in registers one group ahead
constexpr int kWidePfGroups = 4;      // L2 prefetch distance of the A values, in groups
constexpr int kWideGroupCols = 8;     // columns per group (one lane group each in the reciprocal pass)
constexpr int kWideGroupA = 48;       // A entries per group of more than one column
constexpr int kWideColHeader = 48;     // group record header
constexpr int kWideChunkHeader = 16;
constexpr int kWideProgStages = 8;      // ring slots of the program stream (see WideStream in lu_wide.cu)
}  // namespace csp3

// ---------------------------------------------------------------------------------------------------------
// Wide triangular sweeps (lu_wide.cu, wide_solve.cpp).  Same lane mapping and value layout as the wide refactor
// kernel; one program per sweep (forward: cs_lsolve column order, backward: cs_usolve column order), executed by
// one warp per bundle.  The live part of the solution vector sits in host-assigned shared-memory slots.
//
// Fixed-size records, E = lane groups of the kernel:
//   +0    16-byte header: u16 flags (bits 1-2 stages entered, bit 3 wrap), rest reserved
//   +16   E/2 x { i32 gidx, u16 slot_off, u16 0 }            right-hand-side loads: slot <- zin[gidx]   (gidx < 0: none)
//   +..   2E  x i32                                          factor entries the update operations of the record
//                                                            kSweepLookahead records AHEAD will read (< 0: none)
//   +..   E/2 x i32                                          divisors (factor entries) of that record's finalisations
//   +..   E/2 x { i32 out_idx, u16 slot_off, u16 divide }    finalisations: [slot /= divisor;] zout[out_idx] <- slot
//                                                            (out_idx < 0: none)
//   +..   2E  x { u16 mult_off, u16 tgt_off }                updates: slot[tgt] -= value * slot[mult]
//                                                            (tgt_off == 0xffff: none)
// (224 bytes for E = 8: two records per 512-byte program stage).  slot_off / mult_off / tgt_off are in 16-byte units.
// Factor values are gathered with cp.async into a landing area of kSweepLookahead + 1 record-sized sets (2E update
// values, then E/2 divisors in the backward sweep), record r uses set r mod (kSweepLookahead + 1).  Right-hand sides
// are copied straight into their slot at least kSweepLookahead records before the first operation that touches it.
namespace csp3 {
constexpr int kSweepLookahead = 3;      // records between a gather / load and its use in the sweep kernels
constexpr int kSweepProgStages = 6;     // ring slots of the sweep program stream
constexpr int kWideSolveHeader = 16;
inline constexpr int wide_solve_record_bytes(int groups) { return kWideSolveHeader + groups * (4 + 8 + 2 + 4 + 8); }
}  // namespace csp3
