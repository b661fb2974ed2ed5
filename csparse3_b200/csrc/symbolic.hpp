// symbolic.hpp -- host-side symbolic phase of the B200 LU path (product code).
//
// Runs once per sparsity pattern and is cached by the caller (north-star): fill-reducing ordering
// (CSparse cs_amd semantics), first factorisation with threshold partial pivoting (CSparse cs_lu
// semantics: fixes pinv and the L/U patterns), level sets, and the device schedule the CUDA kernels in
// lu_kernels.cu execute.  The reference (SanPen/CSparse3) has no counterpart for any of this
// (SURVEY.md section 0.1); integer outputs are contract-equal to oracle/csp3_oracle.c.
#pragma once
#include <cstdint>
#include <vector>

namespace csp3 {

using i32 = int32_t;
using i64 = int64_t;

// ---- pattern utilities (CSparse cs_transpose / cs_add / cs_multiply on patterns) ------------------------
struct Pattern {
    i64 m = 0, n = 0;
    std::vector<i32> p, i;
};
Pattern transpose_pattern(i64 m, i64 n, const i32 *Ap, const i32 *Ai);

// ---- orderings / trees ---------------------------------------------------------------------------------
std::vector<i32> amd_order(i64 order, i64 m, i64 n, const i32 *Ap, const i32 *Ai);
std::vector<i32> etree(i64 m, i64 n, const i32 *Ap, const i32 *Ai, bool ata);
std::vector<i32> postorder(i64 n, const i32 *parent);

// ---- first factorisation -----------------------------------------------------------------------------
struct Factor {
    std::vector<i32> Lp, Li, Up, Ui, pinv;
    std::vector<double> Lx, Ux;
};
// returns 0 or k+1 when step k has no non-zero pivot
int lu_factor(i64 n, const i32 *Ap, const i32 *Ai, const double *Ax, const i32 *q, double tol, Factor &F);

struct LevelSet {
    std::vector<i32> level, order, lptr;   // lptr.size() == nlev + 1
    i64 nlev() const { return (i64)lptr.size() - 1; }
};
// kind 0: refactor (columns of U), 1: L solve (rows), 2: U solve (rows)
LevelSet build_levels(i64 n, const std::vector<i32> &Gp, const std::vector<i32> &Gi, int kind);

// ---- device schedule -------------------------------------------------------------------------------------
// Column k of the permuted matrix is computed in a per-warp accumulator of `len = ucnt + lcnt - 1` slots:
//   slot t < ucnt            -> U(:,k) entry t in stored order (diagonal is slot ucnt-1)
//   slot ucnt + t            -> L(:,k) entry t+1 in stored order (the unit diagonal has no slot)
struct ColDesc {            // 8 x int32, one per column, indexed by column number
    i32 up, lp;             // Up[k], Lp[k]
    i32 ucnt, lcnt;         // entries in U(:,k) (incl. diag), L(:,k) (incl. unit diag)
    i32 a_ptr, a_cnt;       // range in a_src / a_off
    i32 pair_ptr, pair_cnt; // range in pair arrays
};
struct PairDesc {           // 4 x int32, one per off-diagonal U entry, grouped by column, ascending pivot
    i32 moff;               // accumulator slot of the multiplier U(j,k)
    i32 lstart;             // Lp[j] + 1
    i32 llen;               // Lp[j+1] - Lp[j] - 1
    i32 mapstart;           // first entry of upd_map for this pair
};

// One triangular sweep in column order.  Column j owns the value range [start, start+len) of its factor array
// (strictly lower part of L(:,j), or strictly upper part of U(:,j)); every entry updates the live row kept in
// shared-memory slot `slot[p]`.  A row is given a slot the first time it is touched (alloc list of that
// column: the slot is initialised with the row's right-hand side) and releases it when its own column is
// reached.  slot_of_col[j] < 0: the row was never touched, its value is still the right-hand side.
struct SolveCol {                  // 4 x int32
    i32 slot;                      // slot holding y[j] when column j is reached, or -1
    i32 start;                     // first value of the column's off-diagonal range
    i32 len_alloc;                 // len | (number of allocs << 16)
    i32 alloc_ptr;                 // first alloc of this column
};
struct SolveStream {
    std::vector<SolveCol> cols;            // indexed by column
    std::vector<uint16_t> slot;            // per factor entry (same indexing as Lx / Ux)
    std::vector<i32> alloc_row;            // row (pivot numbering) whose right-hand side initialises the slot
    std::vector<uint16_t> alloc_slot;
    i32 nslots = 0;
};

// A compiled program: one byte stream of variable-length records, consumed strictly sequentially by every
// warp through a shared-memory ring.  `stage` (a power of two) is >= the largest single record, `bytes` is a
// multiple of `stage`.
struct Program {
    std::vector<uint8_t> bytes;
    i32 stage = 0;
};

struct Schedule {
    // refactor
    std::vector<ColDesc> cols;
    std::vector<i32> a_src;            // index into Ax (original CSC order)
    std::vector<uint16_t> a_off;       // accumulator slot
    std::vector<PairDesc> pairs;
    std::vector<uint16_t> upd_map;     // accumulator slot of row Li[lstart+t] in column k
    i32 max_col_len = 0;
    LevelSet lev_refactor;
    // solves: row views (CSR of the strict parts), entries reference positions in Lx / Ux
    std::vector<i32> lrow_ptr, lrow_col, lrow_pos;
    std::vector<i32> urow_ptr, urow_col, urow_pos;
    LevelSet lev_lsolve, lev_usolve;
    i64 flops = 0;
    // streaming (column-oriented) solves: cs_lsolve / cs_usolve order, live rows of y in host-assigned slots
    SolveStream ls, us;
    std::vector<i32> prow;             // prow[i] = original row r with pinv[r] == i  (y = P b: y[i] = b[prow[i]])
    // compiled programs: what the kernels actually execute (see program.hpp for the record formats)
    Program rf_prog, ls_prog, us_prog, ur_prog;
    i32 ur_nslots = 0, ur_max_len = 0;     // row-oriented backward sweep: live x values, longest row
};

// Wide (lane = system) refactor program, see program.hpp and wide_program.cpp.
struct WideProgram {
    bool ok = false;
    i32 width = 0;                         // systems per bundle
    i32 groups = 0;                        // lane groups of the kernel (entries of a column processed at once)
    i32 acc_slots = 0, ring_entries = 0, stage_entries = 0;
    size_t smem_bytes = 0;
    Program prog;
    i32 records = 0, immediate_fetches = 0, ngroups = 0;
    i64 near_fma = 0, far_fma = 0;         // update operations whose source column is cached / fetched
    i64 chunks = 0, chunk_ops = 0;         // chunk records and the update operations they hold
    i64 bank_clashes = 0;                  // shared-memory wavefront collisions left after the slot assignment (model)
};
struct Factor;
struct Schedule;
bool compile_wide_refactor(const Schedule &S, const Factor &F, i32 width, i32 groups, size_t smem_budget,
                           i32 ring_override, i32 stage_override, WideProgram &W, const char **why);

// Wide triangular sweep program (program.hpp, wide_solve.cpp)
struct WideSweep {
    bool ok = false;
    i32 width = 0, groups = 0, nslots = 0, landing_entries = 0, records = 0;
    size_t smem_bytes = 0;
    i64 ops = 0;
    Program prog;
};
bool compile_wide_sweep(const Factor &F, bool lower, i32 width, i32 groups, size_t smem_budget, WideSweep &W, const char **why);

// Builds the schedule; returns false (with message) when a limit is exceeded (column longer than 65535
// entries or more than 2^31-1 update slots).
bool build_schedule(i64 n, const i32 *Ap, const i32 *Ai, const std::vector<i32> &q, const Factor &F,
                    Schedule &S, const char **why);

}  // namespace csp3
