/*
 * csparse3_b200.h -- C ABI of libcsparse3_b200.so, the B200 (sm_100a) drop-in for the numeric hot
 * path of SanPen/CSparse3.
 *
 * Every entry point replaces one flat kernel of the reference's backend seam
 *     src/CSparse3/csc.py:33-41   (sptools.* and `from CSparse3.csc_numba import *`)
 * and is cited against it below.  LU / triangular solve / ordering entry points have no reference
 * counterpart (the reference ships none); they follow the flat signatures SURVEY.md
 * section 8(a11) proposes in the reference's idiom and the semantics of upstream CSparse (cs_amd, cs_lu,
 * cs_lsolve, cs_usolve, cs_ipvec).
 *
 * Conventions
 *   - indices int32_t, values double, sizes int64_t (reference: i4[:] / f8[:] / i8, csc_numba.py:36-744).
 *   - `_host` suffix: all array arguments are HOST pointers; the call copies in, runs the CUDA kernels and
 *     copies out before returning (this is what the numpy-facing drop-in module binds).
 *   - no suffix: all array arguments are DEVICE pointers, the call is stream-ordered on `stream`
 *     (a cudaStream_t passed as void*; NULL = legacy default stream), never synchronises and never retains
 *     caller pointers.  The caller owns every buffer.
 *   - return value: 0 ok; negative = argument / CUDA error (text in csp3_last_error_string());
 *     positive k = zero or non-finite pivot at column k-1 (KLU style) where documented.
 *   - there is NO CPU fallback: without a CUDA device every compute call returns CSP3_ERR_CUDA.
 */
#ifndef CSPARSE3_B200_H
#define CSPARSE3_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CSP3_OK 0
#define CSP3_ERR_ARG (-1)
#define CSP3_ERR_CUDA (-2)
#define CSP3_ERR_ALLOC (-3)
#define CSP3_ERR_OVERFLOW (-4) /* nnz of a product exceeds int32 (sparsetools csr.h:591-596) */
#define CSP3_ERR_SINGULAR (-5)

/* ---- library ------------------------------------------------------------------------------------- */
int csp3_version(void);
const char *csp3_last_error_string(void);
int csp3_device_count(void);             /* 0 when no CUDA device is usable */
int csp3_set_device(int device);

/* ---- SpMV / SpMM ----------------------------------------------------------------------------------- */
/* y = A*x.  Replaces csc_mat_vec_ff(m,n,Ap,Ai,Ax,x)->y, src/CSparse3/csc_numba.py:309-328. */
int csp3_csc_mat_vec_ff_host(int64_t m, int64_t n, const int32_t *Ap, const int32_t *Ai, const double *Ax,
                             const double *x, double *y);
/* Y += A*X.  Replaces sptools.csc_matvec, call site src/CSparse3/csc.py:374-379, source mirror
 * src/sparsetools/csc.h:27-45. */
int csp3_csc_matvec_host(int64_t n_row, int64_t n_col, const int32_t *Ap, const int32_t *Ai, const double *Ax,
                         const double *Xx, double *Yx);
/* Y[n_row,n_vecs] += A*X[n_col,n_vecs], row-major.  Replaces sptools.csc_matvecs, csc.py:409-415,
 * src/sparsetools/csc.h:68-84. */
int csp3_csc_matvecs_host(int64_t n_row, int64_t n_col, int64_t n_vecs, const int32_t *Ap, const int32_t *Ai,
                          const double *Ax, const double *Xx, double *Yx);

/* Device SpMV plan: CSC -> row-split CSR view built once per pattern (counting sort on device, as
 * csc_to_csr csc_numba.py:360-397), then y = beta*y + A*x for `batch` value sets sharing the pattern.
 *   Ax[batch, nnz] (stride_ax doubles between systems; 0 = one shared value set), x[batch, n], y[batch, m]. */
typedef struct csp3_spmv_plan csp3_spmv_plan;
int csp3_spmv_plan_create(int64_t m, int64_t n, const int32_t *Ap_dev, const int32_t *Ai_dev, void *stream,
                          csp3_spmv_plan **plan);
int csp3_spmv_plan_destroy(csp3_spmv_plan *plan);
int csp3_spmv_batched(const csp3_spmv_plan *plan, int64_t batch, const double *Ax, int64_t stride_ax,
                      const double *x, double *y, double beta, void *stream);

/* ---- format conversion ------------------------------------------------------------------------------ */
/* Replaces csc_transpose(m,n,Ap,Ai,Ax)->(n,m,Cp,Ci,Cx), csc_numba.py:400-436.  Cp[m+1], Ci/Cx[nnz]. */
int csp3_csc_transpose_host(int64_t m, int64_t n, const int32_t *Ap, const int32_t *Ai, const double *Ax,
                            int32_t *Cp, int32_t *Ci, double *Cx);
/* Replaces csc_to_csr(m,n,Ap,Ai,Ax,Bp,Bi,Bx), csc_numba.py:360-397 (Bp need not be pre-zeroed). */
int csp3_csc_to_csr_host(int64_t m, int64_t n, const int32_t *Ap, const int32_t *Ai, const double *Ax,
                         int32_t *Bp, int32_t *Bi, double *Bx);
/* Device form of both (the two are the same counting sort). */
int csp3_csc_transpose(int64_t m, int64_t n, const int32_t *Ap, const int32_t *Ai, const double *Ax,
                       int32_t *Cp, int32_t *Ci, double *Cx, void *stream);

/* ---- SpGEMM ------------------------------------------------------------------------------------------- */
/* C = A*B, CSC, hash accumulation, two phases mirroring scipy's pass1/pass2 contract.
 * symbolic: fills Cp[Bn+1], returns nnz(C) in *nnz (structural: explicit zeros kept, like csc_multiply_ff).
 * numeric: fills Ci (row indices SORTED inside each column) and Cx.
 * Replaces csc_multiply_ff, csc_numba.py:222-306 (caller CscMat.dot csc.py:483-500) and
 * sptools.csc_matmat_pass1/pass2, csc.py:356-370, src/sparsetools/csc.h:115-137. */
int csp3_spgemm_symbolic_host(int64_t Am, int64_t An, const int32_t *Ap, const int32_t *Ai, int64_t Bm,
                              int64_t Bn, const int32_t *Bp, const int32_t *Bi, int32_t *Cp, int64_t *nnz);
int csp3_spgemm_numeric_host(int64_t Am, int64_t An, const int32_t *Ap, const int32_t *Ai, const double *Ax,
                             int64_t Bm, int64_t Bn, const int32_t *Bp, const int32_t *Bi, const double *Bx,
                             const int32_t *Cp, int32_t *Ci, double *Cx);
int csp3_spgemm_symbolic(int64_t Am, int64_t An, const int32_t *Ap, const int32_t *Ai, int64_t Bm, int64_t Bn,
                         const int32_t *Bp, const int32_t *Bi, int32_t *Cp, int64_t *nnz_host, void *stream);
int csp3_spgemm_numeric(int64_t Am, int64_t An, const int32_t *Ap, const int32_t *Ai, const double *Ax,
                        int64_t Bm, int64_t Bn, const int32_t *Bp, const int32_t *Bi, const double *Bx,
                        const int32_t *Cp, int32_t *Ci, double *Cx, void *stream);

/* ---- A + B, A - B ---------------------------------------------------------------------------------------- */
/* C = A + sign*B (sign = +1 or -1).  Replaces sptools.csc_plus_csc / csc_minus_csc, call sites
 * src/CSparse3/csc.py:312-315, :336-339; source mirror src/sparsetools/csc.h:203-219 -> csr.h:692-908
 * (duplicates summed per operand, exact zeros dropped).  Cp[n+1] is filled, Ci/Cx (capacity nnzA + nnzB, as the
 * reference allocates them) receive Cp[n] entries with row indices sorted inside each column. */
int csp3_csc_plusminus_host(int64_t m, int64_t n, const int32_t *Ap, const int32_t *Ai, const double *Ax,
                            const int32_t *Bp, const int32_t *Bi, const double *Bx, double sign, int32_t *Cp,
                            int32_t *Ci, double *Cx);

/* C = alpha*A + beta*B in the form of the reference's own kernel csc_add_ff, src/CSparse3/csc_numba.py:183-219
 * (csc_scatter_f :125-151): per column alpha*A(:,j) then beta*B(:,j) are scattered, rows come out in FIRST-TOUCH
 * order, explicit zeros are kept, duplicates are summed in traversal order.  Cp[n+1] is filled; Ci/Cx (capacity
 * nnzA + nnzB, the reference's allocation) receive Cp[n] entries. */
int csp3_csc_add_ff_host(int64_t m, int64_t n, const int32_t *Ap, const int32_t *Ai, const double *Ax,
                         const int32_t *Bp, const int32_t *Bi, const double *Bx, double alpha, double beta, int32_t *Cp,
                         int32_t *Ci, double *Cx);

/* ---- [[A, B], [C, D]]: Jacobian assembly ------------------------------------------------------------------- */
/* Replaces csc_stack_4_by_4_ff, src/CSparse3/csc_numba.py:640-720 (caller pack_4_by_4, src/CSparse3/csc.py:588-606).
 * Argument order as in the reference: (m, n, indices, indptr[, data]) per block.  Column-wise concatenation, the
 * upper block's entries before the lower block's inside a column, lower row indices shifted by the upper block's
 * row count; nothing is sorted or merged.  Shapes must satisfy am == bm, cm == dm, an == cn, bn == dn (the
 * reference's assertions, csc_numba.py:679-682) or CSP3_ERR_ARG is returned.
 *
 * A plan holds the pattern work (host arrays in, done once per pattern); csp3_stack4_batched is the numeric step
 * on the device for `batch` value sets sharing the four patterns:
 *     out[s * ldo + p] = X_block(p)[s * ld_block + pos(p)],   p < nnz = nnzA + nnzB + nnzC + nnzD
 * Ax..Dx, out are DEVICE pointers; ld* are the distances (in doubles) between the value sets of a block, 0 = one
 * value set shared by the whole batch (e.g. a constant block).  The call is stream-ordered. */
typedef struct csp3_stack4 csp3_stack4;
int csp3_stack4_create(int64_t am, int64_t an, const int32_t *Ai, const int32_t *Ap,
                       int64_t bm, int64_t bn, const int32_t *Bi, const int32_t *Bp,
                       int64_t cm, int64_t cn, const int32_t *Ci, const int32_t *Cp,
                       int64_t dm, int64_t dn, const int32_t *Di, const int32_t *Dp, csp3_stack4 **plan);
int csp3_stack4_destroy(csp3_stack4 *plan);
/* out[0..2] = m, n, nnz of the stacked matrix; out[4..7] = nnz of A, B, C, D */
int csp3_stack4_sizes(const csp3_stack4 *plan, int64_t out[8]);
/* indices[nnz], indptr[n+1] of the stacked matrix (host arrays) */
int csp3_stack4_get_pattern(const csp3_stack4 *plan, int32_t *indices, int32_t *indptr);
int csp3_stack4_batched(csp3_stack4 *plan, int64_t batch, const double *Ax, int64_t lda, const double *Bx, int64_t ldb,
                        const double *Cx, int64_t ldc, const double *Dx, int64_t ldd, double *out, int64_t ldo,
                        void *stream);
/* One matrix, host arrays in and out (the reference's call): indices[nnz], indptr[an+bn+1], data[nnz]. */
int csp3_csc_stack_4_by_4_host(int64_t am, int64_t an, const int32_t *Ai, const int32_t *Ap, const double *Ax,
                               int64_t bm, int64_t bn, const int32_t *Bi, const int32_t *Bp, const double *Bx,
                               int64_t cm, int64_t cn, const int32_t *Ci, const int32_t *Cp, const double *Cx,
                               int64_t dm, int64_t dn, const int32_t *Di, const int32_t *Dp, const double *Dx,
                               int32_t *indices, int32_t *indptr, double *data);

/* ---- host symbolic phase (runs once per pattern, cached by the caller) ---------------------------------- */
/* q = amd(order, A): CSparse cs_amd.  order 0 natural, 1 A+A', 2 S'S (dense rows dropped), 3 A'A.  q[n]. */
int csp3_csc_amd(int64_t order, int64_t m, int64_t n, const int32_t *Ap, const int32_t *Ai, int32_t *q);
/* parent = etree(A) (ata=0, upper part) or etree(A'A) (ata=1): CSparse cs_etree.  post = cs_post(parent). */
int csp3_csc_etree(int64_t m, int64_t n, const int32_t *Ap, const int32_t *Ai, int ata, int32_t *parent);
int csp3_csc_post(int64_t n, const int32_t *parent, int32_t *post);

/* Opaque symbolic object: ordering q, first factorisation with threshold partial pivoting (pinv and the
 * L/U patterns it fixes, CSparse cs_lu layout: L unit diagonal first, U diagonal last, reach order inside
 * columns), level sets and the device schedule derived from them.  Host arrays only until
 * csp3_lu_upload(); plain data, safe to destroy at any time after the last kernel using it completed. */
typedef struct csp3_lu_symbolic csp3_lu_symbolic;

/* Analyse + first factorisation on the host.  All pointers are HOST pointers.  q_in may be NULL (then
 * q = amd(order)).  Returns 0, or k+1 (>0) if no non-zero pivot exists in step k. */
int csp3_lu_analyze(int64_t order, int64_t n, const int32_t *Ap, const int32_t *Ai, const double *Ax,
                    const int32_t *q_in, double tol, csp3_lu_symbolic **sym);
/* Same object from a KNOWN ordering, pivot sequence and L/U patterns (e.g. returned earlier by
 * csp3_lu_get_pattern and cached by the caller): no factorisation is performed, values are not needed. */
int csp3_lu_analyze_fixed(int64_t n, const int32_t *Ap, const int32_t *Ai, const int32_t *q, const int32_t *pinv,
                          const int32_t *Lp, const int32_t *Li, const int32_t *Up, const int32_t *Ui,
                          csp3_lu_symbolic **sym);
int csp3_lu_destroy(csp3_lu_symbolic *sym);

/* sizes: out[0]=n out[1]=nnzA out[2]=lnz out[3]=unz out[4]=refactor levels out[5]=L-solve levels
 * out[6]=U-solve levels out[7]=refactor flops out[8]=bytes of device schedule */
int csp3_lu_sizes(const csp3_lu_symbolic *sym, int64_t out[16]);
/* Copy the integer results to caller HOST arrays (any may be NULL): q[n], pinv[n], Lp[n+1], Li[lnz],
 * Up[n+1], Ui[unz]; and the first factorisation's values Lx[lnz], Ux[unz]. */
int csp3_lu_get_pattern(const csp3_lu_symbolic *sym, int32_t *q, int32_t *pinv, int32_t *Lp, int32_t *Li,
                        int32_t *Up, int32_t *Ui, double *Lx, double *Ux);
/* Level sets.  kind 0 refactor (from U), 1 L-solve, 2 U-solve.  level[n], order[n], lptr[nlev+1]. */
int csp3_lu_get_levels(const csp3_lu_symbolic *sym, int kind, int32_t *level, int32_t *order, int32_t *lptr);
/* Fundamental supernodes of L (maximal runs of columns j, j+1, ... whose patterns nest: rows(L(:,j)) minus {j+1} equals
 * rows(L(:,j+1))): sn_ptr[0..count] are the first columns of the supernodes, sn_ptr[count] = n (room for n + 1 entries).
 * The dense trailing update of a wide supernode is what csp3_dense_update_batched computes on the FP64 tensor cores. */
int csp3_lu_supernodes(const csp3_lu_symbolic *sym, int32_t *sn_ptr, int64_t *count);
/* Introspection of the compiled device programs (csparse3_b200/csrc/program.hpp): which = 0 refactor, 1 forward
 * sweep, 2 backward sweep (row-oriented), 3 wide refactor, 4 / 5 wide forward / backward sweep.  Returns the size in bytes (copied into buf when
 * capacity allows), or a negative error.  geometry (optional, 8 values): [0] stage bytes; wide refactor only:
 * [1] bundle width, [2] accumulator slots, [3] L cache entries, [4] landing entries, [5] records, [6] smem bytes,
 * [7] lane groups; wide sweeps: [1] bundle width, [2] slots, [3] landing entries, [5] records, [6] smem bytes,
 * [7] lane groups. */
int64_t csp3_lu_get_program(const csp3_lu_symbolic *sym, int which, uint8_t *buf, int64_t capacity,
                            int64_t geometry[8]);
/* (which = 6: panel refactor, 7: row-lane refactor -- 76 words per quad, csparse3_b200/csrc/rowlane_program.hpp;
 * geometry [0] update quads, [1] quads per stage, [2] accumulator slots, [3] late quads, [4] update operations,
 * [5] quads, [6] smem bytes, [7] update records.  which = 9 / 10: forward / backward row-sweep program,
 * csparse3_b200/csrc/rowsweep_program.hpp; geometry [0] levels, [1] warps, [2] panels, [3] chunks, [4] terms.) */
/* Name of the refactorisation kernel the workspace path (csp3_lu_refactor_ws) runs for `batch` systems on the
 * current device (the symbolic object must have been uploaded): "lu_refactor_wide_kernel", "lu_refactor_rowlane_kernel",
 * "lu_refactor_panel_kernel", "lu_refactor_tmem_kernel" or "lu_refactor_kernel".  NULL on error. */
const char *csp3_lu_refactor_kernel_name(const csp3_lu_symbolic *sym, int64_t batch);

/* Replicate the schedule onto the current device (idempotent per device). */
int csp3_lu_upload(csp3_lu_symbolic *sym, void *stream);

/* ---- device numeric phase: batched refactor / solve on one pattern --------------------------------------- */
/* Refactor `batch` systems: Ax[batch, nnzA] (original CSC entry order) -> Lx[batch, lnz], Ux[batch, unz]
 * in the cs_lu layout of csp3_lu_get_pattern().  status[batch] (device int32): 0 ok, k+1 = zero or
 * non-finite pivot in column k (that system's factors are unusable; other systems are unaffected). */
int csp3_lu_refactor_batched(const csp3_lu_symbolic *sym, int64_t batch, const double *Ax, double *Lx,
                             double *Ux, int32_t *status, void *stream);
/* Solve with existing factors: x = Q (U \ (L \ (P b))).  b[batch, n], x[batch, n] (must not alias). */
int csp3_lu_solve_batched(const csp3_lu_symbolic *sym, int64_t batch, const double *Lx, const double *Ux,
                          const double *b, double *x, void *stream);
/* Fused refactor + solve.  With Lx and Ux non-NULL the factors are returned in the API layout; with both
 * NULL they stay in `work` (csp3_lu_workspace_bytes(sym, batch) bytes of device scratch) in the internal
 * bundle-interleaved layout, which is the fast path. */
int64_t csp3_lu_workspace_bytes(const csp3_lu_symbolic *sym, int64_t batch);
int csp3_lu_refactor_solve_batched(const csp3_lu_symbolic *sym, int64_t batch, const double *Ax,
                                   const double *b, double *x, double *Lx, double *Ux, int32_t *status,
                                   void *work, void *stream);

/* The two halves of the fused path, with the factors kept in `work` in the library's internal
 * bundle-interleaved layout (opaque to the caller; only csp3_lu_solve_ws with the same batch reads it).
 * This is the fast path: each factor entry of a bundle of systems is one contiguous run in HBM. */
/* status[s]: 0, or k + 1 when column k of system s has a zero / non-finite pivot (the oracle's code), or -2 when the
 * warps of a bundle lost their synchronisation (a defect, never seen: the kernel gives up instead of hanging).
 * The kernel is chosen from `batch` (csp3_lu_refactor_kernel_name); the program of a geometry that has not been used
 * on this device yet is compiled and uploaded inside the first call (synchronously: call csp3_lu_prepare first when
 * the call must stay asynchronous, e.g. under stream capture). */
int csp3_lu_refactor_ws(const csp3_lu_symbolic *sym, int64_t batch, const double *Ax, void *work,
                        int32_t *status, void *stream);
/* Makes everything csp3_lu_refactor_ws / csp3_lu_solve_ws need for batches of `batch` systems resident on the current
 * device (compiles the kernel program of that batch size on first use).  Idempotent. */
int csp3_lu_prepare(const csp3_lu_symbolic *sym, int64_t batch);
int csp3_lu_solve_ws(const csp3_lu_symbolic *sym, int64_t batch, void *work, const double *b, double *x,
                     void *stream);

/* Pivot-growth indicator of the factors csp3_lu_refactor_ws left in `work`: growth[batch] (device) = max |L(i,j)|,
 * i > j.  The pivot sequence is frozen at the first factorisation (there |L| <= 1/tol); values that drift can make a
 * frozen pivot tiny, which shows here long before `status` reports a zero pivot.  Callers
 * compare it with a limit and re-analyse the flagged systems (csp3_lu_analyze on that system's values). */
int csp3_lu_growth_ws(const csp3_lu_symbolic *sym, int64_t batch, const void *work, double *growth, void *stream);

/* Host-buffer batched refactor+solve: chunks the batch, overlapping H2D copies, kernels and D2H copies on
 * internal streams.  Ax[batch, nnzA], b[batch, n], x[batch, n], status[batch] are HOST pointers (pinned
 * memory gives full PCIe rate).  This is the end-to-end call the e2e benchmark times. */
int csp3_lu_refactor_solve_host(csp3_lu_symbolic *sym, int64_t batch, const double *Ax, const double *b,
                                double *x, int32_t *status);

/* Host-buffer refactor / solve (synchronous; copy in, run the kernels, copy out). */
int csp3_lu_refactor_host(csp3_lu_symbolic *sym, int64_t batch, const double *Ax, double *Lx, double *Ux,
                          int32_t *status);
int csp3_lu_solve_host(csp3_lu_symbolic *sym, int64_t batch, const double *Lx, const double *Ux,
                       const double *b, double *x);

/* One-shot: CSparse cs_lusol(order, A, b, tol) on the device path (analyse on host, refactor+solve on GPU).
 * b is overwritten with x.  Host pointers. */
int csp3_csc_lusol_host(int64_t order, int64_t n, const int32_t *Ap, const int32_t *Ai, const double *Ax,
                        double *b, double tol);

/* ---- FP64 tensor-core building blocks (groundwork for supernodal trailing updates, BASELINE.json configs[4]) -------- */
/* Measured DMMA (mma.sync.m8n8k4.f64) throughput of the current device in TFLOP/s: `iters` instructions per warp per
 * accumulator chain, best of 5 launches, CUDA events. */
int csp3_dmma_peak(int64_t iters, double *tflops);
/* C_s -= A_s * B_s for s < batch: column-major blocks A[m x k] (lda), B[k x n] (ldb), C[m x n] (ldc), system s at
 * offset s * stride (doubles).  This is the trailing update L21 * U12 of one supernode for every system of a
 * same-pattern batch; a warp owns a 32 x 32 tile of C and issues DMMA.8x8x4.  DEVICE pointers, stream-ordered. */
int csp3_dense_update_batched(int64_t batch, int64_t m, int64_t n, int64_t k, const double *A, int64_t lda, int64_t strideA,
                              const double *B, int64_t ldb, int64_t strideB, double *C, int64_t ldc, int64_t strideC, void *stream);

/* ---- topology operations on the device ------------------------------------------- */
/* Replaces find_islands(node_number, indptr, indices), src/CSparse3/csc_numba.py:743-808 (caller CscMat.islands,
 * csc.py:515-521): the islands of an undirected graph given as a symmetric adjacency pattern.  Output in the
 * reference's order: islands by ascending smallest node, nodes inside an island in the order of its traversal
 * (first appearance in its front-popped list = breadth-first from the smallest node, neighbours in adjacency order).
 * order[n] = node ids, island k = order[island_ptr[k] .. island_ptr[k+1]); island_ptr has room for n + 1 entries. */
int csp3_find_islands_host(int64_t n, const int32_t *indptr, const int32_t *indices, int32_t *order, int32_t *island_ptr,
                           int64_t *n_islands);
/* The N-1 form: `batch` cases share the adjacency pattern, case c has the edge between nodes out_from[c] and out_to[c]
 * removed (both NULL, or -1: nothing removed).  label[batch, n] = smallest node id of the node's island (equal labels
 * <=> same island; islands sorted by label are the reference's island order), islands[batch] = number of islands
 * (> 1 for an outage that splits the grid: a bridge).  One CTA per case.  DEVICE pointers, stream-ordered. */
int csp3_islands_batched(int64_t n, const int32_t *indptr, const int32_t *indices, int64_t batch, const int32_t *out_from,
                         const int32_t *out_to, int32_t *label, int32_t *islands, void *stream);
int csp3_islands_batched_host(int64_t n, const int32_t *indptr, const int32_t *indices, int64_t batch, const int32_t *out_from,
                              const int32_t *out_to, int32_t *label, int32_t *islands);
/* Replaces csc_sub_matrix / csc_sub_matrix_cols / csc_sub_matrix_rows, src/CSparse3/csc_numba.py:463-578 (callers
 * CscMat.__getitem__, csc.py:150-292).  rows == NULL: every row, original row indices (the _cols kernel); cols == NULL:
 * every column (the _rows kernel).  With `rows` the entries of a column come out ordered by the position of their row
 * in `rows` and are numbered by the reference's running counter (csc_numba.py:485-493).  Bp[ncols+1] (or An+1), and
 * Bi / Bx with room for nnz(A) entries; *nnz = entries written.  `rows` must not repeat an index. */
int csp3_csc_sub_matrix_host(int64_t Am, int64_t An, const int32_t *Ap, const int32_t *Ai, const double *Ax, int64_t nrows,
                             const int32_t *rows, int64_t ncols, const int32_t *cols, int32_t *Bp, int32_t *Bi, double *Bx,
                             int64_t *nnz);

/* ---- Newton-Raphson power-flow iteration on the device -------------------- */
/* The consumer loop the reference's pieces exist for: J = pack_4_by_4(H, N, M, L)
 * (src/CSparse3/csc.py:588-606 -> csc_stack_4_by_4_ff csc_numba.py:640-720), mismatch from Ybus * V
 * (CscMat.__mul__ on a vector, csc.py:374-379 -> csc_matvec), then factor, solve, update.  The reference runs each
 * piece on the host for one case; this plan keeps a batch of same-topology cases on the GPU for whole iterations:
 * evaluate I = Ybus V and the mismatch, write the four polar Jacobian blocks straight into the refactorisation's
 * value array (the block stacking is a precomputed entry map), refactor, solve, update (vm, va).
 *
 *   Ybus        CSR pattern y_rowptr[n_bus+1], y_col[nnz_y], complex values y_val[nnz_y] (re, im interleaved)
 *   pvpq, pq    bus indices of the angle unknowns / magnitude unknowns; n = npvpq + npq equations:
 *               P mismatch of pvpq buses, then Q mismatch of pq buses (the order of the Jacobian's rows and columns)
 *   Jacobian    entry p of the CSC value array (the pattern `sym` was analysed with) is block j_block[p]
 *               (0 dP/dVa, 1 dP/dVm, 2 dQ/dVa, 3 dQ/dVm) evaluated at Ybus entry j_ent[p]
 *   branches    optional N-1 support: case c with out_branch[c] = k >= 0 subtracts br_val[q][k] from Ybus entry
 *               br_slot[q][k], q = 0..3 (ff, ft, tf, tt contributions of branch k); arrays are [4][n_branch]
 * All pointers of csp3_nr_create are HOST pointers; the plan keeps device copies on the current device and uses `sym`
 * (which must outlive it). */
typedef struct csp3_nr_plan csp3_nr_plan;
int csp3_nr_create(int64_t n_bus, int64_t nnz_y, const int32_t *y_rowptr, const int32_t *y_col, const double *y_val,
                   int64_t npvpq, const int32_t *pvpq, int64_t npq, const int32_t *pq, int64_t jnnz, const int32_t *j_ent,
                   const int32_t *j_block, int64_t n_branch, const int32_t *br_slot, const double *br_val,
                   csp3_lu_symbolic *sym, csp3_nr_plan **plan);
int csp3_nr_destroy(csp3_nr_plan *plan);
int64_t csp3_nr_workspace_bytes(const csp3_nr_plan *plan, int64_t batch);
/* Jacobian values Ax[batch, jnnz] and right-hand side b[batch, n] = -(S_calc - S_spec) at the state (vm, va)
 * [batch, n_bus]; fnorm[batch] = max |mismatch|.  sspec[batch, n]: specified P of the pvpq buses, then Q of the pq
 * buses.  out_branch[batch] may be NULL.  DEVICE pointers, stream-ordered. */
int csp3_nr_jacobian(const csp3_nr_plan *plan, int64_t batch, const double *vm, const double *va, const int32_t *out_branch,
                     const double *sspec, double *Ax, double *b, double *fnorm, void *work, void *stream);
/* `iters` Newton iterations (each: evaluate, refactor, solve, update) on (vm, va) in place; fnorm = max |mismatch|
 * at the returned state; status = the last refactorisation's (0 ok, k+1 = bad pivot in column k).  DEVICE pointers. */
int csp3_nr_solve(const csp3_nr_plan *plan, int64_t batch, int64_t iters, const double *sspec, const int32_t *out_branch,
                  double *vm, double *va, double *fnorm, int32_t *status, void *work, void *stream);
/* Host-buffer form (chunked H2D -> iterations -> D2H pipeline on internal streams).  The start state is vm0 / va0:
 * one vector shared by all cases (start_stride = 0, e.g. a flat start) or one per case (start_stride = n_bus).
 * Per case only sspec travels to the device and (vm, va, fnorm, status) come back. */
int csp3_nr_solve_host(csp3_nr_plan *plan, int64_t batch, int64_t iters, const double *sspec, const int32_t *out_branch,
                       const double *vm0, const double *va0, int64_t start_stride, double *vm, double *va, double *fnorm,
                       int32_t *status);

#ifdef __cplusplus
}
#endif
#endif /* CSPARSE3_B200_H */
