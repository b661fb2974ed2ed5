// csc_kernels.cuh -- launchers of csc_kernels.cu
#pragma once
#include "common.cuh"

namespace csp3 {

// keeps freed stream-ordered allocations cached in the device's default pool (called by the allocating helpers)
void keep_device_pool();

struct SpmvPlanData {
    i64 m = 0, n = 0, nnz = 0;
    i32 *rp = nullptr, *rc = nullptr, *perm = nullptr;   // CSR view: row ptr, column, source CSC entry
    uint32_t *pk = nullptr;                              // perm | column << 16 when nnz, n < 65536 (one index load per entry)
};
int spmv_plan_pack(SpmvPlanData &P, cudaStream_t st);    // fills pk when the pattern qualifies

// C = A' (== CSC->CSR of A).  Any of Ci, Cx, perm may be nullptr.  perm[t] = source entry of output entry t.
int transpose_device(i64 m, i64 n, const i32 *Ap, const i32 *Ai, const double *Ax, i32 nnz, i32 *Cp, i32 *Ci,
                     double *Cx, i32 *perm, cudaStream_t st);
int spmv_device(const SpmvPlanData &P, i64 batch, const double *Ax, i64 stride_ax, const double *x, double *y,
                double beta, cudaStream_t st);
int spmm_device(const SpmvPlanData &P, i64 nv, const double *Ax, const double *X, double *Y, cudaStream_t st);
int spgemm_device(bool numeric, i64 Am, i64 An, const i32 *Ap, const i32 *Ai, const double *Ax, i64 Bm, i64 Bn,
                  const i32 *Bp, const i32 *Bi, const double *Bx, i32 *Cp, i32 *Ci, double *Cx, i64 *nnz_out,
                  cudaStream_t st);

int csc_add_device(i64 m, i64 n, const i32 *Ap, const i32 *Ai, const double *Ax, const i32 *Bp, const i32 *Bi,
                   const double *Bx, double sign, i32 *Cp, i32 *Ci, double *Cx, cudaStream_t st);

int csc_add_ff_device(i64 m, i64 n, const i32 *Ap, const i32 *Ai, const double *Ax, const i32 *Bp, const i32 *Bi,
                      const double *Bx, double alpha, double beta, i32 *Cp, i32 *Ci, double *Cx, cudaStream_t st);

}  // namespace csp3
