// ref_shim.cpp -- C-ABI thunks over the REFERENCE's own vendored scipy
// sparsetools templates, included where they lie under
// /root/reference/src/sparsetools (csc.h, csr.h, dense.h, util.h).  Built by
// `make -C oracle ref` into oracle/_ref/libsptools_ref.so (git-ignored).  No
// reference source is copied into this repo: this file only instantiates the
// <int32, double> templates.  Test infrastructure, same rules as the oracle.
//
// The typedef shim below replaces the numpy headers the templates expect
// (SURVEY.md section 8c); complex_ops.h is not needed for <int, double>.
#include <stdint.h>
#include <algorithm>
#include <complex>
#include <functional>
#include <stdexcept>
#include <vector>

typedef intptr_t npy_intp;
#define NPY_MAX_INTP INTPTR_MAX
typedef std::complex<float> npy_cfloat_wrapper;
typedef std::complex<double> npy_cdouble_wrapper;
typedef std::complex<long double> npy_clongdouble_wrapper;

#include "csc.h"

extern "C" {

void ref_csc_matvec(int n_row, int n_col, const int *Ap, const int *Ai,
                    const double *Ax, const double *Xx, double *Yx)
{ csc_matvec<int, double>(n_row, n_col, Ap, Ai, Ax, Xx, Yx); }

void ref_csc_matvecs(int n_row, int n_col, int n_vecs, const int *Ap,
                     const int *Ai, const double *Ax, const double *Xx, double *Yx)
{ csc_matvecs<int, double>(n_row, n_col, n_vecs, Ap, Ai, Ax, Xx, Yx); }

int ref_csc_matmat_pass1(int n_row, int n_col, const int *Ap, const int *Ai,
                         const int *Bp, const int *Bi, int *Cp)
{
    try { csc_matmat_pass1<int>(n_row, n_col, Ap, Ai, Bp, Bi, Cp); }
    catch (const std::overflow_error &) { return -3; }
    return 0;
}

void ref_csc_matmat_pass2(int n_row, int n_col, const int *Ap, const int *Ai,
                          const double *Ax, const int *Bp, const int *Bi,
                          const double *Bx, int *Cp, int *Ci, double *Cx)
{ csc_matmat_pass2<int, double>(n_row, n_col, Ap, Ai, Ax, Bp, Bi, Bx, Cp, Ci, Cx); }

void ref_csc_tocsr(int n_row, int n_col, const int *Ap, const int *Ai,
                   const double *Ax, int *Bp, int *Bj, double *Bx)
{ csc_tocsr<int, double>(n_row, n_col, Ap, Ai, Ax, Bp, Bj, Bx); }

void ref_csc_plus_csc(int n_row, int n_col, const int *Ap, const int *Ai,
                      const double *Ax, const int *Bp, const int *Bi,
                      const double *Bx, int *Cp, int *Ci, double *Cx)
{ csc_plus_csc<int, double>(n_row, n_col, Ap, Ai, Ax, Bp, Bi, Bx, Cp, Ci, Cx); }

void ref_csc_minus_csc(int n_row, int n_col, const int *Ap, const int *Ai,
                       const double *Ax, const int *Bp, const int *Bi,
                       const double *Bx, int *Cp, int *Ci, double *Cx)
{ csc_minus_csc<int, double>(n_row, n_col, Ap, Ai, Ax, Bp, Bi, Bx, Cp, Ci, Cx); }

}  // extern "C"
