"""B200 backend with the names and positional signatures of the reference's flat kernels.

This module is what plugs into the reference's backend seam (src/CSparse3/csc.py:33-41):

    if __config__.NATIVE: from CSparse3.csc_native import *   ->   from csparse3_b200.csc_b200 import *
    import scipy.sparse.sparsetools as sptools                ->   from csparse3_b200.csc_b200 import sptools

Every numeric function forwards to libcsparse3_b200.so (CUDA, sm_100a) through ctypes with HOST numpy
buffers, exactly like the numba kernels take them; device-resident batched work goes through
csparse3_b200.lu / csparse3_b200.spmv instead.  dtype strictness follows the reference's eager numba
signatures (i8 scalars, i4[:] indices, f8[:] values).

The last block of helpers (dense conversion, diagonals, sub-matrices, islands) is host-side
assembly glue that the reference also runs outside the numeric loop; it is plain numpy and never used by
the refactor / solve / multiply path.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import as_f64, as_i32, check, ptr

__all__ = [
    "csc_mat_vec_ff", "csc_multiply_ff", "csc_transpose", "csc_to_csr", "csc_cumsum_i", "sptools",
    "csc_add_ff", "csc_scatter_f", "csc_scatter_ff", "csc_spalloc_f", "csc_sprealloc_f", "ialloc", "xalloc",
    "csc_amd", "csc_etree", "csc_post", "csc_lu", "csc_lu_refactor", "csc_lu_solve", "csc_lusol",
    "csc_to_dense", "csc_diagonal", "csc_diagonal_from_array", "csc_stack_4_by_4_ff", "csc_sub_matrix",
    "csc_sub_matrix_cols", "csc_sub_matrix_rows", "csc_norm", "find_islands", "find_islands_batched", "coo_to_csc",
]


# ---- numeric hot path (CUDA) -----------------------------------------------------------------------------

def csc_mat_vec_ff(m, n, Ap, Ai, Ax, x):
    """y = A * x.  Reference: csc_numba.py:309-328 (same signature, returns a fresh y)."""
    Ap, Ai, Ax, x = as_i32(Ap, "Ap"), as_i32(Ai, "Ai"), as_f64(Ax, "Ax"), as_f64(x, "x")
    assert n == x.shape[0]
    y = np.empty(m, dtype=np.float64)
    check(_lib.lib().csp3_csc_mat_vec_ff_host(m, n, ptr(Ap), ptr(Ai), ptr(Ax), ptr(x), ptr(y)), "csc_mat_vec_ff")
    return y


def csc_multiply_ff(Am, An, Ap, Ai, Ax, Bm, Bn, Bp, Bi, Bx):
    """C = A * B -> (Cm, Cn, Cp, Ci, Cx, nnz).  Reference: csc_numba.py:222-306.
    Cp and nnz are identical to the reference (explicit zeros kept); row indices inside each column come
    out SORTED (the reference's order is first-touch)."""
    assert An == Bm
    Ap, Ai, Ax = as_i32(Ap, "Ap"), as_i32(Ai, "Ai"), as_f64(Ax, "Ax")
    Bp, Bi, Bx = as_i32(Bp, "Bp"), as_i32(Bi, "Bi"), as_f64(Bx, "Bx")
    L = _lib.lib()
    Cp = np.empty(Bn + 1, dtype=np.int32)
    nnz = C.c_int64(0)
    check(L.csp3_spgemm_symbolic_host(Am, An, ptr(Ap), ptr(Ai), Bm, Bn, ptr(Bp), ptr(Bi), ptr(Cp), C.byref(nnz)),
          "csc_multiply_ff (symbolic)")
    Ci = np.empty(nnz.value, dtype=np.int32)
    Cx = np.empty(nnz.value, dtype=np.float64)
    check(L.csp3_spgemm_numeric_host(Am, An, ptr(Ap), ptr(Ai), ptr(Ax), Bm, Bn, ptr(Bp), ptr(Bi), ptr(Bx),
                                     ptr(Cp), ptr(Ci), ptr(Cx)), "csc_multiply_ff (numeric)")
    return Am, Bn, Cp, Ci, Cx, int(nnz.value)


def csc_transpose(m, n, Ap, Ai, Ax):
    """C = A' -> (Cm, Cn, Cp, Ci, Cx).  Reference: csc_numba.py:400-436."""
    Ap, Ai, Ax = as_i32(Ap, "Ap"), as_i32(Ai, "Ai"), as_f64(Ax, "Ax")
    nnz = int(Ap[n])
    Cp = np.empty(m + 1, dtype=np.int32)
    Ci = np.empty(nnz, dtype=np.int32)
    Cx = np.empty(nnz, dtype=np.float64)
    check(_lib.lib().csp3_csc_transpose_host(m, n, ptr(Ap), ptr(Ai), ptr(Ax[:max(nnz, 0)]) if nnz else None,
                                             ptr(Cp), ptr(Ci), ptr(Cx)), "csc_transpose")
    return n, m, Cp, Ci, Cx


def csc_to_csr(m, n, Ap, Ai, Ax, Bp, Bi, Bx):
    """Fill caller-allocated CSR arrays.  Reference: csc_numba.py:360-397 (void, outputs preallocated)."""
    Ap, Ai, Ax = as_i32(Ap, "Ap"), as_i32(Ai, "Ai"), as_f64(Ax, "Ax")
    for a, dt in ((Bp, np.int32), (Bi, np.int32), (Bx, np.float64)):
        if a.dtype != dt or not a.flags.c_contiguous:
            raise TypeError("csc_to_csr outputs must be contiguous int32/int32/float64 arrays")
    check(_lib.lib().csp3_csc_to_csr_host(m, n, ptr(Ap), ptr(Ai), ptr(Ax), ptr(Bp), ptr(Bi), ptr(Bx)), "csc_to_csr")


def csc_cumsum_i(p, c, n):
    """p[0..n] = cumsum(c), c <- p[0..n-1]; returns sum(c).  Reference: csc_numba.py:75-94 (host helper)."""
    np.cumsum(c[:n], out=p[1:n + 1])
    p[0] = 0
    c[:n] = p[:n]
    return int(p[n])


def ialloc(n):
    """csc_numba.py:36-38"""
    return np.zeros(n, dtype=np.int32)


def xalloc(n):
    """csc_numba.py:41-43"""
    return np.zeros(n, dtype=np.float64)


def csc_spalloc_f(m, n, nzmax):
    """Allocate an empty CSC matrix -> (m, n, indptr, indices, data, nzmax).  csc_numba.py:46-60 (cs_spalloc);
    allocation bookkeeping on the host, there is nothing to run on the device."""
    nzmax = max(int(nzmax), 1)
    return m, n, ialloc(n + 1), ialloc(nzmax), xalloc(nzmax), nzmax


def csc_sprealloc_f(An, Aindptr, Aindices, Adata, nzmax):
    """Change the capacity of a CSC matrix -> (indices, data, nzmax).  csc_numba.py:97-122 (cs_sprealloc); the entries
    beyond the old length are uninitialised in the reference (np.empty), zero here."""
    if nzmax <= 0:
        nzmax = int(Aindptr[An])
    Ai = np.zeros(nzmax, dtype=np.int32)
    Ax = np.zeros(nzmax, dtype=np.float64)
    k = min(nzmax, len(Aindices)); Ai[:k] = Aindices[:k]
    k = min(nzmax, len(Adata)); Ax[:k] = Adata[:k]
    return Ai, Ax, nzmax


def csc_scatter_f(Ap, Ai, Ax, j, beta, w, x, mark, Ci, nz):
    """x += beta * A(:,j) with pattern tracking (cs_scatter).  csc_numba.py:125-151: a one-column host helper of the
    reference's own add / multiply loops (w, x, Ci are updated in place; returns the new nz).  The batched work those
    loops do runs on the device in csc_add_ff / csc_multiply_ff; this helper is kept for callers that drive it by hand."""
    Ap, Ai, Ax = as_i32(Ap, "Ap"), as_i32(Ai, "Ai"), as_f64(Ax, "Ax")
    for p in range(int(Ap[j]), int(Ap[j + 1])):
        i = int(Ai[p])
        if w[i] < mark:
            w[i] = mark
            Ci[nz] = i
            nz += 1
            x[i] = beta * Ax[p]
        else:
            x[i] += beta * Ax[p]
    return nz


csc_scatter_ff = csc_scatter_f          # csc_numba.py:154-180: the same kernel under its second name


def csc_add_ff(Am, An, Aindptr, Aindices, Adata, Bm, Bn, Bindptr, Bindices, Bdata, alpha, beta):
    """C = alpha*A + beta*B -> (Cm, Cn, Cp, Ci, Cx).  Reference: csc_numba.py:183-219 -- first-touch row order inside a
    column, explicit zeros kept, Ci / Cx returned at their allocated length nnz(A) + nnz(B) (>= 1) with a zero tail,
    exactly as the reference hands them back."""
    Ap, Ai, Ax = as_i32(Aindptr, "Aindptr"), as_i32(Aindices, "Aindices"), as_f64(Adata, "Adata")
    Bp, Bi, Bx = as_i32(Bindptr, "Bindptr"), as_i32(Bindices, "Bindices"), as_f64(Bdata, "Bdata")
    m, n = int(Am), int(Bn)
    cap = max(int(Ap[An]) + int(Bp[n]), 1)
    Cp = np.zeros(n + 1, dtype=np.int32)
    Ci = np.zeros(cap, dtype=np.int32)
    Cx = np.zeros(cap, dtype=np.float64)
    check(_lib.lib().csp3_csc_add_ff_host(m, n, ptr(Ap), ptr(Ai), ptr(Ax), ptr(Bp), ptr(Bi), ptr(Bx), float(alpha), float(beta),
                                          ptr(Cp), ptr(Ci), ptr(Cx)), "csc_add_ff")
    return m, n, Cp, Ci, Cx


class _SpTools:
    """Stand-in for the `sptools` module the reference imports at csc.py:33 (scipy.sparse.sparsetools):
    same names, argument order and in-place output convention (src/sparsetools/csc.h)."""

    @staticmethod
    def csc_matvec(n_row, n_col, Ap, Ai, Ax, Xx, Yx):
        """Yx += A * Xx.  csc.h:27-45, call site csc.py:374-379."""
        Ap, Ai, Ax, Xx = as_i32(Ap, "Ap"), as_i32(Ai, "Ai"), as_f64(Ax, "Ax"), as_f64(Xx, "Xx")
        if Yx.dtype != np.float64 or not Yx.flags.c_contiguous:
            raise ValueError("Yx must be a contiguous float64 array")
        check(_lib.lib().csp3_csc_matvec_host(n_row, n_col, ptr(Ap), ptr(Ai), ptr(Ax), ptr(Xx), ptr(Yx)), "csc_matvec")

    @staticmethod
    def csc_matvecs(n_row, n_col, n_vecs, Ap, Ai, Ax, Xx, Yx):
        """Yx[n_row,n_vecs] += A * Xx[n_col,n_vecs].  csc.h:68-84, call site csc.py:409-415."""
        Ap, Ai, Ax, Xx = as_i32(Ap, "Ap"), as_i32(Ai, "Ai"), as_f64(Ax, "Ax"), as_f64(Xx, "Xx")
        if Yx.dtype != np.float64 or not Yx.flags.c_contiguous:
            raise ValueError("Yx must be a contiguous float64 array")
        check(_lib.lib().csp3_csc_matvecs_host(n_row, n_col, n_vecs, ptr(Ap), ptr(Ai), ptr(Ax), ptr(Xx), ptr(Yx)),
              "csc_matvecs")

    @staticmethod
    def csc_matmat_pass1(n_row, n_col, Ap, Ai, Bp, Bi, Cp):
        """Column pointer of C = A*B.  csc.h:115-123; n_row = rows of A, n_col = columns of B.
        (Structural count, like pass 1 of SMMP: cancellations are not anticipated.)"""
        Ap, Ai, Bp, Bi = as_i32(Ap, "Ap"), as_i32(Ai, "Ai"), as_i32(Bp, "Bp"), as_i32(Bi, "Bi")
        An = len(Ap) - 1
        if len(Bp) != n_col + 1 or (len(Bi) and int(Bi.max()) >= An) or (len(Ai) and int(Ai.max()) >= n_row):
            raise ValueError("csc_matmat_pass1: inner dimensions do not match (A has %d columns)" % An)
        nnz = C.c_int64(0)
        rc = _lib.lib().csp3_spgemm_symbolic_host(n_row, An, ptr(Ap), ptr(Ai), An, n_col, ptr(Bp), ptr(Bi), ptr(Cp),
                                                  C.byref(nnz))
        if rc == -4:
            raise RuntimeError("nnz of the result is too large")       # std::overflow_error -> RuntimeError
        check(rc, "csc_matmat_pass1")

    @staticmethod
    def csc_matmat_pass2(n_row, n_col, Ap, Ai, Ax, Bp, Bi, Bx, Cp, Ci, Cx):
        """Entries of C = A*B given pass 1's Cp.  csc.h:125-137 -> csr.h:608-670: exact zeros are dropped and
        Cp is rewritten accordingly (row indices come out sorted; scipy emits reverse first-touch order)."""
        Ap, Ai, Ax = as_i32(Ap, "Ap"), as_i32(Ai, "Ai"), as_f64(Ax, "Ax")
        Bp, Bi, Bx = as_i32(Bp, "Bp"), as_i32(Bi, "Bi"), as_f64(Bx, "Bx")
        An = len(Ap) - 1
        if len(Bp) != n_col + 1 or (len(Bi) and int(Bi.max()) >= An) or (len(Ai) and int(Ai.max()) >= n_row):
            raise ValueError("csc_matmat_pass2: inner dimensions do not match (A has %d columns)" % An)
        check(_lib.lib().csp3_spgemm_numeric_host(n_row, An, ptr(Ap), ptr(Ai), ptr(Ax), An, n_col, ptr(Bp), ptr(Bi),
                                                  ptr(Bx), ptr(Cp), ptr(Ci), ptr(Cx)), "csc_matmat_pass2")
        nnz = int(Cp[n_col])
        keep = Cx[:nnz] != 0
        if not keep.all():                                               # output-format step only (csr.h:655)
            col = np.repeat(np.arange(n_col), np.diff(Cp[:n_col + 1]))[keep]
            k = int(keep.sum())
            Ci[:k] = Ci[:nnz][keep]
            Cx[:k] = Cx[:nnz][keep]
            Cp[:] = np.concatenate([[0], np.cumsum(np.bincount(col, minlength=n_col))]).astype(np.int32)


    @staticmethod
    def _plusminus(sign, n_row, n_col, Ap, Ai, Ax, Bp, Bi, Bx, Cp, Ci, Cx):
        Ap, Ai, Ax = as_i32(Ap, "Ap"), as_i32(Ai, "Ai"), as_f64(Ax, "Ax")
        Bp, Bi, Bx = as_i32(Bp, "Bp"), as_i32(Bi, "Bi"), as_f64(Bx, "Bx")
        if len(Ci) < int(Ap[n_col]) + int(Bp[n_col]) or len(Cx) < int(Ap[n_col]) + int(Bp[n_col]):
            raise ValueError("output arrays must hold nnz(A) + nnz(B) entries")
        check(_lib.lib().csp3_csc_plusminus_host(n_row, n_col, ptr(Ap), ptr(Ai), ptr(Ax), ptr(Bp), ptr(Bi), ptr(Bx),
                                                 float(sign), ptr(Cp), ptr(Ci), ptr(Cx)), "csc_plus/minus_csc")

    @staticmethod
    def csc_plus_csc(n_row, n_col, Ap, Ai, Ax, Bp, Bi, Bx, Cp, Ci, Cx):
        """C = A + B into caller-allocated arrays.  csc.h:203-210, call site csc.py:312-315."""
        _SpTools._plusminus(1.0, n_row, n_col, Ap, Ai, Ax, Bp, Bi, Bx, Cp, Ci, Cx)

    @staticmethod
    def csc_minus_csc(n_row, n_col, Ap, Ai, Ax, Bp, Bi, Bx, Cp, Ci, Cx):
        """C = A - B into caller-allocated arrays.  csc.h:212-219, call site csc.py:336-339."""
        _SpTools._plusminus(-1.0, n_row, n_col, Ap, Ai, Ax, Bp, Bi, Bx, Cp, Ci, Cx)


sptools = _SpTools()


# ---- LU / solve (no reference counterpart; CSparse semantics) -----------------------------

def csc_amd(order, m, n, Ap, Ai):
    """q = amd(order, A) -> int32[n] (CSparse cs_amd; host symbolic phase)."""
    Ap, Ai = as_i32(Ap, "Ap"), as_i32(Ai, "Ai")
    q = np.empty(max(n, 1), dtype=np.int32)
    check(_lib.lib().csp3_csc_amd(order, m, n, ptr(Ap), ptr(Ai), ptr(q)), "csc_amd")
    return q[:n]


def csc_etree(m, n, Ap, Ai, ata=False):
    """parent = etree(A) or etree(A'A) -> int32[n] (CSparse cs_etree)."""
    Ap, Ai = as_i32(Ap, "Ap"), as_i32(Ai, "Ai")
    parent = np.empty(max(n, 1), dtype=np.int32)
    check(_lib.lib().csp3_csc_etree(m, n, ptr(Ap), ptr(Ai), int(bool(ata)), ptr(parent)), "csc_etree")
    return parent[:n]


def csc_post(n, parent):
    """post = postorder(parent) -> int32[n] (CSparse cs_post)."""
    parent = as_i32(parent, "parent")
    post = np.empty(max(n, 1), dtype=np.int32)
    check(_lib.lib().csp3_csc_post(n, ptr(parent), ptr(post)), "csc_post")
    return post[:n]


def csc_lu(n, Ap, Ai, Ax, q, tol):
    """First factorisation -> (Lp, Li, Lx, Up, Ui, Ux, pinv) in the CSparse cs_lu layout.  The pivot
    sequence and patterns come from the host symbolic phase; the returned VALUES are recomputed by the
    CUDA refactor kernel."""
    from .lu import LuSymbolic
    sym = LuSymbolic(n, Ap, Ai, Ax, q=q, tol=tol)
    Lx, Ux, status = sym.refactor_host(np.asarray(Ax, dtype=np.float64)[None, :])
    if status[0]:
        raise ArithmeticError("zero or non-finite pivot in column %d" % (status[0] - 1))
    return sym.Lp, sym.Li, Lx[0], sym.Up, sym.Ui, Ux[0], sym.pinv


def csc_lu_refactor(n, Ap, Ai, Ax, q, pinv, Lp, Li, Up, Ui):
    """Refactor with frozen pattern and pivots -> (Lx, Ux)."""
    from .lu import LuSymbolic
    sym = LuSymbolic.from_pattern(n, Ap, Ai, q, pinv, Lp, Li, Up, Ui)
    Lx, Ux, status = sym.refactor_host(np.asarray(Ax, dtype=np.float64)[None, :])
    if status[0]:
        raise ArithmeticError("zero or non-finite pivot in column %d" % (status[0] - 1))
    return Lx[0], Ux[0]


def csc_lu_solve(n, Ap, Ai, q, pinv, Lp, Li, Lx, Up, Ui, Ux, b):
    """x = Q (U \\ (L \\ (P b))) with existing factors (tail of CSparse cs_lusol)."""
    from .lu import LuSymbolic
    sym = LuSymbolic.from_pattern(n, Ap, Ai, q, pinv, Lp, Li, Up, Ui)
    return sym.solve_host(as_f64(Lx)[None, :], as_f64(Ux)[None, :], as_f64(b)[None, :])[0]


def csc_lusol(order, n, Ap, Ai, Ax, b, tol):
    """x = A \\ b (CSparse cs_lusol): analyse on the host, refactor + solve on the GPU."""
    Ap, Ai, Ax = as_i32(Ap, "Ap"), as_i32(Ai, "Ai"), as_f64(Ax, "Ax")
    x = as_f64(b, "b").copy()
    check(_lib.lib().csp3_csc_lusol_host(order, n, ptr(Ap), ptr(Ai), ptr(Ax), ptr(x), float(tol)), "csc_lusol")
    return x


# ---- host-side assembly glue (numpy; outside the numeric loop, reference marks these OUT of the hot path) ----

def csc_to_dense(m, n, indptr, indices, data):
    """csc_numba.py:581-597 (last duplicate wins, like the reference's assignment loop)."""
    val = np.zeros((m, n), dtype=np.float64)
    cols = np.repeat(np.arange(n), np.diff(np.asarray(indptr)[:n + 1]))
    nnz = int(indptr[n])
    val[np.asarray(indices)[:nnz], cols] = np.asarray(data)[:nnz]
    return val


def csc_diagonal(m, value=1.0):
    """csc_numba.py:600-617 -> (indices, indptr, data)"""
    return np.arange(m, dtype=np.int32), np.arange(m + 1, dtype=np.int32), np.full(m, value, dtype=np.float64)


def csc_diagonal_from_array(m, array):
    """csc_numba.py:620-637 -> (indices, indptr, data)"""
    return np.arange(m, dtype=np.int32), np.arange(m + 1, dtype=np.int32), np.array(array[:m], dtype=np.float64)


def csc_stack_4_by_4_ff(am, an, Ai, Ap, Ax, bm, bn, Bi, Bp, Bx, cm, cn, Ci, Cp, Cx, dm, dn, Di, Dp, Dx):
    """[[A, B], [C, D]] -> (m, n, indices, indptr, data).  csc_numba.py:640-720 (argument order indices,
    indptr, data; column-wise concatenation, A's entries before C's inside a column).  CUDA: the pattern is
    laid out by libcsparse3_b200's plan, the values are gathered on the device (csp3_csc_stack_4_by_4_host);
    device-resident batches use csparse3_b200.assemble.Stack4Plan."""
    assert am == bm and cm == dm and an == cn and bn == dn
    Ai, Ap, Ax = as_i32(Ai, "Ai"), as_i32(Ap, "Ap"), as_f64(Ax, "Ax")
    Bi, Bp, Bx = as_i32(Bi, "Bi"), as_i32(Bp, "Bp"), as_f64(Bx, "Bx")
    Ci, Cp, Cx = as_i32(Ci, "Ci"), as_i32(Cp, "Cp"), as_f64(Cx, "Cx")
    Di, Dp, Dx = as_i32(Di, "Di"), as_i32(Dp, "Dp"), as_f64(Dx, "Dx")
    nnz = int(Ap[an]) + int(Bp[bn]) + int(Cp[cn]) + int(Dp[dn])
    indices = np.empty(nnz, dtype=np.int32)
    indptr = np.empty(an + bn + 1, dtype=np.int32)
    data = np.empty(nnz, dtype=np.float64)
    check(_lib.lib().csp3_csc_stack_4_by_4_host(am, an, ptr(Ai), ptr(Ap), ptr(Ax), bm, bn, ptr(Bi), ptr(Bp), ptr(Bx),
                                                cm, cn, ptr(Ci), ptr(Cp), ptr(Cx), dm, dn, ptr(Di), ptr(Dp), ptr(Dx),
                                                ptr(indices), ptr(indptr), ptr(data)), "csc_stack_4_by_4_ff")
    return am + cm, an + bn, indices, indptr, data


def _sub_matrix(Am, An, Ap, Ai, Ax, rows, cols):
    Ap, Ai, Ax = as_i32(Ap, "Ap"), as_i32(Ai, "Ai"), as_f64(Ax, "Ax")
    rows = None if rows is None else np.ascontiguousarray(rows, dtype=np.int32)
    cols = None if cols is None else np.ascontiguousarray(cols, dtype=np.int32)
    nc = An if cols is None else len(cols)
    nz = int(Ap[An])
    Bp = np.zeros(nc + 1, dtype=np.int32)
    Bi = np.empty(max(nz, 1), dtype=np.int32)
    Bx = np.empty(max(nz, 1), dtype=np.float64)
    nnz = C.c_int64(0)
    check(_lib.lib().csp3_csc_sub_matrix_host(Am, An, ptr(Ap), ptr(Ai), ptr(Ax), 0 if rows is None else len(rows), ptr(rows),
                                              0 if cols is None else len(cols), ptr(cols), ptr(Bp), ptr(Bi), ptr(Bx),
                                              C.byref(nnz)), "csc_sub_matrix")
    n = int(nnz.value)
    return n, Bp, Bi[:n], Bx[:n]


def csc_sub_matrix_cols(Am, Anz, Ap, Ai, Ax, cols):
    """Selected columns, all rows -> (n, Bp, Bi, Bx).  csc_numba.py:505-537; on the device (topo_kernels.cu)."""
    return _sub_matrix(Am, len(Ap) - 1, Ap, Ai, Ax, None, cols)


def csc_sub_matrix(Am, Anz, Ap, Ai, Ax, rows, cols):
    """Arbitrary sub-matrix -> (n, Bp, Bi, Bx).  csc_numba.py:464-502: entries ordered by the position of their row in
    `rows`, numbered by the reference's running counter; on the device (topo_kernels.cu).  `rows` must not repeat."""
    return _sub_matrix(Am, len(Ap) - 1, Ap, Ai, Ax, rows, cols)


def csc_sub_matrix_rows(An, Anz, Ap, Ai, Ax, rows):
    """Selected rows, all columns -> (n, Bp, Bi, Bx).  csc_numba.py:541-578; on the device."""
    Ai = as_i32(Ai, "Ai")
    Am = max(int(Ai.max()) + 1 if len(Ai) else 0, int(np.max(rows)) + 1 if len(rows) else 0)
    return _sub_matrix(Am, An, Ap, Ai, Ax, rows, None)


def csc_norm(n, Ap, Ax):
    """1-norm.  csc_numba.py:723-739"""
    Ap = np.asarray(Ap)
    if n == 0 or Ap[n] == 0:
        return 0.0
    col = np.repeat(np.arange(n), np.diff(Ap[:n + 1]))
    return float(np.bincount(col, weights=np.abs(np.asarray(Ax)[:Ap[n]]), minlength=n).max())


def find_islands(node_number, indptr, indices):
    """Islands of a graph -> list of lists of node ids, in the reference's order (islands by smallest node, nodes in
    the order of its front-popped traversal).  csc_numba.py:743-808; on the device (topo_kernels.cu: hook + pointer
    jumping for the components, then a level-synchronous breadth-first numbering)."""
    indptr, indices = as_i32(indptr, "indptr"), as_i32(indices, "indices")
    n = int(node_number)
    order = np.empty(max(n, 1), dtype=np.int32)
    iptr = np.zeros(n + 1, dtype=np.int32)
    cnt = C.c_int64(0)
    check(_lib.lib().csp3_find_islands_host(n, ptr(indptr), ptr(indices), ptr(order), ptr(iptr), C.byref(cnt)), "find_islands")
    return [order[iptr[k]:iptr[k + 1]].tolist() for k in range(int(cnt.value))]


def find_islands_batched(node_number, indptr, indices, out_from=None, out_to=None, batch=None):
    """The N-1 form of find_islands: case c is the graph without the edge (out_from[c], out_to[c]) (-1: nothing removed).
    -> (label[batch, n], n_islands[batch]); label = smallest node id of the node's island.  One CTA per case."""
    indptr, indices = as_i32(indptr, "indptr"), as_i32(indices, "indices")
    n = int(node_number)
    if out_from is not None:
        out_from, out_to = as_i32(out_from, "out_from"), as_i32(out_to, "out_to")
        batch = len(out_from)
    batch = 1 if batch is None else int(batch)
    label = np.empty((batch, n), dtype=np.int32)
    cnt = np.empty(batch, dtype=np.int32)
    check(_lib.lib().csp3_islands_batched_host(n, ptr(indptr), ptr(indices), batch, ptr(out_from), ptr(out_to), ptr(label), ptr(cnt)),
          "find_islands_batched")
    return label, cnt


def coo_to_csc(m, n, Ti, Tj, Tx, nz):
    """csc_numba.py:331-357 -> (Cm, Cn, Cp, Ci, Cx): stable counting sort by column, duplicates kept."""
    Tj = np.asarray(Tj)[:nz]
    order = np.argsort(Tj, kind="stable")
    Cp = np.concatenate([[0], np.cumsum(np.bincount(Tj, minlength=n))]).astype(np.int32)
    return m, n, Cp, np.asarray(Ti)[:nz][order].astype(np.int32), np.asarray(Tx)[:nz][order].astype(np.float64)
