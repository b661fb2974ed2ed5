"""CscMat -- the reference's CSC matrix type, kept as the drop-in container.

Interface mirror of src/CSparse3/csc.py:44-606: same constructor, same public fields (m, n, nzmax, indptr
int32[n+1], indices int32[nnz], data float64[nnz]), same operators and free functions.  The arithmetic
behind the operators goes to the B200 backend (csparse3_b200.csc_b200) instead of numba / scipy sparsetools.
New on top of the reference: `lu()` and `solve()`.
"""
from collections.abc import Iterable

import numpy as np

from .csc_b200 import (csc_diagonal, csc_diagonal_from_array, csc_multiply_ff, csc_stack_4_by_4_ff,
                       csc_sub_matrix, csc_sub_matrix_cols, csc_sub_matrix_rows, csc_to_csr, csc_to_dense,
                       csc_transpose, find_islands, sptools)


def _index_array(key):
    if isinstance(key, (int, np.integer)):
        return np.array([key], dtype=np.int32)
    return np.asarray(key, dtype=np.int32)


class CscMat:
    """Matrix in compressed-column form (reference csc.py:44-141)."""

    def __init__(self, m=0, n=0, nz_max=0, indptr=None, indices=None, data=None, zeros=False):
        self.m = m
        self.n = n
        if indptr is None:
            alloc = np.zeros if zeros else np.empty
            self.nzmax = max(nz_max, 1)
            self.indptr = alloc(n + 1, dtype=np.int32)
            self.indices = alloc(nz_max, dtype=np.int32)
            self.data = alloc(nz_max, dtype=np.float64)
        else:
            self.indptr, self.indices, self.data = indptr, indices, data
            self.nzmax = len(self.data)

    # ---- slicing (reference csc.py:143-290; same eight key forms) ---------------------------------------------
    def __getitem__(self, key):
        if not isinstance(key, tuple):
            raise Exception('The indices must be a tuple :/')
        r, c = key
        r_int, c_int = isinstance(r, (int, np.integer)), isinstance(c, (int, np.integer))
        r_all, c_all = isinstance(r, slice), isinstance(c, slice)
        if r_int and c_int:
            raise NotImplementedError('Single value extraction not implemented')
        if r_all and c_all:
            return self
        if r_all:                                   # (:, b) or (:, list_b)
            cols = _index_array(c)
            _, Bp, Bi, Bx = csc_sub_matrix_cols(self.m, self.nzmax, self.indptr, self.indices, self.data, cols)
            return CscMat(m=self.m, n=len(cols), indptr=Bp, indices=Bi, data=Bx)
        if c_all:                                   # (a, :) or (list_a, :)
            rows = _index_array(r)
            _, Bp, Bi, Bx = csc_sub_matrix_rows(self.n, self.nzmax, self.indptr, self.indices, self.data, rows)
            return CscMat(m=len(rows), n=self.n, indptr=Bp, indices=Bi, data=Bx)
        if (r_int or isinstance(r, Iterable)) and (c_int or isinstance(c, Iterable)):
            rows, cols = _index_array(r), _index_array(c)
            _, Bp, Bi, Bx = csc_sub_matrix(self.m, self.nzmax, self.indptr, self.indices, self.data, rows, cols)
            return CscMat(m=len(rows), n=len(cols), indptr=Bp, indices=Bi, data=Bx)
        raise Exception('The indices must be a tuple :/')

    def __setitem__(self, key, value):
        raise Exception('Setting values is not allowed in a CSC Matrix, use a Lil Matrix instead and convert it to CSC')

    def __str__(self):
        return str(self.todense())

    # ---- arithmetic -------------------------------------------------------------------------------------------
    def _plusminus(self, other, fn):
        assert other.m == self.m and other.n == self.n
        na, nb = int(self.indptr[self.n]), int(other.indptr[other.n])
        C = CscMat(m=self.m, n=self.n, nz_max=na + nb, zeros=True)
        fn(self.m, self.n, self.indptr, self.indices[:na], self.data[:na],
           other.indptr, other.indices[:nb], other.data[:nb], C.indptr, C.indices, C.data)
        return C

    def __add__(self, other):
        """Reference csc.py:301-323 (sptools.csc_plus_csc): result arrays are sized nnz(A)+nnz(B) like the
        reference's, the valid part is indptr[n] entries."""
        if isinstance(other, CscMat):
            return self._plusminus(other, sptools.csc_plus_csc)
        if isinstance(other, (float, int)):
            raise NotImplementedError('Adding a nonzero scalar to a sparse matrix would make it a dense matrix.')
        raise NotImplementedError('Type not supported')

    def __sub__(self, other):
        """Reference csc.py:325-346 (sptools.csc_minus_csc)."""
        if isinstance(other, CscMat):
            return self._plusminus(other, sptools.csc_minus_csc)
        if isinstance(other, (float, int)):
            raise NotImplementedError('Adding a non-zero scalar to a sparse matrix would make it a dense matrix.')
        raise NotImplementedError('Type not supported')

    def __mul__(self, other):
        """Reference csc.py:348-423: matrix -> two-pass SpGEMM, 1-D array -> SpMV, 2-D array -> SpMM,
        scalar -> scaled copy."""
        if isinstance(other, CscMat):
            if self.n != other.m:
                raise ValueError("dimension mismatch: (%d, %d) * (%d, %d)" % (self.m, self.n, other.m, other.n))
            Cp = np.empty(other.n + 1, dtype=np.int32)
            sptools.csc_matmat_pass1(self.m, other.n, self.indptr, self.indices, other.indptr, other.indices, Cp)
            nnz = int(Cp[-1])
            Ci = np.empty(nnz, dtype=np.int32)
            Cx = np.empty(nnz, dtype=np.float64)
            sptools.csc_matmat_pass2(self.m, other.n, self.indptr, self.indices, self.data,
                                     other.indptr, other.indices, other.data, Cp, Ci, Cx)
            nnz = int(Cp[-1])
            return CscMat(m=self.m, n=other.n, indptr=Cp, indices=Ci[:nnz], data=Cx[:nnz])
        if isinstance(other, np.ndarray):
            if other.shape[0] != self.n:
                raise ValueError("dimension mismatch: (%d, %d) * %s" % (self.m, self.n, other.shape))
            if other.ndim == 1:
                y = np.zeros(self.m, dtype=np.float64)
                sptools.csc_matvec(self.m, self.n, self.indptr, self.indices, self.data, other, y)
                return y
            if other.ndim == 2:
                n_vecs = other.shape[1]
                y = np.zeros((self.m, n_vecs), dtype=np.float64)
                sptools.csc_matvecs(self.m, self.n, n_vecs, self.indptr, self.indices, self.data,
                                    np.ascontiguousarray(other), y)
                return y
        if isinstance(other, (float, int)):
            out = self.copy()
            out.data *= other
            return out
        raise Exception('Type not supported')

    def __neg__(self):
        return self.__mul__(-1.0)

    def __eq__(self, other):
        """Reference csc.py:432-457: exact equality of indices, indptr and data."""
        if self.shape != other.shape:
            return False
        return (np.array_equal(self.indices, other.indices) and np.array_equal(self.indptr, other.indptr)
                and np.array_equal(self.data, other.data))

    __hash__ = None

    # ---- conversions ------------------------------------------------------------------------------------------
    def todense(self):
        return csc_to_dense(self.m, self.n, self.indptr, self.indices, self.data)

    def to_csr(self):
        """-> (Bp, Bi, Bx).  Reference csc.py:466-478."""
        nnz = int(self.indptr[self.n])
        Bp = np.zeros(self.m + 1, dtype=np.int32)
        Bi = np.empty(nnz, dtype=np.int32)
        Bx = np.empty(nnz, dtype=np.float64)
        csc_to_csr(self.m, self.n, self.indptr, self.indices, self.data, Bp, Bi, Bx)
        return Bp, Bi, Bx

    def get_nnz(self):
        return self.indptr[self.n]

    def dot(self, o):
        """C = self * o via csc_multiply_ff.  Reference csc.py:483-500."""
        out = CscMat()
        out.m, out.n, out.indptr, out.indices, out.data, out.nzmax = csc_multiply_ff(
            self.m, self.n, self.indptr, self.indices, self.data, o.m, o.n, o.indptr, o.indices, o.data)
        return out

    def t(self):
        """Transpose.  Reference csc.py:502-513."""
        out = CscMat()
        out.m, out.n, out.indptr, out.indices, out.data = csc_transpose(self.m, self.n, self.indptr, self.indices,
                                                                         self.data)
        out.nzmax = len(out.data)
        return out

    def islands(self):
        return [np.sort(np.array(isl)) for isl in find_islands(self.n, self.indptr, self.indices)]

    @property
    def shape(self):
        return self.m, self.n

    def copy(self):
        out = CscMat()
        out.m, out.n, out.nzmax = self.m, self.n, self.nzmax
        out.data, out.indices, out.indptr = self.data.copy(), self.indices.copy(), self.indptr.copy()
        return out

    # ---- LU / solve (new; CSparse cs_lusol semantics) -----------------------------------------------------------
    def lu(self, order=1, tol=1e-3):
        """Symbolic analysis + first factorisation -> csparse3_b200.lu.LuSymbolic (cache it per pattern)."""
        from .lu import LuSymbolic
        assert self.m == self.n
        nnz = int(self.indptr[self.n])
        return LuSymbolic(self.n, self.indptr, self.indices[:nnz], self.data[:nnz], order=order, tol=tol)

    def solve(self, b, order=1, tol=1e-3, sym=None):
        """x = A \\ b on the GPU; pass a cached `sym` (from lu()) to skip the host symbolic phase."""
        sym = self.lu(order, tol) if sym is None else sym
        nnz = int(self.indptr[self.n])
        b2 = np.ascontiguousarray(b, dtype=np.float64).reshape(1, -1)
        x, status = sym.refactor_solve_host(np.ascontiguousarray(self.data[:nnz]).reshape(1, -1), b2)
        if status[0]:
            raise ArithmeticError('zero or non-finite pivot in column %d' % (status[0] - 1))
        return x[0]


def scipy_to_mat(scipy_mat):
    """Alias a scipy CSC matrix without copying (reference csc.py:541-553)."""
    mat = CscMat()
    mat.m, mat.n = scipy_mat.shape
    mat.data, mat.indices, mat.indptr = scipy_mat.data, scipy_mat.indices, scipy_mat.indptr
    mat.nzmax = scipy_mat.nnz
    return mat


def Diag(m, n, value=1.0):
    """Diagonal matrix of `value` (reference csc.py:556-569)."""
    A = CscMat(m, n)
    A.indices, A.indptr, A.data = csc_diagonal(A.m, value)
    A.n = A.m
    A.nzmax = A.indptr[A.n]
    return A


def Diags(array):
    """Diagonal matrix from an array (reference csc.py:572-585)."""
    m = array.shape[0]
    A = CscMat(m, m)
    A.indices, A.indptr, A.data = csc_diagonal_from_array(A.m, array)
    A.nzmax = A.indptr[A.n]
    return A


def pack_4_by_4(A11, A12, A21, A22):
    """[[A11, A12], [A21, A22]] (reference csc.py:588-606)."""
    m, n, Pi, Pp, Px = csc_stack_4_by_4_ff(A11.m, A11.n, A11.indices, A11.indptr, A11.data,
                                           A12.m, A12.n, A12.indices, A12.indptr, A12.data,
                                           A21.m, A21.n, A21.indices, A21.indptr, A21.data,
                                           A22.m, A22.n, A22.indices, A22.indptr, A22.data)
    P = CscMat(m, n)
    P.indptr, P.indices, P.data = Pp, Pi, Px
    P.nzmax = len(Px)
    return P
