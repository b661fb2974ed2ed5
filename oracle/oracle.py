"""ctypes front-end of the CPU oracle (oracle/csp3_oracle.c) -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product package never does.

Function names, argument order and return shapes follow the reference's flat
kernels (src/CSparse3/csc_numba.py) so that parity tests read like the
reference's own tests.  The LU half restates upstream CSparse (PARITY UNPINNED,
see the header of csp3_oracle.c).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libcsp3_oracle.so")
_REF_PATH = os.path.join(_HERE, "_ref", "libsptools_ref.so")

i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
i64 = C.c_int64


def build(force=False):
    """Compile the C restatement (and oracle/_ref when /root/reference exists)."""
    if force or not os.path.exists(_LIB_PATH) or \
            os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "csp3_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "all"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference/src/sparsetools") and (force or not os.path.exists(_REF_PATH)):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_csc_multiply_ff.restype = i64
        _lib.orc_csc_matmat_pass1.restype = i64
        _lib.orc_csc_matmat_pass2.restype = i64
        _lib.orc_csc_add_ff.restype = i64
        _lib.orc_csc_plusminus_csc.restype = i64
        _lib.orc_csc_stack_4_by_4_ff.restype = i64
        _lib.orc_csc_norm.restype = C.c_double
        _lib.orc_lu_levels.restype = i64
        _lib.orc_lu_refactor_flops.restype = i64
    return _lib


def ref():
    """The reference's vendored sparsetools, compiled where it lies (or None)."""
    global _ref
    if _ref is None and os.path.exists(_REF_PATH):
        _ref = C.CDLL(_REF_PATH)
    return _ref


def _i(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _pi(a):
    return a.ctypes.data_as(C.c_void_p)


# ---- kernels that exist in the reference ---------------------------------------------------

def csc_mat_vec_ff(m, n, Ap, Ai, Ax, x):
    """csc_numba.py:309-328"""
    assert n == x.shape[0]
    y = np.empty(m, dtype=np.float64)
    lib().orc_csc_mat_vec_ff(i64(m), i64(n), _pi(_i(Ap)), _pi(_i(Ai)), _pi(_f(Ax)), _pi(_f(x)), _pi(y))
    return y


def csc_matvec(n_row, n_col, Ap, Ai, Ax, Xx, Yx):
    """sparsetools csc.h:27-45 (Yx += A*Xx in place)"""
    assert Yx.dtype == np.float64 and Yx.flags.c_contiguous
    lib().orc_csc_matvec(i64(n_row), i64(n_col), _pi(_i(Ap)), _pi(_i(Ai)), _pi(_f(Ax)), _pi(_f(Xx)), _pi(Yx))


def csc_matvecs(n_row, n_col, n_vecs, Ap, Ai, Ax, Xx, Yx):
    """sparsetools csc.h:68-84"""
    assert Yx.dtype == np.float64 and Yx.flags.c_contiguous
    lib().orc_csc_matvecs(i64(n_row), i64(n_col), i64(n_vecs), _pi(_i(Ap)), _pi(_i(Ai)), _pi(_f(Ax)),
                          _pi(_f(Xx)), _pi(Yx))


def csc_multiply_ff(Am, An, Ap, Ai, Ax, Bm, Bn, Bp, Bi, Bx):
    """csc_numba.py:222-306 -> (Cm, Cn, Cp, Ci, Cx, nnz)"""
    assert An == Bm
    Ap, Ai, Ax, Bp, Bi, Bx = _i(Ap), _i(Ai), _f(Ax), _i(Bp), _i(Bi), _f(Bx)
    Cp = np.empty(Bn + 1, dtype=np.int32)
    args = (i64(Am), i64(An), _pi(Ap), _pi(Ai), _pi(Ax), i64(Bm), i64(Bn), _pi(Bp), _pi(Bi), _pi(Bx))
    nnz = lib().orc_csc_multiply_ff(*args, _pi(Cp), None, None)
    Ci = np.empty(nnz, dtype=np.int32)
    Cx = np.empty(nnz, dtype=np.float64)
    lib().orc_csc_multiply_ff(*args, _pi(Cp), _pi(Ci), _pi(Cx))
    return Am, Bn, Cp, Ci, Cx, int(nnz)


def csc_matmat_pass1(n_row, n_col, Ap, Ai, Bp, Bi, Cp):
    """sparsetools csc.h:115-123"""
    nnz = lib().orc_csc_matmat_pass1(i64(n_row), i64(n_col), _pi(_i(Ap)), _pi(_i(Ai)), _pi(_i(Bp)), _pi(_i(Bi)),
                                     _pi(Cp))
    if nnz == -3:
        raise RuntimeError("nnz of the result is too large")


def csc_matmat_pass2(n_row, n_col, Ap, Ai, Ax, Bp, Bi, Bx, Cp, Ci, Cx):
    """sparsetools csc.h:125-137"""
    return int(lib().orc_csc_matmat_pass2(i64(n_row), i64(n_col), _pi(_i(Ap)), _pi(_i(Ai)), _pi(_f(Ax)),
                                          _pi(_i(Bp)), _pi(_i(Bi)), _pi(_f(Bx)), _pi(Cp), _pi(Ci), _pi(Cx)))


def csc_transpose(m, n, Ap, Ai, Ax):
    """csc_numba.py:400-436 -> (Cm, Cn, Cp, Ci, Cx)"""
    Ap, Ai, Ax = _i(Ap), _i(Ai), _f(Ax)
    nnz = int(Ap[n])
    Cp = np.empty(m + 1, dtype=np.int32)
    Ci = np.empty(max(nnz, 1), dtype=np.int32)
    Cx = np.empty(max(nnz, 1), dtype=np.float64)
    lib().orc_csc_transpose(i64(m), i64(n), _pi(Ap), _pi(Ai), _pi(Ax), _pi(Cp), _pi(Ci), _pi(Cx))
    return n, m, Cp, Ci[:nnz], Cx[:nnz]


def csc_to_csr(m, n, Ap, Ai, Ax, Bp, Bi, Bx):
    """csc_numba.py:360-397 (fills Bp, Bi, Bx)"""
    lib().orc_csc_to_csr(i64(m), i64(n), _pi(_i(Ap)), _pi(_i(Ai)), _pi(_f(Ax)), _pi(Bp), _pi(Bi), _pi(Bx))


def csc_add_ff(Am, An, Ap, Ai, Ax, Bm, Bn, Bp, Bi, Bx, alpha, beta):
    """csc_numba.py:183-219 -> (Cm, Cn, Cp, Ci, Cx)"""
    Ap, Ai, Ax, Bp, Bi, Bx = _i(Ap), _i(Ai), _f(Ax), _i(Bp), _i(Bi), _f(Bx)
    cap = max(int(Ap[An]) + int(Bp[Bn]), 1)
    Cp = np.empty(Bn + 1, dtype=np.int32)
    Ci = np.empty(cap, dtype=np.int32)
    Cx = np.empty(cap, dtype=np.float64)
    nnz = lib().orc_csc_add_ff(i64(Am), i64(Bn), _pi(Ap), _pi(Ai), _pi(Ax), _pi(Bp), _pi(Bi), _pi(Bx),
                               C.c_double(alpha), C.c_double(beta), _pi(Cp), _pi(Ci), _pi(Cx))
    return Am, Bn, Cp, Ci[:nnz], Cx[:nnz]


def csc_plusminus_csc(n_row, n_col, Ap, Ai, Ax, Bp, Bi, Bx, sign, Cp, Ci, Cx):
    """sparsetools csc.h:203-219 (csc_plus_csc sign=+1 / csc_minus_csc sign=-1)"""
    return int(lib().orc_csc_plusminus_csc(i64(n_row), i64(n_col), _pi(_i(Ap)), _pi(_i(Ai)), _pi(_f(Ax)),
                                           _pi(_i(Bp)), _pi(_i(Bi)), _pi(_f(Bx)), C.c_double(sign),
                                           _pi(Cp), _pi(Ci), _pi(Cx)))


def csc_stack_4_by_4_ff(am, an, Ai, Ap, Ax, bm, bn, Bi, Bp, Bx, cm, cn, Ci, Cp, Cx, dm, dn, Di, Dp, Dx):
    """csc_numba.py:640-720 -> (m, n, indices, indptr, data); NB argument order indices, indptr, data."""
    assert am == bm and cm == dm and an == cn and bn == dn
    arrs = [_i(Ai), _i(Ap), _f(Ax), _i(Bi), _i(Bp), _f(Bx), _i(Ci), _i(Cp), _f(Cx), _i(Di), _i(Dp), _f(Dx)]
    nnz = int(arrs[1][an] + arrs[4][bn] + arrs[7][cn] + arrs[10][dn])
    indices = np.zeros(nnz, dtype=np.int32)
    indptr = np.zeros(an + bn + 1, dtype=np.int32)
    data = np.zeros(nnz, dtype=np.float64)
    lib().orc_csc_stack_4_by_4_ff(i64(am), i64(an), _pi(arrs[0]), _pi(arrs[1]), _pi(arrs[2]),
                                  i64(bm), i64(bn), _pi(arrs[3]), _pi(arrs[4]), _pi(arrs[5]),
                                  i64(cm), i64(cn), _pi(arrs[6]), _pi(arrs[7]), _pi(arrs[8]),
                                  i64(dm), i64(dn), _pi(arrs[9]), _pi(arrs[10]), _pi(arrs[11]),
                                  _pi(indices), _pi(indptr), _pi(data))
    return am + cm, an + bn, indices, indptr, data


def csc_norm(n, Ap, Ax):
    """csc_numba.py:723-739"""
    return float(lib().orc_csc_norm(i64(n), _pi(_i(Ap)), _pi(_f(Ax))))


# ---- CSparse restatement (parity unpinned) -------------------------------------------------

def csc_amd(order, m, n, Ap, Ai):
    """CSparse cs_amd -> q int32[n]"""
    P = np.empty(n + 1, dtype=np.int32)
    st = lib().orc_csc_amd(i64(order), i64(m), i64(n), _pi(_i(Ap)), _pi(_i(Ai)), _pi(P))
    assert st == 0
    return P[:n].copy()


def csc_etree(m, n, Ap, Ai, ata=False):
    """CSparse cs_etree -> parent int32[n]"""
    parent = np.empty(n, dtype=np.int32)
    lib().orc_csc_etree(i64(m), i64(n), _pi(_i(Ap)), _pi(_i(Ai)), C.c_int(int(ata)), _pi(parent))
    return parent


def csc_post(n, parent):
    """CSparse cs_post -> post int32[n]"""
    post = np.empty(n, dtype=np.int32)
    lib().orc_csc_post(i64(n), _pi(_i(parent)), _pi(post))
    return post


def _take(ptr, count, dtype):
    ct = {np.int32: C.c_int32, np.float64: C.c_double}[dtype]
    arr = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(max(count, 1),))[:count].copy()
    lib().orc_free(ptr)
    return arr


def csc_lu(n, Ap, Ai, Ax, q, tol):
    """CSparse cs_lu -> (Lp, Li, Lx, Up, Ui, Ux, pinv).  Raises ZeroDivisionError-like
    ArithmeticError on a structurally/numerically singular step."""
    Ap, Ai, Ax = _i(Ap), _i(Ai), _f(Ax)
    qa = None if q is None else _i(q)
    outs = [C.c_void_p() for _ in range(7)]
    st = lib().orc_csc_lu(i64(n), _pi(Ap), _pi(Ai), _pi(Ax), None if qa is None else _pi(qa), C.c_double(tol),
                          *[C.byref(o) for o in outs])
    if st != 0:
        raise ArithmeticError("singular matrix: no pivot in step %d" % (st - 1))
    Lp = _take(outs[0], n + 1, np.int32)
    lnz = int(Lp[n])
    Li = _take(outs[1], lnz, np.int32)
    Lx = _take(outs[2], lnz, np.float64)
    Up = _take(outs[3], n + 1, np.int32)
    unz = int(Up[n])
    Ui = _take(outs[4], unz, np.int32)
    Ux = _take(outs[5], unz, np.float64)
    pinv = _take(outs[6], n, np.int32)
    return Lp, Li, Lx, Up, Ui, Ux, pinv


def csc_lu_refactor(n, Ap, Ai, Ax, q, pinv, Lp, Li, Up, Ui):
    """Frozen-pattern, frozen-pivot refactorization -> (Lx, Ux)."""
    Lx = np.empty(int(Lp[n]), dtype=np.float64)
    Ux = np.empty(int(Up[n]), dtype=np.float64)
    qa = None if q is None else _i(q)
    st = lib().orc_csc_lu_refactor(i64(n), _pi(_i(Ap)), _pi(_i(Ai)), _pi(_f(Ax)), None if qa is None else _pi(qa),
                                   _pi(_i(pinv)), _pi(_i(Lp)), _pi(_i(Li)), _pi(_i(Up)), _pi(_i(Ui)), _pi(Lx), _pi(Ux))
    if st != 0:
        raise ArithmeticError("zero or non-finite pivot in column %d" % (st - 1))
    return Lx, Ux


def csc_lsolve(n, Lp, Li, Lx, x):
    lib().orc_csc_lsolve(i64(n), _pi(_i(Lp)), _pi(_i(Li)), _pi(_f(Lx)), _pi(x))


def csc_usolve(n, Up, Ui, Ux, x):
    lib().orc_csc_usolve(i64(n), _pi(_i(Up)), _pi(_i(Ui)), _pi(_f(Ux)), _pi(x))


def csc_lu_solve(n, Lp, Li, Lx, Up, Ui, Ux, pinv, q, b):
    """x = Q * (U \\ (L \\ (P*b)))  (tail of CSparse cs_lusol)"""
    out = np.empty(n, dtype=np.float64)
    work = np.empty(n, dtype=np.float64)
    qa = None if q is None else _i(q)
    lib().orc_csc_lu_solve(i64(n), _pi(_i(Lp)), _pi(_i(Li)), _pi(_f(Lx)), _pi(_i(Up)), _pi(_i(Ui)), _pi(_f(Ux)),
                           _pi(_i(pinv)), None if qa is None else _pi(qa), _pi(_f(b)), _pi(out), _pi(work))
    return out


def csc_lusol(order, n, Ap, Ai, Ax, b, tol):
    """CSparse cs_lusol -> x"""
    x = _f(b).copy()
    st = lib().orc_csc_lusol(i64(order), i64(n), _pi(_i(Ap)), _pi(_i(Ai)), _pi(_f(Ax)), _pi(x), C.c_double(tol))
    if st != 0:
        raise ArithmeticError("singular matrix: no pivot in step %d" % (st - 1))
    return x


def lu_levels(n, Gp, Gi, kind):
    """kind 0: refactor levels from U; 1: L-solve levels from L; 2: U-solve levels from U.
    -> (level[n], order[n], lptr[nlev+1])"""
    level = np.empty(n, dtype=np.int32)
    order = np.empty(n, dtype=np.int32)
    lptr = np.empty(n + 2, dtype=np.int32)
    nlev = lib().orc_lu_levels(i64(n), _pi(_i(Gp)), _pi(_i(Gi)), C.c_int(kind), _pi(level), _pi(order), _pi(lptr))
    return level, order, lptr[:nlev + 1].copy()


def lu_refactor_flops(n, Lp, Up, Ui):
    return int(lib().orc_lu_refactor_flops(i64(n), _pi(_i(Lp)), _pi(_i(Up)), _pi(_i(Ui))))


def csc_lu_refactor_solve_batch(n, Ap, Ai, q, pinv, Lp, Li, Up, Ui, Ax, b, threads=1):
    """Refactor + solve a batch on one pattern with `threads` OpenMP threads -> (x[batch, n], n_bad)."""
    Ax = np.ascontiguousarray(Ax, dtype=np.float64).reshape(-1, int(Ap[n]))
    b = np.ascontiguousarray(b, dtype=np.float64).reshape(-1, n)
    x = np.empty_like(b)
    qa = None if q is None else _i(q)
    bad = lib().orc_csc_lu_refactor_solve_batch(i64(n), _pi(_i(Ap)), _pi(_i(Ai)), None if qa is None else _pi(qa),
                                                _pi(_i(pinv)), _pi(_i(Lp)), _pi(_i(Li)), _pi(_i(Up)), _pi(_i(Ui)),
                                                i64(Ax.shape[0]), _pi(Ax), _pi(b), _pi(x), C.c_int(int(threads)))
    return x, int(bad)
