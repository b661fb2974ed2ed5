"""Pin the oracle's half (1) -- the kernels that exist in the reference -- against:
  * the reference's own golden vectors (cscs_to_csr_test.py:13-25, docs/connectivity_matrix.rst:93-105),
  * fixtures recorded from the reference's numba kernels (tests/golden/make_golden.py),
  * scipy's operators, with the exact-equality criterion of src/test/test1_operations.py:55-61,
  * oracle/_ref (the reference's vendored sparsetools C++ compiled where it lies), when present.
"""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import CONNECTIVITY, SIX_BY_THREE, sort_columns
from oracle import oracle as orc


def test_csc_to_csr_reference_golden():
    g = SIX_BY_THREE
    Bp = np.zeros(g["m"] + 1, dtype=np.int32); Bi = np.empty(10, dtype=np.int32); Bx = np.empty(10)
    orc.csc_to_csr(g["m"], g["n"], g["indptr"], g["indices"], g["data"], Bp, Bi, Bx)
    assert (Bp == g["csr_indptr"]).all() and (Bi == g["csr_indices"]).all() and (Bx == g["csr_data"]).all()


def test_connectivity_docs_golden():
    g = CONNECTIVITY
    Tm, Tn, Tp, Ti, Tx = orc.csc_transpose(g["m"], g["n"], g["indptr"], g["indices"], g["data"])
    assert (Tm, Tn) == (5, 3)
    y = np.zeros(5)
    orc.csc_matvec(5, 3, Tp, Ti, Tx, g["p"], y)
    assert (y == g["injections"]).all()
    assert (orc.csc_mat_vec_ff(5, 3, Tp, Ti, Tx, g["p"]) == g["injections"]).all()


def test_test1_operations_fixture(golden_test1):
    d = golden_test1
    m, n = d["Ashape"]
    Ap, Ai, Ax, Bp, Bi, Bx = d["Ap"], d["Ai"], d["Ax"], d["Bp"], d["Bi"], d["Bx"]
    # reference numba kernels, bit for bit (order inside columns included)
    assert np.array_equal(orc.csc_mat_vec_ff(m, n, Ap, Ai, Ax, d["x"]), d["ref_matvec"])
    Cm, Cn, Cp, Ci, Cx, nnz = orc.csc_multiply_ff(m, n, Ap, Ai, Ax, m, n, Bp, Bi, Bx)
    assert np.array_equal(Cp, d["ref_mul_p"]) and np.array_equal(Ci, d["ref_mul_i"]) and np.array_equal(Cx, d["ref_mul_x"])
    Tm, Tn, Tp, Ti, Tx = orc.csc_transpose(m, n, Ap, Ai, Ax)
    assert np.array_equal(Tp, d["ref_t_p"]) and np.array_equal(Ti, d["ref_t_i"]) and np.array_equal(Tx, d["ref_t_x"])
    Rp = np.zeros(m + 1, dtype=np.int32); Ri = np.empty(Ap[n], dtype=np.int32); Rx = np.empty(Ap[n])
    orc.csc_to_csr(m, n, Ap, Ai, Ax, Rp, Ri, Rx)
    assert np.array_equal(Rp, d["ref_csr_p"]) and np.array_equal(Ri, d["ref_csr_i"]) and np.array_equal(Rx, d["ref_csr_x"])
    # scipy operators, exact dense equality as in test1_operations.py:55-61
    y = np.zeros(m); orc.csc_matvec(m, n, Ap, Ai, Ax, d["x"], y)
    assert (y == d["scipy_Ax"]).all()
    Y = np.zeros((m, 5)); orc.csc_matvecs(m, n, 5, Ap, Ai, Ax, d["xx"], Y)
    assert (Y == d["scipy_Axx"]).all()
    Cp2 = np.empty(n + 1, dtype=np.int32); orc.csc_matmat_pass1(m, n, Ap, Ai, Bp, Bi, Cp2)
    Ci2 = np.empty(Cp2[-1], dtype=np.int32); Cx2 = np.empty(Cp2[-1])
    k = orc.csc_matmat_pass2(m, n, Ap, Ai, Ax, Bp, Bi, Bx, Cp2, Ci2, Cx2)
    assert (sp.csc_matrix((Cx2[:k], Ci2[:k], Cp2), shape=(m, n)).toarray() == d["scipy_AB"]).all()
    assert (sp.csc_matrix((Tx, Ti, Tp), shape=(n, m)).toarray() == d["scipy_AT"]).all()
    for sign, key in ((1.0, "scipy_ApB"), (-1.0, "scipy_AmB")):
        Sp_ = np.empty(n + 1, dtype=np.int32); Si = np.empty(Ap[n] + Bp[n], dtype=np.int32); Sx = np.empty(Ap[n] + Bp[n])
        k = orc.csc_plusminus_csc(m, n, Ap, Ai, Ax, Bp, Bi, Bx, sign, Sp_, Si, Sx)
        assert (sp.csc_matrix((Sx[:k], Si[:k], Sp_), shape=(m, n)).toarray() == d[key]).all()
    Am, An, Sp2, Si2, Sx2 = orc.csc_add_ff(m, n, Ap, Ai, Ax, m, n, Bp, Bi, Bx, 1.0, 1.0)
    assert (sp.csc_matrix((Sx2, Si2, Sp2), shape=(m, n)).toarray() == d["scipy_ApB"]).all()


def test_reference_kernel_fixture(golden_ref):
    d = golden_ref
    assert np.array_equal(orc.csc_mat_vec_ff(37, 53, d["Rp"], d["Ri"], d["Rx"], d["xr"]), d["ref_R_matvec"])
    Cm, Cn, Cp, Ci, Cx, nnz = orc.csc_multiply_ff(37, 53, d["Rp"], d["Ri"], d["Rx"], 53, 53, d["Sp"], d["Si"], d["Sx"])
    assert np.array_equal(Cp, d["ref_RS_p"]) and np.array_equal(Ci, d["ref_RS_i"]) and np.array_equal(Cx, d["ref_RS_x"])
    Tm, Tn, Tp, Ti, Tx = orc.csc_transpose(37, 53, d["Rp"], d["Ri"], d["Rx"])
    assert np.array_equal(Tp, d["ref_Rt_p"]) and np.array_equal(Ti, d["ref_Rt_i"]) and np.array_equal(Tx, d["ref_Rt_x"])
    # unsorted row indices inside columns
    assert np.array_equal(orc.csc_mat_vec_ff(53, 53, d["U_p"], d["U_i"], d["U_x"], d["xr"]), d["ref_U_matvec"])
    Tm, Tn, Tp, Ti, Tx = orc.csc_transpose(53, 53, d["U_p"], d["U_i"], d["U_x"])
    assert np.array_equal(Tp, d["ref_Ut_p"]) and np.array_equal(Ti, d["ref_Ut_i"]) and np.array_equal(Tx, d["ref_Ut_x"])
    st = [d["st_" + c + k] for c in "abcd" for k in ("shape", "i", "p", "x")]
    args = []
    for t in range(4):
        sh, ii, pp, xx = st[4 * t:4 * t + 4]
        args += [int(sh[0]), int(sh[1]), ii, pp, xx]
    mm, nn, Pi, Pp, Px = orc.csc_stack_4_by_4_ff(*args)
    assert (mm, nn) == tuple(d["ref_st_shape"])
    assert np.array_equal(Pi, d["ref_st_i"]) and np.array_equal(Pp, d["ref_st_p"]) and np.array_equal(Px, d["ref_st_x"])


def test_empty_and_ragged():
    # empty matrix, empty columns, m != n
    Ap = np.zeros(4, dtype=np.int32); Ai = np.zeros(0, dtype=np.int32); Ax = np.zeros(0)
    assert (orc.csc_mat_vec_ff(5, 3, Ap, Ai, Ax, np.ones(3)) == 0).all()
    Tm, Tn, Tp, Ti, Tx = orc.csc_transpose(5, 3, Ap, Ai, Ax)
    assert (Tp == 0).all() and len(Ti) == 0
    Cm, Cn, Cp, Ci, Cx, nnz = orc.csc_multiply_ff(5, 3, Ap, Ai, Ax, 3, 4, np.zeros(5, dtype=np.int32), Ai, Ax)
    assert nnz == 0 and (Cp == 0).all()


@pytest.mark.skipif(orc.ref() is None, reason="oracle/_ref not built (needs /root/reference at build time)")
def test_against_compiled_reference_sparsetools(golden_test1):
    """The restatement vs the reference's own C++ (src/sparsetools/csc.h) compiled in oracle/_ref."""
    R = orc.ref()
    rng = np.random.default_rng(3)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    for trial in range(6):
        m, k, n = (int(v) for v in rng.integers(1, 60, 3))
        A = sp.csc_matrix(sp.random(m, k, density=0.15, random_state=int(rng.integers(1 << 30))))
        B = sp.csc_matrix(sp.random(k, n, density=0.15, random_state=int(rng.integers(1 << 30))))
        Ap, Ai, Ax = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data
        Bp, Bi, Bx = B.indptr.astype(np.int32), B.indices.astype(np.int32), B.data
        x = rng.standard_normal(k); X = rng.standard_normal((k, 3))
        y0 = rng.standard_normal(m); y1 = y0.copy(); y2 = y0.copy()
        R.ref_csc_matvec(m, k, vp(Ap), vp(Ai), vp(Ax), vp(x), vp(y1)); orc.csc_matvec(m, k, Ap, Ai, Ax, x, y2)
        assert np.array_equal(y1, y2)
        Y1 = np.zeros((m, 3)); Y2 = np.zeros((m, 3))
        R.ref_csc_matvecs(m, k, 3, vp(Ap), vp(Ai), vp(Ax), vp(X), vp(Y1)); orc.csc_matvecs(m, k, 3, Ap, Ai, Ax, X, Y2)
        assert np.array_equal(Y1, Y2)
        Cp1 = np.empty(n + 1, dtype=np.int32); Cp2 = np.empty(n + 1, dtype=np.int32)
        assert R.ref_csc_matmat_pass1(m, n, vp(Ap), vp(Ai), vp(Bp), vp(Bi), vp(Cp1)) == 0
        orc.csc_matmat_pass1(m, n, Ap, Ai, Bp, Bi, Cp2)
        assert np.array_equal(Cp1, Cp2)
        Ci1 = np.empty(Cp1[-1], dtype=np.int32); Cx1 = np.empty(Cp1[-1]); Ci2 = Ci1.copy(); Cx2 = Cx1.copy()
        R.ref_csc_matmat_pass2(m, n, vp(Ap), vp(Ai), vp(Ax), vp(Bp), vp(Bi), vp(Bx), vp(Cp1), vp(Ci1), vp(Cx1))
        orc.csc_matmat_pass2(m, n, Ap, Ai, Ax, Bp, Bi, Bx, Cp2, Ci2, Cx2)
        kk = Cp1[-1]
        assert np.array_equal(Cp1, Cp2) and np.array_equal(Ci1[:kk], Ci2[:kk]) and np.array_equal(Cx1[:kk], Cx2[:kk])
        Tp1 = np.empty(m + 1, dtype=np.int32); Ti1 = np.empty(Ap[k], dtype=np.int32); Tx1 = np.empty(Ap[k])
        R.ref_csc_tocsr(m, k, vp(Ap), vp(Ai), vp(Ax), vp(Tp1), vp(Ti1), vp(Tx1))
        Tp2 = np.zeros(m + 1, dtype=np.int32); Ti2 = Ti1.copy(); Tx2 = Tx1.copy()
        orc.csc_to_csr(m, k, Ap, Ai, Ax, Tp2, Ti2, Tx2)
        assert np.array_equal(Tp1, Tp2) and np.array_equal(Ti1, Ti2) and np.array_equal(Tx1, Tx2)
        A2 = sp.csc_matrix(sp.random(m, k, density=0.15, random_state=int(rng.integers(1 << 30))))
        A2p, A2i, A2x = A2.indptr.astype(np.int32), A2.indices.astype(np.int32), A2.data
        for fn, sign in ((R.ref_csc_plus_csc, 1.0), (R.ref_csc_minus_csc, -1.0)):
            cap = Ap[k] + A2p[k]
            Sp1 = np.empty(k + 1, dtype=np.int32); Si1 = np.empty(cap, dtype=np.int32); Sx1 = np.empty(cap)
            Sp2 = Sp1.copy(); Si2 = Si1.copy(); Sx2 = Sx1.copy()
            fn(m, k, vp(Ap), vp(Ai), vp(Ax), vp(A2p), vp(A2i), vp(A2x), vp(Sp1), vp(Si1), vp(Sx1))
            kk = orc.csc_plusminus_csc(m, k, Ap, Ai, Ax, A2p, A2i, A2x, sign, Sp2, Si2, Sx2)
            assert np.array_equal(Sp1, Sp2) and np.array_equal(Si1[:kk], Si2[:kk]) and np.array_equal(Sx1[:kk], Sx2[:kk])
