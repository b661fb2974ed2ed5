"""Generate tests/golden/*.npz by running the REFERENCE itself in the build container.

    python tests/golden/make_golden.py          (needs /root/reference; not available on the GPU box)

Imports the reference's numba kernels (src/CSparse3/csc_numba.py) unmodified and records their outputs on
seeded inputs, plus scipy's answers for the operators the reference's tests compare against
(src/test/test1_operations.py:13-61).  The fixtures are committed; the tests never read /root/reference.
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, "/root/reference/src")
HERE = os.path.dirname(os.path.abspath(__file__))


def ref_kernels():
    import warnings
    warnings.filterwarnings("ignore")
    import CSparse3.csc_numba as K      # eager numba compile, ~10 s
    return K


def case_test1():
    """The matrices of src/test/test1_operations.py:13-20 (np.random.seed(0), 100x100, density 0.01 + I)."""
    np.random.seed(0)
    k = 100
    A = sp.csc_matrix(sp.random(k, k, density=0.01)) + sp.diags(np.ones(k))
    B = sp.csc_matrix(sp.random(k, k, density=0.01)) + sp.diags(np.ones(k))
    x = np.random.random(k)
    xx = np.random.random((k, 5))
    A, B = sp.csc_matrix(A), sp.csc_matrix(B)
    return A, B, x, xx


def pack(prefix, M):
    return {prefix + "p": M.indptr.astype(np.int32), prefix + "i": M.indices.astype(np.int32),
            prefix + "x": M.data.astype(np.float64), prefix + "shape": np.array(M.shape)}


def main():
    K = ref_kernels()
    out = {}
    # ---- test1_operations matrices: reference numba kernels + scipy operators ----
    A, B, x, xx = case_test1()
    d = {}
    d.update(pack("A", A)); d.update(pack("B", B)); d["x"] = x; d["xx"] = xx
    m, n = A.shape
    Ap, Ai, Ax = d["Ap"], d["Ai"], d["Ax"]
    Bp, Bi, Bx = d["Bp"], d["Bi"], d["Bx"]
    d["ref_matvec"] = K.csc_mat_vec_ff(m, n, Ap, Ai, Ax, x)
    Cm, Cn, Cp, Ci, Cx, nnz = K.csc_multiply_ff(m, n, Ap, Ai, Ax, m, n, Bp, Bi, Bx)
    d["ref_mul_p"], d["ref_mul_i"], d["ref_mul_x"] = Cp, np.array(Ci), np.array(Cx)
    Tm, Tn, Tp, Ti, Tx = K.csc_transpose(m, n, Ap, Ai, Ax)
    d["ref_t_p"], d["ref_t_i"], d["ref_t_x"] = Tp, Ti, Tx
    Rp = np.zeros(m + 1, dtype=np.int32); Ri = np.empty(Ap[n], dtype=np.int32); Rx = np.empty(Ap[n])
    K.csc_to_csr(m, n, Ap, Ai, Ax, Rp, Ri, Rx)
    d["ref_csr_p"], d["ref_csr_i"], d["ref_csr_x"] = Rp, Ri, Rx
    d["scipy_AB"] = (A @ B).toarray(); d["scipy_Ax"] = A @ x; d["scipy_Axx"] = A @ xx
    d["scipy_AT"] = A.T.toarray(); d["scipy_A5"] = (A * 5).toarray()
    d["scipy_ApB"] = (A + B).toarray(); d["scipy_AmB"] = (A - B).toarray()
    np.savez_compressed(os.path.join(HERE, "test1_operations.npz"), **d)

    # ---- rectangular, unsorted-index and duplicate cases through the reference kernels ----
    rng = np.random.default_rng(42)
    d = {}
    R = sp.csc_matrix(sp.random(37, 53, density=0.08, random_state=7))
    S_ = sp.csc_matrix(sp.random(53, 53, density=0.06, random_state=8))      # Am <= Bn keeps the reference in bounds
    d.update(pack("R", R)); d.update(pack("S", S_))
    xr = rng.standard_normal(53)
    d["xr"] = xr
    d["ref_R_matvec"] = K.csc_mat_vec_ff(37, 53, d["Rp"], d["Ri"], d["Rx"], xr)
    Cm, Cn, Cp, Ci, Cx, nnz = K.csc_multiply_ff(37, 53, d["Rp"], d["Ri"], d["Rx"], 53, 53, d["Sp"], d["Si"], d["Sx"])
    d["ref_RS_p"], d["ref_RS_i"], d["ref_RS_x"] = Cp, np.array(Ci), np.array(Cx)
    Tm, Tn, Tp, Ti, Tx = K.csc_transpose(37, 53, d["Rp"], d["Ri"], d["Rx"])
    d["ref_Rt_p"], d["ref_Rt_i"], d["ref_Rt_x"] = Tp, Ti, Tx
    # the reference's own product is first-touch ordered: transpose twice to get unsorted-but-valid input
    Cm, Cn, Up, Ui, Ux, nnz = K.csc_multiply_ff(53, 53, d["Sp"], d["Si"], d["Sx"], 53, 53, d["Sp"], d["Si"], d["Sx"])
    d["U_p"], d["U_i"], d["U_x"] = Up, np.array(Ui), np.array(Ux)                # unsorted row indices
    d["ref_U_matvec"] = K.csc_mat_vec_ff(53, 53, Up, np.array(Ui), np.array(Ux), xr)
    Tm, Tn, Tp, Ti, Tx = K.csc_transpose(53, 53, Up, np.array(Ui), np.array(Ux))
    d["ref_Ut_p"], d["ref_Ut_i"], d["ref_Ut_x"] = Tp, Ti, Tx
    # stacking (test_matrix_stacking.py, small and seeded here)
    k = 20
    Q = [sp.csc_matrix(sp.random(*s, density=0.2, random_state=10 + t)) for t, s in
         enumerate([(k, 4 * k), (k, k), (6 * k, 4 * k), (6 * k, k)])]
    for name, M in zip("abcd", Q):
        d.update(pack("st_" + name, M))
    mm, nn, Pi, Pp, Px = K.csc_stack_4_by_4_ff(
        Q[0].shape[0], Q[0].shape[1], d["st_ai"], d["st_ap"], d["st_ax"],
        Q[1].shape[0], Q[1].shape[1], d["st_bi"], d["st_bp"], d["st_bx"],
        Q[2].shape[0], Q[2].shape[1], d["st_ci"], d["st_cp"], d["st_cx"],
        Q[3].shape[0], Q[3].shape[1], d["st_di"], d["st_dp"], d["st_dx"])
    d["ref_st_i"], d["ref_st_p"], d["ref_st_x"], d["ref_st_shape"] = Pi, Pp, Px, np.array([mm, nn])
    np.savez_compressed(os.path.join(HERE, "reference_kernels.npz"), **d)
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
