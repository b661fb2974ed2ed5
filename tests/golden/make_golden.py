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


def helpers():
    """Round 2: the reference's add kernel (alpha / beta form), its cs_* helpers and the topology kernels."""
    K = ref_kernels()
    d = {}
    S_ = sp.csc_matrix(sp.random(53, 53, density=0.06, random_state=8))
    T_ = sp.csc_matrix(sp.random(53, 53, density=0.05, random_state=9))
    d.update(pack("S", S_)); d.update(pack("T", T_))
    # unsorted input with duplicates: the column-wise concatenation [S; S] folded back onto 53 rows
    Dp = (2 * d["Sp"]).astype(np.int32)
    Di = np.concatenate([np.r_[d["Si"][a:b], d["Si"][a:b][::-1]] for a, b in zip(d["Sp"][:-1], d["Sp"][1:])]).astype(np.int32)
    Dx = np.concatenate([np.r_[d["Sx"][a:b], 0.5 * d["Sx"][a:b][::-1]] for a, b in zip(d["Sp"][:-1], d["Sp"][1:])])
    d["Dp"], d["Di"], d["Dx"] = Dp, Di, Dx
    for name, (Ap, Ai, Ax, Bp, Bi, Bx, al, be) in {
            "add1": (d["Sp"], d["Si"], d["Sx"], d["Tp"], d["Ti"], d["Tx"], 2.5, -0.75),
            "add2": (Dp, Di, Dx, d["Tp"], d["Ti"], d["Tx"], 1.0, 1.0),
            "add3": (d["Sp"], d["Si"], d["Sx"], d["Sp"], d["Si"], d["Sx"], 1.0, -1.0)}.items():      # exact zeros are kept
        Cm, Cn, Cp, Ci, Cx = K.csc_add_ff(53, 53, Ap, Ai, Ax, 53, 53, Bp, Bi, Bx, al, be)
        d[name + "_p"], d[name + "_i"], d[name + "_x"] = Cp, np.array(Ci), np.array(Cx)
    c = np.array([3, 0, 5, 1, 0, 7], dtype=np.int32); pcs = np.zeros(7, dtype=np.int32)
    d["cumsum_in"] = c.copy()
    d["cumsum_ret"] = np.array([K.csc_cumsum_i(pcs, c, 6)]); d["cumsum_p"], d["cumsum_c"] = pcs, c
    m, n, Pp, Pi, Px, nzmax = K.csc_spalloc_f(4, 3, 0)
    d["spalloc"] = np.array([m, n, len(Pp), len(Pi), len(Px), nzmax])
    Ri, Rx, rz = K.csc_sprealloc_f(53, d["Sp"], d["Si"], d["Sx"], 0)
    d["sprealloc0_i"], d["sprealloc0_x"], d["sprealloc0_nz"] = Ri, Rx, np.array([rz])
    Ri, Rx, rz = K.csc_sprealloc_f(53, d["Sp"], d["Si"], d["Sx"], 40)
    d["sprealloc40_i"], d["sprealloc40_x"] = Ri, Rx
    w = np.zeros(53, dtype=np.int32); x = np.zeros(53); Ci = np.zeros(200, dtype=np.int32)
    nz = K.csc_scatter_f(Dp, Di, Dx, 7, 2.0, w, x, 8, Ci, 0)
    nz = K.csc_scatter_f(d["Tp"], d["Ti"], d["Tx"], 7, -1.5, w, x, 8, Ci, nz)
    d["scatter_nz"], d["scatter_w"], d["scatter_x"], d["scatter_ci"] = np.array([nz]), w, x, Ci
    # topology: islands of a graph with several components (symmetric adjacency in CSC), sub-matrices
    rng = np.random.default_rng(5)
    nn = 60
    ei = rng.integers(0, nn, 45); ej = (ei + rng.integers(1, 4, 45)) % nn
    keep = (ei // 12) == (ej // 12)                                   # edges only inside blocks of 12 nodes -> >= 5 islands
    G = sp.coo_matrix((np.ones(keep.sum() * 2), (np.r_[ei[keep], ej[keep]], np.r_[ej[keep], ei[keep]])), shape=(nn, nn)).tocsc()
    G.sum_duplicates()
    d.update(pack("G", G))
    isl = K.find_islands(nn, d["Gp"], d["Gi"])
    d["islands_flat"] = np.array([v for isle in isl for v in isle], dtype=np.int64)
    d["islands_len"] = np.array([len(isle) for isle in isl], dtype=np.int64)
    rows = np.array([5, 2, 40, 41, 7, 3], dtype=np.int32); cols = np.array([9, 0, 33, 2], dtype=np.int32)
    d["sub_rows"], d["sub_cols"] = rows, cols
    nS = int(d["Sp"][-1])
    for name, res in {"sub": K.csc_sub_matrix(53, nS, d["Sp"], d["Si"], d["Sx"], rows, cols),
                      "subc": K.csc_sub_matrix_cols(53, nS, d["Sp"], d["Si"], d["Sx"], cols),
                      "subr": K.csc_sub_matrix_rows(53, nS, d["Sp"], d["Si"], d["Sx"], rows)}.items():
        d[name + "_n"], d[name + "_p"], d[name + "_i"], d[name + "_x"] = np.array([res[0]]), res[1], np.array(res[2]), np.array(res[3])
    np.savez_compressed(os.path.join(HERE, "reference_helpers.npz"), **d)
    print("wrote reference_helpers.npz")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "helpers":
        helpers()
    else:
        main()
        helpers()
