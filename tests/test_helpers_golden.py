"""Round-2 fixtures recorded from the reference's own numba kernels (tests/golden/make_golden.py helpers):
csc_add_ff (alpha / beta form, csc_numba.py:183-219), the cs_* helpers (csc_cumsum_i :75-94, csc_spalloc_f :46-60,
csc_sprealloc_f :97-122, csc_scatter_f :125-151), find_islands (:743-808) and the sub-matrix kernels (:463-578).
CPU part: the oracle and the host-side helpers of the drop-in module reproduce them bit for bit.  GPU part: so does
the device path of csc_add_ff."""
import numpy as np
import pytest

from csparse3_b200 import csc_b200 as B
from oracle import oracle as orc

ADD_CASES = {"add1": ("S", "T", 2.5, -0.75), "add2": ("D", "T", 1.0, 1.0), "add3": ("S", "S", 1.0, -1.0)}


def _mat(d, k):
    return d[k + "p"], d[k + "i"], d[k + "x"]


def test_oracle_csc_add_ff_matches_reference(golden_helpers):
    d = golden_helpers
    for name, (a, b, al, be) in ADD_CASES.items():
        Ap, Ai, Ax = _mat(d, a); Bp, Bi, Bx = _mat(d, b)
        Cm, Cn, Cp, Ci, Cx = orc.csc_add_ff(53, 53, Ap, Ai, Ax, 53, 53, Bp, Bi, Bx, al, be)
        nz = int(d[name + "_p"][-1])
        assert np.array_equal(Cp, d[name + "_p"]) and np.array_equal(Ci[:nz], d[name + "_i"][:nz]) and np.array_equal(Cx[:nz], d[name + "_x"][:nz])
    assert (d["add3_x"][:int(d["add3_p"][-1])] == 0).all()          # explicit zeros are kept by this kernel


def test_cs_helpers_match_reference(golden_helpers):
    d = golden_helpers
    c = d["cumsum_in"].copy(); p = np.zeros(7, dtype=np.int32)
    assert B.csc_cumsum_i(p, c, 6) == int(d["cumsum_ret"][0])
    assert np.array_equal(p, d["cumsum_p"]) and np.array_equal(c, d["cumsum_c"])
    m, n, Pp, Pi, Px, nzmax = B.csc_spalloc_f(4, 3, 0)
    assert [m, n, len(Pp), len(Pi), len(Px), nzmax] == d["spalloc"].tolist()
    assert Pp.dtype == np.int32 and Pi.dtype == np.int32 and Px.dtype == np.float64 and not Pp.any() and not Px.any()
    Sp, Si, Sx = _mat(d, "S")
    Ri, Rx, nz = B.csc_sprealloc_f(53, Sp, Si, Sx, 0)
    assert nz == int(d["sprealloc0_nz"][0]) and np.array_equal(Ri, d["sprealloc0_i"]) and np.array_equal(Rx, d["sprealloc0_x"])
    Ri, Rx, nz = B.csc_sprealloc_f(53, Sp, Si, Sx, 40)
    assert nz == 40 and np.array_equal(Ri, d["sprealloc40_i"]) and np.array_equal(Rx, d["sprealloc40_x"])
    w = np.zeros(53, dtype=np.int32); x = np.zeros(53); Ci = np.zeros(200, dtype=np.int32)
    nz = B.csc_scatter_f(d["Dp"], d["Di"], d["Dx"], 7, 2.0, w, x, 8, Ci, 0)
    nz = B.csc_scatter_ff(d["Tp"], d["Ti"], d["Tx"], 7, -1.5, w, x, 8, Ci, nz)
    assert nz == int(d["scatter_nz"][0]) and np.array_equal(w, d["scatter_w"]) and np.array_equal(x, d["scatter_x"]) and np.array_equal(Ci, d["scatter_ci"])


@pytest.mark.gpu
def test_topology_kernels_match_reference(golden_helpers):
    """find_islands and the three sub-matrix kernels on the device (topo_kernels.cu) against the reference's outputs,
    order of the nodes / entries included."""
    d = golden_helpers
    isl = B.find_islands(60, d["Gp"], d["Gi"])
    assert [len(i) for i in isl] == d["islands_len"].tolist()
    assert np.array_equal(np.concatenate([np.asarray(i) for i in isl]), d["islands_flat"])
    Sp, Si, Sx = _mat(d, "S")
    nS = int(Sp[-1])
    for name, res in {"sub": B.csc_sub_matrix(53, nS, Sp, Si, Sx, d["sub_rows"], d["sub_cols"]),
                      "subc": B.csc_sub_matrix_cols(53, nS, Sp, Si, Sx, d["sub_cols"]),
                      "subr": B.csc_sub_matrix_rows(53, nS, Sp, Si, Sx, d["sub_rows"])}.items():
        assert res[0] == int(d[name + "_n"][0]), name
        assert np.array_equal(res[1], d[name + "_p"]) and np.array_equal(res[2], d[name + "_i"]) and np.array_equal(res[3], d[name + "_x"]), name


@pytest.mark.gpu
def test_csc_add_ff_device_matches_reference(golden_helpers):
    d = golden_helpers
    for name, (a, b, al, be) in ADD_CASES.items():
        Ap, Ai, Ax = _mat(d, a); Bp, Bi, Bx = _mat(d, b)
        Cm, Cn, Cp, Ci, Cx = B.csc_add_ff(53, 53, Ap, Ai, Ax, 53, 53, Bp, Bi, Bx, al, be)
        assert (Cm, Cn) == (53, 53)
        assert np.array_equal(Cp, d[name + "_p"]) and np.array_equal(Ci, d[name + "_i"]) and np.array_equal(Cx, d[name + "_x"]), name
    # larger random case with unsorted columns and duplicates against the oracle
    rng = np.random.default_rng(1)
    n = 3000
    cnt = rng.integers(0, 9, n)
    Ap = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int32)
    Ai = rng.integers(0, n, Ap[-1]).astype(np.int32); Ax = rng.standard_normal(Ap[-1])
    cnt = rng.integers(0, 7, n)
    Bp = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int32)
    Bi = rng.integers(0, n, Bp[-1]).astype(np.int32); Bx = rng.standard_normal(Bp[-1])
    o = orc.csc_add_ff(n, n, Ap, Ai, Ax, n, n, Bp, Bi, Bx, -1.25, 3.0)
    r = B.csc_add_ff(n, n, Ap, Ai, Ax, n, n, Bp, Bi, Bx, -1.25, 3.0)
    nz = int(o[2][-1])
    assert np.array_equal(r[2], o[2]) and np.array_equal(r[3][:nz], o[3][:nz]) and np.array_equal(r[4][:nz], o[4][:nz])


@pytest.mark.gpu
def test_islands_batched_n_minus_1_vs_scipy_and_bridges():
    """Every branch of the 118-bus and the 10,000-bus synthetic grids removed in turn (one CTA per case): island counts
    and labels against scipy's connected components; an outage splits the grid exactly when the branch is a bridge."""
    import scipy.sparse as sp
    from scipy.sparse.csgraph import connected_components
    from csparse3_b200 import synth
    for nb, check_all in ((118, True), (10000, False)):
        case = synth.GridCase(nb)
        N, f, t = case.n_bus, case.f, case.t
        Adj = sp.coo_matrix((np.ones(2 * len(f)), (np.r_[f, t], np.r_[t, f])), shape=(N, N)).tocsc()
        Adj.sum_duplicates(); Adj.sort_indices()
        mult = np.asarray(Adj[f, t]).ravel()                      # parallel branches: removing one does not cut the pair
        of = np.where(mult > 1, -1, f).astype(np.int32); ot = np.where(mult > 1, -1, t).astype(np.int32)
        label, cnt = B.find_islands_batched(N, Adj.indptr.astype(np.int32), Adj.indices.astype(np.int32), of, ot)
        bridges = synth._bridges(N, f, t)
        assert np.array_equal(cnt > 1, bridges)
        assert (cnt[~bridges] == 1).all() and (label[~bridges] == 0).all()
        ids = np.arange(len(f)) if check_all else np.r_[np.where(bridges)[0][:40], np.where(~bridges)[0][:10]]
        for k in ids:
            A2 = Adj.tolil(copy=True)
            if of[k] >= 0:
                A2[f[k], t[k]] = 0; A2[t[k], f[k]] = 0
            ncomp, lab = connected_components(A2.tocsr(), directed=False)
            assert ncomp == cnt[k]
            # same partition; our label is the smallest node of the island
            first = np.full(ncomp, N); np.minimum.at(first, lab, np.arange(N))
            assert np.array_equal(label[k], first[lab])
