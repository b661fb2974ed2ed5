"""Oracle half (2): the CSparse restatement.  PARITY UNPINNED (no reference LU exists); checked through
algebraic identities and an independent solver (scipy SuperLU)."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from csparse3_b200 import synth
from oracle import oracle as orc


def _cases():
    out = [("lap2d_20", *synth.laplacian_2d(20)), ("lap3d_6", *synth.laplacian_3d(6))]
    g = synth.GridCase(118)
    out.append(("grid118", *g.base_jacobian()))
    rng = np.random.default_rng(11)
    for t in range(4):
        n = int(rng.integers(5, 90))
        A = sp.random(n, n, density=0.08, random_state=int(rng.integers(1 << 30)), format="csc") + \
            sp.diags(rng.uniform(0.5, 2.0, n))
        A = sp.csc_matrix(A)
        out.append(("rand%d" % n, n, A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.copy()))
    return out


@pytest.mark.parametrize("order,tol", [(0, 1.0), (1, 1e-3), (2, 1.0), (3, 0.1)])
def test_lu_identities(order, tol):
    for name, n, Ap, Ai, Ax in _cases():
        q = orc.csc_amd(order, n, n, Ap, Ai)
        assert sorted(q.tolist()) == list(range(n)), name
        Lp, Li, Lx, Up, Ui, Ux, pinv = orc.csc_lu(n, Ap, Ai, Ax, q, tol)
        assert sorted(pinv.tolist()) == list(range(n))
        A = sp.csc_matrix((Ax, Ai, Ap), shape=(n, n))
        L = sp.csc_matrix((Lx, Li, Lp), shape=(n, n)); U = sp.csc_matrix((Ux, Ui, Up), shape=(n, n))
        # P A Q = L U : row i of A is row pinv[i] of PAQ, column k of PAQ is column q[k] of A
        PAQ = sp.csc_matrix((Ax, pinv[Ai], Ap), shape=(n, n))[:, q]
        assert abs(PAQ - L @ U).max() <= 1e-11 * max(1.0, abs(A).max()), name
        # layout: L unit diagonal first, U diagonal last, strictly triangular otherwise
        for k in range(n):
            assert Li[Lp[k]] == k and Lx[Lp[k]] == 1.0 and (Li[Lp[k] + 1:Lp[k + 1]] > k).all()
            assert Ui[Up[k + 1] - 1] == k and (Ui[Up[k]:Up[k + 1] - 1] < k).all()
        # refactor reproduces the first factorisation bit for bit
        Lx2, Ux2 = orc.csc_lu_refactor(n, Ap, Ai, Ax, q, pinv, Lp, Li, Up, Ui)
        assert np.array_equal(Lx, Lx2) and np.array_equal(Ux, Ux2), name
        # solve vs SuperLU
        b = np.random.default_rng(1).standard_normal(n)
        x = orc.csc_lu_solve(n, Lp, Li, Lx, Up, Ui, Ux, pinv, q, b)
        x_ref = spla.splu(A).solve(b)
        assert np.linalg.norm(x - x_ref) <= 1e-9 * np.linalg.norm(x_ref), name
        assert np.linalg.norm(A @ x - b) <= 1e-10 * np.linalg.norm(b), name
        assert np.array_equal(orc.csc_lusol(order, n, Ap, Ai, Ax, b, tol), x)


def test_levels_are_valid_schedules():
    g = synth.GridCase(300)
    n, Ap, Ai, Ax = g.base_jacobian()
    q = orc.csc_amd(1, n, n, Ap, Ai)
    Lp, Li, Lx, Up, Ui, Ux, pinv = orc.csc_lu(n, Ap, Ai, Ax, q, 1e-3)
    for kind, (Gp, Gi) in enumerate(((Up, Ui), (Lp, Li), (Up, Ui))):
        level, order, lptr = orc.lu_levels(n, Gp, Gi, kind)
        assert sorted(order.tolist()) == list(range(n)) and lptr[0] == 0 and lptr[-1] == n
        for l in range(len(lptr) - 1):
            seg = order[lptr[l]:lptr[l + 1]]
            assert (level[seg] == l).all() and (np.diff(seg) > 0).all()
        for k in range(n):
            rows = Gi[Gp[k]:Gp[k + 1]]
            if kind == 0:
                deps = rows[rows < k]; assert (level[deps] < level[k]).all()
                assert level[k] == (level[deps].max() + 1 if len(deps) else 0)
            elif kind == 1:
                tgt = rows[rows > k]; assert (level[tgt] > level[k]).all()
            else:
                tgt = rows[rows < k]; assert (level[tgt] > level[k]).all()


def test_etree_post():
    n, Ap, Ai, Ax = synth.laplacian_2d(12)
    parent = orc.csc_etree(n, n, Ap, Ai, False)
    assert ((parent > np.arange(n)) | (parent == -1)).all()
    post = orc.csc_post(n, parent)
    assert sorted(post.tolist()) == list(range(n))
    rank = np.empty(n, dtype=int); rank[post] = np.arange(n)
    has_parent = parent >= 0
    assert (rank[np.where(has_parent)[0]] < rank[parent[has_parent]]).all()     # children before parents
    parent_ata = orc.csc_etree(n, n, Ap, Ai, True)
    assert ((parent_ata > np.arange(n)) | (parent_ata == -1)).all()


def test_singular_reports_step():
    Ap = np.array([0, 1, 2, 2], dtype=np.int32); Ai = np.array([0, 1], dtype=np.int32); Ax = np.array([1.0, 2.0])
    with pytest.raises(ArithmeticError):
        orc.csc_lu(3, Ap, Ai, Ax, None, 1.0)


def test_config_sizes_match_survey():
    """SURVEY.md section 8d quotes these sizes for the seeded generators."""
    g = synth.GridCase(118); assert (g.n, g.nnz, g.n_branch) == (211, 1393, 162)
    n, Ap, Ai, Ax = synth.laplacian_2d(100); assert (n, Ap[n]) == (10000, 49600)
