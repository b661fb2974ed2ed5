"""Host-side assembly glue and the CscMat container (no GPU): reference fixtures + scipy."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import SIX_BY_THREE
from csparse3_b200 import CscMat, Diag, Diags, scipy_to_mat
from csparse3_b200 import csc_b200 as B
from csparse3_b200 import synth


@pytest.mark.gpu
def test_slices_and_islands_run_on_the_device():
    """CscMat.__getitem__ (reference csc.py:150-292) and CscMat.islands (csc.py:515-521) through topo_kernels.cu."""
    g = SIX_BY_THREE
    A = CscMat(g["m"], g["n"], indptr=g["indptr"], indices=g["indices"], data=g["data"])
    S = sp.csc_matrix((g["data"], g["indices"], g["indptr"]), shape=(6, 3))
    assert (A[:, 1].todense() == S[:, [1]].toarray()).all()
    assert (A[:, [0, 2]].todense() == S[:, [0, 2]].toarray()).all()
    adj = sp.csc_matrix(np.array([[1, 1, 0], [1, 1, 0], [0, 0, 1.0]]))
    M = scipy_to_mat(adj)
    assert [i.tolist() for i in M.islands()] == [[0, 1], [2]]


def test_dense_diag_norm():
    g = SIX_BY_THREE
    A = CscMat(g["m"], g["n"], indptr=g["indptr"], indices=g["indices"], data=g["data"])
    S = sp.csc_matrix((g["data"], g["indices"], g["indptr"]), shape=(6, 3))
    assert (A.todense() == S.toarray()).all()
    assert (Diag(4, 4, 2.5).todense() == 2.5 * np.eye(4)).all()
    assert (Diags(np.array([1.0, 2.0, 3.0])).todense() == np.diag([1.0, 2.0, 3.0])).all()
    assert (A * 5).data.tolist() == (g["data"] * 5).tolist() and A.shape == (6, 3) and A.get_nnz() == 10
    assert A == A.copy() and not (A == (A * 2))
    assert B.csc_norm(3, g["indptr"], g["data"]) == 28.0


def test_grid_generator_is_consistent_with_dense_formula():
    """The vectorised Jacobian generator vs a direct dense evaluation of dS/dV (MATPOWER formulas)."""
    g = synth.GridCase(30, seed=5)
    V = g.voltages([123])[0]
    Y = np.zeros((30, 30), dtype=complex); Y[g.yi, g.yk] = g.ybus_values()
    I = Y @ V
    dVm = np.diag(V) @ np.conj(Y @ np.diag(V / abs(V))) + np.conj(np.diag(I)) @ np.diag(V / abs(V))
    dVa = 1j * np.diag(V) @ np.conj(np.diag(I) - Y @ np.diag(V))
    J = np.block([[dVa.real[np.ix_(g.pvpq, g.pvpq)], dVm.real[np.ix_(g.pvpq, g.pq)]],
                  [dVa.imag[np.ix_(g.pq, g.pvpq)], dVm.imag[np.ix_(g.pq, g.pq)]]])
    Jx = g.jacobian_values(V[None, :])[0]
    Js = sp.csc_matrix((Jx, g.Ai, g.Ap), shape=(g.n, g.n)).toarray()
    assert np.allclose(J, Js, rtol=1e-13, atol=1e-13)
    # outages keep the pattern and only change values; non-bridge outages stay non-singular
    Ax, b = g.outage_batch(0, 4)
    assert Ax.shape == (4, g.nnz) and np.isfinite(Ax).all()
    assert all(np.linalg.cond(sp.csc_matrix((Ax[k], g.Ai, g.Ap), shape=(g.n, g.n)).toarray()) < 1e8 for k in range(4))
