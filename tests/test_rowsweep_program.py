"""CPU validation of the row-sweep programs (csparse3_b200/csrc/rowsweep_program.cpp): the interpreter of
tests/rowsweep_interp.py must reproduce the oracle's triangular solves bit for bit."""
import numpy as np
import pytest
import scipy.sparse as sp

from csparse3_b200 import synth
from csparse3_b200.lu import LuSymbolic
from oracle import oracle as orc

import rowsweep_interp as rs


def _check(n, Ap, Ai, Axb, order=1, tol=1e-3, seed=0):
    sym = LuSymbolic(n, Ap, Ai, Axb[0], order=order, tol=tol)
    rng = np.random.default_rng(seed)
    b = rng.standard_normal((Axb.shape[0], n))
    Lx, Ux, xo = [], [], []
    for k in range(Axb.shape[0]):
        L, U = orc.csc_lu_refactor(n, Ap, Ai, Axb[k], sym.q, sym.pinv, sym.Lp, sym.Li, sym.Up, sym.Ui)
        Lx.append(L); Ux.append(U)
        xo.append(orc.csc_lu_solve(n, sym.Lp, sym.Li, L, sym.Up, sym.Ui, U, sym.pinv, sym.q, b[k]))
    for s in (0, 1):
        x = rs.solve(sym, np.array(Lx), np.array(Ux), b, seed=s)
        assert np.array_equal(x, np.array(xo))
    return sym


def test_rowsweep_grid118_and_config3_pattern():
    for nb, cnt in ((118, 3), (2000, 2)):
        g = synth.GridCase(nb)
        n, Ap, Ai, Ax0 = g.base_jacobian()
        Axb, _ = g.jacobian_batch(0, cnt)
        sym = _check(n, Ap, Ai, Axb)
        geo = rs.get_program(sym, 9)[1]
        assert geo[0] == sym.levels(1)[2].shape[0] - 1          # one barrier per level of the L-solve graph


@pytest.mark.parametrize("order,tol", [(1, 1e-3), (2, 1.0), (0, 1.0), (3, 0.1)])
def test_rowsweep_small_matrices(order, tol):
    rng = np.random.default_rng(order)
    cases = [synth.laplacian_2d(9), synth.laplacian_3d(5), (1, np.array([0, 1], dtype=np.int32), np.array([0], dtype=np.int32), np.array([2.0]))]
    for t in range(5):
        n = int(rng.integers(2, 150))
        A = sp.csc_matrix(sp.random(n, n, density=min(1.0, 3.0 / n + 0.03), random_state=int(rng.integers(1 << 30)),
                                    format="csc") + sp.diags(rng.uniform(0.5, 2.0, n)))
        cases.append((n, A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.copy()))
    for n, Ap, Ai, Ax in cases:
        Axb = Ax[None, :] * rng.uniform(0.9, 1.1, (3, len(Ax)))
        Axb[0] = Ax
        _check(n, Ap, Ai, Axb, order=order, tol=tol)
