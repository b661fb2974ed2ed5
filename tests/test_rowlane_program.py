"""CPU validation of the compiled "row-lane" refactor program (csparse3_b200/csrc/rowlane_program.cpp): the
interpreter of tests/rowlane_interp.py (which models the kernel's operand look-ahead) must reproduce the oracle's
factors bit for bit, with every multiply-subtract executed once."""
import os
import subprocess
import sys

import numpy as np
import pytest
import scipy.sparse as sp

from csparse3_b200 import synth
from csparse3_b200.lu import LuSymbolic
from oracle import oracle as orc

import rowlane_interp as ri

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _check(n, Ap, Ai, Axb, order=1, tol=1e-3):
    sym = LuSymbolic(n, Ap, Ai, Axb[0], order=order, tol=tol)
    Lx, Ux, fail, stats = ri.run_refactor(sym, Axb)
    oL, oU = [], []
    for k in range(Axb.shape[0]):
        L, U = orc.csc_lu_refactor(n, Ap, Ai, Axb[k], sym.q, sym.pinv, sym.Lp, sym.Li, sym.Up, sym.Ui)
        oL.append(L); oU.append(U)
    assert (fail == 0).all()
    assert np.array_equal(Ux, np.array(oU)) and np.array_equal(Lx, np.array(oL))
    assert stats["ops"] * 2 == sym.flops                       # every update operation exactly once
    return sym, stats


def test_rowlane_program_grid118():
    g = synth.GridCase(118)
    n, Ap, Ai, Ax0 = g.base_jacobian()
    Axb, _ = g.jacobian_batch(0, 4)
    _check(n, Ap, Ai, Axb)


def test_rowlane_program_config3_pattern():
    g = synth.GridCase(2000)
    n, Ap, Ai, Ax0 = g.base_jacobian()
    Axb, _ = g.jacobian_batch(0, 2)
    sym, stats = _check(n, Ap, Ai, Axb)
    # the column order keeps late operand reads rare and the accumulator is a few KB per bundle
    assert stats["late_quads"] <= 0.05 * stats["update_quads"]
    assert ri.get_program(sym)[1][6] <= 11 * 1024


@pytest.mark.parametrize("order,tol", [(1, 1e-3), (2, 1.0), (0, 1.0), (3, 0.1)])
def test_rowlane_program_small_matrices(order, tol):
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


def test_rowlane_program_reports_bad_pivots():
    Ap = np.array([0, 1, 2, 5], dtype=np.int32); Ai = np.array([0, 1, 0, 1, 2], dtype=np.int32)
    Ax = np.array([2.0, 3.0, 1.0, 1.0, 4.0])
    sym = LuSymbolic(3, Ap, Ai, Ax, order=0, tol=1.0)
    Axb = np.tile(Ax, (3, 1))
    Axb[1, 4] = 0.0
    Axb[2, 4] = np.inf
    _, _, fail, _ = ri.run_refactor(sym, Axb)
    assert fail.tolist() == [0, 3, 3]


@pytest.mark.parametrize("env", [{"CSP3_RL_WINDOW": "1"}, {"CSP3_RL_WINDOW": "4"}, {"CSP3_RL_W": "2", "CSP3_RL_NQ": "1"},
                                 {"CSP3_RL_W": "4", "CSP3_RL_NQ": "2"}, {"CSP3_RL_W": "8"}, {"CSP3_RL_W": "2", "CSP3_RL_MARGIN": "-100000"}])
def test_rowlane_program_other_geometries(env):
    """Other look-ahead depths / the natural column order (every chain reads late): knobs are read per process."""
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r); import numpy as np; "
            "import test_rowlane_program as t; from csparse3_b200 import synth; g = synth.GridCase(118); "
            "n, Ap, Ai, Ax0 = g.base_jacobian(); Axb, bb = g.jacobian_batch(0, 3); t._check(n, Ap, Ai, Axb); "
            "t.test_rowlane_program_small_matrices(1, 1e-3); print('ok')"
            ) % (ROOT, os.path.join(ROOT, "tests"))
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env), capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-2000:]


def test_rowlane_program_laplacians_many_warps():
    """Patterns with wide elimination trees and long columns (2-D / 3-D Laplacians): several warps per bundle, different
    interleavings of the warps, overflow quads (more than 8 entries of a role per column)."""
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r); import numpy as np; "
            "import rowlane_interp as ri; from csparse3_b200 import synth; from csparse3_b200.lu import LuSymbolic; "
            "from oracle import oracle as orc\n"
            "for n, Ap, Ai, Ax in (synth.laplacian_2d(24), synth.laplacian_3d(7)):\n"
            "    rng = np.random.default_rng(3); Axb = Ax[None, :] * rng.uniform(0.9, 1.1, (2, len(Ax)))\n"
            "    sym = LuSymbolic(n, Ap, Ai, Axb[0])\n"
            "    for seed in (0, 1):\n"
            "        Lx, Ux, fail, stats = ri.run_refactor(sym, Axb, seed=seed)\n"
            "        for k in range(2):\n"
            "            L, U = orc.csc_lu_refactor(n, Ap, Ai, Axb[k], sym.q, sym.pinv, sym.Lp, sym.Li, sym.Up, sym.Ui)\n"
            "            assert np.array_equal(Lx[k], L) and np.array_equal(Ux[k], U)\n"
            "        assert (fail == 0).all() and stats['ops'] * 2 == sym.flops\n"
            "print('ok')") % (ROOT, os.path.join(ROOT, "tests"))
    for env in ({"CSP3_RL_W": "8", "CSP3_RL_NQ": "1"}, {"CSP3_RL_W": "4", "CSP3_RL_NQ": "2"}):
        out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env), capture_output=True, text=True)
        assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-2000:]


def _structured_cases():
    rng = np.random.default_rng(7)
    cases = []
    n = 60; cases.append(sp.csc_matrix(rng.uniform(-1, 1, (n, n)) + np.eye(n) * n))                      # dense: overflow quads
    n = 100; A = sp.lil_matrix((n, n)); A.setdiag(4.0); A[0, :] = 1.0; A[:, 0] = 1.0; A[n - 1, :] = 0.5; A[:, n - 1] = 0.5
    A[n - 1, n - 1] = 9.0; cases.append(sp.csc_matrix(A))                                                 # arrow: fills in completely with order 0
    n = 120; cases.append(sp.csc_matrix(sp.diags([rng.uniform(0.1, 1, n - abs(k)) for k in range(-9, 10)], list(range(-9, 10))) + sp.eye(n) * 30))
    out = []
    for A in cases:
        A.sort_indices()
        out.append((A.shape[0], A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.copy()))
    return out


def test_rowlane_and_rowsweep_structured_patterns():
    """Dense, arrow and banded matrices (columns with more than 8 entries of every role: STOREL4 / STOREU4 / LOAD4 quads)
    through the row-lane refactor program (one warp per bundle here; 8 warps in the child process below) and the
    row-sweep programs: bit-identical to the oracle."""
    import rowsweep_interp as rs
    rng = np.random.default_rng(1)
    for n, Ap, Ai, Ax in _structured_cases():
        for order in (0, 1):
            Axb = Ax[None, :] * rng.uniform(0.95, 1.05, (2, len(Ax)))
            sym, stats = _check(n, Ap, Ai, Axb, order=order)
            Lx, Ux, _, _ = ri.run_refactor(sym, Axb)
            b = rng.standard_normal((2, n))
            x = rs.solve(sym, Lx, Ux, b)
            for k in range(2):
                assert np.array_equal(x[k], orc.csc_lu_solve(n, sym.Lp, sym.Li, Lx[k], sym.Up, sym.Ui, Ux[k], sym.pinv, sym.q, b[k]))
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r); import numpy as np; import test_rowlane_program as t; "
            "rng = np.random.default_rng(2)\n"
            "for n, Ap, Ai, Ax in t._structured_cases():\n"
            "    for order in (0, 1): t._check(n, Ap, Ai, Ax[None, :] * rng.uniform(0.95, 1.05, (2, len(Ax))), order=order)\n"
            "print('ok')") % (ROOT, os.path.join(ROOT, "tests"))
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, CSP3_RL_W="8", CSP3_RL_NQ="1"), capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-2000:]
