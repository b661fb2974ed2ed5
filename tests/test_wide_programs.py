"""CPU validation of the compiled "wide" device programs (csparse3_b200/csrc/wide_program.cpp, wide_solve.cpp).

tests/wide_interp.py executes the byte streams exactly as lu_wide.cu does (same slots, same L cache / landing areas,
cp.async modelled at its two extremes), so the host compilers are checked bit for bit against the oracle without
a GPU: every factor entry and every solution entry must equal oracle/csp3_oracle.c's.
"""
import os
import subprocess
import sys

import numpy as np
import pytest
import scipy.sparse as sp

from csparse3_b200 import synth
from csparse3_b200.lu import LuSymbolic
from oracle import oracle as orc

import wide_interp as wi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _check(n, Ap, Ai, Axb, bb, order=1, tol=1e-3):
    sym = LuSymbolic(n, Ap, Ai, Axb[0], order=order, tol=tol)
    if sym.wide_width == 0:
        pytest.skip("wide programs not available for this pattern")
    Lx, Ux, fail, stats = wi.run_refactor(sym, Axb)
    oL, oU, ox = [], [], []
    for k in range(Axb.shape[0]):
        L, U = orc.csc_lu_refactor(n, Ap, Ai, Axb[k], sym.q, sym.pinv, sym.Lp, sym.Li, sym.Up, sym.Ui)
        oL.append(L); oU.append(U)
        ox.append(orc.csc_lu_solve(n, sym.Lp, sym.Li, L, sym.Up, sym.Ui, U, sym.pinv, sym.q, bb[k]))
    assert (fail == 0).all()
    assert np.array_equal(Lx, np.array(oL)) and np.array_equal(Ux, np.array(oU))
    assert stats["ops"] * 2 == sym.flops                       # every update operation exactly once
    if wi.get_program(sym, 4)[0] is not None:
        assert np.array_equal(wi.run_solve(sym, Lx, Ux, bb), np.array(ox))
    return sym, stats


def test_wide_programs_grid118():
    g = synth.GridCase(118)
    n, Ap, Ai, Ax0 = g.base_jacobian()
    Axb, bb = g.jacobian_batch(0, 4)
    _check(n, Ap, Ai, Axb, bb)


def test_wide_programs_config3_pattern():
    g = synth.GridCase(2000)
    n, Ap, Ai, Ax0 = g.base_jacobian()
    Axb, bb = g.jacobian_batch(0, 2)
    sym, stats = _check(n, Ap, Ai, Axb, bb)
    # the bench configuration keeps every bundle of a 10,000-system batch resident: 9 one-warp CTAs per SM
    for which in (3, 4, 5):
        assert wi.get_program(sym, which)[1][6] <= (228 * 1024) // 9 - 1024


@pytest.mark.parametrize("order,tol", [(1, 1e-3), (2, 1.0), (0, 1.0), (3, 0.1)])
def test_wide_programs_small_matrices(order, tol):
    rng = np.random.default_rng(order)                 # same matrices as tests/test_gpu_parity.py::test_lu_small_matrices_bit_exact
    cases = [synth.laplacian_2d(9), synth.laplacian_3d(5), (1, np.array([0, 1], dtype=np.int32), np.array([0], dtype=np.int32), np.array([2.0]))]
    for t in range(5):
        n = int(rng.integers(2, 150))
        A = sp.csc_matrix(sp.random(n, n, density=min(1.0, 3.0 / n + 0.03), random_state=int(rng.integers(1 << 30)),
                                    format="csc") + sp.diags(rng.uniform(0.5, 2.0, n)))
        cases.append((n, A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.copy()))
    for n, Ap, Ai, Ax in cases:
        Axb = Ax[None, :] * rng.uniform(0.9, 1.1, (3, len(Ax)))
        Axb[0] = Ax
        _check(n, Ap, Ai, Axb, rng.standard_normal((3, n)), order=order, tol=tol)


def test_wide_refactor_reports_bad_pivots():
    # 3 x 3 with a dense last column: a zero / NaN pivot in the LAST column poisons nothing downstream
    Ap = np.array([0, 1, 2, 5], dtype=np.int32); Ai = np.array([0, 1, 0, 1, 2], dtype=np.int32)
    Ax = np.array([2.0, 3.0, 1.0, 1.0, 4.0])
    sym = LuSymbolic(3, Ap, Ai, Ax, order=0, tol=1.0)
    if sym.wide_width == 0:
        pytest.skip("wide programs not available")
    Axb = np.tile(Ax, (3, 1))
    Axb[1, 4] = 0.0
    Axb[2, 4] = np.inf
    _, Ux, fail, _ = wi.run_refactor(sym, Axb)
    assert fail.tolist() == [0, 3, 3]
    for k in (1, 2):
        with pytest.raises(ArithmeticError) as e:
            orc.csc_lu_refactor(3, Ap, Ai, Axb[k], sym.q, sym.pinv, sym.Lp, sym.Li, sym.Up, sym.Ui)
        assert int(str(e.value).split()[-1]) + 1 == fail[k]


@pytest.mark.parametrize("env", [{"CSP3_WIDE_S": "16"}, {"CSP3_WIDE_S": "4"}, {"CSP3_WIDE_S": "16", "CSP3_WIDE_LANE": "4"},
                                 {"CSP3_WIDE_F": "32"}, {"CSP3_WIDE_R": "24", "CSP3_WIDE_F": "64"},
                                 {"CSP3_WIDE_SCHED": "0"}, {"CSP3_WIDE_SCHED": "0", "CSP3_WIDE_F": "32"},
                                 {"CSP3_WIDE_PAIRS": "1"}, {"CSP3_WIDE_PAIRS": "16", "CSP3_WIDE_RUN": "64"},
                                 {"CSP3_WIDE_GA": "16"}])
def test_wide_programs_other_geometries(env):
    """Bundle widths, 4 systems per lane, tiny landing area (immediate fetches) and tiny L cache; the in-order chunk
    packer (the list scheduler's fallback), pair windows, long landing runs, small groups: the knobs are read once
    per process, so each setting runs in a child process."""
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r); import numpy as np; "
            "import test_wide_programs as t; from csparse3_b200 import synth; g = synth.GridCase(118); "
            "n, Ap, Ai, Ax0 = g.base_jacobian(); Axb, bb = g.jacobian_batch(0, 3); t._check(n, Ap, Ai, Axb, bb); print('ok')"
            ) % (ROOT, os.path.join(ROOT, "tests"))
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env), capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-2000:]
