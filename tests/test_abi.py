"""The C-ABI library loads, exports every symbol include/csparse3_b200.h declares, and its host-side
(symbolic) entry points agree bit for bit with the oracle.  No CUDA compute is launched here."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import ROOT
from csparse3_b200 import _lib, synth
from csparse3_b200 import csc_b200 as B
from csparse3_b200.lu import LuSymbolic
from oracle import oracle as orc


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "csparse3_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(csp3_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    names = _declared_symbols()
    assert len(names) >= 30
    for name in names:
        assert hasattr(L, name), "not exported: " + name
        assert name in _lib.SIGNATURES, "not bound in _lib.SIGNATURES: " + name
    assert set(_lib.SIGNATURES) == set(names)
    assert L.csp3_version() >= 100


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    n, Ap, Ai, Ax = synth.laplacian_2d(4)
    with pytest.raises(_lib.Csp3Error, match="no CPU fallback"):
        B.csc_mat_vec_ff(n, n, Ap, Ai, Ax, np.ones(n))
    with pytest.raises(_lib.Csp3Error, match="no CPU fallback"):
        B.csc_multiply_ff(n, n, Ap, Ai, Ax, n, n, Ap, Ai, Ax)
    with pytest.raises(_lib.Csp3Error, match="no CPU fallback"):
        B.csc_lusol(1, n, Ap, Ai, Ax, np.ones(n), 1.0)
    sym = LuSymbolic(n, Ap, Ai, Ax)
    with pytest.raises(_lib.Csp3Error, match="no CPU fallback"):
        sym.refactor_solve_host(Ax[None, :].copy(), np.ones((1, n)))
    # the kernel-selection introspection needs an uploaded schedule: a NULL name and an error code, no crash
    L = _lib.lib()
    assert L.csp3_lu_refactor_kernel_name(sym._h, 8) is None
    assert L.csp3_lu_prepare(sym._h, 8) != 0
    assert b"not uploaded" in L.csp3_last_error_string()


def test_dtype_strictness_like_reference():
    n, Ap, Ai, Ax = synth.laplacian_2d(4)
    with pytest.raises(TypeError, match="No matching definition"):
        B.csc_mat_vec_ff(n, n, Ap.astype(np.int64), Ai, Ax, np.ones(n))
    with pytest.raises(TypeError, match="No matching definition"):
        B.csc_transpose(n, n, Ap, Ai, Ax.astype(np.float32))


@pytest.mark.parametrize("order,tol", [(1, 1e-3), (2, 1.0), (3, 0.1), (0, 1.0)])
def test_host_symbolic_bit_exact_vs_oracle(order, tol):
    cases = [synth.laplacian_2d(24), synth.laplacian_3d(7), synth.GridCase(118).base_jacobian(),
             synth.GridCase(400, seed=3).base_jacobian()]
    rng = np.random.default_rng(order)
    for t in range(6):
        n = int(rng.integers(1, 100))
        A = sp.csc_matrix(sp.random(n, n, density=min(1.0, 3.0 / n + 0.03), random_state=int(rng.integers(1 << 30)),
                                    format="csc") + sp.diags(rng.uniform(0.2, 2.0, n)))
        cases.append((n, A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.copy()))
    for n, Ap, Ai, Ax in cases:
        sym = LuSymbolic(n, Ap, Ai, Ax, order=order, tol=tol)
        q = orc.csc_amd(order, n, n, Ap, Ai)
        Lp, Li, Lx, Up, Ui, Ux, pinv = orc.csc_lu(n, Ap, Ai, Ax, q, tol)
        assert np.array_equal(B.csc_amd(order, n, n, Ap, Ai), q)
        for a, b in ((q, sym.q), (pinv, sym.pinv), (Lp, sym.Lp), (Li, sym.Li), (Up, sym.Up), (Ui, sym.Ui)):
            assert np.array_equal(a, b)
        assert np.array_equal(Lx, sym.Lx0) and np.array_equal(Ux, sym.Ux0)
        for kind, (Gp, Gi) in enumerate(((Up, Ui), (Lp, Li), (Up, Ui))):
            for a, b in zip(orc.lu_levels(n, Gp, Gi, kind), sym.levels(kind)):
                assert np.array_equal(a, b)
        assert sym.flops == orc.lu_refactor_flops(n, Lp, Up, Ui)
        assert sym.nnz_lu == Lp[n] + Up[n] - n
        # rebuilding from the cached pattern gives the same schedule sizes
        sym2 = LuSymbolic.from_pattern(n, Ap, Ai, sym.q, sym.pinv, sym.Lp, sym.Li, sym.Up, sym.Ui)
        assert (sym2.flops, sym2.nlev_refactor, sym2.max_col_len) == (sym.flops, sym.nlev_refactor, sym.max_col_len)
        par = orc.csc_etree(n, n, Ap, Ai, False)
        assert np.array_equal(B.csc_etree(n, n, Ap, Ai, False), par)
        assert np.array_equal(B.csc_post(n, par), orc.csc_post(n, par))
        assert np.array_equal(B.csc_etree(n, n, Ap, Ai, True), orc.csc_etree(n, n, Ap, Ai, True))


def test_analyze_reports_singular_step():
    Ap = np.array([0, 1, 2, 2], dtype=np.int32); Ai = np.array([0, 1], dtype=np.int32); Ax = np.array([1.0, 2.0])
    with pytest.raises(ArithmeticError):
        LuSymbolic(3, Ap, Ai, Ax, order=0, tol=1.0)


def test_bytes_model():
    """8*nnzA + 16*nnzLU + 16*n per system (two-call path)."""
    g = synth.GridCase(118)
    n, Ap, Ai, Ax = g.base_jacobian()
    sym = LuSymbolic(n, Ap, Ai, Ax)
    assert sym.bytes_per_system() == 8 * sym.nnz + 16 * sym.nnz_lu + 16 * n
    assert sym.bytes_per_system(fused=True) == 8 * sym.nnz + 8 * sym.nnz_lu + 16 * n
