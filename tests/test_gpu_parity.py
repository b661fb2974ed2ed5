"""Parity proper: the CUDA path, called through the C ABI, against the oracle on the same seeded inputs.

Bit-exact (np.array_equal) for everything -- integer outputs AND floating-point values: the kernels follow
the oracle's operation order with unfused multiply/add, so no tolerance is needed.  The north-star
tolerances (solutions 1e-9 relative, residual 1e-10) are additionally asserted where they apply.
"""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import CONNECTIVITY, SIX_BY_THREE, sort_columns
from csparse3_b200 import CscMat, scipy_to_mat, synth
from csparse3_b200 import csc_b200 as B
from csparse3_b200.lu import LuSymbolic
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _rand_csc(rng, m, n, density):
    A = sp.csc_matrix(sp.random(m, n, density=density, random_state=int(rng.integers(1 << 30))))
    return A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.copy()


# ---- SpMV / SpMM --------------------------------------------------------------------------------------------
def test_reference_golden_vectors():
    g = CONNECTIVITY
    gen = CscMat(g["m"], g["n"], indptr=g["indptr"], indices=g["indices"], data=g["data"])
    assert ((gen.t() * g["p"]) == g["injections"]).all()              # docs/connectivity_matrix.rst:93-105
    s = SIX_BY_THREE
    A = CscMat(s["m"], s["n"], indptr=s["indptr"], indices=s["indices"], data=s["data"])
    Bp, Bi, Bx = A.to_csr()                                             # src/test/cscs_to_csr_test.py:13-31
    assert (Bp == s["csr_indptr"]).all() and (Bi == s["csr_indices"]).all() and (Bx == s["csr_data"]).all()


def test_test1_operations_replay(golden_test1):
    """src/test/test1_operations.py, every operator, with its exact-equality criterion."""
    d = golden_test1
    m, n = (int(v) for v in d["Ashape"])
    A2 = CscMat(m, n, indptr=d["Ap"], indices=d["Ai"], data=d["Ax"])
    B2 = CscMat(m, n, indptr=d["Bp"], indices=d["Bi"], data=d["Bx"])
    C2 = A2 + B2
    assert (C2.todense() == d["scipy_ApB"]).all()
    assert ((A2 - B2).todense() == d["scipy_AmB"]).all()
    assert np.array_equal(C2 * d["x"], sp.csc_matrix(d["scipy_ApB"]) @ d["x"])          # G = (A + B) * x
    assert ((A2 * B2).todense() == d["scipy_AB"]).all()
    assert (A2.dot(B2).todense() == d["scipy_AB"]).all()
    assert ((A2 * d["x"]) == d["scipy_Ax"]).all()
    assert ((A2 * d["xx"]) == d["scipy_Axx"]).all()
    assert ((A2 * 5).todense() == d["scipy_A5"]).all()
    assert (A2.t().todense() == d["scipy_AT"]).all()
    # the reference's numba kernels, bit for bit
    assert np.array_equal(B.csc_mat_vec_ff(m, n, d["Ap"], d["Ai"], d["Ax"], d["x"]), d["ref_matvec"])
    Tm, Tn, Tp, Ti, Tx = B.csc_transpose(m, n, d["Ap"], d["Ai"], d["Ax"])
    assert np.array_equal(Tp, d["ref_t_p"]) and np.array_equal(Ti, d["ref_t_i"]) and np.array_equal(Tx, d["ref_t_x"])
    Cm, Cn, Cp, Ci, Cx, nnz = B.csc_multiply_ff(m, n, d["Ap"], d["Ai"], d["Ax"], m, n, d["Bp"], d["Bi"], d["Bx"])
    ri, rx = sort_columns(n, d["ref_mul_p"], d["ref_mul_i"], d["ref_mul_x"])
    assert np.array_equal(Cp, d["ref_mul_p"]) and np.array_equal(Ci, ri) and np.array_equal(Cx, rx)


def test_reference_kernel_fixture(golden_ref):
    d = golden_ref
    assert np.array_equal(B.csc_mat_vec_ff(37, 53, d["Rp"], d["Ri"], d["Rx"], d["xr"]), d["ref_R_matvec"])
    assert np.array_equal(B.csc_mat_vec_ff(53, 53, d["U_p"], d["U_i"], d["U_x"], d["xr"]), d["ref_U_matvec"])   # unsorted
    for key, (m, n, p, i, x) in {"ref_Rt": (37, 53, d["Rp"], d["Ri"], d["Rx"]),
                                 "ref_Ut": (53, 53, d["U_p"], d["U_i"], d["U_x"])}.items():
        Tm, Tn, Tp, Ti, Tx = B.csc_transpose(m, n, p, i, x)
        assert np.array_equal(Tp, d[key + "_p"]) and np.array_equal(Ti, d[key + "_i"]) and np.array_equal(Tx, d[key + "_x"])
    Cm, Cn, Cp, Ci, Cx, nnz = B.csc_multiply_ff(37, 53, d["Rp"], d["Ri"], d["Rx"], 53, 53, d["Sp"], d["Si"], d["Sx"])
    ri, rx = sort_columns(53, d["ref_RS_p"], d["ref_RS_i"], d["ref_RS_x"])
    assert np.array_equal(Cp, d["ref_RS_p"]) and np.array_equal(Ci, ri) and np.array_equal(Cx, rx)


@pytest.mark.parametrize("seed", range(4))
def test_spmv_spmm_vs_oracle(seed):
    rng = np.random.default_rng(seed)
    for m, n, dens in ((1, 1, 1.0), (17, 5, 0.3), (300, 411, 0.02), (64, 64, 0.9), (2000, 1500, 0.004)):
        Ap, Ai, Ax = _rand_csc(rng, m, n, dens)
        x = rng.standard_normal(n)
        # rows averaging > 48 entries take the 8-lanes-per-row kernel (shuffle reduction): same result to
        # rounding; everything shorter follows the reference's summation order exactly
        same = np.array_equal if Ap[n] <= 48 * m else (lambda u, v: np.allclose(u, v, rtol=1e-13, atol=1e-13))
        assert same(B.csc_mat_vec_ff(m, n, Ap, Ai, Ax, x), orc.csc_mat_vec_ff(m, n, Ap, Ai, Ax, x))
        y0 = rng.standard_normal(m); y1 = y0.copy(); y2 = y0.copy()
        B.sptools.csc_matvec(m, n, Ap, Ai, Ax, x, y1); orc.csc_matvec(m, n, Ap, Ai, Ax, x, y2)
        assert same(y1, y2)
        X = rng.standard_normal((n, 3)); Y1 = rng.standard_normal((m, 3)); Y2 = Y1.copy()
        B.sptools.csc_matvecs(m, n, 3, Ap, Ai, Ax, X, Y1); orc.csc_matvecs(m, n, 3, Ap, Ai, Ax, X, Y2)
        assert np.array_equal(Y1, Y2)


def test_spmv_empty_and_config1():
    Ap = np.zeros(4, dtype=np.int32); Ai = np.zeros(0, dtype=np.int32); Ax = np.zeros(0)
    assert (B.csc_mat_vec_ff(5, 3, Ap, Ai, Ax, np.ones(3)) == 0).all()
    n, Ap, Ai, Ax = synth.laplacian_2d(100)                             # config 1 matrix
    x = np.random.default_rng(0).standard_normal(n)
    assert np.array_equal(B.csc_mat_vec_ff(n, n, Ap, Ai, Ax, x), orc.csc_mat_vec_ff(n, n, Ap, Ai, Ax, x))


def test_spmv_plan_batched_device():
    import torch
    from csparse3_b200.spmv import SpmvPlan
    g = synth.GridCase(118)
    Axb, b = g.jacobian_batch(0, 33)
    plan = SpmvPlan(g.n, g.n, g.Ap, g.Ai)
    y = plan.matvec(torch.as_tensor(Axb).cuda(), torch.as_tensor(b).cuda()).cpu().numpy()
    for k in range(33):
        assert np.array_equal(y[k], orc.csc_mat_vec_ff(g.n, g.n, g.Ap, g.Ai, Axb[k], b[k]))
    y1 = plan.matvec(torch.as_tensor(Axb[0]).cuda(), torch.as_tensor(b).cuda()).cpu().numpy()   # shared values
    assert np.array_equal(y1[5], orc.csc_mat_vec_ff(g.n, g.n, g.Ap, g.Ai, Axb[0], b[5]))


def test_spmv_plan_large_batch():
    """Batches of a few hundred systems (grid.y > 1 row blocks per system on the config-3 pattern): bit-identical to
    the oracle and to a small-batch launch, with beta != 0 and with one value set shared by the batch."""
    import torch
    from csparse3_b200.spmv import SpmvPlan
    for nbus, batch in ((118, 333), (2000, 300)):
        g = synth.GridCase(nbus)
        Axb, b = g.jacobian_batch(0, 16)
        reps = -(-batch // 16)
        Axt = np.tile(Axb, (reps, 1))[:batch] * np.linspace(0.5, 1.5, batch)[:, None]
        bt = np.tile(b, (reps, 1))[:batch] + np.arange(batch)[:, None]
        plan = SpmvPlan(g.n, g.n, g.Ap, g.Ai)
        dA, dx = torch.as_tensor(Axt).cuda(), torch.as_tensor(bt).cuda()
        y = plan.matvec(dA, dx).cpu().numpy()
        small = plan.matvec(dA[:7].contiguous(), dx[:7].contiguous()).cpu().numpy()
        assert np.array_equal(y[:7], small)
        for k in (0, 1, batch // 2, batch - 1):
            assert np.array_equal(y[k], orc.csc_mat_vec_ff(g.n, g.n, g.Ap, g.Ai, Axt[k], bt[k]))
        y0 = torch.as_tensor(bt[::-1].copy()).cuda()
        y2 = plan.matvec(dA, dx, y0.clone(), beta=-0.5).cpu().numpy()
        ref = plan.matvec(dA[:7].contiguous(), dx[:7].contiguous(), y0[:7].clone(), beta=-0.5).cpu().numpy()
        assert np.array_equal(y2[:7], ref)
        y1 = plan.matvec(dA[3].contiguous(), dx).cpu().numpy()                                # shared values
        assert np.array_equal(y1[batch - 1], orc.csc_mat_vec_ff(g.n, g.n, g.Ap, g.Ai, Axt[3], bt[batch - 1]))


# ---- transposition / SpGEMM -----------------------------------------------------------------------------------
@pytest.mark.parametrize("seed", range(3))
def test_transpose_tocsr_spgemm_vs_oracle(seed):
    rng = np.random.default_rng(100 + seed)
    for m, k, n, dens in ((1, 1, 1, 1.0), (23, 31, 19, 0.2), (200, 150, 260, 0.03), (40, 40, 40, 0.7), (600, 600, 600, 0.01)):
        Ap, Ai, Ax = _rand_csc(rng, m, k, dens)
        Bp, Bi, Bx = _rand_csc(rng, k, n, dens)
        for a, b in zip(B.csc_transpose(m, k, Ap, Ai, Ax), orc.csc_transpose(m, k, Ap, Ai, Ax)):
            assert np.array_equal(a, b)
        R1 = [np.zeros(m + 1, dtype=np.int32), np.empty(Ap[k], dtype=np.int32), np.empty(Ap[k])]
        R2 = [np.zeros(m + 1, dtype=np.int32), np.empty(Ap[k], dtype=np.int32), np.empty(Ap[k])]
        B.csc_to_csr(m, k, Ap, Ai, Ax, *R1); orc.csc_to_csr(m, k, Ap, Ai, Ax, *R2)
        assert all(np.array_equal(a, b) for a, b in zip(R1, R2))
        Cm, Cn, Cp, Ci, Cx, nnz = B.csc_multiply_ff(m, k, Ap, Ai, Ax, k, n, Bp, Bi, Bx)
        Om, On, Op, Oi, Ox, onnz = orc.csc_multiply_ff(m, k, Ap, Ai, Ax, k, n, Bp, Bi, Bx)
        oi, ox = sort_columns(n, Op, Oi, Ox)
        assert nnz == onnz and np.array_equal(Cp, Op) and np.array_equal(Ci, oi) and np.array_equal(Cx, ox)
        # scipy two-pass contract (zeros dropped)
        A2 = CscMat(m, k, indptr=Ap, indices=Ai, data=Ax); B2 = CscMat(k, n, indptr=Bp, indices=Bi, data=Bx)
        S = sp.csc_matrix((Ax, Ai, Ap), shape=(m, k)) @ sp.csc_matrix((Bx, Bi, Bp), shape=(k, n))
        assert ((A2 * B2).todense() == S.toarray()).all()


def test_plus_minus_vs_oracle(golden_ref):
    rng = np.random.default_rng(77)
    cases = [(40, 30, 0.2), (1, 1, 1.0), (200, 150, 0.03), (64, 64, 0.6)]
    mats = []
    for m, n, dens in cases:
        mats.append((m, n, _rand_csc(rng, m, n, dens), _rand_csc(rng, m, n, dens)))
    d = golden_ref                                     # unsorted row indices (first-touch ordered product) vs sorted
    mats.append((53, 53, (d["U_p"], d["U_i"], d["U_x"]), (d["Sp"], d["Si"], d["Sx"])))
    dup_p = np.array([0, 3, 5], dtype=np.int32); dup_i = np.array([1, 0, 1, 2, 2], dtype=np.int32)       # duplicates
    dup_x = np.array([1.0, 2.0, 3.0, 4.0, -4.0])
    mats.append((3, 2, (dup_p, dup_i, dup_x), (dup_p, dup_i, dup_x * 0.5)))
    for m, n, (Ap, Ai, Ax), (Bp, Bi, Bx) in mats:
        for sign, fn in ((1.0, B.sptools.csc_plus_csc), (-1.0, B.sptools.csc_minus_csc)):
            cap = int(Ap[n] + Bp[n])
            Cp = np.zeros(n + 1, dtype=np.int32); Ci = np.zeros(max(cap, 1), dtype=np.int32); Cx = np.zeros(max(cap, 1))
            Op = np.zeros(n + 1, dtype=np.int32); Oi = np.zeros(max(cap, 1), dtype=np.int32); Ox = np.zeros(max(cap, 1))
            fn(m, n, Ap, Ai, Ax, Bp, Bi, Bx, Cp, Ci, Cx)
            k = orc.csc_plusminus_csc(m, n, Ap, Ai, Ax, Bp, Bi, Bx, sign, Op, Oi, Ox)
            oi, ox = sort_columns(n, Op, Oi[:k], Ox[:k])
            assert np.array_equal(Cp, Op) and np.array_equal(Ci[:k], oi) and np.array_equal(Cx[:k], ox)


def test_spgemm_big_columns_and_laplacian():
    rng = np.random.default_rng(9)
    Ap, Ai, Ax = _rand_csc(rng, 3000, 300, 0.5)          # columns with > 512 candidate products -> global tables
    Bp, Bi, Bx = _rand_csc(rng, 300, 40, 0.5)
    Cm, Cn, Cp, Ci, Cx, nnz = B.csc_multiply_ff(3000, 300, Ap, Ai, Ax, 300, 40, Bp, Bi, Bx)
    Om, On, Op, Oi, Ox, onnz = orc.csc_multiply_ff(3000, 300, Ap, Ai, Ax, 300, 40, Bp, Bi, Bx)
    oi, ox = sort_columns(40, Op, Oi, Ox)
    assert np.array_equal(Cp, Op) and np.array_equal(Ci, oi) and np.array_equal(Cx, ox)
    n, Ap, Ai, Ax = synth.laplacian_2d(100)              # config 1: A*A, nnz(C) = 128,004
    Cm, Cn, Cp, Ci, Cx, nnz = B.csc_multiply_ff(n, n, Ap, Ai, Ax, n, n, Ap, Ai, Ax)
    Om, On, Op, Oi, Ox, onnz = orc.csc_multiply_ff(n, n, Ap, Ai, Ax, n, n, Ap, Ai, Ax)
    oi, ox = sort_columns(n, Op, Oi, Ox)
    assert nnz == 128004 and np.array_equal(Cp, Op) and np.array_equal(Ci, oi) and np.array_equal(Cx, ox)


# ---- LU refactor / solve ------------------------------------------------------------------------------------------
def _oracle_batch(sym, n, Ap, Ai, Axb, bb):
    Lx = np.empty((len(Axb), sym.lnz)); Ux = np.empty((len(Axb), sym.unz)); x = np.empty((len(Axb), n))
    for k in range(len(Axb)):
        Lx[k], Ux[k] = orc.csc_lu_refactor(n, Ap, Ai, Axb[k], sym.q, sym.pinv, sym.Lp, sym.Li, sym.Up, sym.Ui)
        x[k] = orc.csc_lu_solve(n, sym.Lp, sym.Li, Lx[k], sym.Up, sym.Ui, Ux[k], sym.pinv, sym.q, bb[k])
    return Lx, Ux, x


def _check_accuracy(n, Ap, Ai, Axb, bb, x):
    for k in range(len(Axb)):
        A = sp.csc_matrix((Axb[k], Ai, Ap), shape=(n, n))
        assert np.linalg.norm(A @ x[k] - bb[k]) <= 1e-10 * np.linalg.norm(bb[k])


@pytest.mark.parametrize("order,tol", [(1, 1e-3), (2, 1.0), (0, 1.0), (3, 0.1)])
def test_lu_small_matrices_bit_exact(order, tol):
    rng = np.random.default_rng(order)
    cases = [synth.laplacian_2d(9), synth.laplacian_3d(5), (1, np.array([0, 1], dtype=np.int32), np.array([0], dtype=np.int32), np.array([2.0]))]
    for t in range(5):
        n = int(rng.integers(2, 150))
        A = sp.csc_matrix(sp.random(n, n, density=min(1.0, 3.0 / n + 0.03), random_state=int(rng.integers(1 << 30)),
                                    format="csc") + sp.diags(rng.uniform(0.5, 2.0, n)))
        cases.append((n, A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.copy()))
    for n, Ap, Ai, Ax in cases:
        sym = LuSymbolic(n, Ap, Ai, Ax, order=order, tol=tol)
        nb = 5
        Axb = Ax[None, :] * rng.uniform(0.9, 1.1, (nb, len(Ax)))
        Axb[0] = Ax
        bb = rng.standard_normal((nb, n))
        Lx, Ux, status = sym.refactor_host(Axb)
        oLx, oUx, ox = _oracle_batch(sym, n, Ap, Ai, Axb, bb)
        assert (status == 0).all()
        assert np.array_equal(Lx, oLx) and np.array_equal(Ux, oUx)
        assert np.array_equal(Lx[0], sym.Lx0) and np.array_equal(Ux[0], sym.Ux0)      # == first factorisation
        assert np.array_equal(sym.solve_host(Lx, Ux, bb), ox)
        x, status = sym.refactor_solve_host(Axb, bb)
        assert np.array_equal(x, ox) and (status == 0).all()


@pytest.mark.parametrize("path,S", [("sm", 1), ("sm", 2), ("sm", 4), ("sm", 8), ("sm", 16),
                                    ("ws", 2), ("ws", 4), ("ws", 8), ("ws", 16)])
def test_lu_grid118_bundle_widths_agree(path, S):
    """Every bundle width and both factor layouts (system-major API / interleaved workspace) must give the
    same bits (shard / bundle invariance); ragged batch size; tiny ring and stage sizes to exercise wrap-around."""
    import os, subprocess, sys
    g = synth.GridCase(118)
    n, Ap, Ai, Ax0 = g.base_jacobian()
    sym = LuSymbolic(n, Ap, Ai, Ax0)
    Axb, bb = g.jacobian_batch(0, 37)
    oLx, oUx, ox = _oracle_batch(sym, n, Ap, Ai, Axb, bb)
    # the tuning knobs are read once per process: run the setting under test in a child process
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys, numpy as np, torch; sys.path.insert(0, %r); from csparse3_b200 import synth; "
            "from csparse3_b200.lu import LuSymbolic; g = synth.GridCase(118); n, Ap, Ai, Ax0 = g.base_jacobian(); "
            "sym = LuSymbolic(n, Ap, Ai, Ax0); Axb, bb = g.jacobian_batch(0, 37); "
            "Lx, Ux, st = sym.refactor_host(Axb); x = sym.solve_host(Lx, Ux, bb); "
            "xw, stw = sym.refactor_solve(torch.as_tensor(Axb).cuda(), torch.as_tensor(bb).cuda()); "
            "np.savez(sys.argv[1], Lx=Lx, Ux=Ux, x=x, st=st, xw=xw.cpu().numpy(), stw=stw.cpu().numpy())") % root
    out = os.path.join(os.environ.get("TMPDIR", "/tmp"), "csp3_w%s%d.npz" % (path, S))
    env = dict(os.environ, CSP3_RF_WIN="64", CSP3_SV_STAGE="64")
    if path == "sm":
        env.update(CSP3_RF_S=str(S), CSP3_SV_S=str(S))
    else:
        env.update(CSP3_WS_S=str(S))
    subprocess.check_call([sys.executable, "-c", code, out], env=env)
    r = np.load(out)
    assert np.array_equal(r["Lx"], oLx) and np.array_equal(r["Ux"], oUx) and np.array_equal(r["x"], ox) and (r["st"] == 0).all()
    assert np.array_equal(r["xw"], ox) and (r["stw"] == 0).all()          # fused path through the workspace layout


@pytest.mark.parametrize("env", [{}, {"CSP3_WIDE_S": "16"}, {"CSP3_WIDE_S": "4"}, {"CSP3_WIDE_S": "16", "CSP3_WIDE_LANE": "4"},
                                 {"CSP3_WIDE_F": "32"}, {"CSP3_WIDE_SOLVE": "0"}, {"CSP3_WIDE": "0"},
                                 {"CSP3_WIDE_SCHED": "0"}, {"CSP3_WIDE_PAIRS": "16", "CSP3_WIDE_RUN": "64"}])
def test_lu_wide_geometries_agree(env):
    """The wide (lane = system) kernels for several bundle widths / systems per lane, a tiny landing area (immediate
    fetches), wide refactor + v3 solve, and v3 only: all bit-identical to the oracle on a ragged batch; one system."""
    import os, subprocess, sys
    g = synth.GridCase(118)
    n, Ap, Ai, Ax0 = g.base_jacobian()
    sym = LuSymbolic(n, Ap, Ai, Ax0)
    Axb, bb = g.jacobian_batch(0, 37)
    oLx, oUx, ox = _oracle_batch(sym, n, Ap, Ai, Axb, bb)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys, numpy as np, torch; sys.path.insert(0, %r); from csparse3_b200 import synth; "
            "from csparse3_b200.lu import LuSymbolic; g = synth.GridCase(118); n, Ap, Ai, Ax0 = g.base_jacobian(); "
            "sym = LuSymbolic(n, Ap, Ai, Ax0); Axb, bb = g.jacobian_batch(0, 37); "
            "xw, stw = sym.refactor_solve(torch.as_tensor(Axb).cuda(), torch.as_tensor(bb).cuda()); "
            "x1, st1 = sym.refactor_solve(torch.as_tensor(Axb[:1]).cuda(), torch.as_tensor(bb[:1]).cuda()); "
            "xh, sth = sym.refactor_solve_host(Axb, bb); "
            "np.savez(sys.argv[1], xw=xw.cpu().numpy(), stw=stw.cpu().numpy(), x1=x1.cpu().numpy(), xh=xh, sth=sth, "
            "wide=np.array(sym.wide_width))") % root
    out = os.path.join(os.environ.get("TMPDIR", "/tmp"), "csp3_wide_%s.npz" % "_".join("%s%s" % kv for kv in sorted(env.items())))
    subprocess.check_call([sys.executable, "-c", code, out], env=dict(os.environ, **env))
    r = np.load(out)
    assert np.array_equal(r["xw"], ox) and (r["stw"] == 0).all() and np.array_equal(r["x1"][0], ox[0])
    assert np.array_equal(r["xh"], ox) and (r["sth"] == 0).all()
    if env.get("CSP3_WIDE") != "0":
        assert int(r["wide"]) == int(env.get("CSP3_WIDE_S", 8))


def test_lu_config3_sample_device_api():
    """Config 3 pattern (2,000-bus Jacobian): device-tensor API, 24 systems vs the oracle, bit-exact."""
    import torch
    g = synth.GridCase(2000)
    n, Ap, Ai, Ax0 = g.base_jacobian()
    sym = LuSymbolic(n, Ap, Ai, Ax0)
    assert (n, sym.nnz) == (3598, 24446)
    Axb, bb = g.jacobian_batch(0, 24)
    dA, db = torch.as_tensor(Axb).cuda(), torch.as_tensor(bb).cuda()
    Lx, Ux, status = sym.refactor(dA)
    x = sym.solve(Lx, Ux, db)
    x2, status2 = sym.refactor_solve(dA, db)
    torch.cuda.synchronize()
    oLx, oUx, ox = _oracle_batch(sym, n, Ap, Ai, Axb, bb)
    assert (status.cpu().numpy() == 0).all() and (status2.cpu().numpy() == 0).all()
    assert np.array_equal(Lx.cpu().numpy(), oLx) and np.array_equal(Ux.cpu().numpy(), oUx)
    assert np.array_equal(x.cpu().numpy(), ox) and np.array_equal(x2.cpu().numpy(), ox)
    _check_accuracy(n, Ap, Ai, Axb, bb, ox)
    # north-star tolerance vs an independent solver
    import scipy.sparse.linalg as spla
    xr = spla.splu(sp.csc_matrix((Axb[3], Ai, Ap), shape=(n, n))).solve(bb[3])
    assert np.linalg.norm(ox[3] - xr) <= 1e-9 * np.linalg.norm(xr)


def test_lu_config3_full_batch_properties():
    """BASELINE.json's full size (10,000 systems on the 2,000-bus pattern) through size-independent properties:
    exact homogeneity (A scaled by a power of two scales x by its inverse, bit for bit), exact linearity in b,
    batch invariance (a system's bits do not depend on the batch it is solved in), residual <= 1e-10 for every
    system (batched SpMV on the device), and the oracle on a sample."""
    import torch
    from csparse3_b200.spmv import SpmvPlan
    g = synth.GridCase(2000)
    n, Ap, Ai, Ax0 = g.base_jacobian()
    sym = LuSymbolic(n, Ap, Ai, Ax0)
    nb, reps = 250, 40
    base, bb = g.jacobian_batch(0, nb)
    scale = 2.0 ** (np.arange(reps) - reps // 2)                   # 2^-20 .. 2^19, block reps // 2 is unscaled
    Ax = (base[None, :, :] * scale[:, None, None]).reshape(nb * reps, -1)
    b = np.tile(bb, (reps, 1))
    dA, db = torch.as_tensor(Ax).cuda(), torch.as_tensor(b).cuda()
    x, st = sym.refactor_solve(dA, db)
    assert int(st.abs().max().item()) == 0
    xh = x.cpu().numpy()
    x0 = xh[(reps // 2) * nb:(reps // 2 + 1) * nb]
    for r in range(reps):
        assert np.array_equal(xh[r * nb:(r + 1) * nb] * scale[r], x0), r
    # the oracle on a sample of the unscaled block; the same systems alone in a small batch
    oLx, oUx, ox = _oracle_batch(sym, n, Ap, Ai, base[:6], bb[:6])
    assert np.array_equal(x0[:6], ox)
    xs, _ = sym.refactor_solve(torch.as_tensor(base[:6]).cuda(), torch.as_tensor(bb[:6]).cuda())
    assert np.array_equal(xs.cpu().numpy(), ox)
    # linearity in b (exact for a factor of two)
    x2, _ = sym.refactor_solve(dA, 2.0 * db)
    assert torch.equal(x2, 2.0 * x)
    # residual of every system
    r = SpmvPlan(n, n, Ap, Ai).matvec(dA, x) - db
    rel = (r.norm(dim=1) / db.norm(dim=1)).max().item()
    assert rel <= 1e-10, rel


def test_lu_config4_outages_and_config1():
    g = synth.GridCase(10000)
    n, Ap, Ai, Ax0 = g.base_jacobian()
    sym = LuSymbolic(n, Ap, Ai, Ax0)
    Axb, bb = g.outage_batch(100, 6)                    # N-1 outages: explicit zeros on the shared pattern
    x, status = sym.refactor_solve_host(Axb, bb)
    oLx, oUx, ox = _oracle_batch(sym, n, Ap, Ai, Axb, bb)
    assert (status == 0).all() and np.array_equal(x, ox)
    n, Ap, Ai, Ax = synth.laplacian_2d(100)             # config 1
    sym = LuSymbolic(n, Ap, Ai, Ax, order=1, tol=1.0)
    b = np.ones((1, n))
    x, status = sym.refactor_solve_host(Ax[None, :].copy(), b)
    oLx, oUx, ox = _oracle_batch(sym, n, Ap, Ai, Ax[None, :], b)
    assert status[0] == 0 and np.array_equal(x, ox)
    _check_accuracy(n, Ap, Ai, Ax[None, :], b, x)
    assert np.array_equal(B.csc_lusol(1, n, Ap, Ai, Ax, b[0], 1.0), orc.csc_lusol(1, n, Ap, Ai, Ax, b[0], 1.0))


def test_lu_zero_pivot_is_reported_per_system():
    g = synth.GridCase(118)
    n, Ap, Ai, Ax0 = g.base_jacobian()
    sym = LuSymbolic(n, Ap, Ai, Ax0)
    Axb, bb = g.jacobian_batch(0, 6)
    Axb[2] = 0.0                                       # system 2: every pivot is zero -> first column reported
    Axb[4, :] = np.nan
    x, status = sym.refactor_solve_host(Axb, bb)
    ok = [0, 1, 3, 5]
    oLx, oUx, ox = _oracle_batch(sym, n, Ap, Ai, Axb[ok], bb[ok])
    assert (status[ok] == 0).all() and np.array_equal(x[ok], ox)
    for bad in (2, 4):
        with pytest.raises(ArithmeticError) as e:
            orc.csc_lu_refactor(n, Ap, Ai, Axb[bad], sym.q, sym.pinv, sym.Lp, sym.Li, sym.Up, sym.Ui)
        assert status[bad] == int(str(e.value).split()[-1]) + 1


def test_flat_lu_api_and_cscmat_solve():
    g = synth.GridCase(118)
    n, Ap, Ai, Ax = g.base_jacobian()
    q = B.csc_amd(1, n, n, Ap, Ai)
    Lp, Li, Lx, Up, Ui, Ux, pinv = B.csc_lu(n, Ap, Ai, Ax, q, 1e-3)
    o = orc.csc_lu(n, Ap, Ai, Ax, q, 1e-3)
    assert all(np.array_equal(a, b) for a, b in zip((Lp, Li, Lx, Up, Ui, Ux, pinv), o))
    Ax2 = Ax * 1.01
    Lx2, Ux2 = B.csc_lu_refactor(n, Ap, Ai, Ax2, q, pinv, Lp, Li, Up, Ui)
    oL, oU = orc.csc_lu_refactor(n, Ap, Ai, Ax2, q, pinv, Lp, Li, Up, Ui)
    assert np.array_equal(Lx2, oL) and np.array_equal(Ux2, oU)
    b = np.arange(n, dtype=np.float64)
    assert np.array_equal(B.csc_lu_solve(n, Ap, Ai, q, pinv, Lp, Li, Lx2, Up, Ui, Ux2, b),
                          orc.csc_lu_solve(n, Lp, Li, oL, Up, Ui, oU, pinv, q, b))
    A = CscMat(n, n, indptr=Ap, indices=Ai, data=Ax)
    assert np.array_equal(A.solve(b), orc.csc_lusol(1, n, Ap, Ai, Ax, b, 1e-3))


# ---- [[A, B], [C, D]] assembly ---------------------------------------------------------

def test_stack_matches_reference_fixture(golden_ref):
    """csc_stack_4_by_4_ff against the output of the reference's numba kernel (tests/golden/make_golden.py)."""
    d = golden_ref
    args = []
    for c in "abcd":
        sh = d["st_" + c + "shape"]
        args += [int(sh[0]), int(sh[1]), d["st_" + c + "i"], d["st_" + c + "p"], d["st_" + c + "x"]]
    mm, nn, Pi, Pp, Px = B.csc_stack_4_by_4_ff(*args)
    assert (mm, nn) == tuple(d["ref_st_shape"])
    assert np.array_equal(Pi, d["ref_st_i"]) and np.array_equal(Pp, d["ref_st_p"]) and np.array_equal(Px, d["ref_st_x"])
    oo = orc.csc_stack_4_by_4_ff(*args)
    assert np.array_equal(Pi, oo[2]) and np.array_equal(Pp, oo[3]) and np.array_equal(Px, oo[4])


def test_pack_4_by_4_vs_scipy():
    """src/test/test_matrix_stacking.py:12-40 at a small seeded size."""
    from csparse3_b200 import pack_4_by_4
    k = 15
    Q = [sp.csc_matrix(sp.random(*s, density=0.2, random_state=t)) for t, s in
         enumerate([(k, 4 * k), (k, k), (6 * k, 4 * k), (6 * k, k)])]
    E = sp.hstack((sp.vstack((Q[0], Q[2])), sp.vstack((Q[1], Q[3]))))
    E1 = pack_4_by_4(*[scipy_to_mat(M) for M in Q])
    assert (E.toarray() == E1.todense()).all()


def test_stack_edge_cases():
    """Empty blocks, empty columns, a block without columns; shape mismatch raises like the reference's assert."""
    z = lambda m, n: scipy_to_mat(sp.csc_matrix((m, n)))
    r = lambda m, n, s: scipy_to_mat(sp.csc_matrix(sp.random(m, n, density=0.3, random_state=s)))
    for blocks in ([z(3, 4), z(3, 2), z(5, 4), z(5, 2)], [r(3, 4, 1), z(3, 2), z(5, 4), r(5, 2, 2)],
                   [r(3, 4, 3), r(3, 0, 4), r(2, 4, 5), r(2, 0, 6)], [r(0, 3, 7), r(0, 2, 8), r(4, 3, 9), r(4, 2, 10)]):
        args = []
        for M in blocks:
            args += [M.m, M.n, M.indices, M.indptr, M.data]
        got = B.csc_stack_4_by_4_ff(*args)
        want = orc.csc_stack_4_by_4_ff(*args)
        assert got[:2] == want[:2]
        for a, b in zip(got[2:], want[2:]):
            assert np.array_equal(a, b)
    with pytest.raises(AssertionError):
        B.csc_stack_4_by_4_ff(3, 4, *[np.zeros(0, np.int32), np.zeros(5, np.int32), np.zeros(0)],
                              2, 2, *[np.zeros(0, np.int32), np.zeros(3, np.int32), np.zeros(0)],
                              5, 4, *[np.zeros(0, np.int32), np.zeros(5, np.int32), np.zeros(0)],
                              5, 2, *[np.zeros(0, np.int32), np.zeros(3, np.int32), np.zeros(0)])


def _jacobian_blocks(g, count):
    """The four blocks of the polar Jacobian of `count` value sets, as (pattern CscMat x4, values [count, nnz] x4),
    cut out of the assembled matrices of the generator (so that the stacked result is known)."""
    n, Ap, Ai, Ax0 = g.base_jacobian()
    Axb, bb = g.jacobian_batch(0, count)
    J = sp.csc_matrix((np.arange(1, len(Ai) + 1, dtype=np.float64), Ai, Ap), shape=(n, n))      # entry numbers
    n1 = n // 2 + 3                                     # any split point works for the stacking identity
    blocks, vals = [], []
    for rs, cs in ((slice(0, n1), slice(0, n1)), (slice(0, n1), slice(n1, n)), (slice(n1, n), slice(0, n1)), (slice(n1, n), slice(n1, n))):
        S = sp.csc_matrix(J[rs, cs])
        S.sort_indices()
        src = S.data.astype(np.int64) - 1
        blocks.append(CscMat(S.shape[0], S.shape[1], indptr=S.indptr.astype(np.int32), indices=S.indices.astype(np.int32),
                             data=Ax0[src].copy()))
        vals.append(np.ascontiguousarray(Axb[:, src]))
    return (n, Ap, Ai, Ax0, Axb, bb), blocks, vals


def test_stack4_plan_batched_vs_oracle():
    """Device-resident batch: one gather kernel; every system equals the oracle's stacking of its four blocks, and
    the stacked pattern is the Jacobian's own (sorted blocks of a sorted matrix)."""
    import torch
    from csparse3_b200.assemble import Stack4Plan
    g = synth.GridCase(118)
    (n, Ap, Ai, Ax0, Axb, bb), blocks, vals = _jacobian_blocks(g, 37)
    plan = Stack4Plan(*blocks)
    assert (plan.m, plan.n, plan.nnz) == (n, n, len(Ai))
    assert np.array_equal(plan.indptr, Ap) and np.array_equal(plan.indices, Ai)
    out = plan.assemble(*[torch.as_tensor(v).cuda() for v in vals]).cpu().numpy()
    for s in (0, 5, 36):
        args = []
        for M, v in zip(blocks, vals):
            args += [M.m, M.n, M.indices, M.indptr, v[s]]
        want = orc.csc_stack_4_by_4_ff(*args)
        assert np.array_equal(want[2], plan.indices) and np.array_equal(want[3], plan.indptr)
        assert np.array_equal(out[s], want[4])
    assert np.array_equal(out, Axb)
    # a block shared by the whole batch (ld = 0)
    out2 = plan.assemble(torch.as_tensor(vals[0]).cuda(), torch.as_tensor(vals[1][3]).cuda(), torch.as_tensor(vals[2]).cuda(),
                         torch.as_tensor(vals[3]).cuda()).cpu().numpy()
    want2 = Axb.copy()
    col = np.repeat(np.arange(n), np.diff(Ap))
    is12 = (col >= blocks[0].n) & (Ai < blocks[0].m)    # positions of block 12 in the stacked order
    want2[:, is12] = Axb[3][is12]
    assert np.array_equal(out2, want2)


def test_assemble_refactor_solve_pipeline():
    """Blocks on the device -> Stack4Plan.assemble -> LuSymbolic.refactor_solve, nothing leaves the GPU in between;
    x equals the oracle's solution of the assembled systems bit for bit."""
    import torch
    from csparse3_b200.assemble import Stack4Plan
    g = synth.GridCase(118)
    (n, Ap, Ai, Ax0, Axb, bb), blocks, vals = _jacobian_blocks(g, 19)
    plan = Stack4Plan(*blocks)
    sym = LuSymbolic(plan.n, plan.indptr, plan.indices, plan.assemble(*[torch.as_tensor(v[:1]).cuda() for v in vals]).cpu().numpy()[0])
    Ax_dev = plan.assemble(*[torch.as_tensor(v).cuda() for v in vals])
    x, st = sym.refactor_solve(Ax_dev, torch.as_tensor(bb).cuda())
    oLx, oUx, ox = _oracle_batch(sym, n, Ap, Ai, Axb, bb)
    assert (st.cpu().numpy() == 0).all() and np.array_equal(x.cpu().numpy(), ox)
