"""GPU parity at BASELINE.json's full sizes (round 2): config-5-size CSC kernels against the oracle, the complete
config-4 N-1 sweep (every non-bridge outage of the 10,000-bus grid) with residual <= 1e-10 for every system, pivot
growth detection with the reported host re-pivot, and the experimental panel refactor kernel."""
import os
import subprocess
import sys

import numpy as np
import pytest

from csparse3_b200 import csc_b200 as B
from csparse3_b200 import synth
from csparse3_b200.lu import LuSymbolic
from oracle import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def test_config5_size_spmv_spgemm_transpose_vs_oracle():
    """3-D 7-point Laplacian, n = 1e6, nnz = 6.94e6; A*A has 24.6e6 entries."""
    n, Ap, Ai, Ax = synth.laplacian_3d(100)
    assert n == 1_000_000 and int(Ap[n]) == 6_940_000
    x = np.random.default_rng(0).standard_normal(n)
    assert np.array_equal(B.csc_mat_vec_ff(n, n, Ap, Ai, Ax, x), orc.csc_mat_vec_ff(n, n, Ap, Ai, Ax, x))
    Tm, Tn, Tp, Ti, Tx = B.csc_transpose(n, n, Ap, Ai, Ax)
    o = orc.csc_transpose(n, n, Ap, Ai, Ax)
    assert np.array_equal(Tp, o[2]) and np.array_equal(Ti, o[3]) and np.array_equal(Tx, o[4])
    Cm, Cn, Cp, Ci, Cx, nnz = B.csc_multiply_ff(n, n, Ap, Ai, Ax, n, n, Ap, Ai, Ax)
    Om, On, Op, Oi, Ox, onnz = orc.csc_multiply_ff(n, n, Ap, Ai, Ax, n, n, Ap, Ai, Ax)
    assert nnz == onnz == 24_581_200 and np.array_equal(Cp, Op)
    order = np.lexsort((Oi, np.repeat(np.arange(n), np.diff(Op))))          # the oracle emits first-touch order
    assert np.array_equal(Ci, Oi[order]) and np.array_equal(Cx, Ox[order])
    # multi-vector product through the sptools stand-in (csc_matvecs, row-major X / Y)
    X = np.random.default_rng(1).standard_normal((n, 3))
    Y = np.zeros((n, 3)); Yo = np.zeros((n, 3))
    B.sptools.csc_matvecs(n, n, 3, Ap, Ai, Ax, X, Y)
    orc.csc_matvecs(n, n, 3, Ap, Ai, Ax, X, Yo)
    assert np.array_equal(Y, Yo)


def test_config4_full_outage_sweep_residuals_and_growth():
    """Every non-bridge outage of the 10,000-bus grid (one refactor+solve each, frozen pivots of the base case):
    Jacobians evaluated on the device from Ybus with one branch removed (checked against the numpy generator on a
    sample), bit-exact against the oracle on 64 sampled outages, relative residual <= 1e-10 and bounded pivot growth
    for EVERY system."""
    import torch
    from csparse3_b200.nr import NewtonPlan
    from csparse3_b200.spmv import SpmvPlan
    case = synth.GridCase(10000)
    n, Ap, Ai, Ax0 = case.base_jacobian()
    sym = LuSymbolic(n, Ap, Ai, Ax0)
    plan = NewtonPlan.from_case(case, sym, outages=True)
    nb = case.non_bridge_branches()
    Bt = len(nb)
    assert Bt > 10000
    Vb = case.voltages([case.seed + 7])[0]
    vm = torch.as_tensor(np.broadcast_to(np.abs(Vb), (Bt, case.n_bus)).copy()).cuda()
    va = torch.as_tensor(np.broadcast_to(np.angle(Vb), (Bt, case.n_bus)).copy()).cuda()
    ob = torch.as_tensor(nb.astype(np.int32)).cuda()
    sspec = torch.zeros((Bt, n), dtype=torch.float64, device="cuda")
    Ax, _, _ = plan.jacobian(vm, va, sspec, out_branch=ob)                  # [Bt, nnz] on the device
    del vm, va, sspec
    rng = np.random.default_rng(11)
    sample = np.sort(rng.choice(Bt, 64, replace=False))
    Ax_s = Ax[torch.as_tensor(sample).cuda()].cpu().numpy()
    for t, k in enumerate(sample[:16]):                                     # device generator == numpy generator
        a_np, _ = case.outage_batch(int(k), 1)
        assert np.abs(Ax_s[t] - a_np[0]).max() <= 1e-11 * np.abs(a_np[0]).max()
    b = torch.as_tensor(rng.standard_normal((Bt, n))).cuda()
    x, rep = sym.refactor_solve_checked(Ap, Ai, Ax, b, growth_limit=1e6, resid_tol=1e-10)
    assert int(rep["status"].abs().max().item()) == 0
    growth = rep["growth"].cpu().numpy()
    assert np.isfinite(growth).all() and growth.max() < 1e6, growth.max()
    # residual of every system after the (reported) re-pivots; the sweep needs none or a handful
    r = SpmvPlan(n, n, Ap, Ai).matvec(Ax, x) - b
    rel = (r.norm(dim=1) / b.norm(dim=1)).cpu().numpy()
    assert rel.max() <= 1e-10, (rel.max(), rep["flagged"])
    assert len(rep["flagged"]) <= Bt // 100, len(rep["flagged"])
    # oracle on the sampled outages that were not re-pivoted: same bits
    xs = x[torch.as_tensor(sample).cuda()].cpu().numpy()
    bs = b[torch.as_tensor(sample).cuda()].cpu().numpy()
    for t, k in enumerate(sample):
        if k in rep["flagged"]:
            continue
        Lx, Ux = orc.csc_lu_refactor(n, Ap, Ai, Ax_s[t], sym.q, sym.pinv, sym.Lp, sym.Li, sym.Up, sym.Ui)
        xo = orc.csc_lu_solve(n, sym.Lp, sym.Li, Lx, sym.Up, sym.Ui, Ux, sym.pinv, sym.q, bs[t])
        assert np.array_equal(xs[t], xo), k


def test_pivot_growth_is_detected_and_repivoted():
    """A frozen pivot that the values of one system make tiny: status stays 0, growth explodes, the residual is bad;
    refactor_solve_checked reports the system and re-pivots it on the host."""
    import torch
    rng = np.random.default_rng(5)
    n = 40
    import scipy.sparse as sp
    A = sp.csc_matrix(sp.random(n, n, density=0.15, random_state=3) + sp.diags(rng.uniform(2.0, 3.0, n)))
    A.sort_indices()
    Ap, Ai, Ax0 = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.copy()
    sym = LuSymbolic(n, Ap, Ai, Ax0, order=1, tol=1e-3)
    Axb = np.tile(Ax0, (5, 1)) * rng.uniform(0.95, 1.05, (5, len(Ax0)))
    # system 3: the diagonal entry the first pivot was chosen on becomes 1e-13 (the matrix stays well conditioned:
    # the column has other entries to pivot on)
    col = int(sym.q[0])
    prow = int(np.where(sym.pinv == 0)[0][0])
    p = [t for t in range(Ap[col], Ap[col + 1]) if Ai[t] == prow][0]
    assert Ap[col + 1] - Ap[col] >= 2
    Axb[3, p] = 1e-13
    b = rng.standard_normal((5, n))
    x, rep = sym.refactor_solve_checked(Ap, Ai, torch.as_tensor(Axb).cuda(), torch.as_tensor(b).cuda())
    assert (rep["status"].cpu().numpy() == 0).all()
    assert rep["flagged"].tolist() == [3]
    assert rep["growth"].cpu().numpy()[3] > 1e6 and rep["resid_after"][0] <= 1e-10
    xh = x.cpu().numpy()
    for k in range(5):
        Ak = sp.csc_matrix((Axb[k], Ai, Ap), shape=(n, n))
        assert np.linalg.norm(Ak @ xh[k] - b[k]) <= 1e-10 * np.linalg.norm(b[k]), k


@pytest.mark.parametrize("env", [{"CSP3_PANEL": "1"}, {"CSP3_PANEL": "1", "CSP3_PANEL_FMA": "1"}, {"CSP3_TMEM": "1"},
                                 {"CSP3_ROWLANE": "1"}, {"CSP3_ROWLANE": "1", "CSP3_RL_WINDOW": "1"}])
def test_panel_and_tmem_refactor_kernel_parity(env):
    """The experimental refactor kernels: the panel kernel (lu_panel.cu; bit-exact in exact mode, within 1e-9 with fused
    multiply-add) and the kernel that keeps the accumulator in tensor memory (lu_refactor_tmem_kernel: same program as
    the wide kernel, tcgen05.ld / .st as a per-lane scratchpad, factors in 32-system bundles; used for the 2,000-bus
    pattern, the 118-bus one exceeds its shared-memory budget and takes the wide kernel).  The knobs are read once
    per process, so the check runs in a child process."""
    code = (
        "import sys; sys.path.insert(0, %r); import numpy as np, torch\n"
        "from csparse3_b200 import synth; from csparse3_b200.lu import LuSymbolic; from oracle import oracle as orc\n"
        "for nb, B in ((118, 37), (2000, 71)):\n"
        "    g = synth.GridCase(nb); n, Ap, Ai, Ax0 = g.base_jacobian(); sym = LuSymbolic(n, Ap, Ai, Ax0)\n"
        "    Ax, b = g.jacobian_batch(0, B)\n"
        "    work = sym.workspace(B, 'cuda'); st = sym.refactor_ws(torch.as_tensor(Ax).cuda(), work)\n"
        "    x = sym.solve_ws(work, torch.as_tensor(b).cuda()).cpu().numpy()\n"
        "    assert int(st.abs().max().item()) == 0\n"
        "    worst, exact = 0.0, True\n"
        "    for k in range(B):\n"
        "        Lx, Ux = orc.csc_lu_refactor(n, Ap, Ai, Ax[k], sym.q, sym.pinv, sym.Lp, sym.Li, sym.Up, sym.Ui)\n"
        "        xo = orc.csc_lu_solve(n, sym.Lp, sym.Li, Lx, sym.Up, sym.Ui, Ux, sym.pinv, sym.q, b[k])\n"
        "        exact = exact and np.array_equal(x[k], xo)\n"
        "        worst = max(worst, float(np.linalg.norm(x[k] - xo) / np.linalg.norm(xo)))\n"
        "    assert worst <= 1e-9, worst\n"
        "    assert exact or %r\n"
        "print('ok')\n") % (ROOT, "CSP3_PANEL_FMA" in env)
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env), capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-2000:]


def test_transpose_with_dense_rows_and_spgemm_with_big_columns():
    """Shapes the per-row / per-column path selection exists for: a matrix that is sparse on average
    but holds dense rows (transposition: worklist + CTA sort, 3 length classes), and a product whose columns all
    exceed the shared-memory hash table (one reusable global table per CTA)."""
    import scipy.sparse as sp
    rng = np.random.default_rng(9)
    m, n = 30000, 25000
    A = sp.random(m, n, density=2e-4, random_state=4, format="lil")
    A[7, :] = rng.standard_normal(n)                         # 25,000 entries: beyond the shared-memory sort
    cols = rng.choice(n, 3000, replace=False); A[11, cols] = 1.5          # CTA bitonic sort
    cols = rng.choice(n, 200, replace=False); A[13, cols] = -2.0          # just above the one-thread limit
    A = sp.csc_matrix(A); A.sort_indices()
    Ap, Ai, Ax = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.copy()
    T = B.csc_transpose(m, n, Ap, Ai, Ax)
    o = orc.csc_transpose(m, n, Ap, Ai, Ax)
    assert np.array_equal(T[2], o[2]) and np.array_equal(T[3], o[3]) and np.array_equal(T[4], o[4])
    k = 3000
    M = sp.csc_matrix(sp.random(k, k, density=0.012, random_state=6) + sp.eye(k)); M.sort_indices()   # ~37 entries per column
    Mp, Mi, Mx = M.indptr.astype(np.int32), M.indices.astype(np.int32), M.data.copy()
    assert np.diff(Mp).mean() ** 2 > 512
    Cm, Cn, Cp, Ci, Cx, nnz = B.csc_multiply_ff(k, k, Mp, Mi, Mx, k, k, Mp, Mi, Mx)
    Om, On, Op, Oi, Ox, onnz = orc.csc_multiply_ff(k, k, Mp, Mi, Mx, k, k, Mp, Mi, Mx)
    assert nnz == onnz and np.array_equal(Cp, Op)
    order = np.lexsort((Oi, np.repeat(np.arange(k), np.diff(Op))))
    assert np.array_equal(Ci, Oi[order]) and np.array_equal(Cx, Ox[order])


def test_dmma_dense_update_and_supernodes():
    """Groundwork for config 5's supernodal trailing updates: the batched DMMA update C -= A B against numpy, the
    measured FP64 tensor throughput, and the supernode structure of Laplacian factors (wide supernodes exist only
    near the top of the elimination tree)."""
    import torch
    from csparse3_b200 import dense
    rng = np.random.default_rng(2)
    for (batch, m, n, k) in ((3, 70, 45, 19), (2, 256, 128, 64), (1, 8, 8, 4)):
        A = rng.standard_normal((batch, k, m)); Bm = rng.standard_normal((batch, n, k)); Cm = rng.standard_normal((batch, n, m))
        ref = Cm - np.einsum("skm,snk->snm", A, Bm)            # column-major blocks: (A^T B^T)^T per system
        out = dense.dense_update(torch.as_tensor(A).cuda(), torch.as_tensor(Bm).cuda(), torch.as_tensor(Cm).cuda()).cpu().numpy()
        assert np.abs(out - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max()), (batch, m, n, k)
    tf = dense.dmma_peak(2048)
    assert 5.0 < tf < 200.0, tf
    n_, Ap, Ai, Ax = synth.laplacian_3d(12)
    sym = LuSymbolic(n_, Ap, Ai, Ax, order=1, tol=1.0)
    sn = sym.supernodes()
    w = np.diff(sn)
    assert sn[0] == 0 and sn[-1] == n_ and (w >= 1).all() and w.max() >= 16
    # nesting: inside a supernode every column's pattern is the previous one's minus its own row
    j = int(sn[np.argmax(w)])
    r0 = set(sym.Li[sym.Lp[j] + 1:sym.Lp[j + 1]]); r1 = set(sym.Li[sym.Lp[j + 1] + 1:sym.Lp[j + 2]])
    assert r0 - {j + 1} == r1


@pytest.mark.parametrize("env", [{"CSP3_ROWLANE": "1", "CSP3_RL_W": "1"}, {"CSP3_ROWLANE": "1", "CSP3_RL_W": "2", "CSP3_RL_NQ": "1"},
                                 {"CSP3_ROWLANE": "1", "CSP3_RL_W": "2", "CSP3_RL_NQ": "2"}, {"CSP3_ROWLANE": "1", "CSP3_RL_W": "4", "CSP3_RL_NQ": "2"},
                                 {"CSP3_ROWLANE": "1", "CSP3_RL_W": "8"}, {"CSP3_ROWLANE": "1", "CSP3_RL_W": "4", "CSP3_RL_NQ": "2", "CSP3_RL_MARGIN": "-100000"},
                                 {}, {"CSP3_ROWSWEEP": "1"}])
def test_rowlane_refactor_geometries(env):
    """Row-lane refactor kernel (lu_rowlane.cu) with 1 / 2 / 4 / 8 warps per bundle, forced (CSP3_RL_W) and as the
    automatic choice for small batches: factors and solutions bit-identical to the oracle on the 118-bus, 2,000-bus and
    small irregular patterns, ragged batch sizes; a zero / non-finite pivot is reported with the oracle's code whichever
    warp of the bundle eliminates that column.  CSP3_ROWSWEEP=1 additionally routes the triangular sweeps of these small
    batches through the experimental row-oriented 8-warps-per-bundle kernel (lu_sweep_rows_kernel).  The knobs are read once
    per process: child process."""
    code = (
        "import sys; sys.path.insert(0, %r); import numpy as np, torch, scipy.sparse as sp\n"
        "from csparse3_b200 import synth; from csparse3_b200.lu import LuSymbolic; from oracle import oracle as orc\n"
        "forced = %r\n"
        "def check(n, Ap, Ai, Ax, b, sym):\n"
        "    B = Ax.shape[0]\n"
        "    name = sym.refactor_kernel_name(B)\n"
        "    assert name == 'lu_refactor_rowlane_kernel' or (not forced and B > 444 * 8), name\n"
        "    work = sym.workspace(B, 'cuda'); st = sym.refactor_ws(torch.as_tensor(Ax).cuda(), work)\n"
        "    x = sym.solve_ws(work, torch.as_tensor(b).cuda()).cpu().numpy(); st = st.cpu().numpy()\n"
        "    for k in range(B):\n"
        "        try:\n"
        "            Lx, Ux = orc.csc_lu_refactor(n, Ap, Ai, Ax[k], sym.q, sym.pinv, sym.Lp, sym.Li, sym.Up, sym.Ui); code = 0\n"
        "        except Exception as e:\n"
        "            code = -1\n"
        "        if code == 0:\n"
        "            assert st[k] == 0, (k, st[k])\n"
        "            xo = orc.csc_lu_solve(n, sym.Lp, sym.Li, Lx, sym.Up, sym.Ui, Ux, sym.pinv, sym.q, b[k])\n"
        "            assert np.array_equal(x[k], xo), k\n"
        "        else:\n"
        "            assert st[k] != 0, k\n"
        "for nb, B in ((118, 37), (118, 520), (2000, 71)):\n"
        "    g = synth.GridCase(nb); n, Ap, Ai, Ax0 = g.base_jacobian(); sym = LuSymbolic(n, Ap, Ai, Ax0)\n"
        "    Ax, b = g.jacobian_batch(0, min(B, 71)); reps = -(-B // len(Ax)); Ax = np.tile(Ax, (reps, 1))[:B]; b = np.tile(b, (reps, 1))[:B]\n"
        "    check(n, Ap, Ai, Ax, b, sym)\n"
        "rng = np.random.default_rng(5)\n"
        "for t in range(4):\n"
        "    n = int(rng.integers(2, 200))\n"
        "    A = sp.csc_matrix(sp.random(n, n, density=min(1.0, 3.0 / n + 0.03), random_state=int(rng.integers(1 << 30)), format='csc') + sp.diags(rng.uniform(0.5, 2.0, n)))\n"
        "    Ap, Ai, Ax0 = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.copy()\n"
        "    sym = LuSymbolic(n, Ap, Ai, Ax0, order=1, tol=1e-3)\n"
        "    Ax = Ax0[None, :] * rng.uniform(0.9, 1.1, (13, len(Ax0))); b = rng.standard_normal((13, n))\n"
        "    check(n, Ap, Ai, Ax, b, sym)\n"
        "# status codes: the pivot of the LAST column of system 1 is made zero, of a middle column of system 2 infinite\n"
        "Ap = np.array([0, 1, 2, 5], dtype=np.int32); Ai = np.array([0, 1, 0, 1, 2], dtype=np.int32); Ax0 = np.array([2.0, 3.0, 1.0, 1.0, 4.0])\n"
        "sym = LuSymbolic(3, Ap, Ai, Ax0, order=0, tol=1.0); Ax = np.tile(Ax0, (3, 1)); Ax[1, 4] = 0.0; Ax[2, 1] = np.inf\n"
        "work = sym.workspace(3, 'cuda'); st = sym.refactor_ws(torch.as_tensor(Ax).cuda(), work).cpu().numpy()\n"
        "assert st.tolist() == [0, 3, 2], st\n"
        "print('ok')\n") % (ROOT, bool(env))
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env), capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stderr[-2000:]
