"""Newton-Raphson iteration on the device (csparse3_b200/nr.py, csrc/nr_kernels.cu) against the numpy restatement
oracle/nr_oracle.py: Jacobian values and right-hand sides to 1e-12, iterates to 1e-9 (north-star tolerance for
floating-point results), for time-series batches and N-1 outages."""
import numpy as np
import pytest

from csparse3_b200 import _lib, synth
from csparse3_b200.lu import LuSymbolic
from oracle import nr_oracle as nro


def _case(nb):
    case = synth.GridCase(nb)
    n, Ap, Ai, Ax0 = case.base_jacobian()
    return case, LuSymbolic(n, Ap, Ai, Ax0)


def _targets(case, B, Y=None):
    """Specified injections = S_calc at seeded 'true' voltages (so every case has a solution near the flat start)."""
    Vt = case.voltages(5000 + np.arange(B))
    Vt[:, case.pv] /= np.abs(Vt[:, case.pv]); Vt[:, 0] = 1.0          # consistent with a flat start: |V| = 1 at PV buses and the slack
    Yb = np.broadcast_to(case.ybus_values() if Y is None else Y, (B, case.nnz_y))
    S, _ = nro.s_calc(case, Vt, Yb)
    return np.ascontiguousarray(np.concatenate([S[:, case.pvpq].real, S[:, case.pq].imag], axis=1)), Vt


def test_oracle_newton_converges_quadratically():
    case, sym = _case(118)
    sspec, Vt = _targets(case, 3)
    arrays = (sym.q, sym.pinv, sym.Lp, sym.Li, sym.Up, sym.Ui)
    vm, va, fnorm, hist = nro.newton(case, arrays, sspec, 6)
    assert (fnorm < 1e-11).all()
    h = np.array(hist)
    assert (h[3] < 1e-3 * h[0]).all() and (h[-1] < 1e-9).all()
    # PV buses keep their magnitude, the slack keeps magnitude and angle
    assert np.array_equal(vm[:, case.pv], np.ones((3, len(case.pv)))) and (va[:, 0] == 0).all()


def test_plan_fails_loudly_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    case, sym = _case(118)
    from csparse3_b200.nr import NewtonPlan
    with pytest.raises(_lib.Csp3Error, match="no CPU fallback"):
        NewtonPlan.from_case(case, sym)


@pytest.mark.gpu
@pytest.mark.parametrize("nb,B", [(118, 33), (2000, 9)])
def test_nr_jacobian_and_rhs_vs_numpy(nb, B):
    import torch
    from csparse3_b200.nr import NewtonPlan
    case, sym = _case(nb)
    plan = NewtonPlan.from_case(case, sym)
    sspec, _ = _targets(case, B)
    V = case.voltages(7000 + np.arange(B))
    vm, va = np.abs(V), np.angle(V)
    Ax, b, fnorm = plan.jacobian(torch.as_tensor(vm).cuda(), torch.as_tensor(va).cuda(), torch.as_tensor(sspec).cuda())
    torch.cuda.synchronize()
    Vr = vm * np.exp(1j * va)
    Y = np.broadcast_to(case.ybus_values(), (B, case.nnz_y))
    Ax_o = case.jacobian_values(Vr)
    f_o = nro.mismatch(case, Vr, Y, sspec)
    scale = np.abs(Ax_o).max()
    assert np.abs(Ax.cpu().numpy() - Ax_o).max() <= 1e-12 * scale
    assert np.abs(b.cpu().numpy() + f_o).max() <= 1e-12 * np.abs(f_o).max()
    assert np.allclose(fnorm.cpu().numpy(), np.abs(f_o).max(axis=1), rtol=1e-12, atol=0)


@pytest.mark.gpu
def test_nr_solve_time_series_vs_oracle():
    from csparse3_b200.nr import NewtonPlan
    case, sym = _case(118)
    plan = NewtonPlan.from_case(case, sym)
    B = 21
    sspec, Vt = _targets(case, B)
    vm, va, fnorm, status = plan.solve_host(sspec, iters=5)
    arrays = (sym.q, sym.pinv, sym.Lp, sym.Li, sym.Up, sym.Ui)
    vm_o, va_o, fnorm_o, _ = nro.newton(case, arrays, sspec, 5)
    assert (status == 0).all()
    assert np.abs(vm - vm_o).max() <= 1e-9 and np.abs(va - va_o).max() <= 1e-9
    assert (fnorm < 1e-9).all() and (fnorm_o < 1e-9).all()
    # converged to the voltages the injections were computed from
    assert np.abs(vm * np.exp(1j * va) - Vt).max() < 1e-8
    assert np.array_equal(vm[:, case.pv], np.ones((B, len(case.pv)))) and (va[:, 0] == 0).all()


@pytest.mark.gpu
def test_nr_solve_outages_vs_oracle():
    """N-1: every case removes one non-bridge branch (explicit Ybus edit on the shared pattern)."""
    from csparse3_b200.nr import NewtonPlan
    case, sym = _case(118)
    plan = NewtonPlan.from_case(case, sym, outages=True)
    nb = case.non_bridge_branches()
    B = 12
    ob = nb[:B].astype(np.int32)
    Y = case.ybus_values(ob)
    sspec, Vt = _targets(case, B, Y)
    vm, va, fnorm, status = plan.solve_host(sspec, iters=5, out_branch=ob)
    arrays = (sym.q, sym.pinv, sym.Lp, sym.Li, sym.Up, sym.Ui)
    vm_o, va_o, fnorm_o, _ = nro.newton(case, arrays, sspec, 5, Y=Y)
    assert (status == 0).all()
    assert np.abs(vm - vm_o).max() <= 1e-9 and np.abs(va - va_o).max() <= 1e-9
    assert (fnorm < 1e-9).all()


@pytest.mark.gpu
def test_nr_device_api_and_per_case_start():
    import torch
    from csparse3_b200.nr import NewtonPlan
    case, sym = _case(118)
    plan = NewtonPlan.from_case(case, sym)
    B = 8
    sspec, Vt = _targets(case, B)
    rng = np.random.default_rng(3)
    vm0 = np.ones((B, case.n_bus)); va0 = rng.normal(0, 0.01, (B, case.n_bus)); va0[:, 0] = 0.0
    vm_h, va_h, fn_h, st_h = plan.solve_host(sspec, iters=4, vm0=vm0, va0=va0)
    vm_d, va_d = torch.as_tensor(vm0).cuda(), torch.as_tensor(va0).cuda()
    fn_d, st_d = plan.solve(vm_d, va_d, torch.as_tensor(sspec).cuda(), iters=4)
    torch.cuda.synchronize()
    assert np.array_equal(vm_d.cpu().numpy(), vm_h) and np.array_equal(va_d.cpu().numpy(), va_h)
    assert np.array_equal(fn_d.cpu().numpy(), fn_h) and (st_d.cpu().numpy() == 0).all()
