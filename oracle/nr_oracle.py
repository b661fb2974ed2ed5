"""CPU restatement of the Newton-Raphson power-flow iteration -- TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py's
CPU legs); the product never imports it.

The reference has no power-flow code: it supplies the pieces its consumer (GridCal, README.md:10) assembles -- the
2 x 2 block stacking pack_4_by_4 (src/CSparse3/csc.py:588-606) and the products with Ybus (csc.py:374-379).  This
file states the iteration those pieces serve, in numpy, with the LU of oracle/csp3_oracle.c: PARITY UNPINNED for the
iteration as a whole (no reference test pins it); the Jacobian formulas are MATPOWER's dSbus_dV, the same ones
csparse3_b200/synth.py generates the benchmark matrices with.
"""
import numpy as np

from . import oracle as orc


def s_calc(case, V, Y):
    """S = V conj(Ybus V) for V[B, N], Y[B, nnz_y] -> (S[B, N], I[B, N])"""
    YV = Y * V[:, case.yk]
    I = np.add.reduceat(YV, case.y_rowstart, axis=1)
    return V * np.conj(I), I


def mismatch(case, V, Y, sspec):
    """f[B, n] = (P_calc - P_spec over pvpq, Q_calc - Q_spec over pq)"""
    S, _ = s_calc(case, V, Y)
    return np.concatenate([S[:, case.pvpq].real, S[:, case.pq].imag], axis=1) - sspec


def newton(case, sym_arrays, sspec, iters, Y=None, vm0=None, va0=None):
    """`iters` iterations from (vm0, va0) (flat start by default) -> (vm, va, fnorm, history of max |f|)"""
    q, pinv, Lp, Li, Up, Ui = sym_arrays
    B = sspec.shape[0]
    N, n, npvpq = case.n_bus, case.n, len(case.pvpq)
    Y = case.ybus_values() if Y is None else Y
    Y = np.broadcast_to(Y, (B, case.nnz_y))
    vm = np.ones((B, N)) if vm0 is None else np.broadcast_to(vm0, (B, N)).copy()
    va = np.zeros((B, N)) if va0 is None else np.broadcast_to(va0, (B, N)).copy()
    hist = []
    for _ in range(iters):
        V = vm * np.exp(1j * va)
        f = mismatch(case, V, Y, sspec)
        hist.append(np.abs(f).max(axis=1))
        Ax = case.jacobian_values(V, Y)
        for k in range(B):
            Lx, Ux = orc.csc_lu_refactor(n, case.Ap, case.Ai, Ax[k], q, pinv, Lp, Li, Up, Ui)
            dx = orc.csc_lu_solve(n, Lp, Li, Lx, Up, Ui, Ux, pinv, q, -f[k])
            va[k, case.pvpq] += dx[:npvpq]
            vm[k, case.pq] += dx[npvpq:]
    V = vm * np.exp(1j * va)
    f = mismatch(case, V, Y, sspec)
    return vm, va, np.abs(f).max(axis=1), hist
