"""Randomised CPU campaign for the row-lane program compiler: random sparse patterns (symmetric and unsymmetric), orderings,
pivot tolerances, warps per bundle and stage sizes through tests/rowlane_interp.py (randomly interleaved warps), factors
compared with the oracle bit for bit.  python tools/fuzz_rowlane.py  (6 x 5 runs of 40 matrices, a few minutes)."""
import sys, os, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)
import numpy as np, scipy.sparse as sp
import rowlane_interp as ri
from csparse3_b200.lu import LuSymbolic
from oracle import oracle as orc
rng = np.random.default_rng(int(sys.argv[1]))
bad = 0
for t in range(40):
    n = int(rng.integers(2, 260))
    dens = float(rng.choice([1.5, 3.0, 6.0, 12.0])) / n
    A = sp.random(n, n, density=min(1.0, dens), random_state=int(rng.integers(1 << 30)), format="csc")
    if rng.random() < 0.5: A = A + A.T          # structurally symmetric like power-flow Jacobians
    A = sp.csc_matrix(A + sp.diags(rng.uniform(0.5, 2.0, n)))
    A.sort_indices()
    Ap, Ai, Ax = A.indptr.astype(np.int32), A.indices.astype(np.int32), A.data.copy()
    order = int(rng.choice([0, 1, 2, 3])); tol = float(rng.choice([1e-3, 0.1, 1.0]))
    try:
        sym = LuSymbolic(n, Ap, Ai, Ax, order=order, tol=tol)
    except Exception as e:
        continue
    Axb = Ax[None, :] * rng.uniform(0.9, 1.1, (2, len(Ax)))
    Lx, Ux, fail, stats = ri.run_refactor(sym, Axb, seed=int(rng.integers(1 << 30)))
    for k in range(2):
        try:
            L, U = orc.csc_lu_refactor(n, Ap, Ai, Axb[k], sym.q, sym.pinv, sym.Lp, sym.Li, sym.Up, sym.Ui)
        except ArithmeticError:
            assert fail[k] != 0; continue
        assert fail[k] == 0 and np.array_equal(Lx[k], L) and np.array_equal(Ux[k], U), (n, order, tol)
    assert stats["ops"] * 2 == sym.flops
print("ok")
''' % (ROOT, ROOT + '/tests')
fails = 0
for seed in range(6):
    for W, NQ in ((2, 1), (4, 2), (8, 1), (8, 3), (2, 3)):
        env = dict(os.environ, CSP3_RL_W=str(W), CSP3_RL_NQ=str(NQ))
        if seed % 2: env["CSP3_RL_MARGIN"] = "-100000"
        out = subprocess.run([sys.executable, "-c", code, str(seed * 10 + W)], env=env, capture_output=True, text=True)
        ok = out.returncode == 0 and out.stdout.strip().endswith("ok")
        print(seed, W, NQ, "ok" if ok else "FAIL " + out.stderr[-600:], flush=True)
        fails += not ok
print("fails", fails)
