import sys, torch, numpy as np
sys.path.insert(0, '.')
from csparse3_b200 import synth
from csparse3_b200.lu import LuSymbolic
for nb in (118, 2000):
    g = synth.GridCase(nb); n, Ap, Ai, Ax0 = g.base_jacobian(); sym = LuSymbolic(n, Ap, Ai, Ax0)
    Ax, b = g.jacobian_batch(0, 8)
    for B in (1, 8):
        dA, db = torch.as_tensor(Ax[:B]).cuda(), torch.as_tensor(b[:B]).cuda()
        work = sym.workspace(B, "cuda"); st = torch.empty(B, dtype=torch.int32, device="cuda"); x = torch.empty((B, n), dtype=torch.float64, device="cuda")
        for _ in range(3): sym.refactor_ws(dA, work, st); sym.solve_ws(work, db, x)
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        best = (1e9, 1e9)
        for _ in range(10):
            e[0].record(); sym.refactor_ws(dA, work, st); e[1].record(); sym.solve_ws(work, db, x); e[2].record(); torch.cuda.synchronize()
            t = (e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]))
            if sum(t) < sum(best): best = t
        print("LAT n_bus=%d n=%d batch=%d refactor %.1f us solve %.1f us" % (nb, n, B, best[0]*1e3, best[1]*1e3))
