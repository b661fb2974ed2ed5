import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from csparse3_b200 import synth
from csparse3_b200.lu import LuSymbolic
rng = np.random.default_rng(1)
n, Ap, Ai, Ax = synth.laplacian_2d(int(sys.argv[1]) if len(sys.argv) > 1 else 9)
sym = LuSymbolic(n, Ap, Ai, Ax, order=1, tol=1e-3)
Axb = Ax[None, :] * rng.uniform(0.9, 1.1, (5, len(Ax))); bb = rng.standard_normal((5, n))
print("n", n, "wide", sym.wide_info, flush=True)
dA, db = torch.as_tensor(Axb).cuda(), torch.as_tensor(bb).cuda()
work = sym.workspace(5, "cuda"); st = torch.zeros(5, dtype=torch.int32, device="cuda")
try:
    sym.refactor_ws(dA, work, st); torch.cuda.synchronize(); print("refactor ok", st.cpu().numpy(), flush=True)
    x = sym.solve_ws(work, db); torch.cuda.synchronize(); print("solve ok", flush=True)
except Exception as e:
    print("FAILED:", str(e)[-200:], flush=True)
