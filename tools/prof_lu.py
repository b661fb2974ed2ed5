"""Small driver for ncu captures: runs the workspace-path refactor and solve kernels of one workload a few times.

    python tools/prof_lu.py --workload c3 --batch 10000 --iters 3
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--batch", type=int, default=10000)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--distinct", type=int, default=64)
    ap.add_argument("--check", type=int, default=4, help="systems compared with the oracle (0: none)")
    args = ap.parse_args()
    import torch
    import bench
    from csparse3_b200 import synth
    from csparse3_b200.lu import LuSymbolic
    wl = bench.WORKLOADS[args.workload]
    case = synth.GridCase(wl["n_bus"])
    n, Ap, Ai, Ax0 = case.base_jacobian()
    sym = LuSymbolic(n, Ap, Ai, Ax0)
    gen = case.outage_batch if wl["kind"] == "outage" else case.jacobian_batch
    Ax, b = gen(0, args.distinct)
    reps = -(-args.batch // args.distinct)
    dA = torch.as_tensor(np.tile(Ax, (reps, 1))[:args.batch]).cuda()
    db = torch.as_tensor(np.tile(b, (reps, 1))[:args.batch]).cuda()
    x = torch.empty_like(db)
    st = torch.empty(args.batch, dtype=torch.int32, device="cuda")
    work = sym.workspace(args.batch, "cuda")
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    for it in range(args.iters):
        ev[0].record(); sym.refactor_ws(dA, work, st); ev[1].record(); sym.solve_ws(work, db, x); ev[2].record()
        torch.cuda.synchronize()
        print("iter %d: refactor %.3f ms, solve %.3f ms" % (it, ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])), flush=True)
    assert int(st.abs().max().item()) == 0
    if args.check:
        from oracle import oracle as orc
        xs = x[:args.check].cpu().numpy()
        worst = 0.0
        exact = True
        for k in range(args.check):
            Lx, Ux = orc.csc_lu_refactor(n, Ap, Ai, Ax[k], sym.q, sym.pinv, sym.Lp, sym.Li, sym.Up, sym.Ui)
            xo = orc.csc_lu_solve(n, sym.Lp, sym.Li, Lx, sym.Up, sym.Ui, Ux, sym.pinv, sym.q, b[k])
            exact = exact and np.array_equal(xs[k], xo)
            worst = max(worst, float(np.linalg.norm(xs[k] - xo) / np.linalg.norm(xo)))
        print("check vs oracle: bit-exact %s, worst relative difference %.3e" % (exact, worst), flush=True)


if __name__ == "__main__":
    main()
