"""Kernel tuning sweep (GPU box): times the refactor and solve kernels of one workload for several
(bundle width S, warps per CTA) settings.  The knobs are read once per process (env CSP3_RF_S, CSP3_RF_WARPS,
CSP3_SV_WARPS), so every setting runs in a child process; inputs are generated once and cached in /tmp.

    python tools/tune.py --workload c3 --batch 4096 --rf 4x2,4x1,2x2,1x2,8x2 --sv 4,8,2
"""
import argparse
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(args):
    import torch
    from csparse3_b200 import synth
    from csparse3_b200.lu import LuSymbolic
    import bench
    wl = bench.WORKLOADS[args.workload]
    case = synth.GridCase(wl["n_bus"])
    n, Ap, Ai, Ax0 = case.base_jacobian()
    sym = LuSymbolic(n, Ap, Ai, Ax0)
    cache = "/tmp/csp3_tune_%s.npz" % args.workload
    d = np.load(cache)
    reps = -(-args.batch // d["Ax"].shape[0])
    Ax = torch.as_tensor(np.tile(d["Ax"], (reps, 1))[:args.batch]).cuda()
    b = torch.as_tensor(np.tile(d["b"], (reps, 1))[:args.batch]).cuda()
    B = args.batch
    Lx = torch.empty((B, sym.lnz), dtype=torch.float64, device="cuda")
    Ux = torch.empty((B, sym.unz), dtype=torch.float64, device="cuda")
    x = torch.empty((B, n), dtype=torch.float64, device="cuda")
    st = torch.empty(B, dtype=torch.int32, device="cuda")
    work = sym.workspace(B, "cuda")
    ws = os.environ.get("CSP3_PATH", "ws") == "ws"

    def run_rf():
        if ws: sym.refactor_ws(Ax, work, st)
        else: sym.refactor(Ax, Lx, Ux, st)

    def run_sv():
        if ws: sym.solve_ws(work, b, x)
        else: sym.solve(Lx, Ux, b, x)
    for _ in range(2):
        run_rf(); run_sv()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    rf, sv = [], []
    for _ in range(args.iters):
        ev[0].record(); run_rf(); ev[1].record(); run_sv(); ev[2].record()
        torch.cuda.synchronize()
        rf.append(ev[0].elapsed_time(ev[1])); sv.append(ev[1].elapsed_time(ev[2]))
    ok = int(st.abs().max().item()) == 0
    # correctness spot check against cached oracle solutions
    xo = d["x"]
    same = bool(np.array_equal(x[:xo.shape[0]].cpu().numpy(), xo))
    peak = 6535.1
    brf = (8 * sym.nnz + 8 * sym.nnz_lu) * B; bsv = (8 * sym.nnz_lu + 16 * n) * B
    print(json.dumps({"cfg": os.environ.get("CSP3_CFG"), "batch": B,
                      "rf_ms": min(rf), "sv_ms": min(sv), "rf_frac": brf / (min(rf) * 1e-3) / 1e9 / peak,
                      "sv_frac": bsv / (min(sv) * 1e-3) / 1e9 / peak,
                      "sys_per_s": B / ((min(rf) + min(sv)) * 1e-3), "status_ok": ok, "bit_exact": same}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--cfg", default="ws:S=8;ws:S=4;sm:RFS=4,SVS=4")
    ap.add_argument("--child", action="store_true")
    args = ap.parse_args()
    if args.child:
        return child(args)
    cache = "/tmp/csp3_tune_%s.npz" % args.workload
    if not os.path.exists(cache):
        from csparse3_b200 import synth
        from csparse3_b200.lu import LuSymbolic
        from oracle import oracle as orc
        import bench
        wl = bench.WORKLOADS[args.workload]
        case = synth.GridCase(wl["n_bus"])
        n, Ap, Ai, Ax0 = case.base_jacobian()
        sym = LuSymbolic(n, Ap, Ai, Ax0)
        gen = case.outage_batch if wl["kind"] == "outage" else case.jacobian_batch
        Ax, b = gen(0, 256)
        x, bad = orc.csc_lu_refactor_solve_batch(n, Ap, Ai, sym.q, sym.pinv, sym.Lp, sym.Li, sym.Up, sym.Ui, Ax[:32], b[:32], 8)
        np.savez(cache, Ax=Ax, b=b, x=x)
    # --cfg "ws:S=8,WIN=1024,STAGE=128;ws:S=4;sm:RFS=4,SVS=4"
    for cfg in args.cfg.split(";"):
        path, _, kv = cfg.partition(":")
        env = dict(os.environ, CSP3_PATH=path, CSP3_CFG=cfg)
        names = {"S": "CSP3_WS_S", "WIN": "CSP3_RF_WIN", "STAGE": "CSP3_SV_STAGE", "RFS": "CSP3_RF_S", "SVS": "CSP3_SV_S",
                 "WIDE": "CSP3_WIDE", "WS": "CSP3_WIDE_S", "WR": "CSP3_WIDE_R", "WF": "CSP3_WIDE_F", "WB": "CSP3_WIDE_BUDGET", "WL": "CSP3_WIDE_LANE",
                 "GA": "CSP3_WIDE_GA", "PAIRS": "CSP3_WIDE_PAIRS", "RUN": "CSP3_WIDE_RUN", "SCHED": "CSP3_WIDE_SCHED", "ACC": "CSP3_WIDE_ACC"}
        for item in filter(None, kv.split(",")):
            k, v = item.split("=")
            env[names[k]] = v
        subprocess.call([sys.executable, os.path.abspath(__file__), "--child", "--workload", args.workload,
                         "--batch", str(args.batch), "--iters", str(args.iters)], env=env)


if __name__ == "__main__":
    main()
