"""SpMV / SpGEMM / transposition measurements on the GPU box (SURVEY.md section 8d byte models).

    python tools/bench_csc.py [--big]

Reports, per workload, kernel time (CUDA events), achieved algorithmic GB/s and the fraction of the measured HBM
peak, next to the CPU oracle timed on one host thread in the same run; checks the results against the oracle.
  * batched SpMV on the config-3 pattern (10,000 value sets, device plan)        8*nnz + 8*n + 8*m bytes / system
  * single SpMV on config 1 (2-D Laplacian n = 1e4) and, with --big, config 5 (3-D Laplacian n = 1e6)
                                                                                 12*nnz + 4*(n+1) + 8*n + 8*m bytes
  * SpGEMM A*A and transposition on the same matrices (host-buffer C-ABI calls: copies included)
  * batched Jacobian assembly [[A11, A12], [A21, A22]] on the config-3 pattern (Stack4Plan)   16*nnz bytes / system
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def timed(fn, iters=5):
    import torch
    fn(); torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(iters):
        ev0.record(); fn(); ev1.record(); torch.cuda.synchronize()
        best = min(best, ev0.elapsed_time(ev1))
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--big", action="store_true", help="include config 5 (3-D Laplacian, n = 1e6)")
    args = ap.parse_args()
    import torch
    from csparse3_b200 import csc_b200 as B
    from csparse3_b200 import synth
    from csparse3_b200.spmv import SpmvPlan
    from oracle import oracle as orc
    P = peak()
    out = []

    # ---- batched SpMV on the config-3 pattern ---------------------------------------------------------------
    g = synth.GridCase(2000)
    batch = 10000
    Ax, b = g.jacobian_batch(0, 256)
    reps = -(-batch // 256)
    dA = torch.as_tensor(np.tile(Ax, (reps, 1))[:batch]).cuda()
    dx = torch.as_tensor(np.tile(b, (reps, 1))[:batch]).cuda()
    plan = SpmvPlan(g.n, g.n, g.Ap, g.Ai)
    y = torch.empty_like(dx)
    ms = timed(lambda: plan.matvec(dA, dx, y))
    yo = orc.csc_mat_vec_ff(g.n, g.n, g.Ap, g.Ai, Ax[3], b[3])
    ok = bool(np.array_equal(y[3].cpu().numpy(), yo))
    t0 = time.perf_counter()
    for k in range(256):
        orc.csc_mat_vec_ff(g.n, g.n, g.Ap, g.Ai, Ax[k], b[k])
    cpu_ms = (time.perf_counter() - t0) * 1e3 / 256
    bytes_ = plan.bytes_per_system(True) * batch
    out.append({"op": "spmv_batched c3 pattern x10000", "ms": ms, "GBps": bytes_ / ms / 1e6, "frac": bytes_ / ms / 1e6 / P,
                "bit_exact": ok, "cpu_ms_per_system_1thread": cpu_ms, "gpu_ms_per_system": ms / batch})

    # ---- batched Jacobian assembly [[A11, A12], [A21, A22]] on the config-3 pattern -----------------------------
    import scipy.sparse as sp
    from csparse3_b200 import CscMat
    from csparse3_b200.assemble import Stack4Plan
    n = g.n
    J = sp.csc_matrix((np.arange(1, len(g.Ai) + 1, dtype=np.float64), g.Ai, g.Ap), shape=(n, n))
    n1 = n // 2
    blocks, vals = [], []
    for rs, cs in ((slice(0, n1), slice(0, n1)), (slice(0, n1), slice(n1, n)), (slice(n1, n), slice(0, n1)), (slice(n1, n), slice(n1, n))):
        S = sp.csc_matrix(J[rs, cs]); S.sort_indices()
        src = S.data.astype(np.int64) - 1
        blocks.append(CscMat(S.shape[0], S.shape[1], indptr=S.indptr.astype(np.int32), indices=S.indices.astype(np.int32), data=Ax[0][src].copy()))
        vals.append(np.ascontiguousarray(Ax[:, src]))
    sp4 = Stack4Plan(*blocks)
    dv = [torch.as_tensor(np.tile(v, (reps, 1))[:batch]).cuda() for v in vals]
    dout = torch.empty((batch, sp4.nnz), dtype=torch.float64, device="cuda")
    ms = timed(lambda: sp4.assemble(*dv, out=dout))
    ok = bool(np.array_equal(dout[:256].cpu().numpy(), Ax))
    t0 = time.perf_counter()
    for k in range(64):
        a = []
        for M, v in zip(blocks, vals):
            a += [M.m, M.n, M.indices, M.indptr, v[k]]
        orc.csc_stack_4_by_4_ff(*a)
    cpu_ms = (time.perf_counter() - t0) * 1e3 / 64
    bytes_ = sp4.bytes_per_system() * batch
    out.append({"op": "stack4 (Jacobian assembly) c3 pattern x10000", "ms": ms, "GBps": bytes_ / ms / 1e6, "frac": bytes_ / ms / 1e6 / P,
                "bit_exact": ok, "cpu_ms_per_system_1thread": cpu_ms, "gpu_ms_per_system": ms / batch})

    # ---- single-matrix SpMV / SpGEMM / transpose --------------------------------------------------------------
    cases = [("c1 lap2d n=1e4", synth.laplacian_2d(100))]
    if args.big:
        cases.append(("c5 lap3d n=1e6", synth.laplacian_3d(100)))
    for name, (n, Ap, Ai, Axm) in cases:
        nnz = int(Ap[n])
        x = np.random.default_rng(0).standard_normal(n)
        plan = SpmvPlan(n, n, Ap, Ai)
        dAx, dxx = torch.as_tensor(Axm).cuda(), torch.as_tensor(x).cuda()
        yy = torch.empty(n, dtype=torch.float64, device="cuda")
        ms = timed(lambda: plan.matvec(dAx, dxx, yy))
        t0 = time.perf_counter(); yo = orc.csc_mat_vec_ff(n, n, Ap, Ai, Axm, x); cpu_ms = (time.perf_counter() - t0) * 1e3
        bytes_ = plan.bytes_per_system(False)
        out.append({"op": "spmv " + name, "ms": ms, "GBps": bytes_ / ms / 1e6, "frac": bytes_ / ms / 1e6 / P,
                    "bit_exact": bool(np.array_equal(yy.cpu().numpy(), yo)), "cpu_ms_1thread": cpu_ms})
        # host-buffer C-ABI calls (copies included)
        t0 = time.perf_counter(); Cm, Cn, Cp, Ci, Cx, nnzC = B.csc_multiply_ff(n, n, Ap, Ai, Axm, n, n, Ap, Ai, Axm); gpu_ms = (time.perf_counter() - t0) * 1e3
        t0 = time.perf_counter(); Cm, Cn, Cp, Ci, Cx, nnzC = B.csc_multiply_ff(n, n, Ap, Ai, Axm, n, n, Ap, Ai, Axm); gpu_ms = min(gpu_ms, (time.perf_counter() - t0) * 1e3)
        t0 = time.perf_counter(); Om, On, Op, Oi, Ox, onnz = orc.csc_multiply_ff(n, n, Ap, Ai, Axm, n, n, Ap, Ai, Axm); cpu_ms = (time.perf_counter() - t0) * 1e3
        same_p = bool(np.array_equal(Cp, Op))
        # values after per-column sort of the oracle's first-touch order
        order = np.lexsort((Oi, np.repeat(np.arange(n), np.diff(Op))))
        same_v = bool(np.array_equal(Ci, Oi[order]) and np.array_equal(Cx, Ox[order]))
        bytes_ = 12 * (2 * nnz + nnzC) + 4 * (3 * n + 3)
        out.append({"op": "spgemm A*A " + name + " (host call incl. copies)", "ms": gpu_ms, "nnzC": nnzC, "GBps": bytes_ / gpu_ms / 1e6,
                    "Cp_exact": same_p, "values_exact": same_v, "cpu_ms_1thread": cpu_ms})
        # device-resident SpGEMM (the two-phase C-ABI with device pointers): numeric phase alone and both phases
        import ctypes as C
        from csparse3_b200 import _lib
        L = _lib.lib()
        dAp, dAi = torch.as_tensor(Ap).cuda(), torch.as_tensor(Ai).cuda()
        dCp = torch.empty(n + 1, dtype=torch.int32, device="cuda")
        nz = C.c_int64(0)
        stream = torch.cuda.current_stream().cuda_stream

        def symbolic():
            _lib.check(L.csp3_spgemm_symbolic(n, n, dAp.data_ptr(), dAi.data_ptr(), n, n, dAp.data_ptr(), dAi.data_ptr(),
                                              dCp.data_ptr(), C.byref(nz), stream), "spgemm symbolic")
        symbolic()
        dCi = torch.empty(nz.value, dtype=torch.int32, device="cuda")
        dCx = torch.empty(nz.value, dtype=torch.float64, device="cuda")

        def numeric():
            _lib.check(L.csp3_spgemm_numeric(n, n, dAp.data_ptr(), dAi.data_ptr(), dAx.data_ptr(), n, n, dAp.data_ptr(), dAi.data_ptr(),
                                             dAx.data_ptr(), dCp.data_ptr(), dCi.data_ptr(), dCx.data_ptr(), stream), "spgemm numeric")
        ms_num = timed(numeric)
        ms_both = timed(lambda: (symbolic(), numeric()))
        same_dev = bool(np.array_equal(dCp.cpu().numpy(), Cp) and np.array_equal(dCi.cpu().numpy(), Ci) and np.array_equal(dCx.cpu().numpy(), Cx))
        out.append({"op": "spgemm A*A " + name + " (device resident)", "ms_numeric": ms_num, "ms_symbolic_plus_numeric": ms_both, "nnzC": int(nz.value),
                    "GBps_numeric": bytes_ / ms_num / 1e6, "frac_numeric": bytes_ / ms_num / 1e6 / P, "GBps_both": bytes_ / ms_both / 1e6,
                    "equals_host_call": same_dev, "cpu_ms_1thread": cpu_ms})
        dTp = torch.empty(n + 1, dtype=torch.int32, device="cuda")
        dTi = torch.empty(nnz, dtype=torch.int32, device="cuda")
        dTx = torch.empty(nnz, dtype=torch.float64, device="cuda")
        ms_t = timed(lambda: _lib.check(L.csp3_csc_transpose(n, n, dAp.data_ptr(), dAi.data_ptr(), dAx.data_ptr(), dTp.data_ptr(),
                                                            dTi.data_ptr(), dTx.data_ptr(), stream), "transpose"))
        tbytes = 2 * (12 * nnz + 4 * (n + 1))
        t0 = time.perf_counter(); T = B.csc_transpose(n, n, Ap, Ai, Axm); gpu_ms = (time.perf_counter() - t0) * 1e3
        t0 = time.perf_counter(); To = orc.csc_transpose(n, n, Ap, Ai, Axm); cpu_ms = (time.perf_counter() - t0) * 1e3
        out.append({"op": "transpose " + name + " (device resident)", "ms": ms_t, "GBps": tbytes / ms_t / 1e6, "frac": tbytes / ms_t / 1e6 / P,
                    "exact": bool(np.array_equal(dTp.cpu().numpy(), T[2]) and np.array_equal(dTi.cpu().numpy(), T[3]) and np.array_equal(dTx.cpu().numpy(), T[4]))})
        out.append({"op": "transpose " + name + " (host call incl. copies)", "ms": gpu_ms,
                    "exact": bool(all(np.array_equal(u, v) for u, v in zip(T[2:], To[2:]))), "cpu_ms_1thread": cpu_ms})
    for r in out:
        print(json.dumps(r))


if __name__ == "__main__":
    main()
