#!/usr/bin/env python
"""Benchmark of the north-star hot path: batched fp64 LU refactor + solve of same-pattern CSC systems.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c4|c2]
                    [--scaling strong|weak] [--batch B] [--nr-iters I] [--no-cpu] [--no-secondary]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One JSON line on stdout (rank 0).  A "step" = one refactor+solve pass over the WHOLE batch of the workload
(config 3 of BASELINE.json by default: 10,000 value sets on the 2,000-bus Jacobian pattern in total, sharded
contiguously over the ranks -- strong scaling, as BASELINE.json's configs[2] words it), followed at N > 1 by the
gather of the solution shards on rank 0 (NCCL over NVLink), which is inside the timed region.  `--scaling weak` keeps
the workload's batch per GPU instead.  See DESIGN.md "Measurement" for the byte model.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "fp64 LU refactor+solve systems/sec"
UNIT = "systems/s"
WORKLOADS = {
    "c3": dict(n_bus=2000, batch=10000, kind="timeseries",
               name="config3: 10,000 same-pattern 2,000-bus NR Jacobians (time series), refactor+solve"),
    "c4": dict(n_bus=10000, batch=0, kind="outage",
               name="config4: N-1 sweep of the 10,000-bus NR Jacobian, one refactor+solve per non-bridge outage"),
    "c2": dict(n_bus=118, batch=16384, kind="timeseries",
               name="config2 pattern batched: 16,384 IEEE-118-shaped NR Jacobians, refactor+solve"),
}
ORDER, TOL = 1, 1e-3      # amd(A + A'), threshold partial pivoting: the product's defaults (csparse3_b200/lu.py)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def total_batch(case, wl):
    return len(case.non_bridge_branches()) if wl["kind"] == "outage" and not wl["batch"] else wl["batch"]


def gen_values(case, wl, start, count):
    """Synthetic value sets / right-hand sides of the global systems [start, start+count) (numpy)."""
    if wl["kind"] == "outage":
        nb = len(case.non_bridge_branches())
        ids = (start + np.arange(count)) % nb
        parts = [case.outage_batch(int(i), 1) for i in ids] if count <= 8 else None
        if parts is None:
            # contiguous runs of outage ids
            out_a, out_b, s = [], [], 0
            while s < count:
                i0 = int(ids[s]); c = min(count - s, nb - i0, 256)
                a, b = case.outage_batch(i0, c)
                out_a.append(a); out_b.append(b); s += c
            return np.concatenate(out_a), np.concatenate(out_b)
        return np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts])
    return case.jacobian_batch(start, count)


def make_inputs(case, wl, start, count, pinned):
    import torch
    Ax = torch.empty((count, case.nnz), dtype=torch.float64, pin_memory=pinned)
    b = torch.empty((count, case.n), dtype=torch.float64, pin_memory=pinned)
    for s in range(0, count, 256):
        c = min(256, count - s)
        a_np, b_np = gen_values(case, wl, start + s, c)
        Ax[s:s + c] = torch.from_numpy(a_np)
        b[s:s + c] = torch.from_numpy(b_np)
    return Ax, b


def nr_base_state(case):
    """Operating point the time series / the outage sweep moves around: the shared start vector of every case."""
    Vb = case.voltages([case.seed + 7])[0]
    return np.abs(Vb), np.angle(Vb)


def nr_targets(case, wl, start, count):
    """Specified injections of the global cases [start, start+count): S_calc at a seeded perturbation of the base state
    (angles +- N(0, 0.01) rad, PQ magnitudes x (1 + N(0, 0.005)); PV / slack magnitudes and the slack angle are the
    specified ones), so every case has a solution a few Newton steps away from the shared start.
    -> (sspec[count, n], out_branch[count] or None)   (numpy only)"""
    ids = start + np.arange(count)
    N = case.n_bus
    vmb, vab = nr_base_state(case)
    ob = None
    Y = case.ybus_values()
    if wl["kind"] == "outage":
        nb = case.non_bridge_branches()
        ob = nb[ids % len(nb)].astype(np.int32)
    sspec = np.empty((count, case.n))
    for s in range(0, count, 512):
        sl = slice(s, min(s + 512, count))
        c = sl.stop - sl.start
        vm = np.empty((c, N)); va = np.empty((c, N))
        for r, k in enumerate(ids[sl]):
            rng = np.random.default_rng(int(5000 + k))
            va[r] = vab + rng.normal(0.0, 0.01, N)
            vm[r] = vmb * (1.0 + rng.normal(0.0, 0.005, N))
        vm[:, case.pv] = vmb[case.pv]; vm[:, 0] = vmb[0]; va[:, 0] = vab[0]
        Vt = vm * np.exp(1j * va)
        Yb = case.ybus_values(ob[sl]) if ob is not None else np.broadcast_to(Y, (c, case.nnz_y))
        I = np.add.reduceat(Yb * Vt[:, case.yk], case.y_rowstart, axis=1)
        S = Vt * np.conj(I)
        sspec[sl] = np.concatenate([S[:, case.pvpq].real, S[:, case.pq].imag], axis=1)
    return sspec, ob


def cpu_refactor_solve(orc, sym_arrays, n, Ap, Ai, Ax, b, threads):
    """The oracle port (reference-style CPU path) over a sample: one C call, `threads` threads over independent
    systems.  Returns (seconds, x)."""
    q, pinv, Lp, Li, Up, Ui = sym_arrays
    t0 = time.perf_counter()
    x, bad = orc.csc_lu_refactor_solve_batch(n, Ap, Ai, q, pinv, Lp, Li, Up, Ui, Ax, b, threads)
    dt = time.perf_counter() - t0
    assert bad == 0
    return dt, x


def oracle_symbolic(orc, n, Ap, Ai, Ax0):
    """Symbolic phase with the ORACLE only (no product code): q = amd(A + A'), first factorisation with partial pivoting."""
    q = orc.csc_amd(ORDER, n, n, Ap, Ai)
    Lp, Li, Lx, Up, Ui, Ux, pinv = orc.csc_lu(n, Ap, Ai, Ax0, q, TOL)
    return (q, pinv, Lp, Li, Up, Ui), int(len(Li) + len(Ui) - n), int(orc.lu_refactor_flops(n, Lp, Up, Ui))


def run_reference(args, wl, rank, world):
    """--impl reference: the reference-style CPU path on the box's host cores (the reference has no LU, so this is the
    oracle port -- kind "port"), all host threads, bounded sample per step.  Imports nothing of the product but the
    synthetic generator (numpy only): the symbolic data comes from oracle.csc_amd + oracle.csc_lu."""
    if rank != 0:
        return
    from csparse3_b200 import synth
    from oracle import oracle as orc
    case = synth.GridCase(wl["n_bus"])
    n, Ap, Ai, Ax0 = case.base_jacobian()
    arrays, nnz_lu, flops = oracle_symbolic(orc, n, Ap, Ai, Ax0)
    cores = os.cpu_count() or 1
    per_sys = 1.1e-3 * (flops / 490190.0)
    total = total_batch(case, wl)
    sample = int(max(cores * 8, min(total, 6.0 * cores / max(per_sys, 1e-6))))
    Ax, b = gen_values(case, wl, 0, sample)
    sample = Ax.shape[0]
    for _ in range(args.warmup):
        cpu_refactor_solve(orc, arrays, n, Ap, Ai, Ax[:cores * 4], b[:cores * 4], cores)
    t = 0.0
    for _ in range(args.steps):
        dt, _x = cpu_refactor_solve(orc, arrays, n, Ap, Ai, Ax, b, cores)
        t += dt
    value = sample * args.steps / t
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["name"], "n": n, "nnz": int(Ap[n]), "nnz_lu": nnz_lu},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d systems per step, %d threads over the C oracle port of the "
                                       "reference-style CSparse refactor+solve (the reference ships no LU)" % (sample, cores)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def cuda_timed(fn, iters, flush=None):
    """Best-of-`iters` CUDA-event time of fn() in ms; flush() runs (untimed) before every timed call."""
    import torch
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(iters):
        if flush is not None:
            flush()
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def secondary_ops(peak):
    """The CSC kernels around the LU in the Newton loop, in the same run (SURVEY.md 8d byte models).  GPU: CUDA events,
    device resident, L2 flushed (256 MiB write) before every single-matrix iteration.  CPU: the oracle's C restatement
    of the reference kernel on one host thread (kind "port": /root/reference is not on the GPU box)."""
    import ctypes as C
    import torch
    from csparse3_b200 import _lib, synth
    from csparse3_b200.spmv import SpmvPlan
    from oracle import oracle as orc
    L = _lib.lib()
    out = []
    junk = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    flush = lambda: junk.fill_(1)
    stream = torch.cuda.current_stream().cuda_stream

    g = synth.GridCase(2000)
    batch = 10000
    Ax, b = g.jacobian_batch(0, 128)
    reps = -(-batch // 128)
    dA = torch.as_tensor(np.tile(Ax, (reps, 1))[:batch]).cuda()
    dx = torch.as_tensor(np.tile(b, (reps, 1))[:batch]).cuda()
    plan = SpmvPlan(g.n, g.n, g.Ap, g.Ai)
    y = torch.empty_like(dx)
    ms = cuda_timed(lambda: plan.matvec(dA, dx, y), 5)
    t0 = time.perf_counter()
    for k in range(128):
        yo = orc.csc_mat_vec_ff(g.n, g.n, g.Ap, g.Ai, Ax[k], b[k])
    cpu_ms = (time.perf_counter() - t0) * 1e3 / 128 * batch
    by = plan.bytes_per_system(True) * batch
    out.append({"op": "spmv batched (config-3 pattern x 10,000 value sets)", "ms": ms, "GBps": by / ms / 1e6, "frac": by / ms / 1e6 / peak,
                "bit_exact": bool(np.array_equal(y[127].cpu().numpy(), yo)), "cpu_ms": cpu_ms, "cpu_kind": "port, 1 thread, extrapolated from 128 systems"})
    del dA, dx, y

    n, Ap, Ai, Axm = synth.laplacian_3d(100)
    nnz = int(Ap[n])
    xv = np.random.default_rng(0).standard_normal(n)
    plan = SpmvPlan(n, n, Ap, Ai)
    dAx, dxx = torch.as_tensor(Axm).cuda(), torch.as_tensor(xv).cuda()
    yy = torch.empty(n, dtype=torch.float64, device="cuda")
    ms = cuda_timed(lambda: plan.matvec(dAx, dxx, yy), 5, flush)
    t0 = time.perf_counter(); yo = orc.csc_mat_vec_ff(n, n, Ap, Ai, Axm, xv); cpu_ms = (time.perf_counter() - t0) * 1e3
    by = plan.bytes_per_system(False)
    out.append({"op": "spmv (config 5: 3-D Laplacian n=1e6, L2 flushed)", "ms": ms, "GBps": by / ms / 1e6, "frac": by / ms / 1e6 / peak,
                "bit_exact": bool(np.array_equal(yy.cpu().numpy(), yo)), "cpu_ms": cpu_ms, "cpu_kind": "port, 1 thread"})
    dAp, dAi = torch.as_tensor(Ap).cuda(), torch.as_tensor(Ai).cuda()
    dCp = torch.empty(n + 1, dtype=torch.int32, device="cuda")
    nz = C.c_int64(0)

    def symbolic():
        _lib.check(L.csp3_spgemm_symbolic(n, n, dAp.data_ptr(), dAi.data_ptr(), n, n, dAp.data_ptr(), dAi.data_ptr(),
                                          dCp.data_ptr(), C.byref(nz), stream), "spgemm symbolic")
    symbolic()
    dCi = torch.empty(nz.value, dtype=torch.int32, device="cuda")
    dCx = torch.empty(nz.value, dtype=torch.float64, device="cuda")

    def numeric():
        _lib.check(L.csp3_spgemm_numeric(n, n, dAp.data_ptr(), dAi.data_ptr(), dAx.data_ptr(), n, n, dAp.data_ptr(), dAi.data_ptr(),
                                         dAx.data_ptr(), dCp.data_ptr(), dCi.data_ptr(), dCx.data_ptr(), stream), "spgemm numeric")
    ms_num = cuda_timed(numeric, 3, flush)
    ms_both = cuda_timed(lambda: (symbolic(), numeric()), 3, flush)
    t0 = time.perf_counter(); Om, On, Op, Oi, Ox, onnz = orc.csc_multiply_ff(n, n, Ap, Ai, Axm, n, n, Ap, Ai, Axm); cpu_ms = (time.perf_counter() - t0) * 1e3
    order = np.lexsort((Oi, np.repeat(np.arange(n), np.diff(Op))))
    exact = bool(np.array_equal(dCp.cpu().numpy(), Op) and np.array_equal(dCi.cpu().numpy(), Oi[order]) and np.array_equal(dCx.cpu().numpy(), Ox[order]))
    by = 12 * (2 * nnz + int(nz.value)) + 4 * (3 * n + 3)
    out.append({"op": "spgemm A*A (config 5, nnzC=%d, L2 flushed)" % nz.value, "ms": ms_both, "ms_numeric": ms_num, "GBps": by / ms_both / 1e6,
                "frac": by / ms_both / 1e6 / peak, "bit_exact": exact, "cpu_ms": cpu_ms, "cpu_kind": "port, 1 thread"})
    del dCi, dCx
    dTp = torch.empty(n + 1, dtype=torch.int32, device="cuda")
    dTi = torch.empty(nnz, dtype=torch.int32, device="cuda")
    dTx = torch.empty(nnz, dtype=torch.float64, device="cuda")
    ms = cuda_timed(lambda: _lib.check(L.csp3_csc_transpose(n, n, dAp.data_ptr(), dAi.data_ptr(), dAx.data_ptr(), dTp.data_ptr(),
                                                           dTi.data_ptr(), dTx.data_ptr(), stream), "transpose"), 3, flush)
    t0 = time.perf_counter(); To = orc.csc_transpose(n, n, Ap, Ai, Axm); cpu_ms = (time.perf_counter() - t0) * 1e3
    by = 2 * (12 * nnz + 4 * (n + 1))
    out.append({"op": "transpose / csc_to_csr (config 5, L2 flushed)", "ms": ms, "GBps": by / ms / 1e6, "frac": by / ms / 1e6 / peak,
                "bit_exact": bool(np.array_equal(dTp.cpu().numpy(), To[2]) and np.array_equal(dTi.cpu().numpy(), To[3]) and np.array_equal(dTx.cpu().numpy(), To[4])),
                "cpu_ms": cpu_ms, "cpu_kind": "port, 1 thread"})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--batch", type=int, default=0, help="systems in total (strong) or per GPU (weak); default: the workload's")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--nr-iters", type=int, default=4, help="Newton iterations per case in the end-to-end leg")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-secondary", action="store_true", help="skip the SpMV / SpGEMM / transposition lines")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    wl = dict(WORKLOADS[args.workload])
    if args.batch:
        wl["batch"] = args.batch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, wl, rank, world)

    import torch
    import torch.distributed as dist
    from csparse3_b200 import synth
    from csparse3_b200.dist import gather_to_root, shard_range
    from csparse3_b200.lu import LuSymbolic
    from csparse3_b200.nr import NewtonPlan

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    # ---- setup (untimed): pattern, symbolic phase, synthetic values of this rank's shard -------------------------
    t_setup = time.perf_counter()
    case = synth.GridCase(wl["n_bus"])
    n, Ap, Ai, Ax0 = case.base_jacobian()
    sym = LuSymbolic(n, Ap, Ai, Ax0, order=ORDER, tol=TOL)
    per_gpu = total_batch(case, wl)
    total = per_gpu * world if args.scaling == "weak" else per_gpu
    start, stop = shard_range(total, rank, world)
    B = stop - start
    Ax_h, b_h = make_inputs(case, wl, start, B, pinned=True)
    x_h = torch.empty((B, n), dtype=torch.float64, pin_memory=True)
    st_h = torch.empty(B, dtype=torch.int32, pin_memory=True)
    Ax_d, b_d = Ax_h.to(dev), b_h.to(dev)
    work_d = sym.workspace(B, dev)        # factors: written once by the refactor kernel, read once by the solve kernel
    x_d = torch.empty((B, n), dtype=torch.float64, device=dev)
    st_d = torch.empty(B, dtype=torch.int32, device=dev)
    log("[rank %d] setup %.1fs: n=%d nnzA=%d nnzLU=%d flops=%d levels=%d systems [%d, %d) of %d (%s scaling)" %
        (rank, time.perf_counter() - t_setup, n, sym.nnz, sym.nnz_lu, sym.flops, sym.nlev_refactor, start, stop, total, args.scaling))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)      # started before the warm-up: nvidia-smi needs a moment to deliver its first line
    sampler.start()
    x_all = None
    for _ in range(args.warmup):
        sym.refactor_ws(Ax_d, work_d, st_d)
        sym.solve_ws(work_d, b_d, x_d)
        if world > 1:
            x_all = gather_to_root(x_d, total)          # the first call also pays the NCCL communicator set-up
    barrier()

    # ---- timed region: K steps, CUDA events on the launching (torch current) stream ------------------------------
    K = args.steps
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(K)]
    barrier()
    sampler.lines.clear()                   # keep only what is sampled from here on (the timed region)
    for k in range(K):
        ev[k][0].record()
        sym.refactor_ws(Ax_d, work_d, st_d)
        ev[k][1].record()
        sym.solve_ws(work_d, b_d, x_d)
        ev[k][2].record()
        if world > 1:
            x_all = gather_to_root(x_d, total)
        ev[k][3].record()
    barrier()
    clocks = sampler.stop()
    total_ms = ev[0][0].elapsed_time(ev[K - 1][3])
    rf_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in ev]))
    sv_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in ev]))
    ga_ms = float(np.mean([e[2].elapsed_time(e[3]) for e in ev]))
    assert int(st_d.abs().max().item()) == 0, "a system reported a zero pivot"

    # ---- shard invariance on hardware: rank 0 recomputes systems owned by the other ranks, bits must agree ----------
    shard_check = None
    if world > 1 and rank == 0:
        ids = sorted(set(int(i) for r in range(world) for s0, s1 in [shard_range(total, r, world)] if s1 > s0
                         for i in (list(range(s0, min(s0 + 12, s1))) + list(range(max(s0, s1 - 5), s1)))))
        runs, cur = [], [ids[0], ids[0]]
        for i in ids[1:]:
            if i == cur[1] + 1:
                cur[1] = i
            else:
                runs.append(cur); cur = [i, i]
        runs.append(cur)
        ok, cnt = True, 0
        for a, bnd in runs:
            c = bnd - a + 1
            a_np, b_np = gen_values(case, wl, a, c)
            xs, ss = sym.refactor_solve(torch.as_tensor(a_np).to(dev), torch.as_tensor(b_np).to(dev))
            ok = ok and bool(torch.equal(xs, x_all[a:a + c])) and int(ss.abs().max().item()) == 0
            cnt += c
        shard_check = {"systems_recomputed_on_rank0": cnt, "bit_identical_to_owner_rank": ok}
        assert ok, "results depend on the shard a system was computed in"

    # ---- end to end, host buffers -> host buffers ----------------------------------------------------------------------
    # (1) the Newton-Raphson entry point: per case only the specified injections travel to the device (n doubles), the
    #     Jacobian is evaluated there, `nr_iters` refactor+solve passes run per case, (vm, va) come back.
    plan = NewtonPlan.from_case(case, sym, outages=wl["kind"] == "outage")
    sspec_np, ob_np = nr_targets(case, wl, start, B)
    sspec_h = torch.empty((B, n), dtype=torch.float64, pin_memory=True); sspec_h.copy_(torch.from_numpy(sspec_np))
    vm_h = torch.empty((B, case.n_bus), dtype=torch.float64, pin_memory=True)
    va_h = torch.empty((B, case.n_bus), dtype=torch.float64, pin_memory=True)
    fn_h = torch.empty(B, dtype=torch.float64, pin_memory=True)
    nst_h = torch.empty(B, dtype=torch.int32, pin_memory=True)
    vm0, va0 = nr_base_state(case)
    nr_args = dict(iters=args.nr_iters, out_branch=ob_np, vm0=vm0, va0=va0, vm=vm_h.numpy(), va=va_h.numpy(), fnorm=fn_h.numpy(), status=nst_h.numpy())
    warm = min(B, 512)
    plan.solve_host(sspec_h.numpy()[:warm], iters=1, out_branch=None if ob_np is None else ob_np[:warm], vm0=vm0, va0=va0)       # staging set-up
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        plan.solve_host(sspec_h.numpy(), **nr_args)
    torch.cuda.synchronize()
    nr_s = (time.perf_counter() - t0) / args.e2e_steps
    assert (nst_h.numpy() == 0).all(), "Newton-Raphson: a case reported a bad pivot"
    nr_fnorm = float(fn_h.numpy().max())
    assert np.isfinite(nr_fnorm) and nr_fnorm < 0.1, "Newton-Raphson diverged (max mismatch %g)" % nr_fnorm
    # (2) the plain LU host call (values of every Jacobian over PCIe), for comparison
    Ax_np, b_np, x_np, st_np = Ax_h.numpy(), b_h.numpy(), x_h.numpy(), st_h.numpy()
    sym.refactor_solve_host(Ax_np[:warm], b_np[:warm], x_np[:warm], st_np[:warm])
    barrier()
    t0 = time.perf_counter()
    sym.refactor_solve_host(Ax_np, b_np, x_np, st_np)
    torch.cuda.synchronize()
    lu_s = time.perf_counter() - t0
    same = bool(np.array_equal(x_np, x_d.cpu().numpy()))
    assert same and (st_np == 0).all(), "host-buffer path and device path disagree"

    # ---- small batches on this GPU (one system; 1/8 of the workload = one rank's shard at N = 8): the refactorisation picks
    # ---- its kernel from the batch (several warps per bundle for these), device resident, results compared with the big run
    small = []
    if world == 1:
        for nb in sorted({min(B, 1), min(B, 8), max(1, -(-B // 8))}):
            As, bs = Ax_d[:nb].contiguous(), b_d[:nb].contiguous()
            ws_ = sym.workspace(nb, dev); xs = torch.empty((nb, n), dtype=torch.float64, device=dev)
            ss = torch.empty(nb, dtype=torch.int32, device=dev)
            for _ in range(3):
                sym.refactor_ws(As, ws_, ss); sym.solve_ws(ws_, bs, xs)
            ev3 = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            reps = 10
            torch.cuda.synchronize()
            rf_t = sv_t = 0.0
            for _ in range(reps):
                ev3[0].record(); sym.refactor_ws(As, ws_, ss); ev3[1].record(); sym.solve_ws(ws_, bs, xs); ev3[2].record()
                torch.cuda.synchronize()
                rf_t += ev3[0].elapsed_time(ev3[1]); sv_t += ev3[1].elapsed_time(ev3[2])
            small.append({"systems": nb, "refactor_ms": rf_t / reps, "solve_ms": sv_t / reps, "systems_per_s": nb / ((rf_t + sv_t) / reps * 1e-3),
                          "refactor_kernel": sym.refactor_kernel_name(nb),
                          "bit_identical_to_full_batch": bool(torch.equal(xs, x_d[:nb])) and int(ss.abs().max().item()) == 0})
            del As, bs, ws_, xs, ss

    # ---- weak-scaling point in the same run (the workload's batch per GPU), for the scaling curve's other reading -------
    weak = None
    if world > 1 and args.scaling == "strong":
        reps = -(-per_gpu // max(B, 1))
        Aw = Ax_d.repeat(reps, 1)[:per_gpu].contiguous(); bw = b_d.repeat(reps, 1)[:per_gpu].contiguous()
        ww = sym.workspace(per_gpu, dev); xw = torch.empty((per_gpu, n), dtype=torch.float64, device=dev)
        sw = torch.empty(per_gpu, dtype=torch.int32, device=dev)
        for _ in range(2):
            sym.refactor_ws(Aw, ww, sw); sym.solve_ws(ww, bw, xw)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            sym.refactor_ws(Aw, ww, sw); sym.solve_ws(ww, bw, xw)
        e1.record()
        barrier()
        weak = e0.elapsed_time(e1) / 3
        del Aw, bw, ww, xw

    # ---- reduce timings: max over ranks ---------------------------------------------------------------------------
    t = torch.tensor([total_ms, rf_ms, sv_ms, ga_ms, nr_s * 1e3, lu_s * 1e3, weak or 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, rf_ms, sv_ms, ga_ms, nr_ms, lu_ms, weak_ms = t.tolist()

    if rank == 0:
        peak, peak_src = peaks()
        Bmax = -(-total // world)                                   # the largest shard bounds the step
        bytes_rf = (8 * sym.nnz + 8 * sym.nnz_lu) * Bmax           # read A, write L+U
        bytes_sv = (8 * sym.nnz_lu + 16 * n) * Bmax                # read L+U, read b, write x
        wide = sym.wide_width > 0 and os.environ.get("CSP3_WIDE", "1") != "0"
        rf_name = sym.refactor_kernel_name(Bmax)
        sv_name = "lu_sweep_wide_kernel" if wide else "lu_solve_kernel"
        dom = (rf_name, bytes_rf, rf_ms) if rf_ms >= sv_ms else (sv_name, bytes_sv, sv_ms)
        achieved = dom[1] / (dom[2] * 1e-3) / 1e9
        step_bytes = sym.bytes_per_system() * total
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                ent = json.load(open(tpath)).get(args.workload, {}).get(dom[0])
                if ent:
                    traffic = ent["dram_bytes_per_system"] * Bmax
            except Exception:
                pass
        value = total * K / (total_ms * 1e-3)
        cfg = {"workload": wl["name"], "n": n, "nnz": sym.nnz, "nnz_lu": sym.nnz_lu, "batch_total": total, "batch_per_gpu": Bmax,
               "refactor_flops_per_system": sym.flops, "levels": sym.nlev_refactor,
               "l2": "inputs larger than L2 (%.2f GB of values per step, nothing reused across steps)" % (step_bytes / 1e9),
               "parallelism": "batch-sharded x%d, no collective in refactor/solve%s" % (world, "; gather of x to rank 0 inside the timed step" if world > 1 else ""),
               "small_batches": small,
               "kernel_ms": {rf_name: rf_ms, (sv_name + " x2 + rhs_to_bundles_kernel + bundles_to_x_kernel") if wide else sv_name: sv_ms},
               "result_gather_ms": ga_ms if world > 1 else None,
               "bundle": ("%d systems per warp, 2 per lane" % sym.wide_width) if wide else "v3 kernels",
               "bytes_per_system": sym.bytes_per_system(),
               "step_roofline_frac": step_bytes / world / (total_ms / K * 1e-3) / 1e9 / peak,
               "e2e_lu_host": {"value": total / (lu_ms * 1e-3), "unit": UNIT, "ms_per_step": lu_ms,
                               "h2d_bytes_per_step": int(total * (sym.nnz + n) * 8), "d2h_bytes_per_step": int(total * (n * 8 + 4)),
                               "api": "csp3_lu_refactor_solve_host (every Jacobian value over PCIe)"},
               "nr": {"iters_per_case": args.nr_iters, "cases": total, "max_mismatch_after": nr_fnorm, "ms_per_call": nr_ms}}
        if shard_check:
            cfg["shard_invariance"] = shard_check
        if weak_ms:
            cfg["weak_scaling_point"] = {"batch_per_gpu": per_gpu, "ms_per_step": weak_ms, "value": per_gpu * world / (weak_ms * 1e-3)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": cfg,
            "roofline": {"bound": "hbm", "kernel": dom[0], "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": dom[1]},
            # every Newton iteration of every case is one refactor+solve of a new matrix: cases x iterations systems per call
            "e2e": {"value": total * args.nr_iters / (nr_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": int(total * n * 8 + (0 if ob_np is None else total * 4) + world * case.n_bus * 16),
                    "d2h_bytes_per_step": int(total * (case.n_bus * 16 + 12)),
                    "ms_per_step": nr_ms,
                    "api": "csp3_nr_solve_host (NewtonPlan.solve_host): %d Newton iterations per case, each one Jacobian "
                           "evaluation + refactor + solve on the device; host buffers in and out" % args.nr_iters},
            # refactor + (transpose in, forward sweep, backward sweep, transpose out); v3: refactor + solve
            "gpu_launches": (5 if wide else 2) * K,
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu:
            from oracle import oracle as orc
            arrays = (sym.q, sym.pinv, sym.Lp, sym.Li, sym.Up, sym.Ui)
            per_sys = 1.1e-3 * (sym.flops / 490190.0)
            sample = int(max(64, min(B, 12.0 / max(per_sys, 1e-6))))
            cpu_refactor_solve(orc, arrays, n, Ap, Ai, Ax_np[:8], b_np[:8], 1)
            dt, x_cpu = cpu_refactor_solve(orc, arrays, n, Ap, Ai, Ax_np[:sample], b_np[:sample], 1)
            exact = bool(np.array_equal(x_cpu, x_np[:sample]))
            rel = float(np.max(np.linalg.norm(x_cpu - x_np[:sample], axis=1) / np.linalg.norm(x_cpu, axis=1)))
            assert rel <= 1e-9, "GPU result differs from the oracle by more than 1e-9"
            line["cpu_baseline"] = {"value": sample / dt, "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": "first %d systems of the same batch, single thread, C oracle port of the reference-style "
                                              "CSparse refactor+solve (GPU result %s)" %
                                              (sample, "bit-identical" if exact else "within %.1e" % rel)}
        if world == 1 and not args.no_secondary and args.workload == "c3":
            del Ax_d, b_d, work_d, x_d
            torch.cuda.empty_cache()
            try:
                line["config"]["secondary"] = secondary_ops(peak)
            except Exception as e:                                    # never lose the headline line to a side measurement
                line["config"]["secondary"] = [{"error": repr(e)}]
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
