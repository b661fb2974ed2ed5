#!/usr/bin/env python
"""Benchmark of the north-star hot path: batched fp64 LU refactor + solve of same-pattern CSC systems.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c4|c2] [--batch B]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One JSON line on stdout (rank 0).  A "step" = one refactor+solve pass over the rank's batch of synthetic
systems (config 3 of BASELINE.json by default: 10,000 value sets on the 2,000-bus Jacobian pattern PER GPU,
weak scaling, no collective in the data path).  See DESIGN.md "Measurement" for the byte model.
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
               name="config3: 10,000 same-pattern 2,000-bus NR Jacobians (time series) per GPU, refactor+solve"),
    "c4": dict(n_bus=10000, batch=2048, kind="outage",
               name="config4: N-1 outages of the 10,000-bus NR Jacobian, 2,048 systems per GPU, refactor+solve"),
    "c2": dict(n_bus=118, batch=16384, kind="timeseries",
               name="config2 pattern batched: 16,384 IEEE-118-shaped NR Jacobians per GPU, refactor+solve"),
}


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
                                          "--format=csv,noheader,nounits", "-lms", "100"],
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


def make_inputs(case, wl, start, count, pinned):
    """Synthetic value sets / right-hand sides of systems [start, start+count) into (pinned) host tensors."""
    import torch
    Ax = torch.empty((count, case.nnz), dtype=torch.float64, pin_memory=pinned)
    b = torch.empty((count, case.n), dtype=torch.float64, pin_memory=pinned)
    gen = case.outage_batch if wl["kind"] == "outage" else case.jacobian_batch
    nb = len(case.non_bridge_branches()) if wl["kind"] == "outage" else None
    step = 256
    for s in range(0, count, step):
        c = min(step, count - s)
        s0 = (start + s) % (nb - c) if nb else start + s
        a_np, b_np = gen(s0, c)
        Ax[s:s + c] = torch.from_numpy(a_np)
        b[s:s + c] = torch.from_numpy(b_np)
    return Ax, b


def cpu_refactor_solve(orc, sym_arrays, n, Ap, Ai, Ax, b, threads):
    """The oracle port (reference-style CPU path) over a sample: one C call, `threads` OpenMP threads over
    independent systems.  Returns (seconds, x)."""
    q, pinv, Lp, Li, Up, Ui = sym_arrays
    t0 = time.perf_counter()
    x, bad = orc.csc_lu_refactor_solve_batch(n, Ap, Ai, q, pinv, Lp, Li, Up, Ui, Ax, b, threads)
    dt = time.perf_counter() - t0
    assert bad == 0
    return dt, x


def run_reference(args, wl, rank, world):
    """--impl reference: the reference-style CPU path on the box's host cores (the reference has no LU, so
    this is the oracle port -- kind "port"), all host threads, bounded sample per step."""
    if rank != 0:
        return
    from csparse3_b200 import synth
    from csparse3_b200.lu import LuSymbolic
    from oracle import oracle as orc
    case = synth.GridCase(wl["n_bus"])
    n, Ap, Ai, Ax0 = case.base_jacobian()
    sym = LuSymbolic(n, Ap, Ai, Ax0)          # host symbolic phase only (no CUDA call)
    arrays = (sym.q, sym.pinv, sym.Lp, sym.Li, sym.Up, sym.Ui)
    cores = os.cpu_count() or 1
    per_sys = 1.1e-3 * (sym.flops / 490190.0)
    sample = int(max(cores * 8, min(wl["batch"], 6.0 * cores / max(per_sys, 1e-6))))
    gen = case.outage_batch if wl["kind"] == "outage" else case.jacobian_batch
    Ax, b = gen(0, sample)
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
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["name"], "n": n, "nnz": sym.nnz, "nnz_lu": sym.nnz_lu},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d systems per step, %d threads over the C oracle port of the "
                                       "reference-style CSparse refactor+solve (the reference ships no LU)" % (sample, cores)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="systems per GPU (default: the workload's)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
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
    from csparse3_b200.dist import gather_solutions
    from csparse3_b200.lu import LuSymbolic

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    # ---- setup (untimed): pattern, symbolic phase, synthetic values ---------------------------------------------
    t_setup = time.perf_counter()
    case = synth.GridCase(wl["n_bus"])
    n, Ap, Ai, Ax0 = case.base_jacobian()
    sym = LuSymbolic(n, Ap, Ai, Ax0)
    B = wl["batch"]
    Ax_h, b_h = make_inputs(case, wl, rank * B, B, pinned=True)
    x_h = torch.empty((B, n), dtype=torch.float64, pin_memory=True)
    st_h = torch.empty(B, dtype=torch.int32, pin_memory=True)
    Ax_d, b_d = Ax_h.to(dev), b_h.to(dev)
    work_d = sym.workspace(B, dev)        # factors: written once by the refactor kernel, read once by the solve kernel
    x_d = torch.empty((B, n), dtype=torch.float64, device=dev)
    st_d = torch.empty(B, dtype=torch.int32, device=dev)
    log("[rank %d] setup %.1fs: n=%d nnzA=%d nnzLU=%d flops=%d levels=%d batch=%d" %
        (rank, time.perf_counter() - t_setup, n, sym.nnz, sym.nnz_lu, sym.flops, sym.nlev_refactor, B))

    def step():
        sym.refactor_ws(Ax_d, work_d, st_d)
        sym.solve_ws(work_d, b_d, x_d)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()

    # ---- timed region: K steps, CUDA events on the launching (torch current) stream ------------------------------
    K = args.steps
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    for k in range(K):
        ev[k][0].record()
        sym.refactor_ws(Ax_d, work_d, st_d)
        ev[k][1].record()
        sym.solve_ws(work_d, b_d, x_d)
        ev[k][2].record()
    barrier()
    clocks = sampler.stop()
    total_ms = ev[0][0].elapsed_time(ev[K - 1][2])
    rf_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in ev]))
    sv_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in ev]))
    assert int(st_d.abs().max().item()) == 0, "a system reported a zero pivot"

    # ---- end to end through the host-buffer C ABI (pinned host -> H2D -> kernels -> D2H), wall clock -------------
    Ax_np, b_np, x_np, st_np = Ax_h.numpy(), b_h.numpy(), x_h.numpy(), st_h.numpy()
    sym.refactor_solve_host(Ax_np[:min(B, 512)], b_np[:min(B, 512)], x_np[:min(B, 512)], st_np[:min(B, 512)])   # warm staging
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        sym.refactor_solve_host(Ax_np, b_np, x_np, st_np)
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / args.e2e_steps
    same = bool(np.array_equal(x_np, x_d.cpu().numpy()))
    assert same and (st_np == 0).all(), "host-buffer path and device path disagree"

    # ---- final result gather (NCCL all-gather over NVLink), reported separately ----------------------------------
    gather_ms = None
    if world > 1:
        gather_solutions(x_d, B * world)               # first call pays NCCL communicator set-up: not timed
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        xg = gather_solutions(x_d, B * world)
        g1.record()
        torch.cuda.synchronize()
        gather_ms = g0.elapsed_time(g1)
        assert xg.shape[0] == B * world

    # ---- reduce timings: max over ranks ---------------------------------------------------------------------------
    t = torch.tensor([total_ms, rf_ms, sv_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, rf_ms, sv_ms, e2e_ms = t.tolist()

    if rank == 0:
        peak, peak_src = peaks()
        bytes_rf = (8 * sym.nnz + 8 * sym.nnz_lu) * B          # read A, write L+U
        bytes_sv = (8 * sym.nnz_lu + 16 * n) * B               # read L+U, read b, write x
        wide = sym.wide_width > 0 and os.environ.get("CSP3_WIDE", "1") != "0"
        rf_name = "lu_refactor_wide_kernel" if wide else "lu_refactor_kernel"
        sv_name = "lu_sweep_wide_kernel" if wide else "lu_solve_kernel"
        dom = (rf_name, bytes_rf, rf_ms) if rf_ms >= sv_ms else (sv_name, bytes_sv, sv_ms)
        achieved = dom[1] / (dom[2] * 1e-3) / 1e9
        step_bytes = sym.bytes_per_system() * B
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                tj = json.load(open(tpath))
                ent = tj.get(args.workload, {}).get(dom[0])
                if ent:
                    traffic = ent["dram_bytes_per_system"] * B
            except Exception:
                pass
        value = world * B * K / (total_ms * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["name"], "n": n, "nnz": sym.nnz, "nnz_lu": sym.nnz_lu, "batch_per_gpu": B,
                       "refactor_flops_per_system": sym.flops, "levels": sym.nlev_refactor,
                       "l2": "inputs larger than L2 (%.2f GB of values per step, nothing reused across steps)" % (step_bytes / 1e9),
                       "parallelism": "batch-sharded x%d, no data-path collective" % world,
                       "kernel_ms": {rf_name: rf_ms,
                                     (sv_name + " x2 + rhs_to_bundles_kernel + bundles_to_x_kernel") if wide else sv_name: sv_ms},
                       "bundle": ("%d systems per warp, 2 per lane" % sym.wide_width) if wide else "v3 kernels",
                       "bytes_per_system": sym.bytes_per_system(),
                       "step_roofline_frac": step_bytes / (total_ms / K * 1e-3) / 1e9 / peak,
                       "result_gather_ms": gather_ms},
            "roofline": {"bound": "hbm", "kernel": dom[0], "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": dom[1]},
            "e2e": {"value": world * B / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": int(B * (sym.nnz + n) * 8), "d2h_bytes_per_step": int(B * (n * 8 + 4)),
                    "ms_per_step": e2e_ms, "api": "csp3_lu_refactor_solve_host (LuSymbolic.refactor_solve_host)"},
            # wide path: refactor + (transpose in, forward sweep, backward sweep, transpose out); v3: refactor + solve
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
            assert np.array_equal(x_cpu, x_np[:sample]), "GPU result differs from the oracle"
            line["cpu_baseline"] = {"value": sample / dt, "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": "first %d systems of the same batch, single thread, C oracle port of the "
                                              "reference-style CSparse refactor+solve (bit-identical to the GPU result)" % sample}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
