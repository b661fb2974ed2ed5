"""FP64 tensor-core measurements kept under profiles/: DMMA peak and the batched dense update at supernode-like sizes.

    python tools/dmma_probe.py > profiles/dmma_r02.json
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from csparse3_b200 import dense, synth
from csparse3_b200.lu import LuSymbolic

out = {"dmma_peak_tflops": dense.dmma_peak(8192), "dense_update": [], "supernodes": []}
for batch, m, n, k in ((1024, 64, 64, 16), (1024, 128, 128, 32), (256, 256, 256, 64), (64, 512, 512, 128), (8, 2048, 2048, 256)):
    A = torch.randn((batch, k, m), dtype=torch.float64, device="cuda")
    B = torch.randn((batch, n, k), dtype=torch.float64, device="cuda")
    C = torch.randn((batch, n, m), dtype=torch.float64, device="cuda")
    dense.dense_update(A, B, C); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(5):
        e0.record(); dense.dense_update(A, B, C); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out["dense_update"].append({"batch": batch, "m": m, "n": n, "k": k, "ms": best,
                                "tflops": 2.0 * batch * m * n * k / (best * 1e-3) / 1e12})
for k in (12, 16, 20, 24):
    n, Ap, Ai, Ax = synth.laplacian_3d(k)
    try:
        sym = LuSymbolic(n, Ap, Ai, Ax, order=1, tol=1.0)
    except Exception as e:               # the scalar schedule builder refuses factors with more than 2^31 - 1 update slots
        out["supernodes"].append({"grid": "%d^3" % k, "n": n, "error": str(e)})
        continue
    sn = sym.supernodes()
    w = np.diff(sn)
    lcnt = np.diff(sym.Lp) - 1
    # flops of the trailing updates by source supernode width: sum over source columns of (rows below)^2
    fl = np.add.reduceat(lcnt.astype(np.float64) ** 2, sn[:-1])
    wide = w >= 16
    out["supernodes"].append({"grid": "%d^3" % k, "n": n, "nnz_lu": sym.nnz_lu, "supernodes": int(len(w)), "widest": int(w.max()),
                              "columns_in_supernodes_ge16": int(w[wide].sum()),
                              "update_flops_share_from_supernodes_ge16": float(fl[wide].sum() / fl.sum())})
print(json.dumps(out, indent=1))
