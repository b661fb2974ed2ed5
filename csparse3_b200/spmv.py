"""Device SpMV plan: one CSC pattern, many value sets / vectors (the batched `Ybus * V` of a power-flow
sweep).  Replaces the per-call scipy `csc_matvec` of CscMat.__mul__ (reference csc.py:374-379) when the data
already lives on the GPU."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check


class SpmvPlan:
    def __init__(self, m, n, Ap, Ai, device=None):
        import torch
        self.m, self.n = int(m), int(n)
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.device = dev
        self.Ap = torch.as_tensor(np.ascontiguousarray(Ap, dtype=np.int32)).to(dev)
        self.Ai = torch.as_tensor(np.ascontiguousarray(Ai, dtype=np.int32)).to(dev)
        self.nnz = int(Ap[n])
        h = C.c_void_p()
        with torch.cuda.device(dev):
            check(_lib.lib().csp3_spmv_plan_create(self.m, self.n, self.Ap.data_ptr(), self.Ai.data_ptr(),
                                                   torch.cuda.current_stream().cuda_stream, C.byref(h)),
                  "csp3_spmv_plan_create")
        self._h = h

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                _lib.lib().csp3_spmv_plan_destroy(h)
            except Exception:
                pass
            self._h = None

    def bytes_per_system(self, batched=True):
        """Algorithmic bytes."""
        if batched:
            return 8 * self.nnz + 8 * self.n + 8 * self.m
        return 12 * self.nnz + 4 * (self.n + 1) + 8 * self.n + 8 * self.m

    def matvec(self, Ax, x, y=None, beta=0.0):
        """y[b] = beta*y[b] + A_b x[b].  Ax: [B, nnz] or [nnz] (shared values); x: [B, n]; y: [B, m]."""
        import torch
        assert Ax.is_cuda and x.is_cuda and Ax.dtype == torch.float64 and x.dtype == torch.float64
        assert Ax.is_contiguous() and x.is_contiguous()
        B = x.numel() // self.n
        stride = 0 if Ax.numel() == self.nnz else self.nnz
        if stride:
            assert Ax.numel() == B * self.nnz
        if y is None:
            y = torch.empty(x.shape[:-1] + (self.m,), dtype=torch.float64, device=x.device)
            beta = 0.0
        with torch.cuda.device(x.device):
            check(_lib.lib().csp3_spmv_batched(self._h, B, Ax.data_ptr(), stride, x.data_ptr(), y.data_ptr(),
                                               float(beta), torch.cuda.current_stream().cuda_stream), "csp3_spmv_batched")
        return y
