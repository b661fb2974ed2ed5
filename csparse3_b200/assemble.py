"""Jacobian assembly on the device: [[A11, A12], [A21, A22]] for a batch of value sets that share the four block
patterns -- the step right before the refactorisation in every Newton-Raphson iteration.

Reference: pack_4_by_4 (src/CSparse3/csc.py:588-606) -> csc_stack_4_by_4_ff (src/CSparse3/csc_numba.py:640-720),
which builds one matrix on the host per call.  Here the pattern work is done once (`Stack4Plan`), and
`Stack4Plan.assemble` is one gather kernel over device-resident block values whose output `[batch, nnz]` is exactly
the `Ax` operand of `LuSymbolic.refactor / refactor_solve` -- the values never leave the GPU."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import as_i32, check, ptr


class Stack4Plan:
    """plan = Stack4Plan(A11, A12, A21, A22) with CscMat-like blocks (m, n, indices, indptr; values ignored)."""

    def __init__(self, A11, A12, A21, A22):
        args = []
        self._keep = []
        for M in (A11, A12, A21, A22):
            idx, iptr = as_i32(M.indices, "indices"), as_i32(M.indptr, "indptr")
            self._keep += [idx, iptr]
            args += [int(M.m), int(M.n), ptr(idx), ptr(iptr)]
        h = C.c_void_p()
        check(_lib.lib().csp3_stack4_create(*args, C.byref(h)), "csp3_stack4_create")
        self._h = h
        sz = (C.c_int64 * 8)()
        check(_lib.lib().csp3_stack4_sizes(h, sz), "csp3_stack4_sizes")
        self.m, self.n, self.nnz = int(sz[0]), int(sz[1]), int(sz[2])
        self.block_nnz = tuple(int(sz[4 + i]) for i in range(4))
        self.indices = np.empty(self.nnz, dtype=np.int32)
        self.indptr = np.empty(self.n + 1, dtype=np.int32)
        check(_lib.lib().csp3_stack4_get_pattern(h, ptr(self.indices), ptr(self.indptr)), "csp3_stack4_get_pattern")

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                _lib.lib().csp3_stack4_destroy(h)
            except Exception:
                pass
            self._h = None

    def bytes_per_system(self):
        """Algorithmic bytes of the numeric step: every entry read once and written once."""
        return 16 * self.nnz

    def assemble(self, X11, X12, X21, X22, out=None):
        """Block values: CUDA float64 tensors [B, nnz_block] or [nnz_block] (one value set shared by the batch).
        Returns out[B, nnz] in the stacked matrix's entry order (self.indptr / self.indices)."""
        import torch
        blocks = (X11, X12, X21, X22)
        B = 1
        for X, bn in zip(blocks, self.block_nnz):
            assert X.is_cuda and X.dtype == torch.float64 and X.is_contiguous()
            assert X.numel() % max(bn, 1) == 0 and (bn > 0 or X.numel() == 0)
            if X.dim() == 2:
                B = max(B, X.shape[0])
        ld = []
        for X, bn in zip(blocks, self.block_nnz):
            shared = X.dim() == 1 or X.shape[0] == 1
            assert shared or X.shape[0] == B
            ld.append(0 if shared and B > 1 else bn)
        dev = next(X.device for X in blocks)
        if out is None:
            out = torch.empty((B, self.nnz), dtype=torch.float64, device=dev)
        assert out.is_cuda and out.dtype == torch.float64 and out.is_contiguous() and out.shape == (B, self.nnz)
        with torch.cuda.device(dev):
            check(_lib.lib().csp3_stack4_batched(self._h, B, X11.data_ptr(), ld[0], X12.data_ptr(), ld[1],
                                                 X21.data_ptr(), ld[2], X22.data_ptr(), ld[3], out.data_ptr(), self.nnz,
                                                 torch.cuda.current_stream().cuda_stream), "csp3_stack4_batched")
        return out
