"""Batched sparse LU on one pattern: symbolic phase cached on the host, numeric phase on the GPU.

No reference counterpart.  `LuSymbolic` owns the opaque csp3_lu_symbolic handle of
include/csparse3_b200.h; torch tensors are only the carriers of device buffers (data_ptr + current stream).

    sym = LuSymbolic(n, Ap, Ai, Ax0, order=1, tol=1e-3)      # AMD + first pivoted LU on the host, once
    x, status = sym.refactor_solve(Ax_dev, b_dev)             # every later system: CUDA only
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import as_f64, as_i32, check, ptr

DEFAULT_ORDER = 1      # amd(A + A'): power-flow Jacobians are structurally symmetric
DEFAULT_TOL = 1e-3     # threshold partial pivoting with diagonal preference (CSparse cs_lu `tol`)


class LuSymbolic:
    def __init__(self, n, Ap, Ai, Ax, order=DEFAULT_ORDER, tol=DEFAULT_TOL, q=None):
        Ap, Ai, Ax = as_i32(Ap, "Ap"), as_i32(Ai, "Ai"), as_f64(Ax, "Ax")
        qa = None if q is None else as_i32(q, "q")
        h = C.c_void_p()
        check(_lib.lib().csp3_lu_analyze(order, n, ptr(Ap), ptr(Ai), ptr(Ax), ptr(qa), float(tol), C.byref(h)),
              "csp3_lu_analyze")
        self._init_from_handle(h, n)

    @classmethod
    def from_pattern(cls, n, Ap, Ai, q, pinv, Lp, Li, Up, Ui):
        """Rebuild the object from a cached ordering / pivot sequence / pattern (no factorisation)."""
        self = cls.__new__(cls)
        arrs = [as_i32(a) for a in (Ap, Ai)] + [None if q is None else as_i32(q)] + \
               [as_i32(a) for a in (pinv, Lp, Li, Up, Ui)]
        h = C.c_void_p()
        check(_lib.lib().csp3_lu_analyze_fixed(n, *[ptr(a) for a in arrs], C.byref(h)), "csp3_lu_analyze_fixed")
        self._init_from_handle(h, n)
        return self

    def _init_from_handle(self, h, n):
        self._h = h
        L = _lib.lib()
        sz = (C.c_int64 * 16)()
        check(L.csp3_lu_sizes(h, C.byref(sz)), "csp3_lu_sizes")
        self.n, self.nnz, self.lnz, self.unz = int(sz[0]), int(sz[1]), int(sz[2]), int(sz[3])
        self.nlev_refactor, self.nlev_lsolve, self.nlev_usolve = int(sz[4]), int(sz[5]), int(sz[6])
        self.flops, self.schedule_bytes, self.max_col_len = int(sz[7]), int(sz[8]), int(sz[9])
        # wide (lane = system) kernels: bundle width (0: not available for this pattern, the v3 kernels are used)
        self.wide_width = int(sz[10])
        self.wide_info = dict(width=int(sz[10]), cache_entries=int(sz[11]), landing_entries=int(sz[12]),
                              immediate_fetches=int(sz[13]), cached_updates=int(sz[14]), smem_bytes=int(sz[15]))
        n = self.n
        self.q = np.empty(n, dtype=np.int32); self.pinv = np.empty(n, dtype=np.int32)
        self.Lp = np.empty(n + 1, dtype=np.int32); self.Li = np.empty(self.lnz, dtype=np.int32)
        self.Up = np.empty(n + 1, dtype=np.int32); self.Ui = np.empty(self.unz, dtype=np.int32)
        self.Lx0 = np.empty(self.lnz, dtype=np.float64); self.Ux0 = np.empty(self.unz, dtype=np.float64)
        check(L.csp3_lu_get_pattern(h, ptr(self.q), ptr(self.pinv), ptr(self.Lp), ptr(self.Li), ptr(self.Up),
                                    ptr(self.Ui), ptr(self.Lx0), ptr(self.Ux0)), "csp3_lu_get_pattern")
        self._uploaded = set()

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                _lib.lib().csp3_lu_destroy(h)
            except Exception:
                pass
            self._h = None

    @property
    def nnz_lu(self):
        """nnz(L + U) counting the diagonal once."""
        return self.lnz + self.unz - self.n

    def bytes_per_system(self, fused=False):
        """Algorithmic HBM bytes of one refactor+solve."""
        if fused:
            return 8 * self.nnz + 8 * self.nnz_lu + 16 * self.n
        return 8 * self.nnz + 16 * self.nnz_lu + 16 * self.n

    def supernodes(self):
        """First columns of the fundamental supernodes of L -> int32[count + 1] (last entry n)."""
        sn = np.empty(self.n + 1, dtype=np.int32)
        cnt = C.c_int64(0)
        check(_lib.lib().csp3_lu_supernodes(self._h, ptr(sn), C.byref(cnt)), "csp3_lu_supernodes")
        return sn[:int(cnt.value) + 1].copy()

    def levels(self, kind):
        """kind 0 refactor, 1 L-solve, 2 U-solve -> (level[n], order[n], lptr[nlev+1])"""
        nlev = (self.nlev_refactor, self.nlev_lsolve, self.nlev_usolve)[kind]
        level = np.empty(self.n, dtype=np.int32)
        order = np.empty(self.n, dtype=np.int32)
        lptr = np.empty(nlev + 1, dtype=np.int32)
        check(_lib.lib().csp3_lu_get_levels(self._h, kind, ptr(level), ptr(order), ptr(lptr)), "csp3_lu_get_levels")
        return level, order, lptr

    # ---- host-buffer calls (numpy in, numpy out) ------------------------------------------------------------
    def refactor_host(self, Ax):
        Ax = np.ascontiguousarray(Ax, dtype=np.float64).reshape(-1, self.nnz)
        B = Ax.shape[0]
        Lx = np.empty((B, self.lnz)); Ux = np.empty((B, self.unz)); status = np.zeros(B, dtype=np.int32)
        check(_lib.lib().csp3_lu_refactor_host(self._h, B, ptr(Ax), ptr(Lx), ptr(Ux), ptr(status)), "csp3_lu_refactor_host")
        return Lx, Ux, status

    def solve_host(self, Lx, Ux, b):
        Lx = np.ascontiguousarray(Lx, dtype=np.float64).reshape(-1, self.lnz)
        Ux = np.ascontiguousarray(Ux, dtype=np.float64).reshape(-1, self.unz)
        b = np.ascontiguousarray(b, dtype=np.float64).reshape(-1, self.n)
        x = np.empty_like(b)
        check(_lib.lib().csp3_lu_solve_host(self._h, b.shape[0], ptr(Lx), ptr(Ux), ptr(b), ptr(x)), "csp3_lu_solve_host")
        return x

    def refactor_solve_host(self, Ax, b, x=None, status=None):
        """End-to-end batched call on HOST buffers (numpy arrays or pinned torch CPU tensors viewed as numpy):
        chunked H2D -> refactor -> solve -> D2H pipeline inside the library."""
        B = Ax.shape[0]
        assert Ax.dtype == np.float64 and b.dtype == np.float64 and Ax.flags.c_contiguous and b.flags.c_contiguous
        if x is None:
            x = np.empty((B, self.n), dtype=np.float64)
        if status is None:
            status = np.zeros(B, dtype=np.int32)
        check(_lib.lib().csp3_lu_refactor_solve_host(self._h, B, ptr(Ax), ptr(b), ptr(x), ptr(status)),
              "csp3_lu_refactor_solve_host")
        return x, status

    # ---- device calls (torch CUDA tensors carry the buffers) ------------------------------------------------
    def _upload(self, dev_index, stream):
        if dev_index not in self._uploaded:
            check(_lib.lib().csp3_lu_upload(self._h, stream), "csp3_lu_upload")
            self._uploaded.add(dev_index)

    @staticmethod
    def _dev(t, shape_last, name):
        import torch
        if not (t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()):
            raise TypeError("%s must be a contiguous float64 CUDA tensor" % name)
        if t.shape[-1] != shape_last:
            raise ValueError("%s: last dimension must be %d" % (name, shape_last))
        return t

    def refactor(self, Ax, Lx=None, Ux=None, status=None):
        import torch
        Ax = self._dev(Ax, self.nnz, "Ax")
        B = Ax.numel() // self.nnz
        with torch.cuda.device(Ax.device):
            st = torch.cuda.current_stream().cuda_stream
            self._upload(Ax.device.index, st)
            Lx = torch.empty((B, self.lnz), dtype=torch.float64, device=Ax.device) if Lx is None else Lx
            Ux = torch.empty((B, self.unz), dtype=torch.float64, device=Ax.device) if Ux is None else Ux
            status = torch.empty(B, dtype=torch.int32, device=Ax.device) if status is None else status
            check(_lib.lib().csp3_lu_refactor_batched(self._h, B, Ax.data_ptr(), Lx.data_ptr(), Ux.data_ptr(),
                                                      status.data_ptr(), st), "csp3_lu_refactor_batched")
        return Lx, Ux, status

    def solve(self, Lx, Ux, b, x=None):
        import torch
        b = self._dev(b, self.n, "b")
        B = b.numel() // self.n
        with torch.cuda.device(b.device):
            st = torch.cuda.current_stream().cuda_stream
            self._upload(b.device.index, st)
            x = torch.empty_like(b) if x is None else x
            check(_lib.lib().csp3_lu_solve_batched(self._h, B, self._dev(Lx, self.lnz, "Lx").data_ptr(),
                                                   self._dev(Ux, self.unz, "Ux").data_ptr(), b.data_ptr(),
                                                   x.data_ptr(), st), "csp3_lu_solve_batched")
        return x

    def refactor_solve(self, Ax, b, x=None, Lx=None, Ux=None, status=None, work=None):
        """x = A_k \\ b_k for every system of the batch (device tensors).  Factors are kept only if Lx/Ux are
        given; otherwise they live in `work` (allocated here if None)."""
        import torch
        Ax = self._dev(Ax, self.nnz, "Ax")
        b = self._dev(b, self.n, "b")
        B = b.numel() // self.n
        with torch.cuda.device(b.device):
            st = torch.cuda.current_stream().cuda_stream
            self._upload(b.device.index, st)
            x = torch.empty_like(b) if x is None else x
            status = torch.empty(B, dtype=torch.int32, device=b.device) if status is None else status
            if (Lx is None or Ux is None) and work is None:
                nbytes = _lib.lib().csp3_lu_workspace_bytes(self._h, B)
                work = torch.empty(nbytes, dtype=torch.uint8, device=b.device)
            check(_lib.lib().csp3_lu_refactor_solve_batched(
                self._h, B, Ax.data_ptr(), b.data_ptr(), x.data_ptr(),
                None if Lx is None else Lx.data_ptr(), None if Ux is None else Ux.data_ptr(),
                status.data_ptr(), None if work is None else work.data_ptr(), st), "csp3_lu_refactor_solve_batched")
        return x, status

    def refactor_kernel_name(self, batch, device=None):
        """Kernel the workspace path runs for `batch` systems on the current (or given) device."""
        import torch
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self._upload(dev.index if dev.index is not None else torch.cuda.current_device(), None)
        name = _lib.lib().csp3_lu_refactor_kernel_name(self._h, int(batch))
        if name is None:
            raise RuntimeError(_lib.last_error())
        return name.decode()

    def prepare(self, batch, device=None):
        """Upload the schedule and compile / upload the kernel program batches of this size use (otherwise done inside the
        first refactor_ws call of that size, synchronously)."""
        import torch
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        with torch.cuda.device(dev):
            self._upload(dev.index if dev.index is not None else torch.cuda.current_device(), None)
            check(_lib.lib().csp3_lu_prepare(self._h, int(batch)), "csp3_lu_prepare")

    def refactor_ws(self, Ax, work, status=None):
        """Refactor into the internal bundle-interleaved factor workspace (fast path, see workspace())."""
        import torch
        Ax = self._dev(Ax, self.nnz, "Ax")
        B = Ax.numel() // self.nnz
        with torch.cuda.device(Ax.device):
            st = torch.cuda.current_stream().cuda_stream
            self._upload(Ax.device.index, st)
            status = torch.empty(B, dtype=torch.int32, device=Ax.device) if status is None else status
            check(_lib.lib().csp3_lu_refactor_ws(self._h, B, Ax.data_ptr(), work.data_ptr(), status.data_ptr(), st),
                  "csp3_lu_refactor_ws")
        return status

    def solve_ws(self, work, b, x=None):
        """Solve with the factors left in `work` by refactor_ws (same batch)."""
        import torch
        b = self._dev(b, self.n, "b")
        B = b.numel() // self.n
        with torch.cuda.device(b.device):
            st = torch.cuda.current_stream().cuda_stream
            self._upload(b.device.index, st)
            x = torch.empty_like(b) if x is None else x
            check(_lib.lib().csp3_lu_solve_ws(self._h, B, work.data_ptr(), b.data_ptr(), x.data_ptr(), st),
                  "csp3_lu_solve_ws")
        return x

    def growth_ws(self, work, batch, growth=None):
        """max |L(i,j)|, i > j, per system of the factors refactor_ws left in `work` (pivot-growth indicator)."""
        import torch
        with torch.cuda.device(work.device):
            st = torch.cuda.current_stream().cuda_stream
            growth = torch.empty(batch, dtype=torch.float64, device=work.device) if growth is None else growth
            check(_lib.lib().csp3_lu_growth_ws(self._h, batch, work.data_ptr(), growth.data_ptr(), st), "csp3_lu_growth_ws")
        return growth

    def refactor_solve_checked(self, Ap, Ai, Ax, b, growth_limit=1e6, resid_tol=1e-10, work=None):
        """Refactor + solve with the frozen pivot sequence, then VERIFY every system and re-pivot the ones that fail.

        A frozen pivot sequence is only as good as the values it was chosen on: N-1 outages
        zero entries and can make a pivot tiny.  Every system is checked on the device -- bad-pivot status, element
        growth max |L| > growth_limit, relative residual ||A x - b|| / ||b|| > resid_tol -- and the flagged ones are
        analysed again ON THEIR OWN VALUES (host symbolic phase: new pivot sequence and pattern) and solved through the
        same device kernels.  Nothing is silent: the flagged system ids and what became of them are returned.

        Ax [B, nnz], b [B, n]: CUDA tensors; Ap, Ai: the pattern (numpy int32).
        -> (x [B, n], report) with report = dict(status, growth, resid, flagged=ids, resid_after=[...])"""
        import torch
        from .spmv import SpmvPlan
        B = b.shape[0]
        work = self.workspace(B, b.device) if work is None else work
        status = self.refactor_ws(Ax, work)
        growth = self.growth_ws(work, B)
        x = self.solve_ws(work, b)
        if not hasattr(self, "_spmv") or self._spmv.device != b.device:
            self._spmv = SpmvPlan(self.n, self.n, Ap, Ai, device=b.device)
        r = self._spmv.matvec(Ax, x)
        resid = torch.linalg.vector_norm(r - b, dim=1) / torch.linalg.vector_norm(b, dim=1)
        bad = (status != 0) | ~(growth <= growth_limit) | ~(resid <= resid_tol)
        flagged = torch.nonzero(bad).flatten().cpu().numpy()
        after = []
        for k in flagged:
            Axk = Ax[int(k)].cpu().numpy()
            try:
                symk = LuSymbolic(self.n, Ap, Ai, Axk)                   # new pivot search on this system's own values
            except ArithmeticError:
                after.append(float("inf"))                               # structurally / numerically singular: reported, x left as is
                continue
            xk, stk = symk.refactor_solve(Ax[int(k):int(k) + 1].contiguous(), b[int(k):int(k) + 1].contiguous())
            x[int(k)] = xk[0]
            rk = self._spmv.matvec(Ax[int(k):int(k) + 1].contiguous(), xk)
            after.append(float(torch.linalg.vector_norm(rk[0] - b[int(k)]) / torch.linalg.vector_norm(b[int(k)])))
        return x, dict(status=status, growth=growth, resid=resid, flagged=flagged, resid_after=after)

    def workspace(self, batch, device):
        import torch
        return torch.empty(_lib.lib().csp3_lu_workspace_bytes(self._h, batch), dtype=torch.uint8, device=device)
