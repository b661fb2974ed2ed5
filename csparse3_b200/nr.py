"""Newton-Raphson power-flow iteration kept on the GPU.

The consumer loop the reference's kernels exist for: J = pack_4_by_4(H, N, M, L)
(src/CSparse3/csc.py:588-606), mismatch from Ybus * V (CscMat.__mul__, csc.py:374-379), factor, solve, update.
`NewtonPlan` owns a csp3_nr_plan (include/csparse3_b200.h): a batch of same-topology cases -- time-series value
sets or N-1 outages -- iterates on the device; per case only the specified injections go in and (vm, va) come back.

    sym  = LuSymbolic(n, Ap, Ai, Ax0)                        # pattern of the Jacobian, once
    plan = NewtonPlan.from_case(case, sym)                   # case: csparse3_b200.synth.GridCase or anything with its fields
    vm, va, fnorm, status = plan.solve_host(sspec, iters=4)  # flat start, numpy in / numpy out
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import as_f64, as_i32, check, ptr


class NewtonPlan:
    def __init__(self, n_bus, y_rowptr, y_col, y_val, pvpq, pq, j_ent, j_block, sym, br_slot=None, br_val=None):
        self.n_bus, self.sym = int(n_bus), sym
        y_rowptr, y_col = as_i32(y_rowptr, "y_rowptr"), as_i32(y_col, "y_col")
        y_val = np.ascontiguousarray(y_val, dtype=np.complex128)
        pvpq, pq = as_i32(pvpq, "pvpq"), as_i32(pq, "pq")
        j_ent, j_block = as_i32(j_ent, "j_ent"), as_i32(j_block, "j_block")
        self.npvpq, self.npq, self.n, self.jnnz = len(pvpq), len(pq), len(pvpq) + len(pq), len(j_ent)
        nbr = 0
        if br_slot is not None:
            br_slot = as_i32(np.ascontiguousarray(br_slot), "br_slot")
            br_val = np.ascontiguousarray(br_val, dtype=np.complex128)
            assert br_slot.shape == br_val.shape and br_slot.shape[0] == 4
            nbr = br_slot.shape[1]
        h = C.c_void_p()
        check(_lib.lib().csp3_nr_create(self.n_bus, len(y_col), ptr(y_rowptr), ptr(y_col), ptr(y_val.view(np.float64)),
                                        self.npvpq, ptr(pvpq), self.npq, ptr(pq), self.jnnz, ptr(j_ent), ptr(j_block),
                                        nbr, ptr(br_slot), None if br_val is None else ptr(br_val.view(np.float64)),
                                        sym._h, C.byref(h)), "csp3_nr_create")
        self._h = h

    @classmethod
    def from_case(cls, case, sym, outages=False):
        """case: csparse3_b200.synth.GridCase (Ybus entries sorted by (row, col))."""
        rowptr = np.r_[case.y_rowstart, case.nnz_y].astype(np.int32)
        br_slot = br_val = None
        if outages:
            yff = case.ys + 0.5j * case.bsh
            br_slot = case.br_slot.astype(np.int32)
            br_val = np.stack([yff, -case.ys, -case.ys, yff])
        return cls(case.n_bus, rowptr, case.yk.astype(np.int32), case.ybus_values(), case.pvpq.astype(np.int32),
                   case.pq.astype(np.int32), case.j_ent.astype(np.int32), case.j_block.astype(np.int32), sym, br_slot, br_val)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                _lib.lib().csp3_nr_destroy(h)
            except Exception:
                pass
            self._h = None

    def workspace_bytes(self, batch):
        return int(_lib.lib().csp3_nr_workspace_bytes(self._h, batch))

    # ---- host buffers -------------------------------------------------------------------------------------
    def solve_host(self, sspec, iters=4, out_branch=None, vm0=None, va0=None, vm=None, va=None, fnorm=None, status=None):
        """sspec[B, n] (P of the pvpq buses, then Q of the pq buses).  Start: flat (1.0, 0.0) unless vm0 / va0 are
        given as [n_bus] (shared) or [B, n_bus] (per case).  -> (vm[B, n_bus], va[B, n_bus], fnorm[B], status[B])"""
        sspec = as_f64(sspec, "sspec").reshape(-1, self.n)
        B = sspec.shape[0]
        vm0 = np.ones(self.n_bus) if vm0 is None else as_f64(vm0, "vm0")
        va0 = np.zeros(self.n_bus) if va0 is None else as_f64(va0, "va0")
        stride = 0 if vm0.ndim == 1 else self.n_bus
        assert vm0.shape == va0.shape and vm0.shape[-1] == self.n_bus and (stride == 0 or vm0.shape[0] == B)
        ob = None if out_branch is None else as_i32(out_branch, "out_branch")
        vm = np.empty((B, self.n_bus)) if vm is None else vm
        va = np.empty((B, self.n_bus)) if va is None else va
        fnorm = np.empty(B) if fnorm is None else fnorm
        status = np.zeros(B, dtype=np.int32) if status is None else status
        check(_lib.lib().csp3_nr_solve_host(self._h, B, int(iters), ptr(sspec), ptr(ob), ptr(vm0), ptr(va0), stride,
                                            ptr(vm), ptr(va), ptr(fnorm), ptr(status)), "csp3_nr_solve_host")
        return vm, va, fnorm, status

    # ---- device buffers (torch CUDA tensors carry them) ------------------------------------------------------
    def workspace(self, batch, device):
        import torch
        return torch.empty(self.workspace_bytes(batch), dtype=torch.uint8, device=device)

    def jacobian(self, vm, va, sspec, out_branch=None, work=None):
        """-> (Ax[B, jnnz], b[B, n], fnorm[B]) at the state (vm, va): what one iteration hands to the refactorisation."""
        import torch
        B = vm.shape[0]
        with torch.cuda.device(vm.device):
            st = torch.cuda.current_stream().cuda_stream
            work = self.workspace(B, vm.device) if work is None else work
            Ax = torch.empty((B, self.jnnz), dtype=torch.float64, device=vm.device)
            b = torch.empty((B, self.n), dtype=torch.float64, device=vm.device)
            fnorm = torch.empty(B, dtype=torch.float64, device=vm.device)
            check(_lib.lib().csp3_nr_jacobian(self._h, B, vm.data_ptr(), va.data_ptr(),
                                              None if out_branch is None else out_branch.data_ptr(), sspec.data_ptr(),
                                              Ax.data_ptr(), b.data_ptr(), fnorm.data_ptr(), work.data_ptr(), st), "csp3_nr_jacobian")
        return Ax, b, fnorm

    def solve(self, vm, va, sspec, iters=4, out_branch=None, work=None, fnorm=None, status=None):
        """`iters` Newton iterations in place on the device tensors vm, va [B, n_bus].  -> (fnorm[B], status[B])"""
        import torch
        B = vm.shape[0]
        with torch.cuda.device(vm.device):
            st = torch.cuda.current_stream().cuda_stream
            work = self.workspace(B, vm.device) if work is None else work
            fnorm = torch.empty(B, dtype=torch.float64, device=vm.device) if fnorm is None else fnorm
            status = torch.empty(B, dtype=torch.int32, device=vm.device) if status is None else status
            check(_lib.lib().csp3_nr_solve(self._h, B, int(iters), sspec.data_ptr(),
                                           None if out_branch is None else out_branch.data_ptr(), vm.data_ptr(), va.data_ptr(),
                                           fnorm.data_ptr(), status.data_ptr(), work.data_ptr(), st), "csp3_nr_solve")
        return fnorm, status
