"""ctypes binding of libcsparse3_b200.so (the C ABI in include/csparse3_b200.h).

The library is built in-tree by csparse3_b200/csrc/Makefile (nvcc, sm_100a only).  There is no CPU
fallback: if the shared object is missing, or a compute entry point is called without a CUDA device, the
call raises.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcsparse3_b200.so")
CSRC = os.path.join(_HERE, "csrc")

vp = C.c_void_p
i64 = C.c_int64
f64 = C.c_double
cint = C.c_int

# name -> argtypes (restype is int unless listed in _RESTYPE).  Mirrors include/csparse3_b200.h 1:1;
# tests/test_abi.py checks that every symbol the header declares is exported and bound here.
SIGNATURES = {
    "csp3_version": [],
    "csp3_last_error_string": [],
    "csp3_device_count": [],
    "csp3_set_device": [cint],
    "csp3_csc_mat_vec_ff_host": [i64, i64, vp, vp, vp, vp, vp],
    "csp3_csc_matvec_host": [i64, i64, vp, vp, vp, vp, vp],
    "csp3_csc_matvecs_host": [i64, i64, i64, vp, vp, vp, vp, vp],
    "csp3_spmv_plan_create": [i64, i64, vp, vp, vp, C.POINTER(vp)],
    "csp3_spmv_plan_destroy": [vp],
    "csp3_spmv_batched": [vp, i64, vp, i64, vp, vp, f64, vp],
    "csp3_csc_transpose_host": [i64, i64, vp, vp, vp, vp, vp, vp],
    "csp3_csc_to_csr_host": [i64, i64, vp, vp, vp, vp, vp, vp],
    "csp3_csc_transpose": [i64, i64, vp, vp, vp, vp, vp, vp, vp],
    "csp3_spgemm_symbolic_host": [i64, i64, vp, vp, i64, i64, vp, vp, vp, C.POINTER(i64)],
    "csp3_spgemm_numeric_host": [i64, i64, vp, vp, vp, i64, i64, vp, vp, vp, vp, vp, vp],
    "csp3_spgemm_symbolic": [i64, i64, vp, vp, i64, i64, vp, vp, vp, C.POINTER(i64), vp],
    "csp3_spgemm_numeric": [i64, i64, vp, vp, vp, i64, i64, vp, vp, vp, vp, vp, vp, vp],
    "csp3_csc_plusminus_host": [i64, i64, vp, vp, vp, vp, vp, vp, f64, vp, vp, vp],
    "csp3_csc_add_ff_host": [i64, i64, vp, vp, vp, vp, vp, vp, f64, f64, vp, vp, vp],
    "csp3_stack4_create": [i64, i64, vp, vp, i64, i64, vp, vp, i64, i64, vp, vp, i64, i64, vp, vp, C.POINTER(vp)],
    "csp3_stack4_destroy": [vp],
    "csp3_stack4_sizes": [vp, vp],
    "csp3_stack4_get_pattern": [vp, vp, vp],
    "csp3_stack4_batched": [vp, i64, vp, i64, vp, i64, vp, i64, vp, i64, vp, i64, vp],
    "csp3_csc_stack_4_by_4_host": [i64, i64, vp, vp, vp, i64, i64, vp, vp, vp, i64, i64, vp, vp, vp, i64, i64, vp, vp, vp, vp, vp, vp],
    "csp3_csc_amd": [i64, i64, i64, vp, vp, vp],
    "csp3_csc_etree": [i64, i64, vp, vp, cint, vp],
    "csp3_csc_post": [i64, vp, vp],
    "csp3_lu_analyze": [i64, i64, vp, vp, vp, vp, f64, C.POINTER(vp)],
    "csp3_lu_analyze_fixed": [i64, vp, vp, vp, vp, vp, vp, vp, vp, C.POINTER(vp)],
    "csp3_lu_destroy": [vp],
    "csp3_lu_sizes": [vp, C.POINTER(i64 * 16)],
    "csp3_lu_get_pattern": [vp, vp, vp, vp, vp, vp, vp, vp, vp],
    "csp3_lu_supernodes": [vp, vp, C.POINTER(i64)],
    "csp3_dmma_peak": [i64, C.POINTER(f64)],
    "csp3_dense_update_batched": [i64, i64, i64, i64, vp, i64, i64, vp, i64, i64, vp, i64, i64, vp],
    "csp3_lu_get_levels": [vp, cint, vp, vp, vp],
    "csp3_lu_get_program": [vp, cint, vp, i64, vp],
    "csp3_lu_upload": [vp, vp],
    "csp3_lu_refactor_batched": [vp, i64, vp, vp, vp, vp, vp],
    "csp3_lu_solve_batched": [vp, i64, vp, vp, vp, vp, vp],
    "csp3_lu_workspace_bytes": [vp, i64],
    "csp3_lu_refactor_kernel_name": [vp, i64],
    "csp3_lu_prepare": [vp, i64],
    "csp3_lu_refactor_solve_batched": [vp, i64, vp, vp, vp, vp, vp, vp, vp, vp],
    "csp3_lu_refactor_ws": [vp, i64, vp, vp, vp, vp],
    "csp3_lu_solve_ws": [vp, i64, vp, vp, vp, vp],
    "csp3_lu_growth_ws": [vp, i64, vp, vp, vp],
    "csp3_lu_refactor_solve_host": [vp, i64, vp, vp, vp, vp],
    "csp3_lu_refactor_host": [vp, i64, vp, vp, vp, vp],
    "csp3_lu_solve_host": [vp, i64, vp, vp, vp, vp],
    "csp3_csc_lusol_host": [i64, i64, vp, vp, vp, vp, f64],
    "csp3_find_islands_host": [i64, vp, vp, vp, vp, C.POINTER(i64)],
    "csp3_islands_batched": [i64, vp, vp, i64, vp, vp, vp, vp, vp],
    "csp3_islands_batched_host": [i64, vp, vp, i64, vp, vp, vp, vp],
    "csp3_csc_sub_matrix_host": [i64, i64, vp, vp, vp, i64, vp, i64, vp, vp, vp, vp, C.POINTER(i64)],
    "csp3_nr_create": [i64, i64, vp, vp, vp, i64, vp, i64, vp, i64, vp, vp, i64, vp, vp, vp, C.POINTER(vp)],
    "csp3_nr_destroy": [vp],
    "csp3_nr_workspace_bytes": [vp, i64],
    "csp3_nr_jacobian": [vp, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp],
    "csp3_nr_solve": [vp, i64, i64, vp, vp, vp, vp, vp, vp, vp, vp],
    "csp3_nr_solve_host": [vp, i64, i64, vp, vp, vp, vp, i64, vp, vp, vp, vp],
}
_RESTYPE = {"csp3_last_error_string": C.c_char_p, "csp3_lu_refactor_kernel_name": C.c_char_p, "csp3_lu_workspace_bytes": i64, "csp3_lu_get_program": i64,
             "csp3_nr_workspace_bytes": i64}

_lib = None


def build(force=False):
    """Compile libcsparse3_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(_HERE, "..", "include", "csparse3_b200.h")]
    stale = (not os.path.exists(LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", CSRC, "all"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "csparse3_b200: %s is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = _RESTYPE.get(name, cint)
        _lib = L
    return _lib


class Csp3Error(RuntimeError):
    pass


def last_error():
    return lib().csp3_last_error_string().decode("utf-8", "replace")


def check(rc, what="csp3 call"):
    """Map C status codes to the exception classes the reference's backends raise
    (sparsetools.cxx:361-368: bad_alloc -> MemoryError, other -> RuntimeError; bad dtypes -> ValueError)."""
    if rc == 0:
        return
    msg = "%s failed (%d): %s" % (what, rc, last_error())
    if rc == -3:
        raise MemoryError(msg)
    if rc == -1:
        raise ValueError(msg)
    if rc > 0:
        raise ArithmeticError(msg)
    raise Csp3Error(msg)


def ptr(a):
    """Host pointer of a C-contiguous numpy array (None -> NULL)."""
    if a is None:
        return None
    assert a.flags.c_contiguous
    return a.ctypes.data


def as_i32(a, name="index array"):
    """The reference's kernels are dtype-strict (i4[:] indices; int64 raises TypeError)."""
    a = np.asarray(a)
    if a.dtype != np.int32:
        raise TypeError("%s must be int32 (got %s): No matching definition" % (name, a.dtype))
    return np.ascontiguousarray(a)


def as_f64(a, name="value array"):
    a = np.asarray(a)
    if a.dtype != np.float64:
        raise TypeError("%s must be float64 (got %s): No matching definition" % (name, a.dtype))
    return np.ascontiguousarray(a)
