"""FP64 tensor-core (DMMA) building blocks: groundwork for the supernodal trailing updates of BASELINE.json's config 5.

No reference counterpart (the reference has no LU).  `dmma_peak()` measures the device's mma.sync f64 throughput (the
roofline denominator of this kernel class); `dense_update(A, B, C)` computes C_s -= A_s @ B_s for a batch of column-major
blocks -- the update L21 * U12 of one supernode for every system of a same-pattern batch (csrc/dense_kernels.cu)."""
import ctypes as C

from . import _lib
from ._lib import check


def dmma_peak(iters=4096):
    """Measured DMMA.8x8x4 throughput of the current device, TFLOP/s."""
    v = C.c_double(0.0)
    check(_lib.lib().csp3_dmma_peak(int(iters), C.byref(v)), "csp3_dmma_peak")
    return float(v.value)


def dense_update(A, B, Cm):
    """Cm[s] -= A[s] @ B[s] in place.  Tensors are CUDA float64 of shape [batch, cols, rows] holding COLUMN-MAJOR blocks
    (i.e. A[s].T is the m x k matrix): A [batch, k, m], B [batch, n, k], Cm [batch, n, m]."""
    import torch
    batch, k, m = A.shape
    n = B.shape[1]
    assert B.shape == (batch, n, k) and Cm.shape == (batch, n, m)
    for T in (A, B, Cm):
        assert T.is_cuda and T.dtype == torch.float64 and T.is_contiguous()
    with torch.cuda.device(A.device):
        check(_lib.lib().csp3_dense_update_batched(batch, m, n, k, A.data_ptr(), m, k * m, B.data_ptr(), k, n * k,
                                                   Cm.data_ptr(), m, n * m, torch.cuda.current_stream().cuda_stream),
              "csp3_dense_update_batched")
    return Cm
