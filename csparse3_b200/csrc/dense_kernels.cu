// dense_kernels.cu -- FP64 tensor-core (DMMA) building blocks for supernodal trailing updates: groundwork for
// BASELINE.json's config 5 (3-D Laplacian n = 1e6 LU "with supernodal FP64-DMMA trailing updates").
//
// No reference counterpart (the reference has no LU; north_star: "Tensor cores (FP64 DMMA) are
// used only for dense supernodal trailing updates where a supernode is wide enough to be a real dense contraction").
// tcgen05.mma has no f64 kind, so FP64 tensor work on sm_100a is the warp-level mma.sync path: SASS DMMA.8x8x4
// (mma.sync.aligned.m8n8k4.row.col.f64; the m16n8k8 PTX shape lowers to four of them).
//
//   csp3_dmma_peak            measured DMMA throughput of the device (the roofline denominator for this kernel class;
//                             MEASURED_PEAKS.json carries no FP64 tensor number)
//   csp3_dense_update_batched C_s -= A_s * B_s for a batch of column-major blocks (one supernode's trailing update
//                             L21 * U12 for every system of a same-pattern batch), 32 x 32 tile of C per warp
#include "../../include/csparse3_b200.h"
#include "common.cuh"

using namespace csp3;

namespace {

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// 8 independent accumulator pairs per warp, `iters` DMMA each: 2 * 8 * 8 * 4 flops per instruction
__global__ void __launch_bounds__(256) k_dmma_peak(int iters, double *sink)
{
    double d[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) d[i] = threadIdx.x * 1e-3 + i;
    const double a = 1.0 + threadIdx.x * 1e-6, b = 1.0 - threadIdx.x * 1e-6;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma884(d[2 * i], d[2 * i + 1], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += d[i];
    if (s == 123.456) sink[blockIdx.x * blockDim.x + threadIdx.x] = s;       // never true: keeps the chain alive
}

// C[m x n] -= A[m x k] * B[k x n], column-major, per system s at offsets s * stride.  One warp per 32 x 32 tile of C:
// 4 x 4 DMMA tiles of 8 x 8, accumulators in registers (32 doubles per lane); A / B fragments straight from global
// (L2-resident for the block sizes of a supernode).  Fragment layout of mma.m8n8k4.f64: lane = 4 * g + t (g = 0..7,
// t = 0..3): A(row g, col t), B(row t, col g), C(row g, cols 2t, 2t + 1).  Edges are handled by predication (zeros).
__global__ void __launch_bounds__(128)
k_dense_update(int m, int n, int k, const double *__restrict__ A, int lda, i64 strideA, const double *__restrict__ B, int ldb,
               i64 strideB, double *C, int ldc, i64 strideC)
{
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int tiles_m = (m + 31) / 32, tiles_n = (n + 31) / 32;
    const int tile = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (tile >= tiles_m * tiles_n) return;
    const i64 s = blockIdx.y;
    const double *As = A + s * strideA, *Bs = B + s * strideB;
    double *Cs = C + s * strideC;
    const int i0 = (tile % tiles_m) * 32, j0 = (tile / tiles_m) * 32;
    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
    for (int kk = 0; kk < k; kk += 4) {
        double af[4], bf[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int r = i0 + 8 * a + g, c = kk + t;
            af[a] = (r < m && c < k) ? __ldg(As + (i64)c * lda + r) : 0.0;
        }
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int r = kk + t, c = j0 + 8 * b + g;
            bf[b] = (r < k && c < n) ? __ldg(Bs + (i64)c * ldb + r) : 0.0;
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) dmma884(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int r = i0 + 8 * a + g, c = j0 + 8 * b + 2 * t + e;
                if (r < m && c < n) Cs[(i64)c * ldc + r] -= acc[a][b][e];
            }
}

}  // namespace

extern "C" {

int csp3_dmma_peak(int64_t iters, double *tflops)
{
    if (iters <= 0 || !tflops) { set_error("dmma_peak: bad arguments"); return CSP3_ERR_ARG; }
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt == 0) { cudaGetLastError(); set_error("no CUDA device available: libcsparse3_b200 has no CPU fallback"); return CSP3_ERR_CUDA; }
    double *sink = nullptr;
    const int blocks = kNumSMs * 8, threads = 256;
    CSP3_CUDA(cudaMalloc((void **)&sink, (size_t)blocks * threads * 8));
    cudaEvent_t e0, e1;
    CSP3_CUDA(cudaEventCreate(&e0)); CSP3_CUDA(cudaEventCreate(&e1));
    k_dmma_peak<<<blocks, threads>>>(64, sink);                          // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        CSP3_CUDA(cudaEventRecord(e0));
        k_dmma_peak<<<blocks, threads>>>((int)iters, sink);
        CSP3_CUDA(cudaEventRecord(e1));
        CSP3_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        CSP3_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        best = ms < best ? ms : best;
    }
    const double flops = (double)blocks * (threads / 32) * (double)iters * 8.0 * (2.0 * 8 * 8 * 4);
    *tflops = flops / (best * 1e-3) / 1e12;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(sink);
    return 0;
}

int csp3_dense_update_batched(int64_t batch, int64_t m, int64_t n, int64_t k, const double *A, int64_t lda, int64_t strideA,
                              const double *B, int64_t ldb, int64_t strideB, double *C, int64_t ldc, int64_t strideC, void *stream)
{
    if (batch < 0 || m < 0 || n < 0 || k < 0 || !A || !B || !C || lda < m || ldb < k || ldc < m || batch > 65535) {
        set_error("dense_update_batched: bad arguments");
        return CSP3_ERR_ARG;
    }
    if (batch == 0 || m == 0 || n == 0 || k == 0) return 0;
    const int tiles = (int)(((m + 31) / 32) * ((n + 31) / 32));
    k_dense_update<<<dim3((unsigned)((tiles + 3) / 4), (unsigned)batch), 128, 0, (cudaStream_t)stream>>>(
        (int)m, (int)n, (int)k, A, (int)lda, strideA, B, (int)ldb, strideB, C, (int)ldc, strideC);
    CSP3_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"
