// stack4.cu -- Jacobian assembly on the device: [[A, B], [C, D]] of four CSC blocks.
//
// Replaces csc_stack_4_by_4_ff (src/CSparse3/csc_numba.py:640-720, caller pack_4_by_4 src/CSparse3/csc.py:588-606):
// column-wise concatenation, inside a column the entries of the upper block before those of the lower block, row
// indices of the lower blocks shifted by the upper block's row count, no sorting and no merging of duplicates.
//
// The pattern work (output indptr / indices and, for every output entry, the block and position its value comes
// from) is done once per pattern on the host and kept in a plan; the numeric step -- the one a Newton-Raphson loop
// repeats for every value set of a same-pattern batch -- is a pure gather on the device:
//     out[s][p] = block(p)[s][pos(p)]
// 16 algorithmic bytes per entry and system (8 read, 8 written), HBM bound.
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/csparse3_b200.h"
#include "common.cuh"

using namespace csp3;

namespace {

constexpr int kMaxDev = 16;
constexpr int kStackThreads = 256;
constexpr int kStackPerThread = 4;          // output entries of one system per thread, one grid span apart

struct Blocks4 { const double *x[4]; i64 ld[4]; };

// One CTA row per system (blockIdx.y strides over the batch), consecutive threads write consecutive entries: the
// stores are fully coalesced; the loads are coalesced inside every block segment of a column (the map is
// piecewise linear: runs of consecutive positions of one block).
__global__ void __launch_bounds__(kStackThreads) stack4_gather_kernel(const i32 *__restrict__ map, i64 nnz, i64 batch, Blocks4 B,
                                                                       double *__restrict__ out, i64 ldo)
{
    const i64 p0 = ((i64)blockIdx.x * kStackThreads + threadIdx.x);
    i32 m[kStackPerThread];
    const i64 span = (i64)gridDim.x * kStackThreads;
#pragma unroll
    for (int i = 0; i < kStackPerThread; ++i) {
        const i64 p = p0 + i * span;
        m[i] = p < nnz ? __ldg(map + p) : -1;
    }
    // two systems per trip: 2 * kStackPerThread independent loads in flight per thread
    const i64 step = gridDim.y;
    for (i64 s = blockIdx.y; s < batch; s += 2 * step) {
        const bool two = s + step < batch;
        double v[2][kStackPerThread];
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int i = 0; i < kStackPerThread; ++i)
                if (m[i] >= 0 && (u == 0 || two)) {
                    const unsigned w = (unsigned)m[i];
                    const int blk = (int)(w >> 29);
                    v[u][i] = __ldg(B.x[blk] + (s + u * step) * B.ld[blk] + (i64)(w & 0x1fffffffu));
                }
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int i = 0; i < kStackPerThread; ++i)
                if (m[i] >= 0 && (u == 0 || two)) out[(s + u * step) * ldo + p0 + i * span] = v[u][i];
    }
}

}  // namespace

struct csp3_stack4 {
    i64 m = 0, n = 0, nnz = 0;
    i64 bnnz[4] = {0, 0, 0, 0};
    std::vector<i32> indptr, indices, map;             // map: block << 29 | position inside the block
    i32 *dmap[kMaxDev] = {};
    std::mutex mu;
};

extern "C" {

int csp3_stack4_create(int64_t am, int64_t an, const int32_t *Ai, const int32_t *Ap,
                       int64_t bm, int64_t bn, const int32_t *Bi, const int32_t *Bp,
                       int64_t cm, int64_t cn, const int32_t *Ci, const int32_t *Cp,
                       int64_t dm, int64_t dn, const int32_t *Di, const int32_t *Dp, csp3_stack4 **plan)
{
    if (!plan || !Ap || !Bp || !Cp || !Dp || am < 0 || an < 0 || bn < 0 || cm < 0) { set_error("stack4_create: bad arguments"); return CSP3_ERR_ARG; }
    // the reference's assertions (csc_numba.py:679-682)
    if (am != bm || cm != dm || an != cn || bn != dn) { set_error("stack4_create: block shapes do not match"); return CSP3_ERR_ARG; }
    const i64 total = (i64)Ap[an] + Bp[bn] + Cp[cn] + Dp[dn];
    if (total >= (1ll << 29) || am + cm > INT32_MAX) { set_error("stack4_create: more than 2^29 entries"); return CSP3_ERR_OVERFLOW; }
    auto *P = new csp3_stack4();
    P->m = am + cm; P->n = an + bn; P->nnz = total;
    P->bnnz[0] = Ap[an]; P->bnnz[1] = Bp[bn]; P->bnnz[2] = Cp[cn]; P->bnnz[3] = Dp[dn];
    P->indptr.assign((size_t)P->n + 1, 0);
    P->indices.resize((size_t)total);
    P->map.resize((size_t)total);
    i64 cnt = 0;
    auto half = [&](i64 ncols, const i32 *Tp, const i32 *Ti, unsigned tb, const i32 *Lp_, const i32 *Li_, unsigned lb, i64 shift, i64 col0) {
        for (i64 j = 0; j < ncols; ++j) {
            for (i32 k = Tp[j]; k < Tp[j + 1]; ++k) { P->indices[(size_t)cnt] = Ti[k]; P->map[(size_t)cnt++] = (i32)((tb << 29) | (unsigned)k); }
            for (i32 k = Lp_[j]; k < Lp_[j + 1]; ++k) { P->indices[(size_t)cnt] = Li_[k] + (i32)shift; P->map[(size_t)cnt++] = (i32)((lb << 29) | (unsigned)k); }
            P->indptr[(size_t)(col0 + j + 1)] = (i32)cnt;
        }
    };
    half(an, Ap, Ai, 0u, Cp, Ci, 2u, am, 0);
    half(bn, Bp, Bi, 1u, Dp, Di, 3u, bm, an);
    *plan = P;
    return 0;
}

int csp3_stack4_destroy(csp3_stack4 *plan)
{
    if (!plan) return 0;
    int cur = 0;
    const bool have = cudaGetDevice(&cur) == cudaSuccess;
    for (int d = 0; d < kMaxDev; ++d)
        if (plan->dmap[d]) { cudaSetDevice(d); cudaFree(plan->dmap[d]); }
    if (have) cudaSetDevice(cur);
    cudaGetLastError();
    delete plan;
    return 0;
}

int csp3_stack4_sizes(const csp3_stack4 *plan, int64_t out[8])
{
    if (!plan || !out) { set_error("stack4_sizes: bad arguments"); return CSP3_ERR_ARG; }
    out[0] = plan->m; out[1] = plan->n; out[2] = plan->nnz; out[3] = 0;
    for (int i = 0; i < 4; ++i) out[4 + i] = plan->bnnz[i];
    return 0;
}

int csp3_stack4_get_pattern(const csp3_stack4 *plan, int32_t *indices, int32_t *indptr)
{
    if (!plan || !indptr || (plan->nnz > 0 && !indices)) { set_error("stack4_get_pattern: bad arguments"); return CSP3_ERR_ARG; }
    memcpy(indptr, plan->indptr.data(), plan->indptr.size() * 4);
    if (plan->nnz) memcpy(indices, plan->indices.data(), (size_t)plan->nnz * 4);
    return 0;
}

int csp3_stack4_batched(csp3_stack4 *plan, int64_t batch, const double *Ax, int64_t lda, const double *Bx, int64_t ldb,
                        const double *Cx, int64_t ldc, const double *Dx, int64_t ldd, double *out, int64_t ldo, void *stream)
{
    if (!plan || batch < 0 || (plan->nnz > 0 && !out) || ldo < plan->nnz) { set_error("stack4_batched: bad arguments"); return CSP3_ERR_ARG; }
    const double *xs[4] = {Ax, Bx, Cx, Dx};
    const i64 lds[4] = {lda, ldb, ldc, ldd};
    for (int i = 0; i < 4; ++i)
        if (plan->bnnz[i] > 0 && (!xs[i] || (lds[i] != 0 && lds[i] < plan->bnnz[i]))) { set_error("stack4_batched: bad block %d", i); return CSP3_ERR_ARG; }
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt == 0) {
        cudaGetLastError();
        set_error("no CUDA device available: libcsparse3_b200 has no CPU fallback");
        return CSP3_ERR_CUDA;
    }
    if (batch == 0 || plan->nnz == 0) return 0;
    int dev = 0;
    CSP3_CUDA(cudaGetDevice(&dev));
    if (dev >= kMaxDev) { set_error("stack4_batched: device index too large"); return CSP3_ERR_ARG; }
    cudaStream_t st = (cudaStream_t)stream;
    {
        std::lock_guard<std::mutex> lock(plan->mu);
        if (!plan->dmap[dev]) {
            i32 *d = nullptr;
            CSP3_CUDA(cudaMalloc((void **)&d, (size_t)plan->nnz * 4));
            if (cudaMemcpy(d, plan->map.data(), (size_t)plan->nnz * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
                cudaFree(d);
                set_error("stack4_batched: upload failed (%s)", cudaGetErrorString(cudaGetLastError()));
                return CSP3_ERR_CUDA;
            }
            plan->dmap[dev] = d;
        }
    }
    Blocks4 B;
    for (int i = 0; i < 4; ++i) { B.x[i] = xs[i]; B.ld[i] = lds[i]; }
    const i64 per_cta = (i64)kStackThreads * kStackPerThread;
    const unsigned gx = (unsigned)((plan->nnz + per_cta - 1) / per_cta);
    // enough CTA rows to fill the machine a few times over; every row strides over the batch
    const i64 want = std::max<i64>(1, (i64)kNumSMs * 16 / gx);
    const unsigned gy = (unsigned)std::min<i64>(batch, std::min<i64>(want, 65535));
    stack4_gather_kernel<<<dim3(gx, gy), kStackThreads, 0, st>>>(plan->dmap[dev], plan->nnz, batch, B, out, ldo);
    CSP3_CUDA(cudaGetLastError());
    return 0;
}

int csp3_csc_stack_4_by_4_host(int64_t am, int64_t an, const int32_t *Ai, const int32_t *Ap, const double *Ax,
                               int64_t bm, int64_t bn, const int32_t *Bi, const int32_t *Bp, const double *Bx,
                               int64_t cm, int64_t cn, const int32_t *Ci, const int32_t *Cp, const double *Cx,
                               int64_t dm, int64_t dn, const int32_t *Di, const int32_t *Dp, const double *Dx,
                               int32_t *indices, int32_t *indptr, double *data)
{
    csp3_stack4 *P = nullptr;
    int rc = csp3_stack4_create(am, an, Ai, Ap, bm, bn, Bi, Bp, cm, cn, Ci, Cp, dm, dn, Di, Dp, &P);
    if (rc) return rc;
    rc = csp3_stack4_get_pattern(P, indices, indptr);
    const double *hx[4] = {Ax, Bx, Cx, Dx};
    double *dx[4] = {nullptr, nullptr, nullptr, nullptr}, *dout = nullptr;
    auto cleanup = [&]() { for (double *p : dx) if (p) cudaFree(p); if (dout) cudaFree(dout); csp3_stack4_destroy(P); };
    if (rc == 0 && P->nnz > 0) {
        int cnt = 0;
        if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt == 0) {
            cudaGetLastError();
            set_error("no CUDA device available: libcsparse3_b200 has no CPU fallback");
            cleanup();
            return CSP3_ERR_CUDA;
        }
        bool ok = cudaMalloc((void **)&dout, (size_t)P->nnz * 8) == cudaSuccess;
        for (int i = 0; i < 4 && ok; ++i)
            if (P->bnnz[i] > 0)
                ok = cudaMalloc((void **)&dx[i], (size_t)P->bnnz[i] * 8) == cudaSuccess &&
                     cudaMemcpy(dx[i], hx[i], (size_t)P->bnnz[i] * 8, cudaMemcpyHostToDevice) == cudaSuccess;
        if (ok) {
            rc = csp3_stack4_batched(P, 1, dx[0], 0, dx[1], 0, dx[2], 0, dx[3], 0, dout, P->nnz, nullptr);
            if (rc == 0) ok = cudaMemcpy(data, dout, (size_t)P->nnz * 8, cudaMemcpyDeviceToHost) == cudaSuccess;
        }
        if (!ok) { set_error("stack_4_by_4: device allocation or copy failed (%s)", cudaGetErrorString(cudaGetLastError())); rc = CSP3_ERR_ALLOC; }
    }
    cleanup();
    return rc;
}

}  // extern "C"
