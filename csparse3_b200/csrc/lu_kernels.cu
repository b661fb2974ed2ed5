// lu_kernels.cu -- batched fp64 LU refactorisation and triangular solves on one sparsity pattern (sm_100a).
//
// No reference counterpart exists; semantics are the frozen-pattern / frozen-pivot
// refactorisation and the cs_ipvec -> cs_lsolve -> cs_usolve -> cs_ipvec solve defined by
// oracle/csp3_oracle.c (orc_csc_lu_refactor, orc_csc_lu_solve).
//
// Layout in HBM (system-major, the reference's "leading batch dimension" convention):
//   Ax[batch][nnzA]  values in the caller's CSC entry order
//   Lx[batch][lnz], Ux[batch][unz]  factors in the cs_lu column layout of the symbolic object
//   b[batch][n], x[batch][n]
// All integer arrays (schedule) are shared by the whole batch and stay L2-resident.
//
// Refactor kernel: one CTA works on a bundle of S systems.  A warp owns one column of the current level
// for all S systems at once: lanes are split S x E (E = 32/S lanes over the entries of a column), so index
// loads are shared by S systems and every value load is a contiguous run per system.  The column is
// accumulated in a per-warp shared-memory buffer (slot-major, system-minor -> no bank conflicts between
// systems), updated left-looking with the finished columns L(:,j), j in U(:,k), then written ONCE to
// Ux / Lx.  Levels are separated by __syncthreads(); L(:,j) values written in an earlier level are read
// back through L1/L2 with plain (coherent) loads.
#include "common.cuh"

#include <cmath>

namespace csp3 {

namespace {

struct RefactorArgs {
    const int4 *cols;
    const i32 *a_src;
    const uint16_t *a_off;
    const int4 *pairs;
    const uint16_t *upd_map;
    const i32 *order, *lptr;
    i32 nlev, n, nnzA, lnz, unz, acc_stride;
    i64 batch;
    const double *Ax;
    double *Lx, *Ux;
    i32 *status;
};

__device__ __forceinline__ int4 ld_meta(const int4 *p) { return __ldg(p); }

template <int S>
__global__ void __launch_bounds__(1024) lu_refactor_kernel(const RefactorArgs a)
{
    constexpr int E = 32 / S;
    constexpr int NC = (S >= 4) ? 4 : 2;      // register-prefetched chunks of E entries per pair
    extern __shared__ double smem[];
    __shared__ int fail[32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int sys = lane / E, e = lane % E;
    const i64 g_raw = (i64)blockIdx.x * S + sys;
    const bool valid = g_raw < a.batch;
    const i64 g = valid ? g_raw : a.batch - 1;
    const double *Axg = a.Ax + g * a.nnzA;
    double *Lxg = a.Lx + g * a.lnz;
    double *Uxg = a.Ux + g * a.unz;
    double *acc = smem + (size_t)warp * a.acc_stride + sys;   // slot t of this system: acc[t * S]

    if (threadIdx.x < 32) fail[threadIdx.x] = INT32_MAX;
    __syncthreads();

    for (int l = 0; l < a.nlev; ++l) {
        const int lbeg = __ldg(a.lptr + l), lend = __ldg(a.lptr + l + 1);
        for (int c = lbeg + warp; c < lend; c += nwarps) {
            const int k = __ldg(a.order + c);
            const int4 c0 = ld_meta(a.cols + 2 * k), c1 = ld_meta(a.cols + 2 * k + 1);
            const int up = c0.x, lp = c0.y, ucnt = c0.z, lcnt = c0.w;
            const int a_ptr = c1.x, a_cnt = c1.y, pair_ptr = c1.z, pair_cnt = c1.w;
            const int len = ucnt + lcnt - 1;
            // prefetch the first two pair descriptors and the first pair's L entries while the accumulator is set up
            int4 pd0 = make_int4(0, 0, 0, 0), pd1 = pd0;
            if (pair_cnt > 0) pd0 = ld_meta(a.pairs + pair_ptr);
            if (pair_cnt > 1) pd1 = ld_meta(a.pairs + pair_ptr + 1);
            for (int t = e; t < len; t += E) acc[t * S] = 0.0;
            int off0[NC];
            double lv0[NC];
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                const int t = e + c * E;
                off0[c] = 0; lv0[c] = 0.0;
                if (t < pd0.z) { off0[c] = __ldg(a.upd_map + pd0.w + t); lv0[c] = Lxg[pd0.y + t]; }
            }
            __syncwarp();
            for (int t = e; t < a_cnt; t += E) {
                const int src = __ldg(a.a_src + a_ptr + t);
                const int off = __ldg(a.a_off + a_ptr + t);
                acc[off * S] = __ldg(Axg + src);
            }
            __syncwarp();
            for (int pi = 0; pi < pair_cnt; ++pi) {
                int4 pd2 = make_int4(0, 0, 0, 0);
                if (pi + 2 < pair_cnt) pd2 = ld_meta(a.pairs + pair_ptr + pi + 2);
                // issue the loads of the NEXT pair (independent of the accumulator) before touching this one
                int off1[NC];
                double lv1[NC];
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    const int t = e + c * E;
                    off1[c] = 0; lv1[c] = 0.0;
                    if (t < pd1.z) { off1[c] = __ldg(a.upd_map + pd1.w + t); lv1[c] = Lxg[pd1.y + t]; }   // pd1 = 0 past the end
                }
                const double mult = acc[pd0.x * S];
#pragma unroll
                for (int c = 0; c < NC; ++c)
                    if (e + c * E < pd0.z) acc[off0[c] * S] = __dsub_rn(acc[off0[c] * S], __dmul_rn(lv0[c], mult));
                for (int t = e + NC * E; t < pd0.z; t += E) {
                    const int off = __ldg(a.upd_map + pd0.w + t);
                    acc[off * S] = __dsub_rn(acc[off * S], __dmul_rn(Lxg[pd0.y + t], mult));
                }
                __syncwarp();
                pd0 = pd1; pd1 = pd2;
#pragma unroll
                for (int c = 0; c < NC; ++c) { off0[c] = off1[c]; lv0[c] = lv1[c]; }
            }
            // finalize: U(:,k) as accumulated, L(:,k) = x / pivot, unit diagonal first
            const double pivot = acc[(ucnt - 1) * S];
            if (valid) {
                for (int t = e; t < ucnt; t += E) Uxg[up + t] = acc[t * S];
                if (e == 0) Lxg[lp] = 1.0;
                for (int t = e; t < lcnt - 1; t += E) Lxg[lp + 1 + t] = acc[(ucnt + t) * S] / pivot;
                if (e == 0 && !(fabs(pivot) > 0.0 && isfinite(pivot))) atomicMin(&fail[sys], k + 1);
            }
            __syncwarp();
        }
        __syncthreads();
    }
    if (a.status != nullptr && threadIdx.x < S) {
        const i64 gs = (i64)blockIdx.x * S + threadIdx.x;
        if (gs < a.batch) a.status[gs] = (fail[threadIdx.x] == INT32_MAX) ? 0 : fail[threadIdx.x];
    }
}

struct SolveArgs {
    const i32 *pinv, *q, *Up;
    const i32 *lrow_ptr, *lrow_col, *lrow_pos, *urow_ptr, *urow_col, *urow_pos;
    const i32 *ls_order, *ls_lptr, *us_order, *us_lptr;
    i32 ls_nlev, us_nlev, n, lnz, unz;
    i64 batch;
    const double *Lx, *Ux, *b;
    double *x;
    double *scratch;       // global y[ceil(batch/S)][n*S] when shared memory is too small, else nullptr
};

// One CTA per system.  y lives in shared memory (or in an HBM/L2 scratch row when n is too large).  A level's
// rows are split over the warps; inside a warp, E lanes cooperate on one row (E = 4 in wide levels -> 8 rows
// per warp, E = 32 in narrow levels).  The lanes first fetch the row's factor values and y operands in
// PARALLEL and park the unfused products in shared memory; lane 0 of the group then subtracts them in the
// reference's order (cs_lsolve: ascending column; cs_usolve: descending column), which keeps the result
// bit-identical to the sequential algorithm while only one load round-trip is exposed per row.
constexpr int kProdCap = 256;                 // products parked per warp

template <bool UPPER>
__device__ __forceinline__ void solve_phase(const SolveArgs &a, double *y, double *prod, const double *Fx,
                                            const i32 *__restrict__ order, const i32 *__restrict__ lptr, int nlev,
                                            const i32 *__restrict__ rp, const i32 *__restrict__ rc,
                                            const i32 *__restrict__ rx)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int l = UPPER ? 0 : 1; l < nlev; ++l) {
        const int lbeg = __ldg(lptr + l), lend = __ldg(lptr + l + 1);
        const int E = (lend - lbeg > 2 * nwarps) ? 4 : 32;          // warp-uniform (CTA-uniform) per level
        const int R = 32 / E, cap = kProdCap / R;
        const int grp = lane / E, sub = lane - grp * E;
        double *gp = prod + grp * cap;
        for (int c0 = lbeg + warp * R; c0 < lend; c0 += nwarps * R) {
            const int c = c0 + grp;
            const bool active = c < lend;
            int r = 0, pb = 0, pe = 0;
            double s = 0.0, d = 1.0;
            if (active) {
                r = __ldg(order + c);
                pb = __ldg(rp + r); pe = __ldg(rp + r + 1);
                if (sub == 0) {
                    s = y[r];
                    if (UPPER) d = Fx[__ldg(a.Up + r + 1) - 1];
                }
            }
            const int len = pe - pb;
            for (int base = 0; __any_sync(0xffffffffu, base < len); base += cap) {
                const int stop = min(len, base + cap);
                for (int u = base + sub; u < stop; u += E) {
                    const int t = UPPER ? (pe - 1 - u) : (pb + u);       // cs_usolve walks the row right to left
                    gp[u - base] = __dmul_rn(Fx[__ldg(rx + t)], y[__ldg(rc + t)]);
                }
                __syncwarp();
                if (sub == 0)
                    for (int u = base; u < stop; ++u) s = __dsub_rn(s, gp[u - base]);
                __syncwarp();
            }
            if (active && sub == 0) y[r] = UPPER ? s / d : s;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(1024) lu_solve_kernel(const SolveArgs a)
{
    extern __shared__ double smem[];
    const int nwarps = blockDim.x >> 5, warp = threadIdx.x >> 5;
    const i64 g = blockIdx.x;
    double *y = a.scratch ? a.scratch + (size_t)g * a.n : smem + (size_t)nwarps * kProdCap;
    double *prod = smem + (size_t)warp * kProdCap;
    const double *Lxg = a.Lx + g * a.lnz;
    const double *Uxg = a.Ux + g * a.unz;
    const int n = a.n;
    // y = P b   (cs_ipvec: y[pinv[i]] = b[i])
    for (int i = threadIdx.x; i < n; i += blockDim.x) y[__ldg(a.pinv + i)] = __ldg(a.b + g * n + i);
    __syncthreads();
    solve_phase<false>(a, y, prod, Lxg, a.ls_order, a.ls_lptr, a.ls_nlev, a.lrow_ptr, a.lrow_col, a.lrow_pos);
    solve_phase<true>(a, y, prod, Uxg, a.us_order, a.us_lptr, a.us_nlev, a.urow_ptr, a.urow_col, a.urow_pos);
    // x = Q y   (cs_ipvec: x[q[k]] = y[k])
    for (int k = threadIdx.x; k < n; k += blockDim.x) a.x[g * n + __ldg(a.q + k)] = y[k];
}

template <int S>
int launch_refactor_S(const RefactorArgs &a, int warps, size_t smem, cudaStream_t st)
{
    CSP3_CUDA(cudaFuncSetAttribute(lu_refactor_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const i64 grid = (a.batch + S - 1) / S;
    lu_refactor_kernel<S><<<(unsigned)grid, warps * 32, smem, st>>>(a);
    CSP3_CUDA(cudaGetLastError());
    return 0;
}

constexpr size_t kMaxSmem = 200 * 1024;

}  // namespace

int launch_refactor(const DevSchedule &D, i64 batch, const double *Ax, double *Lx, double *Ux, i32 *status,
                    cudaStream_t st)
{
    if (batch <= 0) return 0;
    RefactorArgs a;
    a.cols = D.cols; a.a_src = D.a_src; a.a_off = D.a_off; a.pairs = D.pairs; a.upd_map = D.upd_map;
    a.order = D.rf_order; a.lptr = D.rf_lptr; a.nlev = D.rf_nlev;
    a.n = D.n; a.nnzA = D.nnzA; a.lnz = D.lnz; a.unz = D.unz;
    a.batch = batch; a.Ax = Ax; a.Lx = Lx; a.Ux = Ux; a.status = status;
    // bundle width: enough systems per CTA to fill lanes, bounded by the batch and by shared memory
    int S = tuning().rf_S;
    if (S == 0) S = (batch >= 8 * kNumSMs) ? 4 : (batch >= 2 * kNumSMs ? 2 : 1);
    int warps = tuning().rf_warps ? tuning().rf_warps : 2;
    const int len = D.max_col_len > 0 ? D.max_col_len : 1;
    while (S > 1 && (size_t)len * S * 8 * 2 > kMaxSmem) S >>= 1;
    a.acc_stride = len * S + 2;                              // +2 doubles: stagger warps across banks
    while (warps > 1 && (size_t)a.acc_stride * warps * 8 > kMaxSmem) warps >>= 1;
    const size_t smem = (size_t)a.acc_stride * warps * 8;
    if (smem > kMaxSmem) { set_error("factor column too long for shared memory (%d entries)", len); return -1; }
    switch (S) {
        case 1: return launch_refactor_S<1>(a, warps, smem, st);
        case 2: return launch_refactor_S<2>(a, warps, smem, st);
        case 4: return launch_refactor_S<4>(a, warps, smem, st);
        case 8: return launch_refactor_S<8>(a, warps, smem, st);
        case 16: return launch_refactor_S<16>(a, warps, smem, st);
        case 32: return launch_refactor_S<32>(a, warps, smem, st);
    }
    set_error("invalid refactor bundle width %d", S);
    return -1;
}

int launch_solve(const DevSchedule &D, i64 batch, const double *Lx, const double *Ux, const double *b,
                 double *x, cudaStream_t st)
{
    if (batch <= 0) return 0;
    SolveArgs a;
    a.pinv = D.pinv; a.q = D.q; a.Up = D.Up;
    a.lrow_ptr = D.lrow_ptr; a.lrow_col = D.lrow_col; a.lrow_pos = D.lrow_pos;
    a.urow_ptr = D.urow_ptr; a.urow_col = D.urow_col; a.urow_pos = D.urow_pos;
    a.ls_order = D.ls_order; a.ls_lptr = D.ls_lptr; a.us_order = D.us_order; a.us_lptr = D.us_lptr;
    a.ls_nlev = D.ls_nlev; a.us_nlev = D.us_nlev;
    a.n = D.n; a.lnz = D.lnz; a.unz = D.unz;
    a.batch = batch; a.Lx = Lx; a.Ux = Ux; a.b = b; a.x = x; a.scratch = nullptr;
    int warps = tuning().sv_warps ? tuning().sv_warps : 4;
    size_t smem = (size_t)(D.n + warps * kProdCap) * 8;
    double *scratch = nullptr;
    if (smem > kMaxSmem) {                                   // y does not fit on chip: keep it in HBM/L2
        CSP3_CUDA(cudaMallocAsync((void **)&scratch, (size_t)batch * D.n * 8, st));
        a.scratch = scratch;
        if (!tuning().sv_warps) warps = 16;
        smem = (size_t)warps * kProdCap * 8;
    } else if (smem > 96 * 1024 && !tuning().sv_warps) {
        warps = 16;                                          // one or two CTAs per SM: more warps each
        smem = (size_t)(D.n + warps * kProdCap) * 8;
    }
    CSP3_CUDA(cudaFuncSetAttribute(lu_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lu_solve_kernel<<<(unsigned)batch, warps * 32, smem, st>>>(a);
    CSP3_CUDA(cudaGetLastError());
    if (scratch) CSP3_CUDA(cudaFreeAsync(scratch, st));
    return 0;
}

}  // namespace csp3
