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
            // prefetch the first two pair descriptors and the first L chunk while the accumulator is set up
            int4 pd0 = make_int4(0, 0, 0, 0), pd1 = pd0;
            if (pair_cnt > 0) pd0 = ld_meta(a.pairs + pair_ptr);
            if (pair_cnt > 1) pd1 = ld_meta(a.pairs + pair_ptr + 1);
            for (int t = e; t < len; t += E) acc[t * S] = 0.0;
            __syncwarp();
            for (int t = e; t < a_cnt; t += E) {
                const int src = __ldg(a.a_src + a_ptr + t);
                const int off = __ldg(a.a_off + a_ptr + t);
                acc[off * S] = __ldg(Axg + src);
            }
            int off0 = 0;
            double lv0 = 0.0;
            if (e < pd0.z) {
                off0 = __ldg(a.upd_map + pd0.w + e);
                lv0 = Lxg[pd0.y + e];
            }
            __syncwarp();
            for (int pi = 0; pi < pair_cnt; ++pi) {
                int4 pd2 = make_int4(0, 0, 0, 0);
                if (pi + 2 < pair_cnt) pd2 = ld_meta(a.pairs + pair_ptr + pi + 2);
                int off1 = 0;
                double lv1 = 0.0;
                if (e < pd1.z) {                       // pd1 is all-zero past the end
                    off1 = __ldg(a.upd_map + pd1.w + e);
                    lv1 = Lxg[pd1.y + e];
                }
                const double mult = acc[pd0.x * S];
                if (e < pd0.z) acc[off0 * S] = __dsub_rn(acc[off0 * S], __dmul_rn(lv0, mult));
                for (int t = e + E; t < pd0.z; t += E) {
                    const int off = __ldg(a.upd_map + pd0.w + t);
                    acc[off * S] = __dsub_rn(acc[off * S], __dmul_rn(Lxg[pd0.y + t], mult));
                }
                __syncwarp();
                pd0 = pd1; pd1 = pd2; off0 = off1; lv0 = lv1;
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

// Lanes: S systems x R rows x E entry-lanes (S*R*E == 32).  y lives slot-major / system-minor.
template <int S, int E>
__global__ void __launch_bounds__(1024) lu_solve_kernel(const SolveArgs a)
{
    constexpr int R = 32 / (S * E);
    extern __shared__ double smem[];
    double *y = a.scratch ? a.scratch + (size_t)blockIdx.x * a.n * S : smem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int sys = lane / (R * E), rsub = (lane / E) % R, e = lane % E;
    const i64 g0 = (i64)blockIdx.x * S;
    const i64 g_raw = g0 + sys;
    const bool valid = g_raw < a.batch;
    const i64 g = valid ? g_raw : a.batch - 1;
    const double *Lxg = a.Lx + g * a.lnz;
    const double *Uxg = a.Ux + g * a.unz;
    const int n = a.n;

    // y = P b   (cs_ipvec: y[pinv[i]] = b[i])
    for (int t = threadIdx.x; t < n * S; t += blockDim.x) {
        const int s = t / n, i = t - s * n;
        const i64 gs = (g0 + s < a.batch) ? g0 + s : a.batch - 1;
        y[(size_t)__ldg(a.pinv + i) * S + s] = __ldg(a.b + gs * n + i);
    }
    __syncthreads();
    // forward: y[r] -= sum_j L(r,j) y[j], rows level by level (level 0 rows have no entries)
    for (int l = 1; l < a.ls_nlev; ++l) {
        const int lbeg = __ldg(a.ls_lptr + l), lend = __ldg(a.ls_lptr + l + 1);
        for (int c = lbeg + warp * R + rsub; c < lend; c += nwarps * R) {
            const int r = __ldg(a.ls_order + c);
            const int pb = __ldg(a.lrow_ptr + r), pe = __ldg(a.lrow_ptr + r + 1);
            if (E == 1) {       // sequential, unfused: bit-identical to cs_lsolve's per-row update order
                double s = y[(size_t)r * S + sys];
                for (int t = pb; t < pe; ++t)
                    s = __dsub_rn(s, __dmul_rn(Lxg[__ldg(a.lrow_pos + t)], y[(size_t)__ldg(a.lrow_col + t) * S + sys]));
                y[(size_t)r * S + sys] = s;
            } else {
                double sum = 0.0;
                for (int t = pb + e; t < pe; t += E)
                    sum += Lxg[__ldg(a.lrow_pos + t)] * y[(size_t)__ldg(a.lrow_col + t) * S + sys];
#pragma unroll
                for (int o = E / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
                if (e == 0) y[(size_t)r * S + sys] -= sum;
            }
        }
        __syncthreads();
    }
    // backward: y[r] = (y[r] - sum_{j>r} U(r,j) y[j]) / U(r,r)
    for (int l = 0; l < a.us_nlev; ++l) {
        const int lbeg = __ldg(a.us_lptr + l), lend = __ldg(a.us_lptr + l + 1);
        for (int c = lbeg + warp * R + rsub; c < lend; c += nwarps * R) {
            const int r = __ldg(a.us_order + c);
            const int pb = __ldg(a.urow_ptr + r), pe = __ldg(a.urow_ptr + r + 1);
            const double d = Uxg[__ldg(a.Up + r + 1) - 1];
            if (E == 1) {       // cs_usolve visits columns in DESCENDING order: row entries right to left
                double s = y[(size_t)r * S + sys];
                for (int t = pe - 1; t >= pb; --t)
                    s = __dsub_rn(s, __dmul_rn(Uxg[__ldg(a.urow_pos + t)], y[(size_t)__ldg(a.urow_col + t) * S + sys]));
                y[(size_t)r * S + sys] = s / d;
            } else {
                double sum = 0.0;
                for (int t = pb + e; t < pe; t += E)
                    sum += Uxg[__ldg(a.urow_pos + t)] * y[(size_t)__ldg(a.urow_col + t) * S + sys];
#pragma unroll
                for (int o = E / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
                if (e == 0) y[(size_t)r * S + sys] = (y[(size_t)r * S + sys] - sum) / d;
            }
        }
        __syncthreads();
    }
    // x = Q y   (cs_ipvec: x[q[k]] = y[k])
    for (int t = threadIdx.x; t < n * S; t += blockDim.x) {
        const int s = t / n, k = t - s * n;
        if (g0 + s < a.batch) a.x[(g0 + s) * n + __ldg(a.q + k)] = y[(size_t)k * S + s];
    }
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

template <int S, int E>
int launch_solve_SE(const SolveArgs &a, int warps, size_t smem, cudaStream_t st)
{
    CSP3_CUDA(cudaFuncSetAttribute(lu_solve_kernel<S, E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const i64 grid = (a.batch + S - 1) / S;
    lu_solve_kernel<S, E><<<(unsigned)grid, warps * 32, smem, st>>>(a);
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
    int warps = tuning().rf_warps ? tuning().rf_warps : 8;
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
    int S = tuning().sv_S;
    if (S == 0) S = (batch >= 8 * kNumSMs) ? 2 : 1;
    while (S > 1 && (size_t)D.n * S * 8 > kMaxSmem / 2) S >>= 1;
    int warps = tuning().sv_warps ? tuning().sv_warps : 8;
    size_t smem = (size_t)D.n * S * 8;
    double *scratch = nullptr;
    if (smem > kMaxSmem) {                                   // y does not fit on chip: keep it in HBM/L2
        const i64 grid = (batch + S - 1) / S;
        CSP3_CUDA(cudaMallocAsync((void **)&scratch, (size_t)grid * D.n * S * 8, st));
        a.scratch = scratch;
        smem = 0;
        warps = 32;
    } else if (smem > 64 * 1024 && !tuning().sv_warps) {
        warps = 16;                                          // few CTAs per SM: more warps each
    }
    int rc;
    switch (S) {
        case 1: rc = launch_solve_SE<1, 1>(a, warps, smem, st); break;
        case 2: rc = launch_solve_SE<2, 1>(a, warps, smem, st); break;
        case 4: rc = launch_solve_SE<4, 1>(a, warps, smem, st); break;
        case 8: rc = launch_solve_SE<8, 1>(a, warps, smem, st); break;
        default: set_error("invalid solve bundle width %d", S); rc = -1;
    }
    if (scratch) CSP3_CUDA(cudaFreeAsync(scratch, st));
    return rc;
}

}  // namespace csp3
