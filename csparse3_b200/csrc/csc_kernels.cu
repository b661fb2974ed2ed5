// csc_kernels.cu -- CSC transposition / CSC->CSR, SpMV, SpMM and hash SpGEMM for sm_100a.
//
// Reference semantics (cited per kernel): src/CSparse3/csc_numba.py and src/sparsetools/csc.h, csr.h.
// All kernels are deterministic: no floating-point atomics on a contended address, so results do not depend
// on grid shape or on how a batch is sharded across GPUs.
#include "csc_kernels.cuh"

#include <algorithm>
#include <vector>

namespace csp3 {

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ int column_of(const i32 *__restrict__ Ap, int n, int p)
{
    // largest c with Ap[c] <= p  (empty columns are skipped because Ap[c+1] > p is required)
    int lo = 0, hi = n;                       // invariant: Ap[lo] <= p < Ap[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(Ap + mid) <= p) lo = mid; else hi = mid;
    }
    return lo;
}

// ---- counting sort: histogram -> scan -> bucket fill -> per-row rank sort ---------------------------------
__global__ void k_row_hist(int nnz, const i32 *__restrict__ Ai, i32 *cnt)
{
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < nnz; p += gridDim.x * blockDim.x)
        atomicAdd(cnt + __ldg(Ai + p), 1);
}

// single-CTA exclusive scan of cnt[0..m) into ptr[0..m]; cnt is overwritten with the running cursor (= ptr)
__global__ void k_scan(int m, i32 *cnt, i32 *ptr)
{
    __shared__ i32 warp_sum[32];
    __shared__ i32 carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int base = 0; base < m; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const i32 v = (i < m) ? cnt[i] : 0;
        i32 s = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const i32 t = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += t;
        }
        if (lane == 31) warp_sum[warp] = s;
        __syncthreads();
        if (warp == 0) {
            i32 ws = (lane < nw) ? warp_sum[lane] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const i32 t = __shfl_up_sync(0xffffffffu, ws, o);
                if (lane >= o) ws += t;
            }
            warp_sum[lane] = ws;
        }
        __syncthreads();
        const i32 excl = carry + (warp ? warp_sum[warp - 1] : 0) + s - v;
        if (i < m) { ptr[i] = excl; cnt[i] = excl; }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) ptr[m] = carry;
}

// same scan on 64-bit values, in place: v[0..m) -> exclusive prefix sums, v[m] = total
__global__ void k_scan64(int m, i64 *v)
{
    __shared__ i64 warp_sum[32];
    __shared__ i64 carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int base = 0; base < m; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const i64 x = (i < m) ? v[i] : 0;
        i64 s = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const i64 t = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += t;
        }
        if (lane == 31) warp_sum[warp] = s;
        __syncthreads();
        if (warp == 0) {
            i64 ws = (lane < nw) ? warp_sum[lane] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const i64 t = __shfl_up_sync(0xffffffffu, ws, o);
                if (lane >= o) ws += t;
            }
            warp_sum[lane] = ws;
        }
        __syncthreads();
        const i64 excl = carry + (warp ? warp_sum[warp - 1] : 0) + s - x;
        if (i < m) v[i] = excl;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = excl + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) v[m] = carry;
}

// Tiled scan for long arrays (the single-CTA kernels above walk 1,024 elements per trip: ~0.3 ms for 1e6 elements):
// tile sums, a single-CTA scan of the (few) tile sums, then every CTA scans its tile with its carry-in.
constexpr int kScanItems = 8, kScanTile = 1024 * kScanItems;
template <class T>
__global__ void __launch_bounds__(1024) k_tile_sums(int m, const T *__restrict__ in, i64 *sums)
{
    __shared__ i64 ws[32];
    const int base = blockIdx.x * kScanTile, end = min(m, base + kScanTile);
    i64 s = 0;
    for (int i = base + threadIdx.x; i < end; i += 1024) s += (i64)in[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = ws[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) sums[blockIdx.x] = s;
    }
}
// out0[i] (and out1[i], when given) = exclusive prefix of in[0..i); element m receives the total
template <class T>
__global__ void __launch_bounds__(1024) k_tile_scan(int m, const T *in, const i64 *__restrict__ tile_off, int ntiles, T *out0, T *out1)
{
    __shared__ i64 ws[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int first = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    T x[kScanItems];
    i64 mine = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) { x[k] = (first + k < m) ? in[first + k] : (T)0; mine += (i64)x[k]; }
    i64 s = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const i64 t = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += t;
    }
    if (lane == 31) ws[warp] = s;
    __syncthreads();
    if (warp == 0) {
        i64 w = ws[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const i64 t = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += t;
        }
        ws[lane] = w;
    }
    __syncthreads();
    i64 run = tile_off[blockIdx.x] + (warp ? ws[warp - 1] : 0) + s - mine;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        if (first + k < m) { out0[first + k] = (T)run; if (out1) out1[first + k] = (T)run; }
        run += (i64)x[k];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { out0[m] = (T)tile_off[ntiles]; }
}

__global__ void k_bucket_fill(int nnz, const i32 *__restrict__ Ai, i32 *cursor, i32 *bucket)
{
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < nnz; p += gridDim.x * blockDim.x)
        bucket[atomicAdd(cursor + __ldg(Ai + p), 1)] = p;
}

// One warp per row: sort the row's source-entry ids ascending (== ascending column, stable inside a column:
// exactly the order the sequential counting sort of csc_to_csr / csc_transpose produces), then emit.
// column of every entry (the CSC pattern expanded): one thread per column
__global__ void k_expand_cols(int n, const i32 *__restrict__ Ap, i32 *col)
{
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x)
        for (int p = __ldg(Ap + j); p < __ldg(Ap + j + 1); ++p) col[p] = j;
}

constexpr int kRowThreadMax = 48;       // rows up to this length are rank-sorted by one thread, longer ones by a CTA

// short rows (a handful of entries): one thread per row instead of one warp per row; a longer row is appended to
// `longrows` (count in longrows[0]) and left to k_row_emit_long, so one dense row never serialises in a thread
__global__ void k_row_emit_thread(int m, int n, const i32 *__restrict__ Ap, const double *__restrict__ Ax,
                                  const i32 *__restrict__ ptr, const i32 *__restrict__ bucket, const i32 *__restrict__ col,
                                  i32 *perm, i32 *Ci, double *Cx, i32 *longrows)
{
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < m; r += gridDim.x * blockDim.x) {
        const int beg = __ldg(ptr + r), len = __ldg(ptr + r + 1) - beg;
        if (len > kRowThreadMax) { longrows[1 + atomicAdd(longrows, 1)] = r; continue; }
        for (int t = 0; t < len; ++t) {
            const int p = __ldg(bucket + beg + t);
            int rank = 0;
            for (int u = 0; u < len; ++u) rank += (__ldg(bucket + beg + u) < p);
            if (perm) perm[beg + rank] = p;
            if (Ci) Ci[beg + rank] = col ? __ldg(col + p) : column_of(Ap, n, p);
            if (Cx) Cx[beg + rank] = __ldg(Ax + p);
        }
    }
}

// long rows of the worklist: one CTA per row.  Up to kLongSort entries are sorted in shared memory (bitonic, keys are
// the distinct source entry ids); longer rows fall back to ranking with the whole CTA (len^2 / 1024 steps per thread).
constexpr int kLongSort = 4096;
__global__ void __launch_bounds__(1024)
k_row_emit_long(int n, const i32 *__restrict__ Ap, const double *__restrict__ Ax, const i32 *__restrict__ ptr,
                const i32 *__restrict__ bucket, const i32 *__restrict__ col, i32 *perm, i32 *Ci, double *Cx,
                const i32 *__restrict__ longrows)
{
    __shared__ i32 key[kLongSort];
    const int nlong = longrows[0];
    for (int w = blockIdx.x; w < nlong; w += gridDim.x) {
        const int r = longrows[1 + w];
        const int beg = __ldg(ptr + r), len = __ldg(ptr + r + 1) - beg;
        if (len <= kLongSort) {
            int pw = 1;
            while (pw < len) pw <<= 1;
            for (int t = threadIdx.x; t < pw; t += blockDim.x) key[t] = t < len ? __ldg(bucket + beg + t) : INT32_MAX;
            __syncthreads();
            for (int k = 2; k <= pw; k <<= 1)
                for (int j = k >> 1; j > 0; j >>= 1) {
                    for (int t = threadIdx.x; t < pw; t += blockDim.x) {
                        const int u = t ^ j;
                        if (u > t) {
                            const i32 a = key[t], b = key[u];
                            if (((t & k) == 0) ? (a > b) : (a < b)) { key[t] = b; key[u] = a; }
                        }
                    }
                    __syncthreads();
                }
            for (int t = threadIdx.x; t < len; t += blockDim.x) {
                const int p = key[t];
                if (perm) perm[beg + t] = p;
                if (Ci) Ci[beg + t] = col ? __ldg(col + p) : column_of(Ap, n, p);
                if (Cx) Cx[beg + t] = __ldg(Ax + p);
            }
            __syncthreads();
        } else {
            for (int t = threadIdx.x; t < len; t += blockDim.x) {
                const int p = __ldg(bucket + beg + t);
                int rank = 0;
                for (int u = 0; u < len; ++u) rank += (__ldg(bucket + beg + u) < p);
                if (perm) perm[beg + rank] = p;
                if (Ci) Ci[beg + rank] = col ? __ldg(col + p) : column_of(Ap, n, p);
                if (Cx) Cx[beg + rank] = __ldg(Ax + p);
            }
        }
    }
}

// ---- SpMV / SpMM (row gather over the CSR view; values stay in the caller's CSC order) --------------------
// y[b][r] = beta*y[b][r] + sum_t Ax[b][perm[t]] * x[b][col[t]]     G lanes per row, batch on blockIdx.y
template <int G>
__global__ void k_spmv(int m, int n, i64 nnz_stride, const i32 *__restrict__ rp, const i32 *__restrict__ rc,
                       const i32 *__restrict__ perm, const double *__restrict__ Ax, const double *__restrict__ x,
                       double *y, double beta)
{
    const i64 b = blockIdx.y;
    const double *Axb = Ax + b * nnz_stride;
    const double *xb = x + b * n;
    double *yb = y + b * m;
    const int sub = threadIdx.x % G;
    for (int r = (blockIdx.x * blockDim.x + threadIdx.x) / G; r < m; r += gridDim.x * blockDim.x / G) {
        const int beg = __ldg(rp + r), end = __ldg(rp + r + 1);
        // G == 1: sequential unfused sum in ascending column order starting from y -- bit-identical to the
        // reference's column-scatter loop (csc_numba.py:325-327, sparsetools csc.h:36-44)
        double s = (G == 1 && beta != 0.0) ? beta * yb[r] : 0.0;
        for (int t = beg + sub; t < end; t += G)
            s = __dadd_rn(s, __dmul_rn(__ldg(Axb + __ldg(perm + t)), __ldg(xb + __ldg(rc + t))));
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (sub == 0) yb[r] = (G == 1 || beta == 0.0) ? s : beta * yb[r] + s;
    }
}

// same as k_spmv<1> with one packed index word per entry (patterns with nnz, n < 65536)
__global__ void k_spmv_packed(int m, int n, i64 nnz_stride, const i32 *__restrict__ rp, const uint32_t *__restrict__ pk,
                              const double *__restrict__ Ax, const double *__restrict__ x, double *y, double beta)
{
    const i64 b = blockIdx.y;
    const double *Axb = Ax + b * nnz_stride;
    const double *xb = x + b * n;
    double *yb = y + b * m;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < m; r += gridDim.x * blockDim.x) {
        const int beg = __ldg(rp + r), end = __ldg(rp + r + 1);
        double s = (beta != 0.0) ? beta * yb[r] : 0.0;
        for (int t = beg; t < end; ++t) {
            const uint32_t w = __ldg(pk + t);
            s = __dadd_rn(s, __dmul_rn(__ldg(Axb + (w & 0xffffu)), __ldg(xb + (w >> 16))));
        }
        yb[r] = s;
    }
}

__global__ void k_pack_index(i32 nnz, const i32 *__restrict__ perm, const i32 *__restrict__ rc, uint32_t *pk)
{
    for (i32 t = blockIdx.x * blockDim.x + threadIdx.x; t < nnz; t += gridDim.x * blockDim.x)
        pk[t] = (uint32_t)perm[t] | ((uint32_t)rc[t] << 16);
}

// Y[r][v] += sum_t Ax[perm[t]] * X[col[t]][v]   (row-major X, Y; thread per (r, v)), sparsetools csc.h:68-84
__global__ void k_spmm(int m, int nv, const i32 *__restrict__ rp, const i32 *__restrict__ rc,
                       const i32 *__restrict__ perm, const double *__restrict__ Ax, const double *__restrict__ X,
                       double *Y)
{
    const i64 total = (i64)m * nv;
    for (i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (i64)gridDim.x * blockDim.x) {
        const int r = (int)(idx / nv), v = (int)(idx - (i64)r * nv);
        const int beg = __ldg(rp + r), end = __ldg(rp + r + 1);
        double s = Y[idx];
        for (int t = beg; t < end; ++t)
            s = __dadd_rn(s, __dmul_rn(__ldg(Ax + __ldg(perm + t)), __ldg(X + (i64)__ldg(rc + t) * nv + v)));
        Y[idx] = s;
    }
}

// ---- SpGEMM (hash accumulation per output column) -----------------------------------------------------------
__global__ void k_spgemm_ub(int Bn, const i32 *__restrict__ Ap, const i32 *__restrict__ Bp,
                            const i32 *__restrict__ Bi, i32 *ub)
{
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < Bn; j += gridDim.x * blockDim.x) {
        i64 s = 0;
        for (int pb = __ldg(Bp + j); pb < __ldg(Bp + j + 1); ++pb) {
            const int k = __ldg(Bi + pb);
            s += __ldg(Ap + k + 1) - __ldg(Ap + k);
        }
        ub[j] = (i32)min(s, (i64)INT32_MAX);
    }
}

__device__ __forceinline__ unsigned hash_row(int r, unsigned mask) { return ((unsigned)r * 2654435761u >> 7) & mask; }

// insert key; returns slot.  *fresh = true when this call created the entry.
__device__ __forceinline__ int hash_insert(i32 *keys, unsigned mask, int row, bool *fresh)
{
    unsigned h = hash_row(row, mask);
    for (;;) {
        const i32 prev = atomicCAS(keys + h, -1, row);
        if (prev == -1) { *fresh = true; return (int)h; }
        if (prev == row) { *fresh = false; return (int)h; }
        h = (h + 1) & mask;
    }
}

constexpr int kSmallSlots = 1024;     // per-warp table for columns with <= 512 candidate products
constexpr int kSmallWarps = 3;

// Process one output column with `nthr` cooperating threads (a warp with a shared-memory table, or a whole
// CTA with a global-memory table).  Products are visited pb-sequentially and pa-parallel, so inside a step
// all target rows are distinct (A has no duplicate row inside a column) and the per-row summation order is
// the reference's loop order (csc_numba.py:284-293).
template <bool NUMERIC, bool BLOCK>
__device__ void spgemm_column(int j, int tid, int nthr, const i32 *__restrict__ Ap, const i32 *__restrict__ Ai,
                              const double *__restrict__ Ax, const i32 *__restrict__ Bp,
                              const i32 *__restrict__ Bi, const double *__restrict__ Bx, i32 *keys, double *vals,
                              unsigned mask, i32 *count_out, const i32 *__restrict__ Cp, i32 *Ci, double *Cx,
                              i32 *compact = nullptr)
{
    const int slots = (int)mask + 1;
    for (int s = tid; s < slots; s += nthr) { keys[s] = -1; if (NUMERIC) vals[s] = 0.0; }
    if (BLOCK) __syncthreads(); else __syncwarp();
    int fresh_cnt = 0;
    if (!BLOCK) {
        // Warp mode.  A step (one entry of B(:,j), i.e. one column of A) depends on three levels of global loads
        // (Bi -> Ap -> Ai / Ax); walked naively that is ~2,000 cycles of latency per step.  The lanes fetch the B
        // entries and the A column ranges of up to 32 steps at once, and the first 32 entries of the next step's A
        // column are loaded before the current step is accumulated.  The order of the sums is unchanged (steps
        // stay sequential).
        const int pb0 = __ldg(Bp + j), pb1 = __ldg(Bp + j + 1);
        for (int base = pb0; base < pb1; base += 32) {
            int a0_l = 0, a1_l = 0;
            double bv_l = 0.0;
            if (base + tid < pb1) {
                const int k = __ldg(Bi + base + tid);
                if (NUMERIC) bv_l = __ldg(Bx + base + tid);
                a0_l = __ldg(Ap + k); a1_l = __ldg(Ap + k + 1);
            }
            const int cnt = min(32, pb1 - base);
            int a0 = __shfl_sync(0xffffffffu, a0_l, 0), a1 = __shfl_sync(0xffffffffu, a1_l, 0);
            double bv = __shfl_sync(0xffffffffu, bv_l, 0);
            int row_n = -1;
            double av_n = 0.0;
            if (a0 + tid < a1) { row_n = __ldg(Ai + a0 + tid); if (NUMERIC) av_n = __ldg(Ax + a0 + tid); }
            for (int q = 0; q < cnt; ++q) {
                const int ca0 = a0, ca1 = a1;
                const double cbv = bv;
                int row = row_n;
                double av = av_n;
                if (q + 1 < cnt) {
                    a0 = __shfl_sync(0xffffffffu, a0_l, q + 1); a1 = __shfl_sync(0xffffffffu, a1_l, q + 1);
                    bv = __shfl_sync(0xffffffffu, bv_l, q + 1);
                    if (a0 + tid < a1) { row_n = __ldg(Ai + a0 + tid); if (NUMERIC) av_n = __ldg(Ax + a0 + tid); }
                }
                for (int pa = ca0 + tid; pa < ca1; pa += 32) {
                    if (pa >= ca0 + 32) { row = __ldg(Ai + pa); if (NUMERIC) av = __ldg(Ax + pa); }
                    bool fresh;
                    const int slot = hash_insert(keys, mask, row, &fresh);
                    fresh_cnt += fresh;
                    if (NUMERIC) atomicAdd(vals + slot, __dmul_rn(cbv, av));
                }
                __syncwarp();
            }
        }
    } else {
        for (int pb = __ldg(Bp + j); pb < __ldg(Bp + j + 1); ++pb) {
            const int k = __ldg(Bi + pb);
            const double bv = NUMERIC ? __ldg(Bx + pb) : 0.0;
            for (int pa = __ldg(Ap + k) + tid; pa < __ldg(Ap + k + 1); pa += nthr) {
                bool fresh;
                const int slot = hash_insert(keys, mask, __ldg(Ai + pa), &fresh);
                fresh_cnt += fresh;
                if (NUMERIC) atomicAdd(vals + slot, __dmul_rn(bv, __ldg(Ax + pa)));
            }
            __syncthreads();
        }
    }
    if (!NUMERIC) {
        if (fresh_cnt) atomicAdd(count_out + j, fresh_cnt);
        return;
    }
    // emit sorted by row: rank of each occupied slot among the occupied slots
    const int base = __ldg(Cp + j);
    if (!BLOCK && compact != nullptr) {
        // warp mode: the occupied keys are compacted first (at most slots / 2 of them), so ranking costs
        // occupied x occupied comparisons instead of occupied x slots
        int occ = 0;
        for (int s0 = 0; s0 < slots; s0 += 32) {
            const int row = keys[s0 + tid];
            const unsigned bal = __ballot_sync(0xffffffffu, row >= 0);
            if (row >= 0) compact[occ + __popc(bal & ((1u << tid) - 1u))] = row;
            occ += __popc(bal);
        }
        __syncwarp();
        for (int s = tid; s < slots; s += nthr) {
            const int row = keys[s];
            if (row < 0) continue;
            int rank = 0;
            for (int u = 0; u < occ; ++u) rank += (compact[u] < row);
            Ci[base + rank] = row;
            Cx[base + rank] = vals[s];
        }
        __syncwarp();
        return;
    }
    if (BLOCK && compact != nullptr) {
        // CTA mode: occupied keys are compacted into the CTA's global scratch (any order: ranks do not depend on it),
        // so ranking costs occupied x occupied / threads comparisons instead of occupied x slots
        __shared__ int s_occ;
        if (tid == 0) s_occ = 0;
        __syncthreads();
        for (int s = tid; s < slots; s += nthr) {
            const int row = keys[s];
            if (row >= 0) compact[atomicAdd(&s_occ, 1)] = row;
        }
        __syncthreads();
        const int occ = s_occ;
        for (int s = tid; s < slots; s += nthr) {
            const int row = keys[s];
            if (row < 0) continue;
            int rank = 0;
            for (int u = 0; u < occ; ++u) rank += (compact[u] < row);
            Ci[base + rank] = row;
            Cx[base + rank] = vals[s];
        }
        __syncthreads();
        return;
    }
    for (int s = tid; s < slots; s += nthr) {
        const int row = keys[s];
        if (row < 0) continue;
        int rank = 0;
        for (int u = 0; u < slots; ++u) { const int o = keys[u]; rank += (o >= 0 && o < row); }
        Ci[base + rank] = row;
        Cx[base + rank] = vals[s];
    }
    if (BLOCK) __syncthreads(); else __syncwarp();
}

// ---- short columns: one THREAD per output column --------------------------------------------------------------
// Power-flow Jacobians and Laplacians multiply into columns with a few dozen candidate products; a warp per column
// keeps 7 of 32 lanes busy and pays shared-memory atomics.  Here every thread owns a column and keeps its row list
// SORTED in shared memory ([slot][thread] layout: conflict-free), inserting the products in the reference's loop
// order (pb outer, pa inner: csc_numba.py:284-293), so every row's sum has the reference's order and no atomics are
// needed.  The lists of a warp's 32 columns are then written out cooperatively (coalesced).
constexpr int kThreadProducts = 64;     // columns with at most this many candidate products take the thread path ...
constexpr int kThreadRows = 32;         // ... in the numeric phase when they have at most this many distinct rows
constexpr int kThreadCols = 96;         // threads (columns) per CTA (36 KB of static shared memory in the numeric phase)

__device__ __forceinline__ bool spgemm_thread_symbolic(i32 ub) { return ub > 0 && ub <= kThreadProducts; }
__device__ __forceinline__ bool spgemm_thread_numeric(i32 ub, i32 cnt) { return ub > 0 && ub <= kThreadProducts && cnt <= kThreadRows; }

template <bool NUMERIC>
__global__ void __launch_bounds__(kThreadCols)
k_spgemm_thread(int Bn, const i32 *__restrict__ ub, const i32 *__restrict__ Ap, const i32 *__restrict__ Ai,
                const double *__restrict__ Ax, const i32 *__restrict__ Bp, const i32 *__restrict__ Bi,
                const double *__restrict__ Bx, i32 *count_out, const i32 *__restrict__ Cp, i32 *Ci, double *Cx)
{
    constexpr int SLOTS = NUMERIC ? kThreadRows : kThreadProducts;
    __shared__ i32 s_key[SLOTS][kThreadCols];
    __shared__ double s_val[NUMERIC ? SLOTS : 1][NUMERIC ? kThreadCols : 1];
    __shared__ i32 s_cp[kThreadCols / 32][33];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    for (int j0 = blockIdx.x * kThreadCols; j0 < Bn; j0 += gridDim.x * kThreadCols) {
        const int j = j0 + t;
        int len = 0;
        bool mine = false;
        if (j < Bn) {
            const i32 u = __ldg(ub + j);
            mine = NUMERIC ? spgemm_thread_numeric(u, __ldg(Cp + j + 1) - __ldg(Cp + j)) : spgemm_thread_symbolic(u);
        }
        if (mine) {
            for (int pb = __ldg(Bp + j); pb < __ldg(Bp + j + 1); ++pb) {
                const int k = __ldg(Bi + pb);
                const double bv = NUMERIC ? __ldg(Bx + pb) : 0.0;
                for (int pa = __ldg(Ap + k); pa < __ldg(Ap + k + 1); ++pa) {
                    const int r = __ldg(Ai + pa);
                    int i = len;
                    while (i > 0 && s_key[i - 1][t] > r) --i;
                    if (i > 0 && s_key[i - 1][t] == r) {
                        if (NUMERIC) s_val[i - 1][t] = __dadd_rn(s_val[i - 1][t], __dmul_rn(bv, __ldg(Ax + pa)));
                    } else {
                        for (int m = len; m > i; --m) { s_key[m][t] = s_key[m - 1][t]; if (NUMERIC) s_val[m][t] = s_val[m - 1][t]; }
                        s_key[i][t] = r;
                        if (NUMERIC) s_val[i][t] = __dmul_rn(bv, __ldg(Ax + pa));
                        ++len;
                    }
                }
            }
        }
        if (!NUMERIC) {
            if (mine) count_out[j] = len;          // (the other kernels add to a zeroed counter; these columns are only ours)
            continue;
        }
        // cooperative write-out: the 32 columns of a warp occupy one contiguous range of Ci / Cx
        __syncwarp();
        const int jw = j0 + warp * 32;
        s_cp[warp][lane] = __ldg(Cp + min(jw + lane, Bn));
        if (lane == 0) s_cp[warp][32] = __ldg(Cp + min(jw + 32, Bn));
        __syncwarp();
        const unsigned have = __ballot_sync(0xffffffffu, mine);
        const int e0 = s_cp[warp][0], e1 = s_cp[warp][32];
        for (int e = e0 + lane; e < e1; e += 32) {
            int lo = 0, hi = 32;                                    // column c with Cp[jw + c] <= e < Cp[jw + c + 1]
            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (s_cp[warp][mid] <= e) lo = mid; else hi = mid; }
            if ((have >> lo) & 1u) {
                const int slot = e - s_cp[warp][lo], tt = warp * 32 + lo;
                Ci[e] = s_key[slot][tt];
                Cx[e] = s_val[slot][tt];
            }
        }
        __syncwarp();
    }
}

// hash-table slots of output column j: 0 = empty column, <= kSmallSlots = a warp with a shared-memory table,
// more = a CTA with a table in global memory.  ub = candidate products of the column (k_spgemm_ub)
__device__ __forceinline__ i64 spgemm_slots(i32 ub, i64 Am)
{
    const i64 cand = min((i64)ub, Am);
    if (cand == 0) return 0;
    i64 slots = 32;
    while (slots < 2 * cand) slots <<= 1;
    return slots;
}

// table offsets of the big columns: off[j] = slots (big) or 0, then k_scan64
__global__ void k_spgemm_bigslots(int Bn, i64 Am, const i32 *__restrict__ ub, i64 *off)
{
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < Bn; j += gridDim.x * blockDim.x) {
        const i64 slots = spgemm_slots(__ldg(ub + j), Am);
        off[j] = slots > kSmallSlots ? slots : 0;
    }
}

__global__ void k_widen(int n, const i32 *__restrict__ in, i64 *out)
{
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) out[j] = in[j];
}
__global__ void k_narrow(int n, const i64 *__restrict__ in, i32 *out)
{
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) out[j] = (i32)min(in[j], (i64)INT32_MAX);
}

// Warp-per-column kernel for the columns whose table has (LO, SLOTS] slots; WARPS warps per CTA, each with its own
// table in shared memory.  Two instances are launched: tiny tables (<= 128 slots: 8 warps per CTA, the whole SM full
// of warps -- the kernel is bound by the latency of its dependent loads) and tables up to kSmallSlots.
template <bool NUMERIC, int LO, int SLOTS, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
k_spgemm_small_all(int Bn, i64 Am, const i32 *__restrict__ ub, const i32 *__restrict__ Ap, const i32 *__restrict__ Ai,
                   const double *__restrict__ Ax, const i32 *__restrict__ Bp, const i32 *__restrict__ Bi,
                   const double *__restrict__ Bx, i32 *count_out, const i32 *__restrict__ Cp, i32 *Ci, double *Cx)
{
    __shared__ i32 s_keys[WARPS][SLOTS];
    __shared__ double s_vals[NUMERIC ? WARPS : 1][NUMERIC ? SLOTS : 1];
    __shared__ i32 s_compact[NUMERIC ? WARPS : 1][NUMERIC ? SLOTS / 2 : 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int j = blockIdx.x * WARPS + warp; j < Bn; j += gridDim.x * WARPS) {
        const i64 slots = spgemm_slots(__ldg(ub + j), Am);
        if (slots <= LO || slots > SLOTS) continue;
        if (NUMERIC ? spgemm_thread_numeric(__ldg(ub + j), __ldg(Cp + j + 1) - __ldg(Cp + j)) : spgemm_thread_symbolic(__ldg(ub + j))) continue;
        spgemm_column<NUMERIC, false>(j, lane, 32, Ap, Ai, Ax, Bp, Bi, Bx, s_keys[warp],
                                      NUMERIC ? (double *)s_vals[warp] : (double *)nullptr, (unsigned)slots - 1u, count_out, Cp, Ci, Cx,
                                      NUMERIC ? (i32 *)s_compact[warp] : (i32 *)nullptr);
    }
}

template <bool NUMERIC>
__global__ void __launch_bounds__(kThreads)
k_spgemm_big_all(int Bn, i64 Am, const i32 *__restrict__ ub, i64 tab_slots, const i32 *__restrict__ Ap,
                 const i32 *__restrict__ Ai, const double *__restrict__ Ax, const i32 *__restrict__ Bp, const i32 *__restrict__ Bi,
                 const double *__restrict__ Bx, i32 *tab_keys, double *tab_vals, i32 *tab_compact, i32 *count_out,
                 const i32 *__restrict__ Cp, i32 *Ci, double *Cx)
{
    // every resident CTA owns ONE table sized for the largest big column and reuses it for all its columns: the
    // memory is O(CTAs x largest column), not O(sum over the big columns)
    const i64 off = (i64)blockIdx.x * tab_slots;
    for (int j = blockIdx.x; j < Bn; j += gridDim.x) {
        const i64 slots = spgemm_slots(__ldg(ub + j), Am);                  // uniform over the CTA
        if (slots <= kSmallSlots) continue;
        spgemm_column<NUMERIC, true>(j, threadIdx.x, blockDim.x, Ap, Ai, Ax, Bp, Bi, Bx, tab_keys + off,
                                     NUMERIC ? tab_vals + off : nullptr, (unsigned)slots - 1u, count_out, Cp, Ci, Cx,
                                     NUMERIC ? tab_compact + off / 2 : nullptr);
    }
}

// largest value of v[0..n) (single CTA; n is the number of columns)
__global__ void k_max64(int n, const i64 *__restrict__ v, i64 *out)
{
    __shared__ i64 part[1024];
    i64 m = 0;
    for (int j = threadIdx.x; j < n; j += blockDim.x) m = max(m, v[j]);
    part[threadIdx.x] = m;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) part[threadIdx.x] = max(part[threadIdx.x], part[threadIdx.x + o]);
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = part[0];
}

// ---- C = A + sign*B (sparsetools csr_binop_csr semantics: duplicates summed per operand, exact zeros dropped) ----
// One thread per column.  Inputs may be unsorted and may contain duplicates (the reference tolerates both); the
// distinct rows of a column are visited in first-touch order (A's entries, then B's), which costs O(len^2) per
// column -- columns of the matrices on this path have a handful of entries.  Output rows are sorted.
template <bool FILL>
__global__ void k_csc_add(int n, const i32 *__restrict__ Ap, const i32 *__restrict__ Ai, const double *__restrict__ Ax,
                          const i32 *__restrict__ Bp, const i32 *__restrict__ Bi, const double *__restrict__ Bx,
                          double sign, i32 *cnt, const i32 *__restrict__ Cp, i32 *Ci, double *Cx)
{
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const int a0 = __ldg(Ap + j), a1 = __ldg(Ap + j + 1), b0 = __ldg(Bp + j), b1 = __ldg(Bp + j + 1);
        const int la = a1 - a0, total = la + (b1 - b0);
        int out = FILL ? __ldg(Cp + j) : 0;
        const int base = out;
        for (int t = 0; t < total; ++t) {
            const int r = (t < la) ? __ldg(Ai + a0 + t) : __ldg(Bi + b0 + t - la);
            bool seen = false;
            for (int u = 0; u < t && !seen; ++u) seen = ((u < la) ? __ldg(Ai + a0 + u) : __ldg(Bi + b0 + u - la)) == r;
            if (seen) continue;
            double sa = 0.0, sb = 0.0;
            for (int u = a0; u < a1; ++u) if (__ldg(Ai + u) == r) sa = __dadd_rn(sa, __ldg(Ax + u));
            for (int u = b0; u < b1; ++u) if (__ldg(Bi + u) == r) sb = __dadd_rn(sb, __ldg(Bx + u));
            const double res = (sign > 0.0) ? __dadd_rn(sa, sb) : __dsub_rn(sa, sb);
            if (res != 0.0) {
                if (FILL) {                       // insertion keeps the column sorted by row
                    int q = out;
                    while (q > base && Ci[q - 1] > r) { Ci[q] = Ci[q - 1]; Cx[q] = Cx[q - 1]; --q; }
                    Ci[q] = r; Cx[q] = res;
                }
                ++out;
            }
        }
        if (!FILL) cnt[j] = out;
    }
}

// ---- C = alpha*A + beta*B in the reference's own kernel's form (csc_add_ff, src/CSparse3/csc_numba.py:183-219) ----
// Column j scatters alpha*A(:,j), then beta*B(:,j) into a dense accumulator: x[i] = coef*v at the first touch of row
// i, x[i] += coef*v afterwards; the pattern is emitted in FIRST-TOUCH order, explicit zeros are kept, duplicates are
// summed.  One thread per column; a row's value is accumulated in exactly that order with unfused multiply / add,
// so the result is bit-identical to the reference's (numba compiles it without fast-math).
template <bool FILL>
__global__ void k_csc_add_ff(int n, const i32 *__restrict__ Ap, const i32 *__restrict__ Ai, const double *__restrict__ Ax,
                             const i32 *__restrict__ Bp, const i32 *__restrict__ Bi, const double *__restrict__ Bx,
                             double alpha, double beta, i32 *cnt, const i32 *__restrict__ Cp, i32 *Ci, double *Cx)
{
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const int a0 = __ldg(Ap + j), a1 = __ldg(Ap + j + 1), b0 = __ldg(Bp + j), b1 = __ldg(Bp + j + 1);
        const int la = a1 - a0, total = la + (b1 - b0);
        int out = FILL ? __ldg(Cp + j) : 0;
        for (int t = 0; t < total; ++t) {
            const int r = (t < la) ? __ldg(Ai + a0 + t) : __ldg(Bi + b0 + t - la);
            bool seen = false;
            for (int u = 0; u < t && !seen; ++u) seen = ((u < la) ? __ldg(Ai + a0 + u) : __ldg(Bi + b0 + u - la)) == r;
            if (seen) continue;
            if (FILL) {
                double x = 0.0;
                bool first = true;
                for (int u = t; u < total; ++u) {
                    const int ru = (u < la) ? __ldg(Ai + a0 + u) : __ldg(Bi + b0 + u - la);
                    if (ru != r) continue;
                    const double term = (u < la) ? __dmul_rn(alpha, __ldg(Ax + a0 + u)) : __dmul_rn(beta, __ldg(Bx + b0 + u - la));
                    x = first ? term : __dadd_rn(x, term);
                    first = false;
                }
                Ci[out] = r; Cx[out] = x;
            }
            ++out;
        }
        if (!FILL) cnt[j] = out;
    }
}

inline int grid_for(i64 work, int per_block, int cap = kNumSMs * 16)
{
    const i64 g = (work + per_block - 1) / per_block;
    return (int)std::max<i64>(1, std::min<i64>(g, cap));
}

struct DevBuf {
    void *p = nullptr;
    cudaStream_t st;
    explicit DevBuf(cudaStream_t s) : st(s) {}
    ~DevBuf() { if (p) cudaFreeAsync(p, st); }
    int alloc(size_t bytes)
    {
        keep_device_pool();
        return cudaMallocAsync(&p, std::max<size_t>(bytes, 16), st) == cudaSuccess ? 0 : -1;
    }
    template <class T> T *as() { return (T *)p; }
};

}  // namespace

// The stream-ordered pool hands freed memory back to the driver at every synchronisation by default; the entry
// points synchronise (they read sizes back), so every call paid for fresh driver allocations (~1 ms).  Keep up to
// 1 GiB cached per device instead.
void keep_device_pool()
{
    static bool done[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || done[dev]) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t threshold = 1ull << 30;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
    }
    cudaGetLastError();
    done[dev] = true;
}

namespace {
// exclusive scans: cnt[0..m) -> ptr[0..m] (and cnt itself), ptr[m] = total; v[0..m) in place, v[m] = total
int scan_i32(int m, i32 *cnt, i32 *ptr, cudaStream_t st)
{
    if (m <= 4 * kScanTile) { k_scan<<<1, 1024, 0, st>>>(m, cnt, ptr); return 0; }
    const int ntiles = (m + kScanTile - 1) / kScanTile;
    DevBuf sums(st);
    if (sums.alloc((size_t)(ntiles + 1) * 8)) return -3;
    k_tile_sums<i32><<<ntiles, 1024, 0, st>>>(m, cnt, sums.as<i64>());
    k_scan64<<<1, 1024, 0, st>>>(ntiles, sums.as<i64>());
    k_tile_scan<i32><<<ntiles, 1024, 0, st>>>(m, cnt, sums.as<i64>(), ntiles, ptr, cnt);
    return 0;
}
int scan_i64(int m, i64 *v, cudaStream_t st)
{
    if (m <= 4 * kScanTile) { k_scan64<<<1, 1024, 0, st>>>(m, v); return 0; }
    const int ntiles = (m + kScanTile - 1) / kScanTile;
    DevBuf sums(st);
    if (sums.alloc((size_t)(ntiles + 1) * 8)) return -3;
    k_tile_sums<i64><<<ntiles, 1024, 0, st>>>(m, v, sums.as<i64>());
    k_scan64<<<1, 1024, 0, st>>>(ntiles, sums.as<i64>());
    k_tile_scan<i64><<<ntiles, 1024, 0, st>>>(m, v, sums.as<i64>(), ntiles, v, nullptr);
    return 0;
}
}  // namespace

// ---------------------------------------------------------------------------------------------------------
int transpose_device(i64 m, i64 n, const i32 *Ap, const i32 *Ai, const double *Ax, i32 nnz, i32 *Cp, i32 *Ci,
                     double *Cx, i32 *perm, cudaStream_t st)
{
    DevBuf cursor(st), bucket(st);
    if (cursor.alloc((size_t)(m + 1) * 4) || bucket.alloc((size_t)nnz * 4)) { set_error("device alloc failed"); return -3; }
    CSP3_CUDA(cudaMemsetAsync(cursor.p, 0, (size_t)(m + 1) * 4, st));
    if (nnz > 0) k_row_hist<<<grid_for(nnz, kThreads), kThreads, 0, st>>>(nnz, Ai, cursor.as<i32>());
    if (scan_i32((int)m, cursor.as<i32>(), Cp, st)) { set_error("device alloc failed"); return -3; }
    if (nnz > 0) {
        // the column of an entry: expanded once (4 bytes per entry) instead of a binary search over Ap per entry
        DevBuf col(st);
        const bool expand = Ci != nullptr && col.alloc((size_t)nnz * 4) == 0;
        if (expand) k_expand_cols<<<grid_for(n, kThreads), kThreads, 0, st>>>((int)n, Ap, col.as<i32>());
        k_bucket_fill<<<grid_for(nnz, kThreads), kThreads, 0, st>>>(nnz, Ai, cursor.as<i32>(), bucket.as<i32>());
        // the emit path is chosen per ROW (a matrix with a few entries per row on average may still hold a dense row):
        // rows up to kRowThreadMax entries by one thread each, the others through a worklist by one CTA each
        DevBuf longrows(st);
        if (longrows.alloc((size_t)(m + 1) * 4)) { set_error("device alloc failed"); return -3; }
        CSP3_CUDA(cudaMemsetAsync(longrows.p, 0, 4, st));
        k_row_emit_thread<<<grid_for(m, kThreads), kThreads, 0, st>>>((int)m, (int)n, Ap, Ax, Cp, bucket.as<i32>(),
                                                                      expand ? col.as<i32>() : nullptr, perm, Ci, Cx, longrows.as<i32>());
        k_row_emit_long<<<kNumSMs, 1024, 0, st>>>((int)n, Ap, Ax, Cp, bucket.as<i32>(), expand ? col.as<i32>() : nullptr, perm, Ci, Cx,
                                                  longrows.as<i32>());
    }
    CSP3_CUDA(cudaGetLastError());
    return 0;
}

int spmv_plan_pack(SpmvPlanData &P, cudaStream_t st)
{
    if (P.pk || P.nnz == 0 || P.nnz >= 65536 || P.n >= 65536) return 0;
    CSP3_CUDA(cudaMalloc((void **)&P.pk, (size_t)P.nnz * 4));
    k_pack_index<<<(unsigned)((P.nnz + 255) / 256), 256, 0, st>>>((i32)P.nnz, P.perm, P.rc, P.pk);
    CSP3_CUDA(cudaGetLastError());
    return 0;
}

int spmv_device(const SpmvPlanData &P, i64 batch, const double *Ax, i64 stride_ax, const double *x, double *y,
                double beta, cudaStream_t st)
{
    if (batch <= 0 || P.m == 0) return 0;
    // short rows (the power-grid / Laplacian case): one lane per row, exact reference summation order;
    // long rows: 8 lanes per row with a shuffle reduction (order differs, results agree to rounding)
    const bool wide = P.nnz > 48 * P.m;
    const i64 rows_per_block = wide ? kThreads / 8 : kThreads;
    i64 gx = (P.m + rows_per_block - 1) / rows_per_block;
    gx = std::max<i64>(1, std::min<i64>(gx, 65535 * 16));
    for (i64 b0 = 0; b0 < batch; b0 += 65535) {
        const i64 nb = std::min<i64>(65535, batch - b0);
        dim3 grid((unsigned)gx, (unsigned)nb);
        if (!wide && P.pk)
            k_spmv_packed<<<grid, kThreads, 0, st>>>((int)P.m, (int)P.n, stride_ax, P.rp, P.pk,
                                                     Ax + b0 * stride_ax, x + b0 * P.n, y + b0 * P.m, beta);
        else if (wide)
            k_spmv<8><<<grid, kThreads, 0, st>>>((int)P.m, (int)P.n, stride_ax, P.rp, P.rc, P.perm,
                                                 Ax + b0 * stride_ax, x + b0 * P.n, y + b0 * P.m, beta);
        else
            k_spmv<1><<<grid, kThreads, 0, st>>>((int)P.m, (int)P.n, stride_ax, P.rp, P.rc, P.perm,
                                                 Ax + b0 * stride_ax, x + b0 * P.n, y + b0 * P.m, beta);
    }
    CSP3_CUDA(cudaGetLastError());
    return 0;
}

int spmm_device(const SpmvPlanData &P, i64 nv, const double *Ax, const double *X, double *Y, cudaStream_t st)
{
    if (P.m == 0 || nv == 0) return 0;
    k_spmm<<<grid_for(P.m * nv, kThreads), kThreads, 0, st>>>((int)P.m, (int)nv, P.rp, P.rc, P.perm, Ax, X, Y);
    CSP3_CUDA(cudaGetLastError());
    return 0;
}

// Host-orchestrated SpGEMM phase.  numeric == false: fills Cp (exclusive scan of per-column distinct-row counts)
// and *nnz_out.  numeric == true: fills Ci (sorted inside columns) and Cx using Cp.
int spgemm_device(bool numeric, i64 Am, i64 An, const i32 *Ap, const i32 *Ai, const double *Ax, i64 Bm, i64 Bn,
                  const i32 *Bp, const i32 *Bi, const double *Bx, i32 *Cp, i32 *Ci, double *Cx, i64 *nnz_out,
                  cudaStream_t st)
{
    if (An != Bm) { set_error("spgemm: inner dimensions differ (%lld vs %lld)", (long long)An, (long long)Bm); return -1; }
    if (Bn == 0) { if (!numeric) { CSP3_CUDA(cudaMemsetAsync(Cp, 0, 4, st)); if (nnz_out) *nnz_out = 0; } return 0; }
    // Planning stays on the device (one 8-byte read-back for the size of the global hash tables): candidate
    // products per column -> table size per column -> offsets of the big columns' tables by a prefix sum.
    DevBuf ub(st), boff(st), tkeys(st), tvals(st), tcomp(st), cnt(st);
    if (ub.alloc((size_t)Bn * 4) || boff.alloc((size_t)(Bn + 1) * 8)) { set_error("device alloc failed"); return -3; }
    k_spgemm_ub<<<grid_for(Bn, kThreads), kThreads, 0, st>>>((int)Bn, Ap, Bp, Bi, ub.as<i32>());
    k_spgemm_bigslots<<<grid_for(Bn, kThreads), kThreads, 0, st>>>((int)Bn, Am, ub.as<i32>(), boff.as<i64>());
    k_max64<<<1, 1024, 0, st>>>((int)Bn, boff.as<i64>(), boff.as<i64>() + Bn);
    i64 tab_total = 0;                                   // slots of the largest big column (0: no big column)
    CSP3_CUDA(cudaMemcpyAsync(&tab_total, boff.as<i64>() + Bn, 8, cudaMemcpyDeviceToHost, st));
    CSP3_CUDA(cudaStreamSynchronize(st));
    // one table per resident CTA; the grid shrinks when the largest column is huge (at most ~1 GiB of tables)
    const int big_grid = tab_total > 0 ? (int)std::max<i64>(1, std::min<i64>(std::min<i64>(Bn, kNumSMs * 4), (1ll << 30) / (tab_total * 14))) : 0;
    if (tab_total > 0 && (tkeys.alloc((size_t)tab_total * big_grid * 4) ||
                          (numeric && (tvals.alloc((size_t)tab_total * big_grid * 8) || tcomp.alloc((size_t)tab_total * big_grid * 2))))) {
        set_error("spgemm: device alloc failed");
        return -3;
    }
    i32 *count = nullptr;
    if (!numeric) {
        if (cnt.alloc((size_t)(Bn + 1) * 4)) { set_error("device alloc failed"); return -3; }
        CSP3_CUDA(cudaMemsetAsync(cnt.p, 0, (size_t)(Bn + 1) * 4, st));
        count = cnt.as<i32>();
    }
    {
        constexpr int kTiny = 128, kTinyWarps = 8;
        const int g1 = grid_for(Bn, kTinyWarps, kNumSMs * 8), g2 = grid_for(Bn, kSmallWarps, kNumSMs * 5);
        const int g0 = grid_for(Bn, kThreadCols, kNumSMs * 16);
        if (numeric) k_spgemm_thread<true><<<g0, kThreadCols, 0, st>>>((int)Bn, ub.as<i32>(), Ap, Ai, Ax, Bp, Bi, Bx, nullptr, Cp, Ci, Cx);
        else k_spgemm_thread<false><<<g0, kThreadCols, 0, st>>>((int)Bn, ub.as<i32>(), Ap, Ai, Ax, Bp, Bi, Bx, count, nullptr, nullptr, nullptr);
        if (numeric) {
            k_spgemm_small_all<true, 0, kTiny, kTinyWarps><<<g1, kTinyWarps * 32, 0, st>>>((int)Bn, Am, ub.as<i32>(), Ap, Ai, Ax, Bp, Bi, Bx, nullptr, Cp, Ci, Cx);
            k_spgemm_small_all<true, kTiny, kSmallSlots, kSmallWarps><<<g2, kSmallWarps * 32, 0, st>>>((int)Bn, Am, ub.as<i32>(), Ap, Ai, Ax, Bp, Bi, Bx, nullptr, Cp, Ci, Cx);
        } else {
            k_spgemm_small_all<false, 0, kTiny, kTinyWarps><<<g1, kTinyWarps * 32, 0, st>>>((int)Bn, Am, ub.as<i32>(), Ap, Ai, Ax, Bp, Bi, Bx, count, nullptr, nullptr, nullptr);
            k_spgemm_small_all<false, kTiny, kSmallSlots, kSmallWarps><<<g2, kSmallWarps * 32, 0, st>>>((int)Bn, Am, ub.as<i32>(), Ap, Ai, Ax, Bp, Bi, Bx, count, nullptr, nullptr, nullptr);
        }
    }
    if (tab_total > 0) {
        const int g = big_grid;
        if (numeric) k_spgemm_big_all<true><<<g, kThreads, 0, st>>>((int)Bn, Am, ub.as<i32>(), tab_total, Ap, Ai, Ax, Bp, Bi, Bx, tkeys.as<i32>(), tvals.as<double>(), tcomp.as<i32>(), nullptr, Cp, Ci, Cx);
        else k_spgemm_big_all<false><<<g, kThreads, 0, st>>>((int)Bn, Am, ub.as<i32>(), tab_total, Ap, Ai, Ax, Bp, Bi, Bx, tkeys.as<i32>(), nullptr, nullptr, count, nullptr, nullptr, nullptr);
    }
    CSP3_CUDA(cudaGetLastError());
    if (!numeric) {
        // Cp = exclusive scan of the counts (64-bit on the device; overflow check as sparsetools csr.h:591-596)
        DevBuf wide(st);
        if (wide.alloc((size_t)(Bn + 1) * 8)) { set_error("device alloc failed"); return -3; }
        k_widen<<<grid_for(Bn, kThreads), kThreads, 0, st>>>((int)Bn, cnt.as<i32>(), wide.as<i64>());
        if (scan_i64((int)Bn, wide.as<i64>(), st)) { set_error("device alloc failed"); return -3; }
        k_narrow<<<grid_for(Bn + 1, kThreads), kThreads, 0, st>>>((int)Bn + 1, wide.as<i64>(), Cp);
        i64 run = 0;
        CSP3_CUDA(cudaMemcpyAsync(&run, wide.as<i64>() + Bn, 8, cudaMemcpyDeviceToHost, st));
        CSP3_CUDA(cudaStreamSynchronize(st));
        if (run > INT32_MAX) { set_error("nnz of the result is too large"); return -4; }
        if (nnz_out) *nnz_out = run;
    }
    return 0;
}

// C = A + sign*B.  Cp[n+1] is filled; Ci/Cx receive Cp[n] entries (caller capacity >= nnzA + nnzB).
int csc_add_device(i64 m, i64 n, const i32 *Ap, const i32 *Ai, const double *Ax, const i32 *Bp, const i32 *Bi,
                   const double *Bx, double sign, i32 *Cp, i32 *Ci, double *Cx, cudaStream_t st)
{
    (void)m;
    if (n == 0) { CSP3_CUDA(cudaMemsetAsync(Cp, 0, 4, st)); return 0; }
    DevBuf cnt(st);
    if (cnt.alloc((size_t)(n + 1) * 4)) { set_error("device alloc failed"); return -3; }
    k_csc_add<false><<<grid_for(n, kThreads), kThreads, 0, st>>>((int)n, Ap, Ai, Ax, Bp, Bi, Bx, sign, cnt.as<i32>(), nullptr, nullptr, nullptr);
    if (scan_i32((int)n, cnt.as<i32>(), Cp, st)) { set_error("device alloc failed"); return -3; }
    k_csc_add<true><<<grid_for(n, kThreads), kThreads, 0, st>>>((int)n, Ap, Ai, Ax, Bp, Bi, Bx, sign, nullptr, Cp, Ci, Cx);
    CSP3_CUDA(cudaGetLastError());
    return 0;
}

// C = alpha*A + beta*B, reference csc_add_ff semantics (first-touch order, zeros kept).  Cp[n+1] is filled; Ci/Cx
// receive Cp[n] entries (caller capacity >= nnzA + nnzB).
int csc_add_ff_device(i64 m, i64 n, const i32 *Ap, const i32 *Ai, const double *Ax, const i32 *Bp, const i32 *Bi,
                      const double *Bx, double alpha, double beta, i32 *Cp, i32 *Ci, double *Cx, cudaStream_t st)
{
    (void)m;
    if (n == 0) { CSP3_CUDA(cudaMemsetAsync(Cp, 0, 4, st)); return 0; }
    DevBuf cnt(st);
    if (cnt.alloc((size_t)(n + 1) * 4)) { set_error("device alloc failed"); return -3; }
    k_csc_add_ff<false><<<grid_for(n, kThreads), kThreads, 0, st>>>((int)n, Ap, Ai, Ax, Bp, Bi, Bx, alpha, beta, cnt.as<i32>(), nullptr, nullptr, nullptr);
    if (scan_i32((int)n, cnt.as<i32>(), Cp, st)) { set_error("device alloc failed"); return -3; }
    k_csc_add_ff<true><<<grid_for(n, kThreads), kThreads, 0, st>>>((int)n, Ap, Ai, Ax, Bp, Bi, Bx, alpha, beta, nullptr, Cp, Ci, Cx);
    CSP3_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace csp3
