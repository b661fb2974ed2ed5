// topo_kernels.cu -- topology operations on the device.
//
// Reference: find_islands (src/CSparse3/csc_numba.py:743-808, caller CscMat.islands csc.py:515-521) and the
// sub-matrix kernels csc_sub_matrix / _cols / _rows (csc_numba.py:463-578, callers CscMat.__getitem__ csc.py:150-292).
// GridCal runs them before every N-1 case: does removing this branch split the grid, and which sub-matrix belongs
// to which island.  The reference does one graph per call on the host; here
//
//   * islands_batched: connected components of `batch` graphs that share one adjacency pattern and differ by one
//     removed edge each (the N-1 sweep), one CTA per case, labels in shared memory (or global memory for very large
//     graphs): hook-to-smaller-root + pointer jumping until nothing changes.  label[v] = smallest node id of v's
//     island, which is also the order in which the reference lists the islands (it scans the nodes in ascending
//     order); islands[case] = number of islands.
//   * sub_matrix: the three extraction kernels with the reference's exact output (order of the entries and its row
//     numbering rule, see k_sub_fill), two passes (count, prefix sum, fill).
#include <algorithm>
#include <vector>

#include "../../include/csparse3_b200.h"
#include "common.cuh"
#include "csc_kernels.cuh"

using namespace csp3;

namespace {

// ---- islands ---------------------------------------------------------------------------------------------------------
// parent[] lives in shared memory when it fits (n * 4 bytes), else in the case's slice of `scratch`.
__global__ void __launch_bounds__(256)
k_islands(int n, const i32 *__restrict__ indptr, const i32 *__restrict__ indices, const i32 *__restrict__ out_from,
          const i32 *__restrict__ out_to, i32 *__restrict__ label, i32 *__restrict__ islands, i32 *scratch, int use_smem)
{
    extern __shared__ i32 sh_parent[];
    __shared__ int changed;
    const i64 c = blockIdx.x;
    i32 *parent = use_smem ? sh_parent : scratch + c * n;
    const int rf = out_from ? out_from[c] : -1, rt = out_to ? out_to[c] : -1;
    for (int v = threadIdx.x; v < n; v += blockDim.x) parent[v] = v;
    __syncthreads();
    for (;;) {
        if (threadIdx.x == 0) changed = 0;
        __syncthreads();
        // hook: for every edge (u, v) the larger of the two current roots is attached under the smaller one
        for (int u = threadIdx.x; u < n; u += blockDim.x) {
            for (int p = __ldg(indptr + u); p < __ldg(indptr + u + 1); ++p) {
                const int v = __ldg(indices + p);
                if (v == u || (u == rf && v == rt) || (u == rt && v == rf)) continue;
                int ru = parent[u], rv = parent[v];
                // a few jumps towards the roots (the arrays are being compressed concurrently: any ancestor is valid)
                ru = parent[ru]; rv = parent[rv];
                if (ru == rv) continue;
                const int hi = max(ru, rv), lo = min(ru, rv);
                if (atomicMin(parent + hi, lo) > lo) changed = 1;
            }
        }
        __syncthreads();
        // compress: full pointer jumping
        for (int v = threadIdx.x; v < n; v += blockDim.x) {
            int r = parent[v];
            while (parent[r] != r) r = parent[r];
            parent[v] = r;
        }
        __syncthreads();
        if (!changed) break;
        __syncthreads();
    }
    int roots = 0;
    for (int v = threadIdx.x; v < n; v += blockDim.x) {
        const int r = parent[v];
        label[c * n + v] = r;
        roots += (r == v);
    }
    __shared__ int total;
    if (threadIdx.x == 0) total = 0;
    __syncthreads();
    if (roots) atomicAdd(&total, roots);
    __syncthreads();
    if (threadIdx.x == 0) islands[c] = total;
}

// Visiting order of the reference's traversal (find_islands pops the FRONT of its list: a breadth-first search from the
// smallest node of every island, neighbours in adjacency order, a node is listed at its FIRST appearance in the list).
// The first appearance of v is appended while its earliest-visited neighbour u is expanded, so inside an island nodes
// are ordered by the key (position of u, index of v in u's adjacency list), level by level.  One CTA; all islands
// advance together.  order[] receives the islands one after the other (ascending smallest node), isl_ptr[] their
// boundaries.  Work arrays (global): key[n] (u64), level[n], pos[n], fa[n], fb[n], isl_size[n], isl_start[n].
__global__ void __launch_bounds__(1024)
k_islands_order(int n, const i32 *__restrict__ indptr, const i32 *__restrict__ indices, const i32 *__restrict__ label,
                unsigned long long *key, i32 *level, i32 *pos, i32 *fa, i32 *fb, i32 *isl_size, i32 *isl_start,
                i32 *order, i32 *isl_ptr)
{
    __shared__ int fcount, ncount;
    const unsigned long long INF = ~0ull;
    for (int v = threadIdx.x; v < n; v += blockDim.x) { key[v] = INF; level[v] = -1; isl_size[v] = 0; }
    __syncthreads();
    for (int v = threadIdx.x; v < n; v += blockDim.x) atomicAdd(isl_size + label[v], 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0, k = 0;
        for (int v = 0; v < n; ++v) {
            isl_start[v] = run;
            if (label[v] == v) { isl_ptr[k++] = run; run += isl_size[v]; }
        }
        isl_ptr[k] = run;
        fcount = 0;
    }
    __syncthreads();
    // level 0: the roots; isl_size is reused as "nodes of the island placed so far"
    for (int v = threadIdx.x; v < n; v += blockDim.x) {
        isl_size[v] = 0;
        if (label[v] == v) { level[v] = 0; pos[v] = 0; fa[atomicAdd(&fcount, 1)] = v; }
    }
    __syncthreads();
    for (int v = threadIdx.x; v < n; v += blockDim.x) if (label[v] == v) isl_size[v] = 1;
    __syncthreads();
    i32 *cur = fa, *nxt = fb;
    for (int lev = 0;; ++lev) {
        const int F = fcount;
        if (F == 0) break;
        __syncthreads();
        if (threadIdx.x == 0) ncount = 0;
        // expand the frontier: every unvisited neighbour keeps its smallest (parent position, adjacency index)
        for (int t = threadIdx.x; t < F; t += blockDim.x) {
            const int u = cur[t];
            const unsigned long long pu = (unsigned long long)(unsigned)pos[u] << 32;
            for (int p = __ldg(indptr + u); p < __ldg(indptr + u + 1); ++p) {
                const int v = __ldg(indices + p);
                if (level[v] < 0) atomicMin(key + v, pu | (unsigned)(p - __ldg(indptr + u)));
            }
        }
        __syncthreads();
        // the nodes reached in this level (found through their parents' adjacency lists again: no scan over all nodes)
        for (int t = threadIdx.x; t < F; t += blockDim.x) {
            const int u = cur[t];
            for (int p = __ldg(indptr + u); p < __ldg(indptr + u + 1); ++p) {
                const int v = __ldg(indices + p);
                if (level[v] < 0 && atomicCAS(level + v, -1, lev + 1) == -1) nxt[atomicAdd(&ncount, 1)] = v;
            }
        }
        __syncthreads();
        const int Nn = ncount;
        // rank inside (island, level) by key
        for (int t = threadIdx.x; t < Nn; t += blockDim.x) {
            const int v = nxt[t];
            const int lb = label[v];
            const unsigned long long kv = key[v];
            int rank = 0;
            for (int w = 0; w < Nn; ++w) { const int x = nxt[w]; rank += (label[x] == lb && key[x] < kv); }
            pos[v] = isl_size[lb] + rank;
        }
        __syncthreads();
        for (int t = threadIdx.x; t < Nn; t += blockDim.x) atomicAdd(isl_size + label[nxt[t]], 1);
        __syncthreads();
        if (threadIdx.x == 0) fcount = Nn;
        i32 *tmp = cur; cur = nxt; nxt = tmp;
        __syncthreads();
    }
    for (int v = threadIdx.x; v < n; v += blockDim.x) order[isl_start[label[v]] + pos[v]] = v;
}

// ---- sub-matrices ------------------------------------------------------------------------------------------------------
// Reference loop (csc_sub_matrix, csc_numba.py:463-502; csc_sub_matrix_rows :541-578 is the same with cols = all):
//     for j in cols: i = 0
//         for r in rows: for k in A(:,j): if Ai[k] == r: emit (row index i, Ax[k]); i += 1
//                        if i == 0: i += 1
// i.e. entries come out ordered by (position of the row in `rows`, k), and the emitted row index is the running
// count of emitted entries of the column, shifted by one when rows[0] does not occur in the column (the reference's
// numbering rule; reproduced as is).  rowpos[r] = position of r in `rows` or -1 (rows must not repeat).
// One thread per selected column; columns on this path hold a handful of entries.
template <bool FILL>
__global__ void k_sub_fill(int ncols, const i32 *__restrict__ cols, const i32 *__restrict__ Ap, const i32 *__restrict__ Ai,
                           const double *__restrict__ Ax, const i32 *__restrict__ rowpos, int first_row, i32 *cnt,
                           const i32 *__restrict__ Bp, i32 *Bi, double *Bx)
{
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < ncols; t += gridDim.x * blockDim.x) {
        const int j = cols ? __ldg(cols + t) : t;
        const int a0 = __ldg(Ap + j), a1 = __ldg(Ap + j + 1);
        if (!FILL) {
            int c = 0;
            for (int k = a0; k < a1; ++k) c += rowpos ? (__ldg(rowpos + __ldg(Ai + k)) >= 0) : 1;
            cnt[t] = c;
            continue;
        }
        int out = __ldg(Bp + t);
        if (!rowpos) {                                                  // all rows, original indices (csc_sub_matrix_cols)
            for (int k = a0; k < a1; ++k) { Bi[out] = __ldg(Ai + k); Bx[out] = __ldg(Ax + k); ++out; }
            continue;
        }
        bool has_first = false;
        for (int k = a0; k < a1; ++k) has_first = has_first || (__ldg(Ai + k) == first_row);
        // selection by increasing (rowpos, k): repeated minimum search (columns are short)
        int last_pos = -1, last_k = -1, i = 0;
        for (;;) {
            int best_pos = INT32_MAX, best_k = -1;
            for (int k = a0; k < a1; ++k) {
                const int ps = __ldg(rowpos + __ldg(Ai + k));
                if (ps < 0) continue;
                if (ps < last_pos || (ps == last_pos && k <= last_k)) continue;
                if (ps < best_pos || (ps == best_pos && k < best_k)) { best_pos = ps; best_k = k; }
            }
            if (best_k < 0) break;
            if (i == 0 && best_pos > 0 && !has_first) i = 1;          // "if i == 0: i += 1" after the first row found nothing
            Bi[out] = i; Bx[out] = __ldg(Ax + best_k);
            ++out; ++i;
            last_pos = best_pos; last_k = best_k;
        }
    }
}

struct DBuf {
    void *p = nullptr;
    ~DBuf() { if (p) cudaFree(p); }
    int alloc(size_t bytes) { return cudaMalloc(&p, std::max<size_t>(bytes, 16)) == cudaSuccess ? 0 : -1; }
    int put(const void *src, size_t bytes)
    {
        if (alloc(bytes)) return -1;
        return bytes == 0 || cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice) == cudaSuccess ? 0 : -1;
    }
    template <class T> T *as() { return (T *)p; }
};

int need_device()
{
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt == 0) {
        cudaGetLastError();
        set_error("no CUDA device available: libcsparse3_b200 has no CPU fallback");
        return CSP3_ERR_CUDA;
    }
    return 0;
}

}  // namespace

extern "C" {

int csp3_islands_batched(int64_t n, const int32_t *indptr, const int32_t *indices, int64_t batch, const int32_t *out_from,
                         const int32_t *out_to, int32_t *label, int32_t *islands, void *stream)
{
    if (n < 0 || batch < 0 || !indptr || (!indices && n > 0) || !label || !islands || (out_from == nullptr) != (out_to == nullptr)) {
        set_error("islands_batched: bad arguments");
        return CSP3_ERR_ARG;
    }
    if (n == 0 || batch == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = (size_t)n * 4;
    const bool use_smem = smem <= 200 * 1024;
    i32 *scratch = nullptr;
    if (!use_smem) CSP3_CUDA(cudaMallocAsync((void **)&scratch, (size_t)batch * n * 4, st));
    else CSP3_CUDA(cudaFuncSetAttribute(k_islands, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_islands<<<(unsigned)batch, 256, use_smem ? smem : 0, st>>>((int)n, indptr, indices, out_from, out_to, label, islands, scratch, use_smem ? 1 : 0);
    CSP3_CUDA(cudaGetLastError());
    if (scratch) CSP3_CUDA(cudaFreeAsync(scratch, st));
    return 0;
}

int csp3_islands_batched_host(int64_t n, const int32_t *indptr, const int32_t *indices, int64_t batch, const int32_t *out_from,
                              const int32_t *out_to, int32_t *label, int32_t *islands)
{
    if (n < 0 || batch < 0 || !indptr || !label || !islands) { set_error("islands_batched: bad arguments"); return CSP3_ERR_ARG; }
    if (int rc = need_device()) return rc;
    if (n == 0 || batch == 0) return 0;
    const i64 nnz = indptr[n];
    for (i64 p = 0; p < nnz; ++p) if (indices[p] < 0 || indices[p] >= n) { set_error("islands_batched: index out of range"); return CSP3_ERR_ARG; }
    DBuf dp, di, df, dt, dl, dc;
    if (dp.put(indptr, (size_t)(n + 1) * 4) || di.put(indices, (size_t)nnz * 4) || dl.alloc((size_t)batch * n * 4) || dc.alloc((size_t)batch * 4) ||
        (out_from && (df.put(out_from, (size_t)batch * 4) || dt.put(out_to, (size_t)batch * 4)))) {
        cudaGetLastError(); set_error("islands_batched: device allocation or copy failed"); return CSP3_ERR_ALLOC;
    }
    if (int rc = csp3_islands_batched(n, dp.as<i32>(), di.as<i32>(), batch, out_from ? df.as<i32>() : nullptr, out_from ? dt.as<i32>() : nullptr,
                                      dl.as<i32>(), dc.as<i32>(), nullptr)) return rc;
    CSP3_CUDA(cudaDeviceSynchronize());
    CSP3_CUDA(cudaMemcpy(label, dl.p, (size_t)batch * n * 4, cudaMemcpyDeviceToHost));
    CSP3_CUDA(cudaMemcpy(islands, dc.p, (size_t)batch * 4, cudaMemcpyDeviceToHost));
    return 0;
}

int csp3_find_islands_host(int64_t n, const int32_t *indptr, const int32_t *indices, int32_t *order, int32_t *island_ptr,
                           int64_t *n_islands)
{
    if (n < 0 || !indptr || !order || !island_ptr || !n_islands) { set_error("find_islands: bad arguments"); return CSP3_ERR_ARG; }
    if (int rc = need_device()) return rc;
    *n_islands = 0;
    island_ptr[0] = 0;
    if (n == 0) return 0;
    const i64 nnz = indptr[n];
    for (i64 p = 0; p < nnz; ++p) if (indices[p] < 0 || indices[p] >= n) { set_error("find_islands: index out of range"); return CSP3_ERR_ARG; }
    DBuf dp, di, dl, dc, dkey, dwork, dord, dptr;
    if (dp.put(indptr, (size_t)(n + 1) * 4) || di.put(indices, (size_t)nnz * 4) || dl.alloc((size_t)n * 4) || dc.alloc(4) ||
        dkey.alloc((size_t)n * 8) || dwork.alloc((size_t)n * 4 * 6) || dord.alloc((size_t)n * 4) || dptr.alloc((size_t)(n + 1) * 4)) {
        cudaGetLastError(); set_error("find_islands: device allocation or copy failed"); return CSP3_ERR_ALLOC;
    }
    if (int rc = csp3_islands_batched(n, dp.as<i32>(), di.as<i32>(), 1, nullptr, nullptr, dl.as<i32>(), dc.as<i32>(), nullptr)) return rc;
    i32 *w = dwork.as<i32>();
    k_islands_order<<<1, 1024>>>((int)n, dp.as<i32>(), di.as<i32>(), dl.as<i32>(), dkey.as<unsigned long long>(), w, w + n, w + 2 * n, w + 3 * n,
                                 w + 4 * n, w + 5 * n, dord.as<i32>(), dptr.as<i32>());
    CSP3_CUDA(cudaGetLastError());
    CSP3_CUDA(cudaDeviceSynchronize());
    i32 cnt = 0;
    CSP3_CUDA(cudaMemcpy(&cnt, dc.p, 4, cudaMemcpyDeviceToHost));
    CSP3_CUDA(cudaMemcpy(order, dord.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
    CSP3_CUDA(cudaMemcpy(island_ptr, dptr.p, (size_t)(cnt + 1) * 4, cudaMemcpyDeviceToHost));
    *n_islands = cnt;
    return 0;
}

int csp3_csc_sub_matrix_host(int64_t Am, int64_t An, const int32_t *Ap, const int32_t *Ai, const double *Ax, int64_t nrows,
                             const int32_t *rows, int64_t ncols, const int32_t *cols, int32_t *Bp, int32_t *Bi, double *Bx, int64_t *nnz)
{
    if (Am < 0 || An < 0 || !Ap || !Bp || !nnz || nrows < 0 || ncols < 0 || (nrows > 0 && !rows && false)) { set_error("csc_sub_matrix: bad arguments"); return CSP3_ERR_ARG; }
    if (int rc = need_device()) return rc;
    const i64 nz = Ap[An];
    const i64 nc = cols ? ncols : An;
    std::vector<i32> rowpos;
    if (rows) {
        rowpos.assign((size_t)std::max<i64>(Am, 1), -1);
        for (i64 t = 0; t < nrows; ++t) {
            if (rows[t] < 0 || rows[t] >= Am) { set_error("csc_sub_matrix: row index out of range"); return CSP3_ERR_ARG; }
            if (rowpos[(size_t)rows[t]] >= 0) { set_error("csc_sub_matrix: repeated row index %d (not supported by the device path)", rows[t]); return CSP3_ERR_ARG; }
            rowpos[(size_t)rows[t]] = (i32)t;
        }
    }
    if (cols) for (i64 t = 0; t < ncols; ++t) if (cols[t] < 0 || cols[t] >= An) { set_error("csc_sub_matrix: column index out of range"); return CSP3_ERR_ARG; }
    Bp[0] = 0;
    if (nc == 0) { *nnz = 0; return 0; }
    DBuf dAp, dAi, dAx, dcols, dpos, dcnt, dBp, dBi, dBx;
    if (dAp.put(Ap, (size_t)(An + 1) * 4) || dAi.put(Ai, (size_t)nz * 4) || dAx.put(Ax, (size_t)nz * 8) || dcnt.alloc((size_t)(nc + 1) * 4) ||
        dBp.alloc((size_t)(nc + 1) * 4) || dBi.alloc((size_t)nz * 4) || dBx.alloc((size_t)nz * 8) ||
        (cols && dcols.put(cols, (size_t)ncols * 4)) || (rows && dpos.put(rowpos.data(), rowpos.size() * 4))) {
        cudaGetLastError(); set_error("csc_sub_matrix: device allocation or copy failed"); return CSP3_ERR_ALLOC;
    }
    const int grid = (int)std::max<i64>(1, std::min<i64>((nc + 255) / 256, kNumSMs * 8));
    const int first_row = rows && nrows > 0 ? rows[0] : -1;
    k_sub_fill<false><<<grid, 256>>>((int)nc, cols ? dcols.as<i32>() : nullptr, dAp.as<i32>(), dAi.as<i32>(), dAx.as<double>(),
                                     rows ? dpos.as<i32>() : nullptr, first_row, dcnt.as<i32>(), nullptr, nullptr, nullptr);
    std::vector<i32> cnt((size_t)nc);
    CSP3_CUDA(cudaMemcpy(cnt.data(), dcnt.p, (size_t)nc * 4, cudaMemcpyDeviceToHost));
    for (i64 t = 0; t < nc; ++t) Bp[t + 1] = Bp[t] + cnt[(size_t)t];          // ncols + 1 integers: pointer bookkeeping on the host
    CSP3_CUDA(cudaMemcpy(dBp.p, Bp, (size_t)(nc + 1) * 4, cudaMemcpyHostToDevice));
    k_sub_fill<true><<<grid, 256>>>((int)nc, cols ? dcols.as<i32>() : nullptr, dAp.as<i32>(), dAi.as<i32>(), dAx.as<double>(),
                                    rows ? dpos.as<i32>() : nullptr, first_row, nullptr, dBp.as<i32>(), dBi.as<i32>(), dBx.as<double>());
    CSP3_CUDA(cudaGetLastError());
    CSP3_CUDA(cudaDeviceSynchronize());
    *nnz = Bp[nc];
    CSP3_CUDA(cudaMemcpy(Bi, dBi.p, (size_t)Bp[nc] * 4, cudaMemcpyDeviceToHost));
    CSP3_CUDA(cudaMemcpy(Bx, dBx.p, (size_t)Bp[nc] * 8, cudaMemcpyDeviceToHost));
    return 0;
}

}  // extern "C"
