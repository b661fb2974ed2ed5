// nr_kernels.cu -- Newton-Raphson power-flow iteration kept on the device.
//
// The consumer loop of the reference (GridCal): J = pack_4_by_4(H, N, M, L)
// (src/CSparse3/csc.py:588-606), mismatch f = S_calc - S_spec with S_calc = V * conj(Ybus * V)
// (CscMat.__mul__ on a vector, csc.py:374-379), then factor + solve + update.  The reference runs every piece on
// the host, one case at a time; here a batch of same-topology cases iterates entirely on the GPU:
//
//   nr_rect_kernel        V = vm * exp(j va)                                          [batch, n_bus] complex
//   nr_current_kernel     I = Ybus * V (row gather over the CSR pattern of Ybus, optional per-case branch outage),
//                         mismatch -> right-hand side b = -(S_calc - S_spec) in the (pvpq | pq) equation order,
//                         fnorm = max |mismatch|
//   nr_jacobian_kernel    the four polar Jacobian blocks written straight into the refactorisation's Ax[batch, nnz]
//                         (the 2 x 2 block stacking of pack_4_by_4 is a precomputed entry map: one thread per entry)
//   lu refactor + solve   (lu_wide.cu / lu_kernels.cu through the workspace path)
//   nr_update_kernel      va[pvpq] += dx, vm[pq] += dx, next V
//
// Formulas (MATPOWER dSbus_dV, the form csparse3_b200/synth.py generates the benchmark Jacobians with):
//   dS/dVa = j diag(V) conj(diag(I) - Ybus diag(V)),   dS/dVm = diag(V) conj(Ybus diag(V/|V|)) + conj(diag(I)) diag(V/|V|)
//   J = [[Re dS/dVa, Re dS/dVm], [Im dS/dVa, Im dS/dVm]] restricted to (pvpq, pq).
#include <algorithm>
#include <cstring>
#include <memory>
#include <vector>

#include "../../include/csparse3_b200.h"
#include "common.cuh"

using namespace csp3;

struct csp3_nr_plan {
    i64 n_bus = 0, nnz_y = 0, npvpq = 0, npq = 0, n = 0, jnnz = 0, n_branch = 0;
    csp3_lu_symbolic *sym = nullptr;
    int device = -1;
    // device arrays
    i32 *y_rowptr = nullptr, *y_col = nullptr, *pos_th = nullptr, *pos_v = nullptr, *br_slot = nullptr;
    int4 *jmap = nullptr;              // per Jacobian entry: (row bus, col bus, Ybus entry, block)
    double2 *y_val = nullptr, *br_val = nullptr;
    void *arena = nullptr;
    // host staging of csp3_nr_solve_host (lazily created)
    struct Stage {
        bool ready = false;
        i64 chunk = 0;
        cudaStream_t st[2] = {nullptr, nullptr};
        double *sspec[2] = {}, *vm[2] = {}, *va[2] = {}, *fnorm[2] = {};
        i32 *status[2] = {}, *outb[2] = {};
        void *work[2] = {};
    } stage;
};

namespace {

__device__ __forceinline__ double2 cmul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double2 cconj(double2 a) { return make_double2(a.x, -a.y); }

// V[s][i] = vm * (cos va, sin va)
__global__ void nr_rect_kernel(i64 total, const double *__restrict__ vm, const double *__restrict__ va, double2 *__restrict__ V)
{
    const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total) return;
    double s, c;
    sincos(va[t], &s, &c);
    const double m = vm[t];
    V[t] = make_double2(m * c, m * s);
}

// Ybus value of entry e for a case with branch `ob` out of service (ob < 0: base case)
__device__ __forceinline__ double2 y_entry(const double2 *__restrict__ y_val, int e, int ob, int n_branch,
                                           const i32 *__restrict__ br_slot, const double2 *__restrict__ br_val)
{
    double2 y = __ldg(y_val + e);
    if (ob >= 0) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (__ldg(br_slot + (size_t)q * n_branch + ob) == e) {
                const double2 c = __ldg(br_val + (size_t)q * n_branch + ob);
                y.x -= c.x; y.y -= c.y;
            }
    }
    return y;
}

// one thread per (case, bus): I = Ybus V, mismatch, right-hand side
__global__ void nr_current_kernel(int n_bus, int npvpq, int n, int n_branch, const i32 *__restrict__ rowptr, const i32 *__restrict__ col,
                                  const double2 *__restrict__ y_val, const i32 *__restrict__ br_slot, const double2 *__restrict__ br_val,
                                  const i32 *__restrict__ out_branch, const i32 *__restrict__ pos_th, const i32 *__restrict__ pos_v,
                                  const double2 *__restrict__ V, const double *__restrict__ sspec, double2 *__restrict__ I,
                                  double *__restrict__ b, unsigned long long *__restrict__ fnorm, i64 case0)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const i64 s = case0 + blockIdx.y;
    double worst = 0.0;
    if (i < n_bus) {
        const int ob = out_branch ? out_branch[s] : -1;
        const double2 *Vs = V + s * n_bus;
        double2 acc = make_double2(0.0, 0.0);
        for (int p = rowptr[i]; p < rowptr[i + 1]; ++p) {
            const double2 yv = cmul(y_entry(y_val, p, ob, n_branch, br_slot, br_val), Vs[col[p]]);
            acc.x += yv.x; acc.y += yv.y;
        }
        I[s * n_bus + i] = acc;
        const double2 sc = cmul(Vs[i], cconj(acc));                    // S_calc = V conj(I)
        const int pt = pos_th[i], pv = pos_v[i];
        if (pt >= 0) { const double f = sc.x - sspec[s * n + pt]; b[s * n + pt] = -f; worst = fmax(worst, fabs(f)); }
        if (pv >= 0) { const double f = sc.y - sspec[s * n + npvpq + pv]; b[s * n + npvpq + pv] = -f; worst = fmax(worst, fabs(f)); }
    }
    // max over the block, then one atomic per block (non-negative doubles order like their bit patterns)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) worst = fmax(worst, __shfl_xor_sync(0xffffffffu, worst, o));
    __shared__ double wmax[8];
    if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = worst;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) worst = fmax(worst, wmax[w]);
        atomicMax(fnorm + s, (unsigned long long)__double_as_longlong(worst));
    }
}

// one thread per (case, Jacobian entry): consecutive threads write consecutive entries of Ax[s][:]
__global__ void nr_jacobian_kernel(int n_bus, int jnnz, int n_branch, const int4 *__restrict__ jmap, const double2 *__restrict__ y_val,
                                   const i32 *__restrict__ br_slot, const double2 *__restrict__ br_val, const i32 *__restrict__ out_branch,
                                   const double2 *__restrict__ V, const double *__restrict__ vm, const double2 *__restrict__ I,
                                   double *__restrict__ Ax, i64 case0)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const i64 s = case0 + blockIdx.y;
    if (p >= jnnz) return;
    const int4 m = __ldg(jmap + p);                                   // (i, k, e, block)
    const int ob = out_branch ? out_branch[s] : -1;
    const double2 y = y_entry(y_val, m.z, ob, n_branch, br_slot, br_val);
    const double2 Vi = V[s * n_bus + m.x], Vk = V[s * n_bus + m.y];
    const double2 Ii = (m.x == m.y) ? I[s * n_bus + m.x] : make_double2(0.0, 0.0);
    double out;
    if ((m.w & 1) == 0) {                                             // dS/dVa = j V_i conj(I_i - Y V_k)
        const double2 yv = cmul(y, Vk);
        const double2 d = cmul(Vi, cconj(make_double2(Ii.x - yv.x, Ii.y - yv.y)));
        out = (m.w == 0) ? -d.y : d.x;                                // j d = (-d.y, d.x)
    } else {                                                          // dS/dVm = V_i conj(Y V_k/|V_k|) + conj(I_i) V_k/|V_k|
        const double r = 1.0 / vm[s * n_bus + m.y];
        const double2 vn = make_double2(Vk.x * r, Vk.y * r);
        const double2 a = cmul(Vi, cconj(cmul(y, vn)));
        const double2 c = cmul(cconj(Ii), vn);
        out = (m.w == 1) ? a.x + c.x : a.y + c.y;
    }
    Ax[s * jnnz + p] = out;
}

// one thread per (case, bus): va[pvpq] += dx, vm[pq] += dx, V for the next iteration
__global__ void nr_update_kernel(int n_bus, int npvpq, int n, const i32 *__restrict__ pos_th, const i32 *__restrict__ pos_v,
                                 const double *__restrict__ dx, double *__restrict__ vm, double *__restrict__ va, double2 *__restrict__ V,
                                 i64 case0)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const i64 s = case0 + blockIdx.y;
    if (i >= n_bus) return;
    const i64 t = s * n_bus + i;
    const int pt = pos_th[i], pv = pos_v[i];
    double a = va[t], m = vm[t];
    if (pt >= 0) { a += dx[s * n + pt]; va[t] = a; }
    if (pv >= 0) { m += dx[s * n + npvpq + pv]; vm[t] = m; }
    double sn, cs;
    sincos(a, &sn, &cs);
    V[t] = make_double2(m * cs, m * sn);
}

// dst[c][i] = src[i] for c < cnt (one shared start vector replicated for every case)
__global__ void nr_broadcast_kernel(i64 total, int n_bus, const double *__restrict__ src, double *__restrict__ dst)
{
    const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < total) dst[t] = src[t % n_bus];
}

struct Carve {
    double2 *V, *I;
    double *Ax, *b, *dx;
    void *lu;
};
size_t a256(size_t v) { return (v + 255) & ~(size_t)255; }
size_t nr_bytes(const csp3_nr_plan *P, i64 batch, Carve *c, char *base)
{
    size_t off = 0;
    auto take = [&](size_t bytes) { char *p = base ? base + off : nullptr; off += a256(bytes); return p; };
    char *v = take((size_t)batch * P->n_bus * 16), *i = take((size_t)batch * P->n_bus * 16);
    char *ax = take((size_t)batch * P->jnnz * 8), *b = take((size_t)batch * P->n * 8), *dx = take((size_t)batch * P->n * 8);
    char *lu = take((size_t)csp3_lu_workspace_bytes(P->sym, batch));
    if (c) { c->V = (double2 *)v; c->I = (double2 *)i; c->Ax = (double *)ax; c->b = (double *)b; c->dx = (double *)dx; c->lu = lu; }
    return off + 256;
}

}  // namespace

extern "C" {

int csp3_nr_create(int64_t n_bus, int64_t nnz_y, const int32_t *y_rowptr, const int32_t *y_col, const double *y_val,
                   int64_t npvpq, const int32_t *pvpq, int64_t npq, const int32_t *pq, int64_t jnnz, const int32_t *j_ent,
                   const int32_t *j_block, int64_t n_branch, const int32_t *br_slot, const double *br_val,
                   csp3_lu_symbolic *sym, csp3_nr_plan **plan)
{
    if (!plan || !sym || n_bus <= 0 || nnz_y <= 0 || !y_rowptr || !y_col || !y_val || !pvpq || !pq || !j_ent || !j_block ||
        npvpq < 0 || npq < 0 || jnnz <= 0 || (n_branch > 0 && (!br_slot || !br_val))) {
        set_error("nr_create: bad arguments");
        return CSP3_ERR_ARG;
    }
    int64_t sz[16];
    if (int rc = csp3_lu_sizes(sym, sz)) return rc;
    if (sz[0] != npvpq + npq || sz[1] != jnnz) { set_error("nr_create: the symbolic object does not belong to this Jacobian pattern"); return CSP3_ERR_ARG; }
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt == 0) { cudaGetLastError(); set_error("no CUDA device available: libcsparse3_b200 has no CPU fallback"); return CSP3_ERR_CUDA; }
    std::unique_ptr<csp3_nr_plan> P(new csp3_nr_plan());
    P->n_bus = n_bus; P->nnz_y = nnz_y; P->npvpq = npvpq; P->npq = npq; P->n = npvpq + npq; P->jnnz = jnnz; P->n_branch = n_branch; P->sym = sym;
    CSP3_CUDA(cudaGetDevice(&P->device));
    // host-side maps
    std::vector<i32> pos_th((size_t)n_bus, -1), pos_v((size_t)n_bus, -1), yrow((size_t)nnz_y);
    for (i64 r = 0; r < npvpq; ++r) { if (pvpq[r] < 0 || pvpq[r] >= n_bus) { set_error("nr_create: pvpq out of range"); return CSP3_ERR_ARG; } pos_th[(size_t)pvpq[r]] = (i32)r; }
    for (i64 r = 0; r < npq; ++r) { if (pq[r] < 0 || pq[r] >= n_bus) { set_error("nr_create: pq out of range"); return CSP3_ERR_ARG; } pos_v[(size_t)pq[r]] = (i32)r; }
    if (y_rowptr[0] != 0 || y_rowptr[n_bus] != nnz_y) { set_error("nr_create: bad Ybus row pointers"); return CSP3_ERR_ARG; }
    for (i64 i = 0; i < n_bus; ++i)
        for (i32 p = y_rowptr[i]; p < y_rowptr[i + 1]; ++p) {
            if (y_col[p] < 0 || y_col[p] >= n_bus) { set_error("nr_create: Ybus column out of range"); return CSP3_ERR_ARG; }
            yrow[(size_t)p] = (i32)i;
        }
    std::vector<int4> jmap((size_t)jnnz);
    for (i64 p = 0; p < jnnz; ++p) {
        const i32 e = j_ent[p], blk = j_block[p];
        if (e < 0 || e >= nnz_y || blk < 0 || blk > 3) { set_error("nr_create: bad Jacobian entry map"); return CSP3_ERR_ARG; }
        jmap[(size_t)p] = make_int4(yrow[(size_t)e], y_col[e], e, blk);
    }
    if (n_branch > 0)
        for (i64 t = 0; t < 4 * n_branch; ++t)
            if (br_slot[t] < 0 || br_slot[t] >= nnz_y) { set_error("nr_create: branch slot out of range"); return CSP3_ERR_ARG; }
    struct Piece { const void *src; size_t bytes, off; };
    std::vector<Piece> pcs;
    size_t total = 0;
    auto add = [&](const void *src, size_t bytes) { total = a256(total); pcs.push_back({src, bytes, total}); total += bytes; return pcs.size() - 1; };
    const size_t i_rp = add(y_rowptr, (size_t)(n_bus + 1) * 4), i_col = add(y_col, (size_t)nnz_y * 4), i_val = add(y_val, (size_t)nnz_y * 16);
    const size_t i_pt = add(pos_th.data(), (size_t)n_bus * 4), i_pv = add(pos_v.data(), (size_t)n_bus * 4), i_jm = add(jmap.data(), (size_t)jnnz * 16);
    const size_t i_bs = add(br_slot, (size_t)n_branch * 16), i_bv = add(br_val, (size_t)n_branch * 64);
    char *arena = nullptr;
    if (cudaMalloc((void **)&arena, a256(total) + 256) != cudaSuccess) { cudaGetLastError(); set_error("nr_create: device allocation failed"); return CSP3_ERR_ALLOC; }
    for (auto &pc : pcs)
        if (pc.bytes && cudaMemcpy(arena + pc.off, pc.src, pc.bytes, cudaMemcpyHostToDevice) != cudaSuccess) {
            cudaFree(arena); set_error("nr_create: copy failed (%s)", cudaGetErrorString(cudaGetLastError())); return CSP3_ERR_CUDA;
        }
    P->arena = arena;
    P->y_rowptr = (i32 *)(arena + pcs[i_rp].off); P->y_col = (i32 *)(arena + pcs[i_col].off); P->y_val = (double2 *)(arena + pcs[i_val].off);
    P->pos_th = (i32 *)(arena + pcs[i_pt].off); P->pos_v = (i32 *)(arena + pcs[i_pv].off); P->jmap = (int4 *)(arena + pcs[i_jm].off);
    P->br_slot = (i32 *)(arena + pcs[i_bs].off); P->br_val = (double2 *)(arena + pcs[i_bv].off);
    if (int rc = csp3_lu_upload(sym, nullptr)) { cudaFree(arena); return rc; }
    *plan = P.release();
    return 0;
}

static void nr_free_stage(csp3_nr_plan *P)
{
    auto &G = P->stage;
    for (int s = 0; s < 2; ++s) {
        cudaFree(G.sspec[s]); cudaFree(G.vm[s]); cudaFree(G.va[s]); cudaFree(G.fnorm[s]); cudaFree(G.status[s]); cudaFree(G.outb[s]); cudaFree(G.work[s]);
        if (G.st[s]) cudaStreamDestroy(G.st[s]);
    }
    G = csp3_nr_plan::Stage();
}

int csp3_nr_destroy(csp3_nr_plan *plan)
{
    if (!plan) return 0;
    int cur = 0;
    const bool have = cudaGetDevice(&cur) == cudaSuccess;
    if (have && plan->device >= 0) cudaSetDevice(plan->device);
    if (plan->stage.ready) nr_free_stage(plan);
    if (plan->arena) cudaFree(plan->arena);
    if (have) cudaSetDevice(cur);
    cudaGetLastError();
    delete plan;
    return 0;
}

int64_t csp3_nr_workspace_bytes(const csp3_nr_plan *plan, int64_t batch)
{
    if (!plan || batch < 0) return -1;
    return (int64_t)nr_bytes(plan, batch, nullptr, nullptr);
}

// Jacobian values and right-hand side at the state (vm, va); V and I are left in `work` for nr_update
static int nr_evaluate(const csp3_nr_plan *P, i64 batch, const double *sspec, const i32 *out_branch, const double *vm, const Carve &c,
                       double *fnorm, bool rect, const double *va, cudaStream_t st)
{
    const int nb = (int)P->n_bus;
    if (rect) {
        const i64 total = batch * P->n_bus;
        nr_rect_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(total, vm, va, c.V);
    }
    CSP3_CUDA(cudaMemsetAsync(fnorm, 0, (size_t)batch * 8, st));
    for (i64 c0 = 0; c0 < batch; c0 += 65535) {
        const unsigned ny = (unsigned)std::min<i64>(65535, batch - c0);
        nr_current_kernel<<<dim3((unsigned)((nb + 255) / 256), ny), 256, 0, st>>>(
            nb, (int)P->npvpq, (int)P->n, (int)P->n_branch, P->y_rowptr, P->y_col, P->y_val, P->br_slot, P->br_val, out_branch, P->pos_th,
            P->pos_v, c.V, sspec, c.I, c.b, reinterpret_cast<unsigned long long *>(fnorm), c0);
        nr_jacobian_kernel<<<dim3((unsigned)((P->jnnz + 255) / 256), ny), 256, 0, st>>>(
            nb, (int)P->jnnz, (int)P->n_branch, P->jmap, P->y_val, P->br_slot, P->br_val, out_branch, c.V, vm, c.I, c.Ax, c0);
    }
    CSP3_CUDA(cudaGetLastError());
    return 0;
}

int csp3_nr_jacobian(const csp3_nr_plan *plan, int64_t batch, const double *vm, const double *va, const int32_t *out_branch,
                     const double *sspec, double *Ax, double *b, double *fnorm, void *work, void *stream)
{
    if (!plan || batch < 0 || !vm || !va || !sspec || !Ax || !b || !fnorm || !work) { set_error("nr_jacobian: bad arguments"); return CSP3_ERR_ARG; }
    if (batch == 0) return 0;
    Carve c;
    nr_bytes(plan, batch, &c, (char *)(((uintptr_t)work + 255) & ~(uintptr_t)255));
    c.Ax = Ax; c.b = b;
    return nr_evaluate(plan, batch, sspec, out_branch, vm, c, fnorm, true, va, (cudaStream_t)stream);
}

int csp3_nr_solve(const csp3_nr_plan *plan, int64_t batch, int64_t iters, const double *sspec, const int32_t *out_branch, double *vm,
                  double *va, double *fnorm, int32_t *status, void *work, void *stream)
{
    if (!plan || batch < 0 || iters < 0 || !sspec || !vm || !va || !fnorm || !status || !work) { set_error("nr_solve: bad arguments"); return CSP3_ERR_ARG; }
    if (batch == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    Carve c;
    nr_bytes(plan, batch, &c, (char *)(((uintptr_t)work + 255) & ~(uintptr_t)255));
    const int nb = (int)plan->n_bus;
    for (i64 it = 0; it < iters; ++it) {
        if (int rc = nr_evaluate(plan, batch, sspec, out_branch, vm, c, fnorm, it == 0, va, st)) return rc;
        if (int rc = csp3_lu_refactor_ws(plan->sym, batch, c.Ax, c.lu, status, st)) return rc;
        if (int rc = csp3_lu_solve_ws(plan->sym, batch, c.lu, c.b, c.dx, st)) return rc;
        for (i64 c0 = 0; c0 < batch; c0 += 65535) {
            const unsigned ny = (unsigned)std::min<i64>(65535, batch - c0);
            nr_update_kernel<<<dim3((unsigned)((nb + 255) / 256), ny), 256, 0, st>>>(nb, (int)plan->npvpq, (int)plan->n, plan->pos_th, plan->pos_v,
                                                                                      c.dx, vm, va, c.V, c0);
        }
        CSP3_CUDA(cudaGetLastError());
    }
    // mismatch at the returned state (one more evaluation of I and f, no factorisation)
    CSP3_CUDA(cudaMemsetAsync(fnorm, 0, (size_t)batch * 8, st));
    if (iters == 0) {
        const i64 total = batch * plan->n_bus;
        nr_rect_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(total, vm, va, c.V);
    }
    for (i64 c0 = 0; c0 < batch; c0 += 65535) {
        const unsigned ny = (unsigned)std::min<i64>(65535, batch - c0);
        nr_current_kernel<<<dim3((unsigned)((nb + 255) / 256), ny), 256, 0, st>>>(
            nb, (int)plan->npvpq, (int)plan->n, (int)plan->n_branch, plan->y_rowptr, plan->y_col, plan->y_val, plan->br_slot, plan->br_val,
            out_branch, plan->pos_th, plan->pos_v, c.V, sspec, c.I, c.b, reinterpret_cast<unsigned long long *>(fnorm), c0);
    }
    CSP3_CUDA(cudaGetLastError());
    return 0;
}

int csp3_nr_solve_host(csp3_nr_plan *plan, int64_t batch, int64_t iters, const double *sspec, const int32_t *out_branch,
                       const double *vm0, const double *va0, int64_t start_stride, double *vm, double *va, double *fnorm, int32_t *status)
{
    if (!plan || batch < 0 || iters < 0 || !sspec || !vm || !va || !fnorm || !status || !vm0 || !va0 || (start_stride != 0 && start_stride != plan->n_bus)) {
        set_error("nr_solve_host: bad arguments (start_stride is 0 for one shared start vector or n_bus)");
        return CSP3_ERR_ARG;
    }
    if (batch == 0) return 0;
    int dev = 0;
    CSP3_CUDA(cudaGetDevice(&dev));
    if (dev != plan->device) { set_error("nr_solve_host: the plan lives on device %d", plan->device); return CSP3_ERR_ARG; }
    auto &G = plan->stage;
    const i64 nb = plan->n_bus, n = plan->n;
    if (!G.ready) {
        // two chunks in flight on two streams: up to 4 GB of Jacobian values each, at most 5,120 cases (the LU kernels are
        // latency-bound below ~5,000 cases per launch, and two co-resident launches fill the GPU)
        i64 chunk = (4096ll << 20) / std::max<i64>(plan->jnnz * 8, 1);
        chunk = std::max<i64>(256, std::min<i64>(chunk, 5120));
        chunk = (chunk + 31) & ~31ll;
        G.chunk = chunk;
        for (int s = 0; s < 2; ++s) {
            bool ok = cudaStreamCreateWithFlags(&G.st[s], cudaStreamNonBlocking) == cudaSuccess &&
                      cudaMalloc((void **)&G.sspec[s], (size_t)chunk * n * 8 + 16) == cudaSuccess &&
                      cudaMalloc((void **)&G.vm[s], (size_t)chunk * nb * 8 + 16) == cudaSuccess &&
                      cudaMalloc((void **)&G.va[s], (size_t)chunk * nb * 8 + 16) == cudaSuccess &&
                      cudaMalloc((void **)&G.fnorm[s], (size_t)chunk * 8 + 16) == cudaSuccess &&
                      cudaMalloc((void **)&G.status[s], (size_t)chunk * 4 + 16) == cudaSuccess &&
                      cudaMalloc((void **)&G.outb[s], (size_t)chunk * 4 + 16) == cudaSuccess &&
                      cudaMalloc(&G.work[s], (size_t)csp3_nr_workspace_bytes(plan, chunk)) == cudaSuccess;
            if (!ok) { cudaGetLastError(); nr_free_stage(plan); set_error("nr_solve_host: staging allocation failed"); return CSP3_ERR_ALLOC; }
        }
        G.ready = true;
    }
    int rc = 0, slot = 0;
    for (i64 s0 = 0; s0 < batch && rc == 0; s0 += G.chunk, slot ^= 1) {
        const i64 cnt = std::min<i64>(G.chunk, batch - s0);
        cudaStream_t st = G.st[slot];
        auto cp = [&](void *d, const void *h, size_t bytes, cudaMemcpyKind k) { if (rc == 0 && cudaMemcpyAsync(d, h, bytes, k, st) != cudaSuccess) { set_error("nr_solve_host: copy failed (%s)", cudaGetErrorString(cudaGetLastError())); rc = CSP3_ERR_CUDA; } };
        cp(G.sspec[slot], sspec + s0 * n, (size_t)cnt * n * 8, cudaMemcpyHostToDevice);
        if (start_stride) {
            cp(G.vm[slot], vm0 + s0 * nb, (size_t)cnt * nb * 8, cudaMemcpyHostToDevice);
            cp(G.va[slot], va0 + s0 * nb, (size_t)cnt * nb * 8, cudaMemcpyHostToDevice);
        } else if (rc == 0) {                                                   // one shared start vector: replicate on the device
            double *tmp = reinterpret_cast<double *>(G.work[slot]);                 // (the workspace is not in use yet)
            cp(tmp, vm0, (size_t)nb * 8, cudaMemcpyHostToDevice);
            cp(tmp + nb, va0, (size_t)nb * 8, cudaMemcpyHostToDevice);
            const i64 total = cnt * nb;
            nr_broadcast_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(total, (int)nb, tmp, G.vm[slot]);
            nr_broadcast_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(total, (int)nb, tmp + nb, G.va[slot]);
        }
        if (out_branch) cp(G.outb[slot], out_branch + s0, (size_t)cnt * 4, cudaMemcpyHostToDevice);
        if (rc == 0) rc = csp3_nr_solve(plan, cnt, iters, G.sspec[slot], out_branch ? G.outb[slot] : nullptr, G.vm[slot], G.va[slot], G.fnorm[slot], G.status[slot], G.work[slot], st);
        cp(vm + s0 * nb, G.vm[slot], (size_t)cnt * nb * 8, cudaMemcpyDeviceToHost);
        cp(va + s0 * nb, G.va[slot], (size_t)cnt * nb * 8, cudaMemcpyDeviceToHost);
        cp(fnorm + s0, G.fnorm[slot], (size_t)cnt * 8, cudaMemcpyDeviceToHost);
        cp(status + s0, G.status[slot], (size_t)cnt * 4, cudaMemcpyDeviceToHost);
    }
    for (int s = 0; s < 2; ++s)                                                 // also on errors: nothing may still target the caller's buffers
        if (cudaStreamSynchronize(G.st[s]) != cudaSuccess && rc == 0) { set_error("nr_solve_host: %s", cudaGetErrorString(cudaGetLastError())); rc = CSP3_ERR_CUDA; }
    return rc;
}

}  // extern "C"
