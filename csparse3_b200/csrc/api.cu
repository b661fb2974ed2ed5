// api.cu -- extern "C" boundary of libcsparse3_b200.so (declared in include/csparse3_b200.h).
#include <algorithm>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <vector>

#include "../../include/csparse3_b200.h"
#include "common.cuh"
#include "csc_kernels.cuh"
#include "panel_program.hpp"
#include "rowlane_program.hpp"
#include "rowsweep_program.hpp"

namespace csp3 {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char *get_error() { return g_err; }

Tuning &tuning()
{
    static Tuning t = [] {
        Tuning v;
        auto env = [](const char *k) { const char *s = getenv(k); return s ? atoi(s) : 0; };
        v.rf_S = env("CSP3_RF_S"); v.sv_S = env("CSP3_SV_S"); v.ws_S = env("CSP3_WS_S");
        v.rf_win = env("CSP3_RF_WIN"); v.sv_stage = env("CSP3_SV_STAGE");
        if (getenv("CSP3_WIDE")) v.wide = env("CSP3_WIDE");
        if (getenv("CSP3_WIDE_SOLVE")) v.wide_solve = env("CSP3_WIDE_SOLVE");
        v.wide_S = env("CSP3_WIDE_S"); v.wide_R = env("CSP3_WIDE_LANE"); v.wide_ring = env("CSP3_WIDE_R"); v.wide_stage = env("CSP3_WIDE_F");
        v.wide_budget = env("CSP3_WIDE_BUDGET");
        if (getenv("CSP3_PANEL")) v.panel = env("CSP3_PANEL");
        if (getenv("CSP3_TMEM")) v.tmem = env("CSP3_TMEM");
        if (getenv("CSP3_ROWLANE")) v.rowlane = env("CSP3_ROWLANE");
        if (getenv("CSP3_ROWSWEEP")) v.rowsweep = env("CSP3_ROWSWEEP");
        v.rl_warps = env("CSP3_RL_W"); v.rl_nq = env("CSP3_RL_NQ");
        v.panel_fma = env("CSP3_PANEL_FMA");
        v.panel_budget = env("CSP3_PANEL_BUDGET");
        return v;
    }();
    return t;
}

namespace {

int require_device()
{
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess || cnt == 0) {
        cudaGetLastError();
        set_error("no CUDA device available: libcsparse3_b200 has no CPU fallback");
        return CSP3_ERR_CUDA;
    }
    return 0;
}

// RAII device buffer for the *_host entry points (synchronous semantics, default stream)
struct Dev {
    void *p = nullptr;
    ~Dev() { if (p) cudaFreeAsync(p, nullptr); }
    int alloc(size_t bytes)
    {
        keep_device_pool();
        return cudaMallocAsync(&p, bytes ? bytes : 16, nullptr) == cudaSuccess ? 0 : -1;
    }
    int put(const void *src, size_t bytes)
    {
        if (alloc(bytes)) return -1;
        return (bytes == 0 || cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice) == cudaSuccess) ? 0 : -1;
    }
    int get(void *dst, size_t bytes) const
    {
        return (bytes == 0 || cudaMemcpy(dst, p, bytes, cudaMemcpyDeviceToHost) == cudaSuccess) ? 0 : -1;
    }
    template <class T> T *as() const { return (T *)p; }
};

#define CSP3_TRY(expr) do { if ((expr) != 0) { if (!*csp3::get_error()) csp3::set_error("device allocation or copy failed (%s)", cudaGetErrorString(cudaGetLastError())); return CSP3_ERR_ALLOC; } } while (0)

constexpr int kMaxDevices = 16;

}  // namespace
}  // namespace csp3

using namespace csp3;

struct csp3_spmv_plan {
    SpmvPlanData d;
};

struct csp3_lu_symbolic {
    i64 n = 0, nnzA = 0;
    std::vector<i32> Ap, Ai, q;
    Factor F;
    Schedule S;
    WideProgram W;                     // wide refactor program (ok == false: pattern does not fit, v3 kernels are used)
    WideSweep WF, WB;                  // wide forward / backward sweep programs
    PanelProgram PP;                   // panel refactor program (lu_panel.cu); ok == false: the wide / v3 kernels are used
    RowlaneProgram RL[kRlVariants];    // row-lane refactor programs (lu_rowlane.cu), one per geometry, compiled on first use
    bool RLtried[kRlVariants] = {};
    RowSweepProgram RSf, RSb;          // row-sweep programs (forward / backward), compiled on first use
    bool RStried = false;
    std::mutex rl_mu;                  // guards RL / RLtried and the per-device variants (taken inside calls that may hold `mu`)
    std::vector<i32> qinv;             // x[c] = x_pivot_order[qinv[c]]
    DevSchedule dev[kMaxDevices];
    // staging for csp3_lu_refactor_solve_host (per device, lazily created)
    struct Stage {
        bool ready = false;
        i64 chunk = 0;
        cudaStream_t st[3] = {nullptr, nullptr, nullptr};
        double *Ax[3] = {}, *b[3] = {}, *x[3] = {}, *Lx[3] = {}, *Ux[3] = {};
        i32 *status[3] = {};
    } stage[kMaxDevices];
    std::mutex mu;
};

// Wide program: bundles of 8 systems by default.  The shared-memory budget of a bundle is chosen so that a
// 10,000-system batch (1,250 one-warp CTAs = 8.4 per SM) is resident at once: 9 CTAs per SM; larger budgets are
// tried for patterns with long columns (fewer resident bundles, more waves).
static void compile_wide(csp3_lu_symbolic &Sy)
{
    const Tuning &t = tuning();
    if (Sy.n > 0) {
        Sy.qinv.assign((size_t)Sy.n, 0);
        for (i64 i = 0; i < Sy.n; ++i) Sy.qinv[(size_t)(Sy.q.empty() ? i : Sy.q[(size_t)i])] = (i32)i;
    }
    if (Sy.n > 0) {          // compiled always (cheap), executed only with CSP3_PANEL=1: the wide kernel is faster today
        const char *why = "";
        // 9 one-warp CTAs per SM keep every bundle of a 10,000-system batch resident (1,250 bundles on 148 SMs)
        const size_t panel_budget = t.panel_budget > 0 ? (size_t)t.panel_budget : (((size_t)227 * 1024 / 9 - 1024) & ~(size_t)63);
        bool ok = false;
        for (int i = 0; i < 4 && !ok; ++i) {                       // patterns with long columns: fewer resident bundles
            ok = compile_panel_refactor(Sy.n, Sy.Ap.data(), Sy.Ai.data(), Sy.q, Sy.F, 8, std::min<size_t>(panel_budget << i, (size_t)200 * 1024), Sy.PP, &why);
            if (!ok && getenv("CSP3_DEBUG")) fprintf(stderr, "csp3: panel refactor unavailable at budget %zu: %s\n", panel_budget << i, why);
            if (t.panel_budget > 0) break;
        }
        if (!ok) Sy.PP = PanelProgram();
    }
    if (t.wide == 0 || Sy.n == 0) return;
    const i32 width = (t.wide_S == 4 || t.wide_S == 16 || t.wide_S == 32) ? t.wide_S : 8;
    // The wide kernel is the FULL-GPU kernel (smaller batches go to the row-lane kernel, lu_kernels.cu::rowlane_variant):
    // its shared-memory budget per bundle is the one that keeps a full-GPU batch resident in one wave; larger budgets
    // are tried below for patterns whose columns do not fit.
    constexpr int kWideFullBatch = 10000;         // systems that should be resident in one wave: 9 bundles of 8 per SM
    const int ctas_per_sm[4] = {(kWideFullBatch / width + kNumSMs - 1) / kNumSMs, 0, 0, 0};
    for (int i = 0; i < 3; ++i) {
        const int per_sm = std::max(1, ctas_per_sm[0] >> i);
        const size_t budget = t.wide_budget > 0 ? (size_t)t.wide_budget
                                                : std::min<size_t>((size_t)200 * 1024, ((size_t)228 * 1024 / per_sm - 1024) & ~(size_t)255);
        const char *why = "";
        const i32 per_lane = (t.wide_R == 4 && width >= 16) ? 4 : 2;
        if (compile_wide_refactor(Sy.S, Sy.F, width, 32 * per_lane / width, budget, t.wide_ring, t.wide_stage, Sy.W, &why)) {
            if (!compile_wide_sweep(Sy.F, true, width, Sy.W.groups, budget, Sy.WF, &why) ||
                !compile_wide_sweep(Sy.F, false, width, Sy.W.groups, budget, Sy.WB, &why) ||
                std::max(Sy.WF.smem_bytes, Sy.WB.smem_bytes) > (size_t)200 * 1024) {
                if (getenv("CSP3_DEBUG")) fprintf(stderr, "csp3: wide sweeps unavailable: %s\n", why);
                Sy.WF = WideSweep(); Sy.WB = WideSweep();
            }
            Sy.qinv.assign((size_t)Sy.n, 0);
            for (i64 i = 0; i < Sy.n; ++i) Sy.qinv[(size_t)(Sy.q.empty() ? i : Sy.q[(size_t)i])] = (i32)i;
            return;
        }
        if (getenv("CSP3_DEBUG")) fprintf(stderr, "csp3: wide refactor unavailable at budget %zu: %s\n", budget, why);
        if (t.wide_budget > 0) break;
    }
    Sy.W = WideProgram();
}

// row-lane program of one geometry: compiled on first use (host), cached in the symbolic object
static RowlaneProgram *rowlane_program(csp3_lu_symbolic &Sy, int variant)
{
    if (variant < 0 || variant >= kRlVariants || Sy.n <= 0) return nullptr;
    if (!Sy.RLtried[variant]) {
        Sy.RLtried[variant] = true;
        const Tuning &t = tuning();
        i32 rw = kRlGeometry[variant][0], rq = kRlGeometry[variant][1];
        if (variant == kRlVariants - 1) {
            rw = (t.rl_warps == 2 || t.rl_warps == 4 || t.rl_warps == 8) ? t.rl_warps : 1;
            rq = (t.rl_nq >= 1 && t.rl_nq <= 3) ? t.rl_nq : 3;
        }
        const char *why = "";
        RowlaneProgram &P = Sy.RL[variant];
        if (!compile_rowlane_refactor(Sy.n, Sy.Ap.data(), Sy.q, Sy.F, Sy.S, rw, rq, P, &why)) {
            if (getenv("CSP3_DEBUG")) fprintf(stderr, "csp3: row-lane refactor unavailable: %s\n", why);
            P = RowlaneProgram();
        } else if (getenv("CSP3_DEBUG")) {
            fprintf(stderr, "csp3: row-lane program: %d warps x %d quads per stage, %d quads (longest stream %d; %lld update quads with %lld records, %lld late, %lld records read another warp's column, %lld empty quads), %lld ops, %d slots, %lld conflict pairs\n",
                    P.warps, P.stage_quads, P.quads, *std::max_element(P.stream_quads, P.stream_quads + kRlMaxWarps), (long long)P.update_quads,
                    (long long)P.update_records, (long long)P.late_quads, (long long)P.cross_records, (long long)P.pad_quads, (long long)P.ops, P.nslots, (long long)P.conflict_pairs);
        }
    }
    return Sy.RL[variant].ok ? &Sy.RL[variant] : nullptr;
}

namespace csp3 {
int ensure_rowlane_variant(const DevSchedule &D, int variant)
{
    if (variant < 0 || variant >= kRlVariants || !D.owner) return -1;
    DevSchedule::RlVariant &R = D.rl[variant];
    if (R.tried.load(std::memory_order_acquire)) return R.ok ? 0 : -1;
    csp3_lu_symbolic &Sy = *static_cast<csp3_lu_symbolic *>(D.owner);
    std::lock_guard<std::mutex> lock(Sy.rl_mu);
    if (R.tried.load(std::memory_order_relaxed)) return R.ok ? 0 : -1;
    const RowlaneProgram *P = rowlane_program(Sy, variant);
    if (P && P->smem_bytes <= (size_t)200 * 1024) {
        uint8_t *dev = nullptr;
        const size_t bytes = P->words.size() * 4;
        if (cudaMalloc((void **)&dev, bytes) == cudaSuccess && cudaMemcpy(dev, P->words.data(), bytes, cudaMemcpyHostToDevice) == cudaSuccess) {
            R.prog = dev; R.quads = P->quads; R.nslots = P->nslots; R.warps = P->warps; R.stage_quads = P->stage_quads; R.smem = P->smem_bytes;
            for (int w = 0; w < kRlMaxWarps; ++w) R.stream_off[w] = (i32)P->stream_off[w];
            R.ok = true;
        } else {
            if (dev) cudaFree(dev);
            cudaGetLastError();
        }
    }
    R.tried.store(true, std::memory_order_release);
    return R.ok ? 0 : -1;
}
int ensure_rowsweep(const DevSchedule &D)
{
    if (!D.owner) return -1;
    DevSchedule::RsPrograms &R = D.rs;
    if (R.tried.load(std::memory_order_acquire)) return R.ok ? 0 : -1;
    csp3_lu_symbolic &Sy = *static_cast<csp3_lu_symbolic *>(D.owner);
    std::lock_guard<std::mutex> lock(Sy.rl_mu);
    if (R.tried.load(std::memory_order_relaxed)) return R.ok ? 0 : -1;
    if (!Sy.RStried) {
        Sy.RStried = true;
        const char *why = "";
        if (!compile_row_sweep(Sy.F, Sy.S, true, Sy.RSf, &why) || !compile_row_sweep(Sy.F, Sy.S, false, Sy.RSb, &why)) {
            if (getenv("CSP3_DEBUG")) fprintf(stderr, "csp3: row sweeps unavailable: %s\n", why);
            Sy.RSf = RowSweepProgram(); Sy.RSb = RowSweepProgram();
        } else if (getenv("CSP3_DEBUG")) {
            fprintf(stderr, "csp3: row sweeps: forward %d levels, %d panels, %lld chunks for %lld terms; backward %d levels, %d panels, %lld chunks for %lld terms\n",
                    Sy.RSf.levels, Sy.RSf.panels, (long long)Sy.RSf.chunks, (long long)Sy.RSf.terms, Sy.RSb.levels, Sy.RSb.panels, (long long)Sy.RSb.chunks, (long long)Sy.RSb.terms);
        }
    }
    if (Sy.RSf.ok && Sy.RSb.ok) {
        uint32_t *df = nullptr, *db = nullptr;
        const size_t bf = Sy.RSf.words.size() * 4, bb = Sy.RSb.words.size() * 4;
        if (cudaMalloc((void **)&df, bf) == cudaSuccess && cudaMalloc((void **)&db, bb) == cudaSuccess &&
            cudaMemcpy(df, Sy.RSf.words.data(), bf, cudaMemcpyHostToDevice) == cudaSuccess &&
            cudaMemcpy(db, Sy.RSb.words.data(), bb, cudaMemcpyHostToDevice) == cudaSuccess) {
            R.prog_f = df; R.prog_b = db; R.levels_f = Sy.RSf.levels; R.levels_b = Sy.RSb.levels;
            for (int w = 0; w < kRsWarps; ++w) { R.stream_off_f[w] = (i32)Sy.RSf.stream_off[w]; R.stream_off_b[w] = (i32)Sy.RSb.stream_off[w]; }
            R.ok = true;
        } else {
            if (df) cudaFree(df);
            if (db) cudaFree(db);
            cudaGetLastError();
        }
    }
    R.tried.store(true, std::memory_order_release);
    return R.ok ? 0 : -1;
}
}  // namespace csp3

extern "C" {

int csp3_version(void) { return 100; }
const char *csp3_last_error_string(void) { return get_error(); }

int csp3_device_count(void)
{
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess) { cudaGetLastError(); return 0; }
    return cnt;
}

int csp3_set_device(int device)
{
    CSP3_CUDA(cudaSetDevice(device));
    return 0;
}

// ---- SpMV plan ------------------------------------------------------------------------------------------
int csp3_spmv_plan_create(int64_t m, int64_t n, const int32_t *Ap_dev, const int32_t *Ai_dev, void *stream,
                          csp3_spmv_plan **plan)
{
    if (!plan || m < 0 || n < 0) { set_error("spmv_plan_create: bad arguments"); return CSP3_ERR_ARG; }
    if (int rc = require_device()) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    i32 nnz = 0;
    CSP3_CUDA(cudaMemcpyAsync(&nnz, Ap_dev + n, 4, cudaMemcpyDeviceToHost, st));
    CSP3_CUDA(cudaStreamSynchronize(st));
    auto *P = new csp3_spmv_plan();
    P->d.m = m; P->d.n = n; P->d.nnz = nnz;
    if (cudaMalloc((void **)&P->d.rp, (size_t)(m + 1) * 4) != cudaSuccess ||
        cudaMalloc((void **)&P->d.rc, (size_t)std::max(nnz, 1) * 4) != cudaSuccess ||
        cudaMalloc((void **)&P->d.perm, (size_t)std::max(nnz, 1) * 4) != cudaSuccess) {
        csp3_spmv_plan_destroy(P);
        set_error("spmv_plan_create: device allocation failed");
        return CSP3_ERR_ALLOC;
    }
    int rc = transpose_device(m, n, Ap_dev, Ai_dev, nullptr, nnz, P->d.rp, P->d.rc, nullptr, P->d.perm, st);
    if (rc == 0) rc = spmv_plan_pack(P->d, st);
    if (rc) { csp3_spmv_plan_destroy(P); return rc; }
    // the plan may be used on any stream afterwards: its index arrays must be complete when this call returns
    if (cudaStreamSynchronize(st) != cudaSuccess) { set_error("spmv_plan_create: %s", cudaGetErrorString(cudaGetLastError())); csp3_spmv_plan_destroy(P); return CSP3_ERR_CUDA; }
    *plan = P;
    return 0;
}

int csp3_spmv_plan_destroy(csp3_spmv_plan *plan)
{
    if (!plan) return 0;
    cudaFree(plan->d.rp); cudaFree(plan->d.rc); cudaFree(plan->d.perm); cudaFree(plan->d.pk);
    delete plan;
    return 0;
}

int csp3_spmv_batched(const csp3_spmv_plan *plan, int64_t batch, const double *Ax, int64_t stride_ax,
                      const double *x, double *y, double beta, void *stream)
{
    if (!plan) { set_error("spmv_batched: null plan"); return CSP3_ERR_ARG; }
    return spmv_device(plan->d, batch, Ax, stride_ax, x, y, beta, (cudaStream_t)stream);
}

static int spmv_host_common(int64_t m, int64_t n, int64_t nv, const int32_t *Ap, const int32_t *Ai,
                            const double *Ax, const double *X, double *Y, double beta)
{
    if (m < 0 || n < 0 || !Ap) { set_error("matvec: bad arguments"); return CSP3_ERR_ARG; }
    if (int rc = require_device()) return rc;
    const i64 nnz = Ap[n];
    Dev dAp, dAi, dAx, dX, dY;
    CSP3_TRY(dAp.put(Ap, (size_t)(n + 1) * 4));
    CSP3_TRY(dAi.put(Ai, (size_t)nnz * 4));
    CSP3_TRY(dAx.put(Ax, (size_t)nnz * 8));
    CSP3_TRY(dX.put(X, (size_t)n * nv * 8));
    if (beta != 0.0) CSP3_TRY(dY.put(Y, (size_t)m * nv * 8)); else CSP3_TRY(dY.alloc((size_t)m * nv * 8));
    csp3_spmv_plan *P = nullptr;
    int rc = csp3_spmv_plan_create(m, n, dAp.as<i32>(), dAi.as<i32>(), nullptr, &P);
    if (rc) return rc;
    if (nv == 1) rc = spmv_device(P->d, 1, dAx.as<double>(), 0, dX.as<double>(), dY.as<double>(), beta, nullptr);
    else rc = spmm_device(P->d, nv, dAx.as<double>(), dX.as<double>(), dY.as<double>(), nullptr);
    if (rc == 0 && cudaDeviceSynchronize() != cudaSuccess) { set_error("matvec kernel failed: %s", cudaGetErrorString(cudaGetLastError())); rc = CSP3_ERR_CUDA; }
    csp3_spmv_plan_destroy(P);
    if (rc) return rc;
    CSP3_TRY(dY.get(Y, (size_t)m * nv * 8));
    return 0;
}

int csp3_csc_mat_vec_ff_host(int64_t m, int64_t n, const int32_t *Ap, const int32_t *Ai, const double *Ax,
                             const double *x, double *y)
{ return spmv_host_common(m, n, 1, Ap, Ai, Ax, x, y, 0.0); }

int csp3_csc_matvec_host(int64_t n_row, int64_t n_col, const int32_t *Ap, const int32_t *Ai, const double *Ax,
                         const double *Xx, double *Yx)
{ return spmv_host_common(n_row, n_col, 1, Ap, Ai, Ax, Xx, Yx, 1.0); }

int csp3_csc_matvecs_host(int64_t n_row, int64_t n_col, int64_t n_vecs, const int32_t *Ap, const int32_t *Ai,
                          const double *Ax, const double *Xx, double *Yx)
{
    if (n_vecs == 1) return spmv_host_common(n_row, n_col, 1, Ap, Ai, Ax, Xx, Yx, 1.0);
    return spmv_host_common(n_row, n_col, n_vecs, Ap, Ai, Ax, Xx, Yx, 1.0);
}

// ---- transposition ----------------------------------------------------------------------------------------
int csp3_csc_transpose(int64_t m, int64_t n, const int32_t *Ap, const int32_t *Ai, const double *Ax,
                       int32_t *Cp, int32_t *Ci, double *Cx, void *stream)
{
    if (int rc = require_device()) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    i32 nnz = 0;
    CSP3_CUDA(cudaMemcpyAsync(&nnz, Ap + n, 4, cudaMemcpyDeviceToHost, st));
    CSP3_CUDA(cudaStreamSynchronize(st));
    return transpose_device(m, n, Ap, Ai, Ax, nnz, Cp, Ci, Cx, nullptr, st);
}

int csp3_csc_transpose_host(int64_t m, int64_t n, const int32_t *Ap, const int32_t *Ai, const double *Ax,
                            int32_t *Cp, int32_t *Ci, double *Cx)
{
    if (m < 0 || n < 0 || !Ap) { set_error("transpose: bad arguments"); return CSP3_ERR_ARG; }
    if (int rc = require_device()) return rc;
    const i64 nnz = Ap[n];
    Dev dAp, dAi, dAx, dCp, dCi, dCx;
    CSP3_TRY(dAp.put(Ap, (size_t)(n + 1) * 4));
    CSP3_TRY(dAi.put(Ai, (size_t)nnz * 4));
    CSP3_TRY(dAx.put(Ax, (size_t)nnz * 8));
    CSP3_TRY(dCp.alloc((size_t)(m + 1) * 4));
    CSP3_TRY(dCi.alloc((size_t)nnz * 4));
    CSP3_TRY(dCx.alloc((size_t)nnz * 8));
    int rc = transpose_device(m, n, dAp.as<i32>(), dAi.as<i32>(), dAx.as<double>(), (i32)nnz, dCp.as<i32>(),
                              dCi.as<i32>(), dCx.as<double>(), nullptr, nullptr);
    if (rc) return rc;
    CSP3_CUDA(cudaDeviceSynchronize());
    CSP3_TRY(dCp.get(Cp, (size_t)(m + 1) * 4));
    CSP3_TRY(dCi.get(Ci, (size_t)nnz * 4));
    CSP3_TRY(dCx.get(Cx, (size_t)nnz * 8));
    return 0;
}

int csp3_csc_to_csr_host(int64_t m, int64_t n, const int32_t *Ap, const int32_t *Ai, const double *Ax,
                         int32_t *Bp, int32_t *Bi, double *Bx)
{ return csp3_csc_transpose_host(m, n, Ap, Ai, Ax, Bp, Bi, Bx); }

// ---- SpGEMM ----------------------------------------------------------------------------------------------
int csp3_spgemm_symbolic(int64_t Am, int64_t An, const int32_t *Ap, const int32_t *Ai, int64_t Bm, int64_t Bn,
                         const int32_t *Bp, const int32_t *Bi, int32_t *Cp, int64_t *nnz_host, void *stream)
{
    if (int rc = require_device()) return rc;
    return spgemm_device(false, Am, An, Ap, Ai, nullptr, Bm, Bn, Bp, Bi, nullptr, Cp, nullptr, nullptr, nnz_host,
                         (cudaStream_t)stream);
}

int csp3_spgemm_numeric(int64_t Am, int64_t An, const int32_t *Ap, const int32_t *Ai, const double *Ax,
                        int64_t Bm, int64_t Bn, const int32_t *Bp, const int32_t *Bi, const double *Bx,
                        const int32_t *Cp, int32_t *Ci, double *Cx, void *stream)
{
    if (int rc = require_device()) return rc;
    return spgemm_device(true, Am, An, Ap, Ai, Ax, Bm, Bn, Bp, Bi, Bx, const_cast<i32 *>(Cp), Ci, Cx, nullptr,
                         (cudaStream_t)stream);
}

int csp3_spgemm_symbolic_host(int64_t Am, int64_t An, const int32_t *Ap, const int32_t *Ai, int64_t Bm,
                              int64_t Bn, const int32_t *Bp, const int32_t *Bi, int32_t *Cp, int64_t *nnz)
{
    if (An != Bm || !Ap || !Bp) { set_error("spgemm: bad arguments"); return CSP3_ERR_ARG; }
    if (int rc = require_device()) return rc;
    Dev dAp, dAi, dBp, dBi, dCp;
    CSP3_TRY(dAp.put(Ap, (size_t)(An + 1) * 4));
    CSP3_TRY(dAi.put(Ai, (size_t)Ap[An] * 4));
    CSP3_TRY(dBp.put(Bp, (size_t)(Bn + 1) * 4));
    CSP3_TRY(dBi.put(Bi, (size_t)Bp[Bn] * 4));
    CSP3_TRY(dCp.alloc((size_t)(Bn + 1) * 4));
    int rc = spgemm_device(false, Am, An, dAp.as<i32>(), dAi.as<i32>(), nullptr, Bm, Bn, dBp.as<i32>(),
                           dBi.as<i32>(), nullptr, dCp.as<i32>(), nullptr, nullptr, nnz, nullptr);
    if (rc) return rc;
    CSP3_CUDA(cudaDeviceSynchronize());
    CSP3_TRY(dCp.get(Cp, (size_t)(Bn + 1) * 4));
    return 0;
}

int csp3_spgemm_numeric_host(int64_t Am, int64_t An, const int32_t *Ap, const int32_t *Ai, const double *Ax,
                             int64_t Bm, int64_t Bn, const int32_t *Bp, const int32_t *Bi, const double *Bx,
                             const int32_t *Cp, int32_t *Ci, double *Cx)
{
    if (An != Bm || !Ap || !Bp || !Cp) { set_error("spgemm: bad arguments"); return CSP3_ERR_ARG; }
    if (int rc = require_device()) return rc;
    const i64 nnzC = Cp[Bn];
    Dev dAp, dAi, dAx, dBp, dBi, dBx, dCp, dCi, dCx;
    CSP3_TRY(dAp.put(Ap, (size_t)(An + 1) * 4));
    CSP3_TRY(dAi.put(Ai, (size_t)Ap[An] * 4));
    CSP3_TRY(dAx.put(Ax, (size_t)Ap[An] * 8));
    CSP3_TRY(dBp.put(Bp, (size_t)(Bn + 1) * 4));
    CSP3_TRY(dBi.put(Bi, (size_t)Bp[Bn] * 4));
    CSP3_TRY(dBx.put(Bx, (size_t)Bp[Bn] * 8));
    CSP3_TRY(dCp.put(Cp, (size_t)(Bn + 1) * 4));
    CSP3_TRY(dCi.alloc((size_t)nnzC * 4));
    CSP3_TRY(dCx.alloc((size_t)nnzC * 8));
    int rc = spgemm_device(true, Am, An, dAp.as<i32>(), dAi.as<i32>(), dAx.as<double>(), Bm, Bn, dBp.as<i32>(),
                           dBi.as<i32>(), dBx.as<double>(), dCp.as<i32>(), dCi.as<i32>(), dCx.as<double>(),
                           nullptr, nullptr);
    if (rc) return rc;
    CSP3_CUDA(cudaDeviceSynchronize());
    CSP3_TRY(dCi.get(Ci, (size_t)nnzC * 4));
    CSP3_TRY(dCx.get(Cx, (size_t)nnzC * 8));
    return 0;
}

// ---- A + B / A - B ---------------------------------------------------------------------------------------
int csp3_csc_plusminus_host(int64_t m, int64_t n, const int32_t *Ap, const int32_t *Ai, const double *Ax,
                            const int32_t *Bp, const int32_t *Bi, const double *Bx, double sign, int32_t *Cp,
                            int32_t *Ci, double *Cx)
{
    if (m < 0 || n < 0 || !Ap || !Bp || !Cp) { set_error("csc_plusminus: bad arguments"); return CSP3_ERR_ARG; }
    if (int rc = require_device()) return rc;
    const i64 na = Ap[n], nb = Bp[n];
    Dev dAp, dAi, dAx, dBp, dBi, dBx, dCp, dCi, dCx;
    CSP3_TRY(dAp.put(Ap, (size_t)(n + 1) * 4)); CSP3_TRY(dAi.put(Ai, (size_t)na * 4)); CSP3_TRY(dAx.put(Ax, (size_t)na * 8));
    CSP3_TRY(dBp.put(Bp, (size_t)(n + 1) * 4)); CSP3_TRY(dBi.put(Bi, (size_t)nb * 4)); CSP3_TRY(dBx.put(Bx, (size_t)nb * 8));
    CSP3_TRY(dCp.alloc((size_t)(n + 1) * 4)); CSP3_TRY(dCi.alloc((size_t)(na + nb) * 4)); CSP3_TRY(dCx.alloc((size_t)(na + nb) * 8));
    int rc = csc_add_device(m, n, dAp.as<i32>(), dAi.as<i32>(), dAx.as<double>(), dBp.as<i32>(), dBi.as<i32>(),
                            dBx.as<double>(), sign, dCp.as<i32>(), dCi.as<i32>(), dCx.as<double>(), nullptr);
    if (rc) return rc;
    CSP3_CUDA(cudaDeviceSynchronize());
    CSP3_TRY(dCp.get(Cp, (size_t)(n + 1) * 4));
    const i64 nc = Cp[n];
    CSP3_TRY(dCi.get(Ci, (size_t)nc * 4));
    CSP3_TRY(dCx.get(Cx, (size_t)nc * 8));
    return 0;
}

int csp3_csc_add_ff_host(int64_t m, int64_t n, const int32_t *Ap, const int32_t *Ai, const double *Ax,
                         const int32_t *Bp, const int32_t *Bi, const double *Bx, double alpha, double beta, int32_t *Cp,
                         int32_t *Ci, double *Cx)
{
    if (m < 0 || n < 0 || !Ap || !Bp || !Cp) { set_error("csc_add_ff: bad arguments"); return CSP3_ERR_ARG; }
    if (int rc = require_device()) return rc;
    const i64 na = Ap[n], nb = Bp[n];
    Dev dAp, dAi, dAx, dBp, dBi, dBx, dCp, dCi, dCx;
    CSP3_TRY(dAp.put(Ap, (size_t)(n + 1) * 4)); CSP3_TRY(dAi.put(Ai, (size_t)na * 4)); CSP3_TRY(dAx.put(Ax, (size_t)na * 8));
    CSP3_TRY(dBp.put(Bp, (size_t)(n + 1) * 4)); CSP3_TRY(dBi.put(Bi, (size_t)nb * 4)); CSP3_TRY(dBx.put(Bx, (size_t)nb * 8));
    CSP3_TRY(dCp.alloc((size_t)(n + 1) * 4)); CSP3_TRY(dCi.alloc((size_t)(na + nb) * 4)); CSP3_TRY(dCx.alloc((size_t)(na + nb) * 8));
    int rc = csc_add_ff_device(m, n, dAp.as<i32>(), dAi.as<i32>(), dAx.as<double>(), dBp.as<i32>(), dBi.as<i32>(),
                               dBx.as<double>(), alpha, beta, dCp.as<i32>(), dCi.as<i32>(), dCx.as<double>(), nullptr);
    if (rc) return rc;
    CSP3_CUDA(cudaDeviceSynchronize());
    CSP3_TRY(dCp.get(Cp, (size_t)(n + 1) * 4));
    const i64 nc = Cp[n];
    CSP3_TRY(dCi.get(Ci, (size_t)nc * 4));
    CSP3_TRY(dCx.get(Cx, (size_t)nc * 8));
    return 0;
}

// ---- host symbolic ----------------------------------------------------------------------------------------
int csp3_csc_amd(int64_t order, int64_t m, int64_t n, const int32_t *Ap, const int32_t *Ai, int32_t *q)
{
    if (order < 0 || order > 3 || m < 0 || n < 0 || !Ap || !q) { set_error("amd: bad arguments"); return CSP3_ERR_ARG; }
    std::vector<i32> p = amd_order(order, m, n, Ap, Ai);
    std::memcpy(q, p.data(), (size_t)n * 4);
    return 0;
}

int csp3_csc_etree(int64_t m, int64_t n, const int32_t *Ap, const int32_t *Ai, int ata, int32_t *parent)
{
    if (!Ap || !parent) { set_error("etree: bad arguments"); return CSP3_ERR_ARG; }
    std::vector<i32> p = etree(m, n, Ap, Ai, ata != 0);
    std::memcpy(parent, p.data(), (size_t)n * 4);
    return 0;
}

int csp3_csc_post(int64_t n, const int32_t *parent, int32_t *post)
{
    if (!parent || !post) { set_error("post: bad arguments"); return CSP3_ERR_ARG; }
    std::vector<i32> p = postorder(n, parent);
    std::memcpy(post, p.data(), (size_t)n * 4);
    return 0;
}

int csp3_lu_analyze(int64_t order, int64_t n, const int32_t *Ap, const int32_t *Ai, const double *Ax,
                    const int32_t *q_in, double tol, csp3_lu_symbolic **sym)
{
    if (!sym || n < 0 || !Ap || !Ai || !Ax || order < 0 || order > 3) { set_error("lu_analyze: bad arguments"); return CSP3_ERR_ARG; }
    std::unique_ptr<csp3_lu_symbolic> Sy(new csp3_lu_symbolic());
    Sy->n = n; Sy->nnzA = Ap[n];
    Sy->Ap.assign(Ap, Ap + n + 1);
    Sy->Ai.assign(Ai, Ai + Ap[n]);
    if (q_in) {
        Sy->q.assign(q_in, q_in + n);
        std::vector<char> seen((size_t)n, 0);
        for (i64 k = 0; k < n; ++k) {
            if (q_in[k] < 0 || q_in[k] >= n || seen[(size_t)q_in[k]]) { set_error("lu_analyze: q is not a permutation of 0..n-1"); return CSP3_ERR_ARG; }
            seen[(size_t)q_in[k]] = 1;
        }
    } else Sy->q = amd_order(order, n, n, Ap, Ai);
    const int st = lu_factor(n, Ap, Ai, Ax, Sy->q.data(), tol, Sy->F);
    if (st != 0) { set_error("lu_analyze: no non-zero pivot in step %d", st - 1); return st; }
    const char *why = "";
    if (!build_schedule(n, Ap, Ai, Sy->q, Sy->F, Sy->S, &why)) { set_error("lu_analyze: %s", why); return CSP3_ERR_ARG; }
    compile_wide(*Sy);
    *sym = Sy.release();
    return 0;
}

int csp3_lu_analyze_fixed(int64_t n, const int32_t *Ap, const int32_t *Ai, const int32_t *q, const int32_t *pinv,
                          const int32_t *Lp, const int32_t *Li, const int32_t *Up, const int32_t *Ui,
                          csp3_lu_symbolic **sym)
{
    if (!sym || n < 0 || !Ap || !Ai || !pinv || !Lp || !Li || !Up || !Ui) { set_error("lu_analyze_fixed: bad arguments"); return CSP3_ERR_ARG; }
    std::unique_ptr<csp3_lu_symbolic> Sy(new csp3_lu_symbolic());
    Sy->n = n; Sy->nnzA = Ap[n];
    Sy->Ap.assign(Ap, Ap + n + 1);
    Sy->Ai.assign(Ai, Ai + Ap[n]);
    Sy->q.resize((size_t)n);
    for (i64 k = 0; k < n; ++k) Sy->q[k] = q ? q[k] : (i32)k;
    Factor &F = Sy->F;
    F.pinv.assign(pinv, pinv + n);
    F.Lp.assign(Lp, Lp + n + 1); F.Li.assign(Li, Li + Lp[n]);
    F.Up.assign(Up, Up + n + 1); F.Ui.assign(Ui, Ui + Up[n]);
    F.Lx.assign((size_t)Lp[n], 0.0); F.Ux.assign((size_t)Up[n], 0.0);
    {   // the compiled device programs index memory with these arrays: check them before anything is built from them
        auto is_perm = [&](const std::vector<i32> &p) {
            std::vector<char> seen((size_t)n, 0);
            for (i64 k = 0; k < n; ++k) { if (p[(size_t)k] < 0 || p[(size_t)k] >= n || seen[(size_t)p[(size_t)k]]) return false; seen[(size_t)p[(size_t)k]] = 1; }
            return true;
        };
        if (!is_perm(Sy->q) || !is_perm(F.pinv)) { set_error("lu_analyze_fixed: q and pinv must be permutations of 0..n-1"); return CSP3_ERR_ARG; }
        if (Lp[0] != 0 || Up[0] != 0 || Ap[0] != 0) { set_error("lu_analyze_fixed: column pointers must start at 0"); return CSP3_ERR_ARG; }
        for (i64 k = 0; k < n; ++k)
            if (Lp[k + 1] < Lp[k] || Up[k + 1] < Up[k] || Ap[k + 1] < Ap[k]) { set_error("lu_analyze_fixed: column pointers must be non-decreasing"); return CSP3_ERR_ARG; }
        for (i64 p = 0; p < Lp[n]; ++p) if (Li[p] < 0 || Li[p] >= n) { set_error("lu_analyze_fixed: Li out of range"); return CSP3_ERR_ARG; }
        for (i64 p = 0; p < Up[n]; ++p) if (Ui[p] < 0 || Ui[p] >= n) { set_error("lu_analyze_fixed: Ui out of range"); return CSP3_ERR_ARG; }
        for (i64 p = 0; p < Ap[n]; ++p) if (Ai[p] < 0 || Ai[p] >= n) { set_error("lu_analyze_fixed: Ai out of range"); return CSP3_ERR_ARG; }
    }
    for (i64 k = 0; k < n; ++k) {
        if (Lp[k + 1] <= Lp[k] || Li[Lp[k]] != k || Up[k + 1] <= Up[k] || Ui[Up[k + 1] - 1] != k) {
            set_error("lu_analyze_fixed: column %lld is not in cs_lu layout (L diagonal first, U diagonal last)", (long long)k);
            return CSP3_ERR_ARG;
        }
    }
    const char *why = "";
    if (!build_schedule(n, Ap, Ai, Sy->q, Sy->F, Sy->S, &why)) { set_error("lu_analyze_fixed: %s", why); return CSP3_ERR_ARG; }
    compile_wide(*Sy);
    *sym = Sy.release();
    return 0;
}

int csp3_lu_sizes(const csp3_lu_symbolic *sym, int64_t out[16])
{
    if (!sym || !out) { set_error("lu_sizes: bad arguments"); return CSP3_ERR_ARG; }
    std::memset(out, 0, 16 * sizeof(int64_t));
    out[0] = sym->n; out[1] = sym->nnzA; out[2] = (i64)sym->F.Li.size(); out[3] = (i64)sym->F.Ui.size();
    out[4] = sym->S.lev_refactor.nlev(); out[5] = sym->S.lev_lsolve.nlev(); out[6] = sym->S.lev_usolve.nlev();
    out[7] = sym->S.flops;
    out[8] = (i64)(sym->S.rf_prog.bytes.size() + sym->S.ls_prog.bytes.size() + sym->S.ur_prog.bytes.size());
    out[9] = sym->S.max_col_len;
    out[10] = sym->W.ok ? sym->W.width : 0; out[11] = sym->W.ring_entries; out[12] = sym->W.stage_entries;
    out[13] = sym->W.immediate_fetches; out[14] = sym->W.near_fma; out[15] = (i64)sym->W.smem_bytes;
    return 0;
}

int csp3_lu_get_pattern(const csp3_lu_symbolic *sym, int32_t *q, int32_t *pinv, int32_t *Lp, int32_t *Li,
                        int32_t *Up, int32_t *Ui, double *Lx, double *Ux)
{
    if (!sym) { set_error("lu_get_pattern: null handle"); return CSP3_ERR_ARG; }
    auto cp = [](void *dst, const void *src, size_t bytes) { if (dst && bytes) std::memcpy(dst, src, bytes); };
    const Factor &F = sym->F;
    cp(q, sym->q.data(), sym->q.size() * 4);
    cp(pinv, F.pinv.data(), F.pinv.size() * 4);
    cp(Lp, F.Lp.data(), F.Lp.size() * 4); cp(Li, F.Li.data(), F.Li.size() * 4);
    cp(Up, F.Up.data(), F.Up.size() * 4); cp(Ui, F.Ui.data(), F.Ui.size() * 4);
    cp(Lx, F.Lx.data(), F.Lx.size() * 8); cp(Ux, F.Ux.data(), F.Ux.size() * 8);
    return 0;
}

int csp3_lu_supernodes(const csp3_lu_symbolic *sym, int32_t *sn_ptr, int64_t *count)
{
    if (!sym || !sn_ptr || !count) { set_error("lu_supernodes: bad arguments"); return CSP3_ERR_ARG; }
    const Factor &F = sym->F;
    const i64 n = sym->n;
    // fundamental supernodes of L: column j+1 continues the supernode of j when rows(L(:,j)) \ {j+1} == rows(L(:,j+1))
    std::vector<i32> a, b;
    i64 k = 0;
    sn_ptr[0] = 0;
    for (i64 j = 0; j + 1 <= n; ++j) {
        bool cont = false;
        if (j + 1 < n) {
            a.assign(F.Li.begin() + F.Lp[(size_t)j] + 1, F.Li.begin() + F.Lp[(size_t)j + 1]);
            b.assign(F.Li.begin() + F.Lp[(size_t)j + 1] + 1, F.Li.begin() + F.Lp[(size_t)j + 2]);
            std::sort(a.begin(), a.end()); std::sort(b.begin(), b.end());
            cont = a.size() == b.size() + 1 && !a.empty() && a[0] == (i32)(j + 1) && std::equal(b.begin(), b.end(), a.begin() + 1);
        }
        if (!cont) sn_ptr[++k] = (i32)(j + 1);
    }
    *count = k;
    return 0;
}

int csp3_lu_get_levels(const csp3_lu_symbolic *sym, int kind, int32_t *level, int32_t *order, int32_t *lptr)
{
    if (!sym || kind < 0 || kind > 2) { set_error("lu_get_levels: bad arguments"); return CSP3_ERR_ARG; }
    const LevelSet &L = kind == 0 ? sym->S.lev_refactor : (kind == 1 ? sym->S.lev_lsolve : sym->S.lev_usolve);
    if (level) std::memcpy(level, L.level.data(), L.level.size() * 4);
    if (order) std::memcpy(order, L.order.data(), L.order.size() * 4);
    if (lptr) std::memcpy(lptr, L.lptr.data(), L.lptr.size() * 4);
    return 0;
}

int64_t csp3_lu_get_program(const csp3_lu_symbolic *sym, int which, uint8_t *buf, int64_t capacity, int64_t geometry[8])
{
    if (!sym) { set_error("lu_get_program: null handle"); return CSP3_ERR_ARG; }
    const Program *P = nullptr;
    switch (which) {
        case 0: P = &sym->S.rf_prog; break;
        case 1: P = &sym->S.ls_prog; break;
        case 2: P = &sym->S.ur_prog; break;
        case 3: P = sym->W.ok ? &sym->W.prog : nullptr; break;
        case 4: P = sym->WF.ok ? &sym->WF.prog : nullptr; break;
        case 5: P = sym->WB.ok ? &sym->WB.prog : nullptr; break;
        case 6: P = sym->PP.ok ? &sym->PP.prog : nullptr; break;
        default: break;
    }
    if (which == 9 || which == 10) {                     // row-sweep programs (rowsweep_program.hpp): 9 forward, 10 backward
        csp3_lu_symbolic &Sy = *const_cast<csp3_lu_symbolic *>(sym);
        std::lock_guard<std::mutex> lock(Sy.rl_mu);
        if (!Sy.RStried) {
            Sy.RStried = true;
            const char *why = "";
            if (!compile_row_sweep(Sy.F, Sy.S, true, Sy.RSf, &why) || !compile_row_sweep(Sy.F, Sy.S, false, Sy.RSb, &why)) { Sy.RSf = RowSweepProgram(); Sy.RSb = RowSweepProgram(); }
        }
        const RowSweepProgram &RS = which == 9 ? Sy.RSf : Sy.RSb;
        if (!RS.ok) { set_error("lu_get_program: program %d not available", which); return CSP3_ERR_ARG; }
        if (geometry) {
            std::memset(geometry, 0, 8 * sizeof(int64_t));
            geometry[0] = RS.levels; geometry[1] = RS.warps; geometry[2] = RS.panels; geometry[3] = RS.chunks; geometry[4] = RS.terms;
            for (int w = 0; w < 3; ++w) geometry[5 + w] = RS.stream_off[w + 1];      // (all offsets follow from the END panels as well)
        }
        if (buf && capacity >= (int64_t)(RS.words.size() * 4)) std::memcpy(buf, RS.words.data(), RS.words.size() * 4);
        return (int64_t)(RS.words.size() * 4);
    }
    if (which == 7) {                                    // row-lane program: 44 words per quad (rowlane_program.hpp)
        // the geometry CSP3_RL_W / CSP3_RL_NQ ask for, otherwise one warp per bundle (compiled here when not yet cached)
        csp3_lu_symbolic &Sy = *const_cast<csp3_lu_symbolic *>(sym);
        const RowlaneProgram *RLp;
        { std::lock_guard<std::mutex> lock(Sy.rl_mu); RLp = rowlane_program(Sy, tuning().rl_warps > 0 ? kRlVariants - 1 : 0); }
        if (!RLp) { set_error("lu_get_program: program %d not available", which); return CSP3_ERR_ARG; }
        const RowlaneProgram &RL = *RLp;
        const std::vector<uint32_t> &v = RL.words;
        if (geometry) {
            std::memset(geometry, 0, 8 * sizeof(int64_t));
            geometry[0] = RL.update_quads; geometry[1] = RL.stage_quads | (RL.warps << 8); geometry[2] = RL.nslots; geometry[3] = RL.late_quads;
            geometry[4] = RL.ops; geometry[5] = RL.quads; geometry[6] = (i64)RL.smem_bytes; geometry[7] = RL.update_records;
        }
        if (buf && capacity >= (int64_t)(v.size() * 4)) std::memcpy(buf, v.data(), v.size() * 4);
        return (int64_t)(v.size() * 4);
    }
    if (!P) { set_error("lu_get_program: program %d not available", which); return CSP3_ERR_ARG; }
    if (geometry) {
        std::memset(geometry, 0, 8 * sizeof(int64_t));
        geometry[0] = P->stage;
        if (which == 3) {
            geometry[1] = sym->W.width; geometry[2] = sym->W.acc_slots; geometry[3] = sym->W.ring_entries;
            geometry[4] = sym->W.stage_entries; geometry[5] = sym->W.records; geometry[6] = (i64)sym->W.smem_bytes;
            geometry[7] = sym->W.groups;
        }
        if (which == 6) {
            geometry[0] = sym->PP.ring; geometry[1] = sym->PP.width; geometry[2] = sym->PP.nslots; geometry[3] = sym->PP.landing; geometry[4] = sym->PP.ops;
            geometry[5] = sym->PP.steps; geometry[6] = (i64)sym->PP.smem_bytes; geometry[7] = sym->PP.groups;
        }
        if (which == 4 || which == 5) {
            const WideSweep &Wsw = which == 4 ? sym->WF : sym->WB;
            geometry[1] = Wsw.width; geometry[2] = Wsw.nslots; geometry[3] = Wsw.landing_entries; geometry[5] = Wsw.records;
            geometry[6] = (i64)Wsw.smem_bytes; geometry[7] = Wsw.groups;
        }
    }
    if (buf && capacity >= (int64_t)P->bytes.size()) std::memcpy(buf, P->bytes.data(), P->bytes.size());
    return (int64_t)P->bytes.size();
}

static void free_stage(csp3_lu_symbolic::Stage &g)
{
    for (int s = 0; s < 3; ++s) {
        cudaFree(g.Ax[s]); cudaFree(g.b[s]); cudaFree(g.x[s]); cudaFree(g.Lx[s]); cudaFree(g.Ux[s]); cudaFree(g.status[s]);
        if (g.st[s]) cudaStreamDestroy(g.st[s]);
    }
    g = csp3_lu_symbolic::Stage();
}

int csp3_lu_destroy(csp3_lu_symbolic *sym)
{
    if (!sym) return 0;
    int cur = 0;
    const bool have = cudaGetDevice(&cur) == cudaSuccess;
    for (int d = 0; d < kMaxDevices; ++d) {
        if (!sym->dev[d].ready && !sym->stage[d].ready) continue;
        if (have) cudaSetDevice(d);
        if (sym->dev[d].arena) cudaFree(sym->dev[d].arena);
        for (auto &R : sym->dev[d].rl) if (R.prog) cudaFree(R.prog);
        if (sym->dev[d].rs.prog_f) cudaFree(sym->dev[d].rs.prog_f);
        if (sym->dev[d].rs.prog_b) cudaFree(sym->dev[d].rs.prog_b);
        if (sym->stage[d].ready) free_stage(sym->stage[d]);
    }
    if (have) cudaSetDevice(cur);
    cudaGetLastError();
    delete sym;
    return 0;
}

int csp3_lu_upload(csp3_lu_symbolic *sym, void *stream)
{
    if (!sym) { set_error("lu_upload: null handle"); return CSP3_ERR_ARG; }
    if (int rc = require_device()) return rc;
    int devid = 0;
    CSP3_CUDA(cudaGetDevice(&devid));
    if (devid >= kMaxDevices) { set_error("lu_upload: device index %d not supported", devid); return CSP3_ERR_ARG; }
    std::lock_guard<std::mutex> lock(sym->mu);
    DevSchedule &D = sym->dev[devid];
    if (D.ready) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const Schedule &S = sym->S;
    const Factor &F = sym->F;
    struct Piece { const void *src; size_t bytes; size_t off; };
    std::vector<Piece> pieces;
    size_t total = 0;
    auto add = [&](const void *src, size_t bytes) { total = (total + 255) & ~(size_t)255; pieces.push_back({src, bytes, total}); total += bytes; return pieces.size() - 1; };
    const size_t i_rf = add(S.rf_prog.bytes.data(), S.rf_prog.bytes.size());
    const size_t i_ls = add(S.ls_prog.bytes.data(), S.ls_prog.bytes.size());
    const size_t i_us = add(S.us_prog.bytes.data(), S.us_prog.bytes.size());
    const size_t i_ur = add(S.ur_prog.bytes.data(), S.ur_prog.bytes.size());
    const size_t i_wrf = add(sym->W.prog.bytes.data(), sym->W.ok ? sym->W.prog.bytes.size() : 0);
    const bool wsolve = sym->W.ok && sym->WF.ok && sym->WB.ok;
    const size_t i_wfs = add(sym->WF.prog.bytes.data(), wsolve ? sym->WF.prog.bytes.size() : 0);
    const size_t i_wbs = add(sym->WB.prog.bytes.data(), wsolve ? sym->WB.prog.bytes.size() : 0);
    const size_t i_prf = add(sym->PP.prog.bytes.data(), sym->PP.ok ? sym->PP.prog.bytes.size() : 0);
    std::vector<uint8_t> ldiag(F.Li.size(), 0);
    for (i64 k = 0; k < sym->n; ++k) ldiag[(size_t)F.Lp[(size_t)k]] = 1;
    const size_t i_ldiag = add(ldiag.data(), ldiag.size());
    const size_t i_pinv = add(F.pinv.data(), wsolve ? F.pinv.size() * 4 : 0);
    const size_t i_qinv = add(sym->qinv.data(), wsolve ? sym->qinv.size() * 4 : 0);
    total = (total + 255) & ~(size_t)255;
    char *arena = nullptr;
    if (cudaMalloc((void **)&arena, total ? total : 256) != cudaSuccess) { set_error("lu_upload: device allocation of %zu bytes failed", total); cudaGetLastError(); return CSP3_ERR_ALLOC; }
    for (auto &pc : pieces)
        if (pc.bytes && cudaMemcpyAsync(arena + pc.off, pc.src, pc.bytes, cudaMemcpyHostToDevice, st) != cudaSuccess) {
            set_error("lu_upload: copy failed (%s)", cudaGetErrorString(cudaGetLastError()));
            cudaFree(arena);
            return CSP3_ERR_CUDA;
        }
    if (cudaStreamSynchronize(st) != cudaSuccess) {
        set_error("lu_upload: %s", cudaGetErrorString(cudaGetLastError()));
        cudaFree(arena);
        return CSP3_ERR_CUDA;
    }
    auto at = [&](size_t idx) { return (const void *)(arena + pieces[idx].off); };
    D.arena = arena; D.arena_bytes = total;
    D.n = (i32)sym->n; D.nnzA = (i32)sym->nnzA; D.lnz = (i32)F.Li.size(); D.unz = (i32)F.Ui.size();
    D.max_col_len = S.max_col_len;
    D.rf_prog = (const uint8_t *)at(i_rf); D.rf_prog_bytes = (i32)S.rf_prog.bytes.size(); D.rf_prog_stage = S.rf_prog.stage;
    D.ls_prog = (const uint8_t *)at(i_ls); D.ls_prog_bytes = (i32)S.ls_prog.bytes.size(); D.ls_prog_stage = S.ls_prog.stage;
    D.us_prog = (const uint8_t *)at(i_us); D.us_prog_bytes = (i32)S.us_prog.bytes.size(); D.us_prog_stage = S.us_prog.stage;
    D.ur_prog = (const uint8_t *)at(i_ur); D.ur_prog_bytes = (i32)S.ur_prog.bytes.size(); D.ur_prog_stage = S.ur_prog.stage;
    D.ls_nslots = S.ls.nslots; D.us_nslots = S.us.nslots; D.ur_nslots = S.ur_nslots; D.ur_max_len = S.ur_max_len;
    if (sym->W.ok) {
        const WideProgram &W = sym->W;
        D.wide_ok = true; D.wide_S = W.width; D.wide_R = W.width * W.groups / 32;
        D.wrf_prog = (const uint8_t *)at(i_wrf); D.wrf_prog_bytes = (i32)W.prog.bytes.size(); D.wrf_prog_stage = W.prog.stage;
        D.wrf_acc_slots = W.acc_slots; D.wrf_lsrc_entries = W.ring_entries + W.stage_entries; D.wrf_smem = W.smem_bytes;
        D.wrf_groups = W.ngroups;
        if (wsolve) {
            D.wide_solve_ok = true;
            D.wfs_prog = (const uint8_t *)at(i_wfs); D.wfs_prog_bytes = (i32)sym->WF.prog.bytes.size(); D.wfs_prog_stage = sym->WF.prog.stage;
            D.wfs_records = sym->WF.records; D.wfs_nslots = sym->WF.nslots; D.wfs_smem = sym->WF.smem_bytes;
            D.wbs_prog = (const uint8_t *)at(i_wbs); D.wbs_prog_bytes = (i32)sym->WB.prog.bytes.size(); D.wbs_prog_stage = sym->WB.prog.stage;
            D.wbs_records = sym->WB.records; D.wbs_nslots = sym->WB.nslots; D.wbs_smem = sym->WB.smem_bytes;
            D.d_pinv = (const i32 *)at(i_pinv); D.d_qinv = (const i32 *)at(i_qinv);
        }
    }
    if (sym->PP.ok && sym->PP.smem_bytes <= (size_t)200 * 1024) {
        D.panel_ok = true;
        D.prf_prog = (const uint8_t *)at(i_prf); D.prf_prog_bytes = (i32)sym->PP.prog.bytes.size();
        D.prf_nslots = sym->PP.nslots; D.prf_lsrc = sym->PP.ring + sym->PP.landing; D.prf_steps = sym->PP.steps; D.prf_smem = sym->PP.smem_bytes;
    }
    D.owner = sym; D.devid = devid;
    D.rl_enabled = D.wide_ok && D.wide_S == 8 && D.wide_solve_ok;      // the sweeps that read 8-system bundles
    D.d_ldiag = (const uint8_t *)at(i_ldiag);
    D.ready = true;
    return 0;
}

static const DevSchedule *current_schedule(const csp3_lu_symbolic *sym)
{
    if (!sym) { set_error("null symbolic handle"); return nullptr; }
    int devid = 0;
    if (cudaGetDevice(&devid) != cudaSuccess || devid >= kMaxDevices || !sym->dev[devid].ready) {
        cudaGetLastError();
        set_error("symbolic object not uploaded to the current device (call csp3_lu_upload)");
        return nullptr;
    }
    return &sym->dev[devid];
}

const char *csp3_lu_refactor_kernel_name(const csp3_lu_symbolic *sym, int64_t batch)
{
    const DevSchedule *D = current_schedule(sym);
    if (!D) return nullptr;
    if (use_panel(*D, batch)) return "lu_refactor_panel_kernel";
    if (use_rowlane(*D, batch)) return "lu_refactor_rowlane_kernel";
    if (use_tmem(*D, batch)) return "lu_refactor_tmem_kernel";
    if (use_wide(*D, batch)) return "lu_refactor_wide_kernel";
    return "lu_refactor_kernel";
}

int csp3_lu_prepare(const csp3_lu_symbolic *sym, int64_t batch)
{
    const DevSchedule *D = current_schedule(sym);
    if (!D) return CSP3_ERR_ARG;
    if (batch < 0) { set_error("lu_prepare: bad arguments"); return CSP3_ERR_ARG; }
    (void)use_rowlane(*D, batch);              // compiles / uploads the row-lane program of this batch size when it is the choice
    return 0;
}

int csp3_lu_refactor_batched(const csp3_lu_symbolic *sym, int64_t batch, const double *Ax, double *Lx,
                             double *Ux, int32_t *status, void *stream)
{
    const DevSchedule *D = current_schedule(sym);
    if (!D) return CSP3_ERR_ARG;
    if (batch < 0 || !Ax || !Lx || !Ux) { set_error("lu_refactor_batched: bad arguments"); return CSP3_ERR_ARG; }
    return launch_refactor(*D, batch, Ax, Lx, Ux, status, false, (cudaStream_t)stream);
}

int csp3_lu_solve_batched(const csp3_lu_symbolic *sym, int64_t batch, const double *Lx, const double *Ux,
                          const double *b, double *x, void *stream)
{
    const DevSchedule *D = current_schedule(sym);
    if (!D) return CSP3_ERR_ARG;
    if (batch < 0 || !Lx || !Ux || !b || !x || b == x) { set_error("lu_solve_batched: bad arguments (b and x must not alias)"); return CSP3_ERR_ARG; }
    return launch_solve(*D, batch, Lx, Ux, b, x, nullptr, false, (cudaStream_t)stream);
}

int64_t csp3_lu_workspace_bytes(const csp3_lu_symbolic *sym, int64_t batch)
{
    if (!sym || batch < 0) return -1;
    const i64 padded = (batch + 31) / 32 * 32;
    return padded * (int64_t)(sym->F.Li.size() + sym->F.Ui.size() + 2 * (size_t)sym->n) * 8 + 512;
}

// workspace carve-up (bundle-interleaved): [ Lw | Uw | z ], each padded to whole bundles
static void carve_workspace(const DevSchedule &D, i64 batch, void *work, double **Lw, double **Uw, double **z)
{
    const i64 padded = (batch + 31) / 32 * 32;
    double *w = (double *)(((uintptr_t)work + 255) & ~(uintptr_t)255);
    *Lw = w; w += padded * D.lnz;
    *Uw = w; w += padded * D.unz;
    *z = w;
}

int csp3_lu_refactor_ws(const csp3_lu_symbolic *sym, int64_t batch, const double *Ax, void *work,
                        int32_t *status, void *stream)
{
    const DevSchedule *D = current_schedule(sym);
    if (!D) return CSP3_ERR_ARG;
    if (batch < 0 || !Ax || !work) { set_error("lu_refactor_ws: bad arguments"); return CSP3_ERR_ARG; }
    double *Lw, *Uw, *z;
    carve_workspace(*D, batch, work, &Lw, &Uw, &z);
    return launch_refactor(*D, batch, Ax, Lw, Uw, status, true, (cudaStream_t)stream);
}

int csp3_lu_solve_ws(const csp3_lu_symbolic *sym, int64_t batch, void *work, const double *b, double *x,
                     void *stream)
{
    const DevSchedule *D = current_schedule(sym);
    if (!D) return CSP3_ERR_ARG;
    if (batch < 0 || !work || !b || !x || b == x) { set_error("lu_solve_ws: bad arguments (b and x must not alias)"); return CSP3_ERR_ARG; }
    double *Lw, *Uw, *z;
    carve_workspace(*D, batch, work, &Lw, &Uw, &z);
    return launch_solve(*D, batch, Lw, Uw, b, x, z, true, (cudaStream_t)stream);
}

int csp3_lu_growth_ws(const csp3_lu_symbolic *sym, int64_t batch, const void *work, double *growth, void *stream)
{
    const DevSchedule *D = current_schedule(sym);
    if (!D) return CSP3_ERR_ARG;
    if (batch < 0 || !work || !growth) { set_error("lu_growth_ws: bad arguments"); return CSP3_ERR_ARG; }
    double *Lw, *Uw, *z;
    carve_workspace(*D, batch, const_cast<void *>(work), &Lw, &Uw, &z);
    return launch_growth(*D, batch, Lw, growth, (cudaStream_t)stream);
}

int csp3_lu_refactor_solve_batched(const csp3_lu_symbolic *sym, int64_t batch, const double *Ax,
                                   const double *b, double *x, double *Lx, double *Ux, int32_t *status,
                                   void *work, void *stream)
{
    const DevSchedule *D = current_schedule(sym);
    if (!D) return CSP3_ERR_ARG;
    if (batch < 0 || !Ax || !b || !x || b == x) { set_error("lu_refactor_solve_batched: bad arguments (b and x must not alias)"); return CSP3_ERR_ARG; }
    if (Lx && Ux) {                                     // caller wants the factors: system-major API layout
        if (int rc = launch_refactor(*D, batch, Ax, Lx, Ux, status, false, (cudaStream_t)stream)) return rc;
        return launch_solve(*D, batch, Lx, Ux, b, x, nullptr, false, (cudaStream_t)stream);
    }
    if (!work) { set_error("lu_refactor_solve_batched: Lx and Ux, or work, must be given"); return CSP3_ERR_ARG; }
    if (int rc = csp3_lu_refactor_ws(sym, batch, Ax, work, status, stream)) return rc;
    return csp3_lu_solve_ws(sym, batch, work, b, x, stream);
}

int csp3_lu_refactor_solve_host(csp3_lu_symbolic *sym, int64_t batch, const double *Ax, const double *b,
                                double *x, int32_t *status)
{
    if (!sym || batch < 0 || !Ax || !b || !x) { set_error("lu_refactor_solve_host: bad arguments"); return CSP3_ERR_ARG; }
    if (int rc = require_device()) return rc;
    if (int rc = csp3_lu_upload(sym, nullptr)) return rc;
    int devid = 0;
    CSP3_CUDA(cudaGetDevice(&devid));
    // The staging buffers and streams belong to the handle: calls on one handle are serialised (two host threads may
    // share a symbolic object; concurrent batches need one handle per thread or the device-pointer entry points).
    std::lock_guard<std::mutex> lock(sym->mu);
    const DevSchedule &D = sym->dev[devid];
    auto &G = sym->stage[devid];
    const i64 n = D.n, nnzA = D.nnzA;
    if (!G.ready) {
        // The kernels are latency-bound: a chunk costs about the same time whether it holds 500 or 5,000
        // systems, so chunks are as large as ~2 GB of values allows (kernels of different chunks overlap on
        // the device, copies overlap with kernels).
        i64 chunk = (2048ll << 20) / std::max<i64>(nnzA * 8, 1);
        chunk = std::max<i64>(256, std::min<i64>(chunk, 4096));
        chunk = (chunk + 31) & ~31ll;
        G.chunk = chunk;
        bool ok = true;
        for (int s = 0; s < 3 && ok; ++s) {
            ok = cudaStreamCreateWithFlags(&G.st[s], cudaStreamNonBlocking) == cudaSuccess &&
                 cudaMalloc((void **)&G.Ax[s], (size_t)chunk * nnzA * 8 + 16) == cudaSuccess &&
                 cudaMalloc((void **)&G.b[s], (size_t)chunk * n * 8 + 16) == cudaSuccess &&
                 cudaMalloc((void **)&G.x[s], (size_t)chunk * n * 8 + 16) == cudaSuccess &&
                 cudaMalloc((void **)&G.Lx[s], (size_t)csp3_lu_workspace_bytes(sym, chunk)) == cudaSuccess &&   // factor workspace
                 cudaMalloc((void **)&G.status[s], (size_t)chunk * 4 + 16) == cudaSuccess;
        }
        if (!ok) {
            set_error("lu_refactor_solve_host: staging allocation failed (%s)", cudaGetErrorString(cudaGetLastError()));
            free_stage(G);                                    // nothing half-built is kept
            return CSP3_ERR_ALLOC;
        }
        G.ready = true;
    }
    int rc = 0, slot = 0;
    for (i64 s0 = 0; s0 < batch && rc == 0; s0 += G.chunk, slot = (slot + 1) % 3) {
        const i64 cnt = std::min<i64>(G.chunk, batch - s0);
        cudaStream_t st = G.st[slot];
        auto cp = [&](void *d, const void *h, size_t bytes, cudaMemcpyKind k) {
            if (rc == 0 && cudaMemcpyAsync(d, h, bytes, k, st) != cudaSuccess) {
                set_error("lu_refactor_solve_host: copy failed (%s)", cudaGetErrorString(cudaGetLastError()));
                rc = CSP3_ERR_CUDA;
            }
        };
        cp(G.Ax[slot], Ax + s0 * nnzA, (size_t)cnt * nnzA * 8, cudaMemcpyHostToDevice);
        cp(G.b[slot], b + s0 * n, (size_t)cnt * n * 8, cudaMemcpyHostToDevice);
        if (rc == 0) rc = csp3_lu_refactor_ws(sym, cnt, G.Ax[slot], G.Lx[slot], G.status[slot], st);
        if (rc == 0) rc = csp3_lu_solve_ws(sym, cnt, G.Lx[slot], G.b[slot], G.x[slot], st);
        cp(x + s0 * n, G.x[slot], (size_t)cnt * n * 8, cudaMemcpyDeviceToHost);
        if (status) cp(status + s0, G.status[slot], (size_t)cnt * 4, cudaMemcpyDeviceToHost);
    }
    // also after an error: nothing queued on the three streams may still target the caller's buffers when we return
    for (int s = 0; s < 3; ++s)
        if (cudaStreamSynchronize(G.st[s]) != cudaSuccess && rc == 0) {
            set_error("lu_refactor_solve_host: %s", cudaGetErrorString(cudaGetLastError()));
            rc = CSP3_ERR_CUDA;
        }
    return rc;
}

int csp3_lu_refactor_host(csp3_lu_symbolic *sym, int64_t batch, const double *Ax, double *Lx, double *Ux,
                          int32_t *status)
{
    if (!sym || batch < 0 || !Ax || !Lx || !Ux) { set_error("lu_refactor_host: bad arguments"); return CSP3_ERR_ARG; }
    if (int rc = require_device()) return rc;
    if (int rc = csp3_lu_upload(sym, nullptr)) return rc;
    const size_t nA = (size_t)sym->nnzA * batch, nL = sym->F.Li.size() * batch, nU = sym->F.Ui.size() * batch;
    Dev dA, dL, dU, dS;
    CSP3_TRY(dA.put(Ax, nA * 8));
    CSP3_TRY(dL.alloc(nL * 8));
    CSP3_TRY(dU.alloc(nU * 8));
    CSP3_TRY(dS.alloc((size_t)batch * 4));
    if (int rc = csp3_lu_refactor_batched(sym, batch, dA.as<double>(), dL.as<double>(), dU.as<double>(), dS.as<i32>(), nullptr)) return rc;
    CSP3_CUDA(cudaDeviceSynchronize());
    CSP3_TRY(dL.get(Lx, nL * 8));
    CSP3_TRY(dU.get(Ux, nU * 8));
    if (status) CSP3_TRY(dS.get(status, (size_t)batch * 4));
    return 0;
}

int csp3_lu_solve_host(csp3_lu_symbolic *sym, int64_t batch, const double *Lx, const double *Ux,
                       const double *b, double *x)
{
    if (!sym || batch < 0 || !Lx || !Ux || !b || !x) { set_error("lu_solve_host: bad arguments"); return CSP3_ERR_ARG; }
    if (int rc = require_device()) return rc;
    if (int rc = csp3_lu_upload(sym, nullptr)) return rc;
    const size_t nL = sym->F.Li.size() * batch, nU = sym->F.Ui.size() * batch, nb = (size_t)sym->n * batch;
    Dev dL, dU, dB, dX;
    CSP3_TRY(dL.put(Lx, nL * 8));
    CSP3_TRY(dU.put(Ux, nU * 8));
    CSP3_TRY(dB.put(b, nb * 8));
    CSP3_TRY(dX.alloc(nb * 8));
    if (int rc = csp3_lu_solve_batched(sym, batch, dL.as<double>(), dU.as<double>(), dB.as<double>(), dX.as<double>(), nullptr)) return rc;
    CSP3_CUDA(cudaDeviceSynchronize());
    CSP3_TRY(dX.get(x, nb * 8));
    return 0;
}

int csp3_csc_lusol_host(int64_t order, int64_t n, const int32_t *Ap, const int32_t *Ai, const double *Ax,
                        double *b, double tol)
{
    if (int rc = require_device()) return rc;
    csp3_lu_symbolic *sym = nullptr;
    int rc = csp3_lu_analyze(order, n, Ap, Ai, Ax, nullptr, tol, &sym);
    if (rc) return rc;
    std::vector<double> x((size_t)std::max<i64>(n, 1));
    i32 status = 0;
    rc = csp3_lu_refactor_solve_host(sym, 1, Ax, b, x.data(), &status);
    csp3_lu_destroy(sym);
    if (rc) return rc;
    if (status) { set_error("lusol: zero or non-finite pivot in column %d", status - 1); return status; }
    std::memcpy(b, x.data(), (size_t)n * 8);
    return 0;
}

}  // extern "C"
