// common.cuh -- shared declarations of libcsparse3_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <string>

#include "symbolic.hpp"

namespace csp3 {

void set_error(const char *fmt, ...);
const char *get_error();

#define CSP3_CUDA(call)                                                                          \
    do {                                                                                         \
        cudaError_t err__ = (call);                                                              \
        if (err__ != cudaSuccess) {                                                              \
            csp3::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(err__)); \
            return -2;                                                                           \
        }                                                                                        \
    } while (0)

constexpr int kNumSMs = 148;   // B200
// row-lane refactor program geometries (warps per bundle, quads per stage); the last one is the CSP3_RL_W / CSP3_RL_NQ override
constexpr int kRlVariants = 5;
constexpr int kRlGeometry[kRlVariants][2] = {{1, 3}, {2, 2}, {4, 2}, {8, 1}, {0, 0}};

// ---- device copy of the LU schedule (one per device) ------------------------------------------------------
struct DevSchedule {
    bool ready = false;
    i32 n = 0, nnzA = 0, lnz = 0, unz = 0, max_col_len = 0;
    // compiled programs (program.hpp), shared by every bundle
    const uint8_t *rf_prog = nullptr, *ls_prog = nullptr, *us_prog = nullptr, *ur_prog = nullptr;
    i32 rf_prog_bytes = 0, rf_prog_stage = 0, ls_prog_bytes = 0, ls_prog_stage = 0, us_prog_bytes = 0, us_prog_stage = 0;
    i32 ur_prog_bytes = 0, ur_prog_stage = 0;
    i32 ls_nslots = 0, us_nslots = 0, ur_nslots = 0, ur_max_len = 0;
    // wide (lane = system) refactor program, lu_wide.cu
    bool wide_ok = false;
    i32 wide_S = 0, wide_R = 2, wrf_groups = 0, wrf_prog_bytes = 0, wrf_prog_stage = 0, wrf_acc_slots = 0, wrf_lsrc_entries = 0;
    const uint8_t *wrf_prog = nullptr;
    size_t wrf_smem = 0;
    // wide triangular sweeps (forward / backward)
    bool wide_solve_ok = false;
    const uint8_t *wfs_prog = nullptr, *wbs_prog = nullptr;
    i32 wfs_prog_bytes = 0, wfs_prog_stage = 0, wfs_records = 0, wfs_nslots = 0;
    i32 wbs_prog_bytes = 0, wbs_prog_stage = 0, wbs_records = 0, wbs_nslots = 0;
    size_t wfs_smem = 0, wbs_smem = 0;
    const i32 *d_pinv = nullptr, *d_qinv = nullptr;
    const uint8_t *d_ldiag = nullptr;     // 1 at the unit-diagonal positions of L (never written in the workspace layout)
    // panel refactor program (lu_panel.cu), bundles of 8 systems
    bool panel_ok = false;
    const uint8_t *prf_prog = nullptr;
    i32 prf_prog_bytes = 0, prf_nslots = 0, prf_steps = 0, prf_lsrc = 0;
    size_t prf_smem = 0;
    // row-lane refactor program (lu_rowlane.cu), bundles of 8 systems
    // (one program per geometry: warps per bundle x quads per stage, compiled and uploaded on first use)
    struct RlVariant {
        bool ok = false;
        std::atomic<bool> tried{false};    // set (release) after ok / prog / ... are final: readers that see it need no lock
        uint8_t *prog = nullptr;           // own device allocation
        i32 quads = 0, nslots = 0, warps = 1, stage_quads = 3, stream_off[8] = {};
        size_t smem = 0;
    };
    mutable RlVariant rl[kRlVariants];
    // row-sweep programs (lu_sweep_rows_kernel in lu_wide.cu): the sweeps of a small batch on 8 warps per bundle
    struct RsPrograms {
        bool ok = false;
        std::atomic<bool> tried{false};
        uint32_t *prog_f = nullptr, *prog_b = nullptr;     // own device allocations
        i32 levels_f = 0, levels_b = 0, stream_off_f[8] = {}, stream_off_b[8] = {};
    };
    mutable RsPrograms rs;
    bool rl_enabled = false;               // the sweeps this kernel's factor layout needs are available
    void *owner = nullptr;                 // csp3_lu_symbolic the schedule belongs to (lazy compilation of the variants)
    int devid = 0;
    void *arena = nullptr;             // single allocation backing all of the above
    size_t arena_bytes = 0;
};

// launchers (lu_kernels.cu).  interleaved == false: Lx/Ux are system-major [batch][lnz|unz] (API layout);
// interleaved == true: Lx/Ux/z point into a workspace in bundle-interleaved layout (see lu_kernels.cu).
int launch_refactor(const DevSchedule &D, i64 batch, const double *Ax, double *Lx, double *Ux, i32 *status,
                    bool interleaved, cudaStream_t st);
// z: interleaved scratch of 2 * n doubles per (padded) system
int launch_solve(const DevSchedule &D, i64 batch, const double *Lx, const double *Ux, const double *b,
                 double *x, double *z, bool interleaved, cudaStream_t st);
int workspace_bundle_width(const DevSchedule &D, i64 batch);
bool use_wide(const DevSchedule &D, i64 batch);
int launch_refactor_wide(const DevSchedule &D, i64 batch, const double *Ax, double *Lw, double *Uw, i32 *status,
                         cudaStream_t st);
bool use_panel(const DevSchedule &D, i64 batch);
bool use_tmem(const DevSchedule &D, i64 batch);
int launch_refactor_tmem(const DevSchedule &D, i64 batch, const double *Ax, double *Lw, double *Uw, i32 *status, cudaStream_t st);
int launch_growth(const DevSchedule &D, i64 batch, const double *Lw, double *growth, cudaStream_t st);
int launch_refactor_panel(const DevSchedule &D, i64 batch, const double *Ax, double *Lw, double *Uw, i32 *status,
                          double *growth, cudaStream_t st);
bool use_rowlane(const DevSchedule &D, i64 batch);
int rowlane_variant(const DevSchedule &D, i64 batch);      // index into DevSchedule::rl (compiled and uploaded), or -1: not the row-lane kernel
int ensure_rowlane_variant(const DevSchedule &D, int variant);   // api.cu: compiles / uploads on first use; 0 when available
int launch_refactor_rowlane(const DevSchedule &D, i64 batch, const double *Ax, double *Lw, double *Uw, i32 *status, cudaStream_t st);
bool use_rowsweep(const DevSchedule &D, i64 batch);
int ensure_rowsweep(const DevSchedule &D);                   // api.cu: compiles / uploads on first use; 0 when available
int launch_solve_rows(const DevSchedule &D, i64 batch, const double *Lw, const double *Uw, const double *b, double *x, double *z1, cudaStream_t st);
int launch_solve_wide(const DevSchedule &D, i64 batch, const double *Lw, const double *Uw, const double *b, double *x,
                      double *z1, double *z2, cudaStream_t st);

// tuning knobs (env CSP3_RF_S, CSP3_SV_S: bundle width of the system-major kernels; CSP3_WS_S: bundle width of
// the workspace path; CSP3_RF_WIN: entries of the recent-L ring; CSP3_SV_STAGE: entries per cp.async stage;
// 0 = automatic)
struct Tuning {
    int rf_S = 0, sv_S = 0, ws_S = 0, rf_win = 0, sv_stage = 0;
    int tmem = 0;                          // CSP3_TMEM=1: experimental refactor kernel with the accumulator in tensor memory (lu_refactor_tmem_kernel)
    int panel = 0, panel_fma = 0, panel_budget = 0;          // CSP3_PANEL=1 selects the experimental panel refactor kernel (lu_panel.cu), CSP3_PANEL_FMA (fused multiply-add, not bit-exact)
    int rowsweep = 0;                      // CSP3_ROWSWEEP=1: row-oriented sweeps on 8 warps per bundle for small batches (lu_sweep_rows_kernel)
    int rowlane = -1, rl_warps = 0, rl_nq = 0;        // row-lane refactor kernel (lu_rowlane.cu): -1 automatic (patterns whose wide program leaves < 5 bundles per SM), CSP3_ROWLANE=0 / 1 never / always; CSP3_RL_W warps per bundle (1, 2, 4, 8), CSP3_RL_NQ quads per stage
    int wide = 1, wide_solve = 1, wide_S = 0, wide_R = 0, wide_ring = 0, wide_stage = 0, wide_budget = 0;   // CSP3_WIDE (0 disables), CSP3_WIDE_S/_R/_F/_BUDGET
};
Tuning &tuning();

}  // namespace csp3
