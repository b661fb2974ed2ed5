// common.cuh -- shared declarations of libcsparse3_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>

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

// ---- device copy of the LU schedule (one per device) ------------------------------------------------------
struct DevSchedule {
    bool ready = false;
    i32 n = 0, nnzA = 0, lnz = 0, unz = 0, max_col_len = 0;
    // refactor
    const int4 *cols = nullptr;        // 2 x int4 per column (ColDesc)
    const i32 *a_src = nullptr;
    const uint16_t *a_off = nullptr;
    const int4 *pairs = nullptr;       // PairDesc
    const uint16_t *upd_map = nullptr;
    const i32 *rf_order = nullptr, *rf_lptr = nullptr;
    i32 rf_nlev = 0;
    // solves
    const i32 *pinv = nullptr, *q = nullptr, *Up = nullptr;
    const i32 *lrow_ptr = nullptr, *lrow_col = nullptr, *lrow_pos = nullptr;
    const i32 *urow_ptr = nullptr, *urow_col = nullptr, *urow_pos = nullptr;
    const i32 *ls_order = nullptr, *ls_lptr = nullptr, *us_order = nullptr, *us_lptr = nullptr;
    i32 ls_nlev = 0, us_nlev = 0;
    void *arena = nullptr;             // single allocation backing all of the above
    size_t arena_bytes = 0;
};

// launchers (lu_kernels.cu)
int launch_refactor(const DevSchedule &D, i64 batch, const double *Ax, double *Lx, double *Ux, i32 *status,
                    cudaStream_t st);
int launch_solve(const DevSchedule &D, i64 batch, const double *Lx, const double *Ux, const double *b,
                 double *x, cudaStream_t st);

// tuning knobs (env CSP3_RF_S, CSP3_RF_WARPS, CSP3_SV_S, CSP3_SV_WARPS; 0 = automatic)
struct Tuning {
    int rf_S = 0, rf_warps = 0, sv_S = 0, sv_warps = 0;
};
Tuning &tuning();

}  // namespace csp3
