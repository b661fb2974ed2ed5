// lu_panel.cu -- "panel" batched LU refactorisation for sm_100a (program: panel_program.cpp / panel_program.hpp).
//
// One warp owns a bundle of 8 systems of the same pattern.  Lane = (row group g = lane / 8, system s = lane % 8):
// every lane works on ONE system, the four row groups take different rows of the source column a task walks.  All
// indices come from the compiled program and are uniform over the systems, every value access of a row group is
// one contiguous 64-byte run (accumulators in shared memory, factors in the bundle-interleaved workspace arrays).
//
// Left-looking elimination in panels of up to two columns; a task applies one or two source columns to the panel
// with its multipliers in registers:
//     acc_x[row] = (acc_x[row] - L(row,j) * U(j,k+x)) - L(row,j+1) * U(j+1,k+x)        x = 0, 1
// i.e. 2 L loads + 2 accumulator loads + 2 stores for 4 multiply-subtracts (the scalar formulation of lu_wide.cu
// needs 4 shared-memory accesses per multiply-subtract and ~1.2 instructions per operation and system; this one
// ~0.2).  The order of the operations on every entry is the order of cs_lu (oracle/csp3_oracle.c
// orc_csc_lu_refactor); in EXACT mode (unfused multiply / subtract, IEEE division) the factors are bit-identical
// to the oracle's, in FMA mode (CSP3_PANEL_FMA=1) they agree to rounding.
//
// Data movement: program words through ld.global.nc two steps ahead (L2 / L1 resident, shared by all bundles);
// L operands of an UPD step through plain ld.global one step ahead (L1 / L2: a column is reused by the next
// columns of its elimination-tree path); A through ld.global.nc; L and U written once.
#include "common.cuh"
#include "lu_arith.cuh"
#include "panel_program.hpp"

namespace csp3 {

namespace {

__device__ __forceinline__ double lds_f64(unsigned a)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_f64(unsigned a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ double ldg_f64(const void *p)
{
    double v;
    asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ldg_nc_f64(const void *p)
{
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void stg_f64(void *p, double v) { asm volatile("st.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }
__device__ __forceinline__ void stg_cs_f64(void *p, double v) { asm volatile("st.global.cs.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }
__device__ __forceinline__ uint2 ldg_word(const uint64_t *p)
{
    uint2 v;
    asm volatile("ld.global.nc.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}

struct PanelArgs {
    const uint64_t *prog;
    i32 nslots, nnzA, lnz, unz;
    i64 batch;
    const double *Ax;
    double *Lw, *Uw;
    i32 *status;
    double *growth;       // optional: max |L(i,j)| per system (pivot-growth indicator)
};

template <bool EXACT>
__device__ __forceinline__ double fnma(double x, double l, double u)
{
    return EXACT ? __dsub_rn(x, __dmul_rn(l, u)) : __fma_rn(-l, u, x);
}

// word fields (panel_program.hpp)
__device__ __forceinline__ unsigned w_op(uint2 w) { return w.y >> 28; }
__device__ __forceinline__ unsigned w_flags(uint2 w) { return (w.y >> 21) & 0x7fu; }
__device__ __forceinline__ unsigned w_c64(uint2 w) { return ((w.y >> 8) & 0x1fffu) << 6; }                 // byte offset of entry c
__device__ __forceinline__ unsigned w_a64(uint2 w) { return (w.x & 0xfffffu) << 6; }
__device__ __forceinline__ unsigned w_b(uint2 w) { return (w.x >> 20) | ((w.y & 0xffu) << 12); }
__device__ __forceinline__ size_t w_ab(uint2 w) { return (size_t)w.x | ((size_t)(w.y & 0xffu) << 32); }

template <bool EXACT>
__global__ void __launch_bounds__(32) lu_refactor_panel_kernel(const PanelArgs a)
{
    constexpr int S = 8;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int lane = threadIdx.x, g = lane >> 3, s = lane & 7;
    const i64 b = blockIdx.x;
    const i64 sys = b * S + s;
    const char *Axs = reinterpret_cast<const char *>(a.Ax + (sys < a.batch ? sys : a.batch - 1) * a.nnzA);
    char *Lb = reinterpret_cast<char *>(a.Lw + (size_t)b * a.lnz * S) + s * 8;
    char *Ub = reinterpret_cast<char *>(a.Uw + (size_t)b * a.unz * S) + s * 8;
    unsigned accb = (unsigned)__cvta_generic_to_shared(smem_raw) + s * 8;
    const unsigned NSB = (unsigned)a.nslots * 64u;                      // acc1 = acc0 + NSB
    for (int t = lane; t < 2 * a.nslots * S; t += 32) reinterpret_cast<double *>(smem_raw)[t] = 0.0;
    __syncwarp();

    const uint64_t *pp = a.prog + g;
    uint2 w0 = ldg_word(pp), w1 = ldg_word(pp + 4);
    pp += 8;
    double u00 = 0.0, u01 = 0.0, u10 = 0.0, u11 = 0.0, piv = 1.0, rcp = 1.0, uk1 = 0.0, lmax = 0.0;
    int fail = INT32_MAX;
    double la0 = 0.0, la1 = 0.0, lb0 = 0.0, lb1 = 0.0;               // L operands of UPD steps, two sets in ping-pong

    // one step: `cur` is executed with the L operands (c0, c1); the L operands of the next step go to (n0, n1)
    auto step = [&](double &c0, double &c1, double &n0, double &n1) -> bool {
        const uint2 cur = w0;
        w0 = w1;
        w1 = ldg_word(pp);
        pp += 4;
        const unsigned op = w_op(cur), fl = w_flags(cur);
        {   // L operands of the next step (it cannot be a step that reads what this one writes: the compiler
            // separates FINL from UPD with at least one other step)
            const unsigned nfl = w_flags(w0);
            if (w_op(w0) == (unsigned)kPanelUpd && (nfl & kPanelValid)) {
                n0 = ldg_f64(Lb + w_a64(w0));
                if (nfl & kPanelWS2) n1 = ldg_f64(Lb + ((size_t)w_b(w0) << 6));
            }
        }
        if (op == (unsigned)kPanelUpd) {
            if (fl & kPanelValid) {
                const unsigned t = accb + w_c64(cur);
                if ((fl & (kPanelM0 | kPanelM1)) == (kPanelM0 | kPanelM1)) {
                    double x0 = lds_f64(t), x1 = lds_f64(t + NSB);
                    x0 = fnma<EXACT>(x0, c0, u00); x1 = fnma<EXACT>(x1, c0, u01);
                    if (fl & kPanelWS2) { x0 = fnma<EXACT>(x0, c1, u10); x1 = fnma<EXACT>(x1, c1, u11); }
                    sts_f64(t, x0); sts_f64(t + NSB, x1);
                } else if (fl & kPanelM0) {
                    double x0 = lds_f64(t);
                    x0 = fnma<EXACT>(x0, c0, u00);
                    if (fl & kPanelWS2) x0 = fnma<EXACT>(x0, c1, u10);
                    sts_f64(t, x0);
                } else {
                    double x1 = lds_f64(t + NSB);
                    x1 = fnma<EXACT>(x1, c0, u01);
                    if (fl & kPanelWS2) x1 = fnma<EXACT>(x1, c1, u11);
                    sts_f64(t + NSB, x1);
                }
            }
        } else if (op == (unsigned)kPanelLoadU) {
            const unsigned tj = accb + w_c64(cur);
            if (fl & kPanelM0) u00 = lds_f64(tj);
            if (fl & kPanelM1) u01 = lds_f64(tj + NSB);
            if (fl & kPanelWS2) {
                const double l = ldg_f64(Lb + w_a64(cur));
                const unsigned tj1 = accb + (w_b(cur) << 6);
                if (fl & kPanelM0) { u10 = fnma<EXACT>(lds_f64(tj1), l, u00); if (g == 0) sts_f64(tj1, u10); }
                if (fl & kPanelM1) { u11 = fnma<EXACT>(lds_f64(tj1 + NSB), l, u01); if (g == 0) sts_f64(tj1 + NSB, u11); }
            }
        } else if (op == (unsigned)kPanelScatter) {
            if (fl & kPanelValid) sts_f64(accb + w_c64(cur), ldg_nc_f64(Axs + w_ab(cur) * 8));
        } else if (op == (unsigned)kPanelFinU) {
            if (fl & kPanelValid) {
                const unsigned t = accb + w_c64(cur);
                const double v = lds_f64(t);
                sts_f64(t, 0.0);
                stg_cs_f64(Ub + w_ab(cur) * 64, v);
            }
        } else if (op == (unsigned)kPanelFinL) {
            if (fl & kPanelValid) {
                const unsigned t = accb + w_c64(cur);
                const double x = lds_f64(t);
                sts_f64(t, 0.0);
                const double qv = EXACT ? div_shared(x, piv, rcp) : x * rcp;
                stg_f64(Lb + w_ab(cur) * 64, qv);
                lmax = fmax(lmax, fabs(qv));
                if (fl & kPanelFused) {
                    const double y = lds_f64(t + NSB);
                    sts_f64(t + NSB, fnma<EXACT>(y, qv, uk1));
                }
            }
        } else if (op == (unsigned)kPanelPiv) {
            const unsigned t = accb + w_c64(cur);
            piv = lds_f64(t);
            rcp = rcp_refined(piv);
            if (!(fabs(piv) > 0.0 && isfinite(piv))) fail = min(fail, (int)cur.x);
            if (fl & kPanelFused) uk1 = lds_f64(t + NSB);
        } else if (op == (unsigned)kPanelEnd) {
            return false;
        }
        __syncwarp();
        return true;
    };
#pragma unroll 1
    for (;;) {
        if (!step(la0, la1, lb0, lb1)) break;
        if (!step(lb0, lb1, la0, la1)) break;
    }
    fail = min(fail, __shfl_xor_sync(0xffffffffu, fail, 8));
    fail = min(fail, __shfl_xor_sync(0xffffffffu, fail, 16));
    lmax = fmax(lmax, __shfl_xor_sync(0xffffffffu, lmax, 8));
    lmax = fmax(lmax, __shfl_xor_sync(0xffffffffu, lmax, 16));
    if (g == 0 && sys < a.batch) {
        if (a.status != nullptr) a.status[sys] = (fail == INT32_MAX) ? 0 : fail;
        if (a.growth != nullptr) a.growth[sys] = lmax;
    }
}

}  // namespace

int launch_refactor_panel(const DevSchedule &D, i64 batch, const double *Ax, double *Lw, double *Uw, i32 *status,
                          double *growth, cudaStream_t st)
{
    if (batch <= 0) return 0;
    if (!D.panel_ok) { set_error("panel refactor program not available for this pattern"); return -1; }
    PanelArgs a;
    a.prog = reinterpret_cast<const uint64_t *>(D.prf_prog);
    a.nslots = D.prf_nslots; a.nnzA = D.nnzA; a.lnz = D.lnz; a.unz = D.unz;
    a.batch = batch; a.Ax = Ax; a.Lw = Lw; a.Uw = Uw; a.status = status; a.growth = growth;
    const i64 grid = (batch + 7) / 8;
    const size_t smem = D.prf_smem;
    if (tuning().panel_fma) {
        CSP3_CUDA(cudaFuncSetAttribute(lu_refactor_panel_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        lu_refactor_panel_kernel<false><<<(unsigned)grid, 32, smem, st>>>(a);
    } else {
        CSP3_CUDA(cudaFuncSetAttribute(lu_refactor_panel_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        lu_refactor_panel_kernel<true><<<(unsigned)grid, 32, smem, st>>>(a);
    }
    CSP3_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace csp3
