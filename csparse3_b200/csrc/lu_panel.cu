// lu_panel.cu -- "panel" batched LU refactorisation for sm_100a (program: panel_program.cpp / panel_program.hpp).
//
// One warp owns a bundle of 8 systems of the same pattern.  Lane = (row group g = lane / 4, h = lane % 4): every
// lane carries the two adjacent systems 2h, 2h + 1 (16-byte accesses), the eight row groups take different words
// (rows) of a step.  All indices come from the compiled program and are uniform over the systems.
//
// Left-looking elimination in panels of up to two columns; a task applies one or two source columns to the panel
// with its multipliers in registers:
//     acc_x[row] = (acc_x[row] - L(row,j) * U(j,k+x)) - L(row,j+1) * U(j+1,k+x)        x = 0, 1
// i.e. 2 L operands + 2 accumulator loads + 2 stores for 4 multiply-subtracts (the scalar formulation of lu_wide.cu
// needs 4 shared-memory accesses and ~1 warp instruction per multiply-subtract and system; this one ~0.2).  The
// order of the operations on every entry is the order of cs_lu (oracle/csp3_oracle.c orc_csc_lu_refactor); in EXACT
// mode (unfused multiply / subtract, IEEE division) the factors are bit-identical to the oracle's, in FMA mode
// (CSP3_PANEL_FMA=1) they agree to rounding.
//
// Data movement: nothing on the dependent chain waits for global memory.  The program, the A values and the L
// operands of columns that have left the shared-memory ring are copied with cp.async kPanelLookahead steps before
// their use (one commit group per step); recently finalised L columns are read from the ring; L and U are written once.
#include "common.cuh"
#include "lu_arith.cuh"
#include "ptx.cuh"
#include "panel_program.hpp"

namespace csp3 {

namespace {

using namespace ptx;

struct PanelArgs {
    const uint8_t *prog;
    i32 prog_bytes, nslots, lsrc_entries, nnzA, lnz, unz;
    i64 batch;
    const double *Ax;
    double *Lw, *Uw;
    i32 *status;
    double *growth;       // optional: max |L(i,j)| per system (pivot-growth indicator)
};

template <bool EXACT>
__device__ __forceinline__ double2 fnma2(double2 x, double2 l, double2 u)
{
    return EXACT ? make_double2(__dsub_rn(x.x, __dmul_rn(l.x, u.x)), __dsub_rn(x.y, __dmul_rn(l.y, u.y)))
                 : make_double2(__fma_rn(-l.x, u.x, x.x), __fma_rn(-l.y, u.y, x.y));
}

// word fields (panel_program.hpp); a word is (lo, hi)
__device__ __forceinline__ unsigned w_op(unsigned hi) { return hi >> 28; }
__device__ __forceinline__ unsigned w_flags(unsigned hi) { return (hi >> 21) & 0x7fu; }
__device__ __forceinline__ unsigned w_c64(unsigned hi) { return ((hi >> 8) & 0x1fffu) << 6; }     // byte offset of entry c
__device__ __forceinline__ unsigned w_a64(unsigned lo) { return (lo & 0xfffffu) << 6; }
__device__ __forceinline__ unsigned w_b64(unsigned lo, unsigned hi) { return ((lo >> 20) | ((hi & 0xffu) << 12)) << 6; }
__device__ __forceinline__ size_t w_ab(unsigned lo, unsigned hi) { return (size_t)lo | ((size_t)(hi & 0xffu) << 32); }

template <bool EXACT>
__global__ void __launch_bounds__(32) lu_refactor_panel_kernel(const PanelArgs a)
{
    constexpr int S = 8, D = kPanelLookahead;
    constexpr unsigned STEP_BYTES = kPanelStepWords * 8, STAGE_BYTES = kPanelStageSteps * STEP_BYTES,
                       RING_BYTES = kPanelProgStages * STAGE_BYTES;
    static_assert(STAGE_BYTES == 512, "one 16-byte piece per lane and stage");
    static_assert((kPanelProgStages - 2) * kPanelStageSteps >= D, "program stages must be requested kPanelLookahead steps ahead");
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int lane = threadIdx.x, g = lane >> 2, h = lane & 3;
    const i64 b = blockIdx.x;
    const i64 sys0 = b * S + 2 * h, sys1 = sys0 + 1;
    const char *Axs0 = reinterpret_cast<const char *>(a.Ax + (sys0 < a.batch ? sys0 : a.batch - 1) * a.nnzA);
    const char *Axs1 = reinterpret_cast<const char *>(a.Ax + (sys1 < a.batch ? sys1 : a.batch - 1) * a.nnzA);
    char *Lb = reinterpret_cast<char *>(a.Lw + (size_t)b * a.lnz * S) + h * 16;
    char *Ub = reinterpret_cast<char *>(a.Uw + (size_t)b * a.unz * S) + h * 16;
    const unsigned smem_s = (unsigned)__cvta_generic_to_shared(smem_raw);
    const unsigned accb = smem_s + h * 16;                              // accumulator entry e: accb + 64 e
    const unsigned NSB = (unsigned)a.nslots * 64u;                      // acc1 = acc0 + NSB
    const unsigned lsb = accb + 2 * NSB;                                // lsrc entry e (ring, then landing): lsb + 64 e
    const unsigned ring_s = smem_s + (2u * a.nslots + a.lsrc_entries) * 64u;    // program ring
    for (int t = lane; t < (2 * a.nslots + a.lsrc_entries) * S; t += 32) reinterpret_cast<double *>(smem_raw)[t] = 0.0;

    // program ring: stage s lives in slot s % kPanelProgStages; kPanelProgStages - 1 stages are loaded up front, then
    // one more is requested every time a stage is entered (it rides in that step's commit group)
    const uint8_t *pnext = a.prog + lane * 16;                          // next stage to request, at this lane's piece
    const uint8_t *pend = a.prog + a.prog_bytes;
    unsigned pdst = 0;                                                  // ring offset the next stage goes to
    auto request_stage = [&]() {
        if (pnext < pend) cp_async16(ring_s + pdst + lane * 16, pnext);
        pnext += STAGE_BYTES;
        pdst = (pdst + STAGE_BYTES == RING_BYTES) ? 0u : pdst + STAGE_BYTES;
    };
#pragma unroll
    for (int t = 0; t < kPanelProgStages - 1; ++t) request_stage();
    cp_async_commit();
    cp_async_wait<0>();
    __syncwarp();

    unsigned rp = 0;                                                    // ring offset of the current step
    uint4 nxt = lds_u4(ring_s + g * 16);
    double2 u00 = make_double2(0.0, 0.0), u01 = u00, u10 = u00, u11 = u00, uk1 = u00;
    double2 piv = make_double2(1.0, 1.0), rcp = piv;
    double lmax0 = 0.0, lmax1 = 0.0;
    int fail0 = INT32_MAX, fail1 = INT32_MAX;
    unsigned tfl = 0;                                                   // flags of the current task
    const double2 zero2 = make_double2(0.0, 0.0);

#pragma unroll 1
    for (;;) {
        const uint4 cur = nxt;
        const unsigned op0 = w_op(cur.y), op1 = w_op(cur.w);
        // ---- fetches of this step, program stage, one commit group ------------------------------------------------
        if (op0 == (unsigned)kPanelFetchL) cp_async16(lsb + w_c64(cur.y), Lb + w_ab(cur.x, cur.y) * 64);
        else if (op0 == (unsigned)kPanelFetchA) {
            const size_t o = w_ab(cur.x, cur.y) * 8;
            cp_async8(lsb + w_c64(cur.y), Axs0 + o); cp_async8(lsb + w_c64(cur.y) + 8, Axs1 + o);
        }
        if (op1 == (unsigned)kPanelFetchL) cp_async16(lsb + w_c64(cur.w), Lb + w_ab(cur.z, cur.w) * 64);
        else if (op1 == (unsigned)kPanelFetchA) {
            const size_t o = w_ab(cur.z, cur.w) * 8;
            cp_async8(lsb + w_c64(cur.w), Axs0 + o); cp_async8(lsb + w_c64(cur.w) + 8, Axs1 + o);
        }
        if ((rp & (STAGE_BYTES - 1)) == 0) request_stage();
        cp_async_commit();
        cp_async_wait<D>();
        __syncwarp();
        rp = (rp + STEP_BYTES == RING_BYTES) ? 0u : rp + STEP_BYTES;
        nxt = lds_u4(ring_s + rp + g * 16);                             // words of the next step
        // ---- header (word 0 of the step, executed by all lanes) ----------------------------------------------------
        const unsigned hx = __shfl_sync(0xffffffffu, cur.x, 0), hy = __shfl_sync(0xffffffffu, cur.y, 0);
        const unsigned hop = w_op(hy);
        if (hop == (unsigned)kPanelHdrU) {
            tfl = w_flags(hy);
            const unsigned tj = accb + w_c64(hy);
            if (tfl & kPanelM0) u00 = lds_d2(tj);
            if (tfl & kPanelM1) u01 = lds_d2(tj + NSB);
            if (tfl & kPanelWS2) {
                const double2 l = (tfl & kPanelX) ? ldg_d2(Lb + (size_t)w_a64(hx)) : lds_d2(lsb + w_a64(hx));
                const unsigned tj1 = accb + w_b64(hx, hy);
                if (tfl & kPanelM0) { u10 = fnma2<EXACT>(lds_d2(tj1), l, u00); if (g == 0) sts_d2(tj1, u10); }
                if (tfl & kPanelM1) { u11 = fnma2<EXACT>(lds_d2(tj1 + NSB), l, u01); if (g == 0) sts_d2(tj1 + NSB, u11); }
            }
        } else if (hop == (unsigned)kPanelHdrP) {
            const unsigned t = accb + w_c64(hy);
            piv = lds_d2(t);
            rcp = make_double2(rcp_refined(piv.x), rcp_refined(piv.y));
            if (!(fabs(piv.x) > 0.0 && isfinite(piv.x))) fail0 = min(fail0, (int)hx);
            if (!(fabs(piv.y) > 0.0 && isfinite(piv.y))) fail1 = min(fail1, (int)hx);
            if (w_flags(hy) & kPanelFused) uk1 = lds_d2(t + NSB);
        } else if (hop == (unsigned)kPanelEnd) {
            break;
        }
        // ---- this lane's two words ------------------------------------------------------------------------------------
        if (op0 == (unsigned)kPanelUpd || op1 == (unsigned)kPanelUpd) {
            // loads of both words first, then the arithmetic and the stores (the two rows are different entries)
            const bool v0 = op0 == (unsigned)kPanelUpd, v1 = op1 == (unsigned)kPanelUpd;
            const bool m0 = tfl & kPanelM0, m1 = tfl & kPanelM1, ws2 = tfl & kPanelWS2;
            const unsigned t0 = accb + w_c64(cur.y), t1 = accb + w_c64(cur.w);
            double2 la0 = zero2, la1 = zero2, lb0 = zero2, lb1 = zero2, xa0 = zero2, xa1 = zero2, xb0 = zero2, xb1 = zero2;
            if (v0) {
                const unsigned f = w_flags(cur.y);
                la0 = (f & kPanelX) ? ldg_d2(Lb + (size_t)w_a64(cur.x)) : lds_d2(lsb + w_a64(cur.x));
                if (ws2) la1 = (f & kPanelY) ? ldg_d2(Lb + (size_t)w_b64(cur.x, cur.y)) : lds_d2(lsb + w_b64(cur.x, cur.y));
                if (m0) xa0 = lds_d2(t0);
                if (m1) xa1 = lds_d2(t0 + NSB);
            }
            if (v1) {
                const unsigned f = w_flags(cur.w);
                lb0 = (f & kPanelX) ? ldg_d2(Lb + (size_t)w_a64(cur.z)) : lds_d2(lsb + w_a64(cur.z));
                if (ws2) lb1 = (f & kPanelY) ? ldg_d2(Lb + (size_t)w_b64(cur.z, cur.w)) : lds_d2(lsb + w_b64(cur.z, cur.w));
                if (m0) xb0 = lds_d2(t1);
                if (m1) xb1 = lds_d2(t1 + NSB);
            }
            if (v0) {
                if (m0) { xa0 = fnma2<EXACT>(xa0, la0, u00); if (ws2) xa0 = fnma2<EXACT>(xa0, la1, u10); sts_d2(t0, xa0); }
                if (m1) { xa1 = fnma2<EXACT>(xa1, la0, u01); if (ws2) xa1 = fnma2<EXACT>(xa1, la1, u11); sts_d2(t0 + NSB, xa1); }
            }
            if (v1) {
                if (m0) { xb0 = fnma2<EXACT>(xb0, lb0, u00); if (ws2) xb0 = fnma2<EXACT>(xb0, lb1, u10); sts_d2(t1, xb0); }
                if (m1) { xb1 = fnma2<EXACT>(xb1, lb0, u01); if (ws2) xb1 = fnma2<EXACT>(xb1, lb1, u11); sts_d2(t1 + NSB, xb1); }
            }
        } else {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const unsigned lo = half ? cur.z : cur.x, hi = half ? cur.w : cur.y;
                const unsigned op = w_op(hi), fl = w_flags(hi);
                if (op == (unsigned)kPanelFinL) {
                    const unsigned t = accb + w_c64(hi);
                    const double2 x = lds_d2(t);
                    sts_d2(t, zero2);
                    const double2 qv = EXACT ? make_double2(div_shared(x.x, piv.x, rcp.x), div_shared(x.y, piv.y, rcp.y))
                                             : make_double2(x.x * rcp.x, x.y * rcp.y);
                    stg_d2(Lb + (size_t)(lo & 0xffffffu) * 64, qv);
                    if (fl & kPanelX) sts_d2(lsb + (((lo >> 24) | ((hi & 0xffu) << 8)) << 6), qv);
                    lmax0 = fmax(lmax0, fabs(qv.x)); lmax1 = fmax(lmax1, fabs(qv.y));
                    if (fl & kPanelFused) {
                        const double2 y = lds_d2(t + NSB);
                        sts_d2(t + NSB, fnma2<EXACT>(y, qv, uk1));
                    }
                } else if (op == (unsigned)kPanelFinU) {
                    const unsigned t = accb + w_c64(hi);
                    const double2 v = lds_d2(t);
                    sts_d2(t, zero2);
                    stg_cs_d2(Ub + w_ab(lo, hi) * 64, v);
                } else if (op == (unsigned)kPanelScatter) {
                    double2 v;
                    if (fl & kPanelX) { const size_t o = w_ab(lo, hi) * 8; v = make_double2(ldg_nc_f64(Axs0 + o), ldg_nc_f64(Axs1 + o)); }
                    else v = lds_d2(lsb + ((unsigned)lo << 6));
                    sts_d2(accb + w_c64(hi), v);
                }
            }
        }
        __syncwarp();
    }
    cp_async_wait<0>();
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) {
        fail0 = min(fail0, __shfl_xor_sync(0xffffffffu, fail0, o));
        fail1 = min(fail1, __shfl_xor_sync(0xffffffffu, fail1, o));
        lmax0 = fmax(lmax0, __shfl_xor_sync(0xffffffffu, lmax0, o));
        lmax1 = fmax(lmax1, __shfl_xor_sync(0xffffffffu, lmax1, o));
    }
    if (g == 0) {
        if (sys0 < a.batch) {
            if (a.status != nullptr) a.status[sys0] = (fail0 == INT32_MAX) ? 0 : fail0;
            if (a.growth != nullptr) a.growth[sys0] = lmax0;
        }
        if (sys1 < a.batch) {
            if (a.status != nullptr) a.status[sys1] = (fail1 == INT32_MAX) ? 0 : fail1;
            if (a.growth != nullptr) a.growth[sys1] = lmax1;
        }
    }
}

}  // namespace

int launch_refactor_panel(const DevSchedule &D, i64 batch, const double *Ax, double *Lw, double *Uw, i32 *status,
                          double *growth, cudaStream_t st)
{
    if (batch <= 0) return 0;
    if (!D.panel_ok) { set_error("panel refactor program not available for this pattern"); return -1; }
    PanelArgs a;
    a.prog = D.prf_prog; a.prog_bytes = D.prf_prog_bytes;
    a.nslots = D.prf_nslots; a.lsrc_entries = D.prf_lsrc; a.nnzA = D.nnzA; a.lnz = D.lnz; a.unz = D.unz;
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
