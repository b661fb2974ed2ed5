// lu_rowlane.cu -- "row-lane" batched LU refactorisation for sm_100a (program: rowlane_program.cpp / rowlane_program.hpp).
//
// One warp owns a bundle of 8 systems of the same pattern.  Lane = (g = lane / 4, h = lane % 4): every lane carries the
// two adjacent systems 2h, 2h + 1 (16-byte accesses), the eight lane groups g take the eight operations of a record.
// Left-looking elimination, one column at a time: the column's accumulator is the only thing in shared memory
// (nslots x 64 bytes per warp); the L operands of a record are loaded from the bundle's factor array (L1 / L2 / HBM)
// straight into registers, a whole program stage (4 quads of 4 records) at a time and one stage ahead; a quad is
// homogeneous, so there is one dispatch per quad and its records run as straight-line predicated code; the program
// goes through a cp.async ring in shared memory three stages ahead, so a multiply-subtract costs one
// accumulator load and one store in shared memory and the dependent chain of a column (multiplier -> update ->
// next multiplier) never waits for global memory.  The order of the operations on every entry is the order of cs_lu
// (oracle/csp3_oracle.c orc_csc_lu_refactor); multiply / subtract unfused, IEEE division: bit-identical factors.
#include "common.cuh"
#include "lu_arith.cuh"
#include "ptx.cuh"
#include "rowlane_program.hpp"

namespace csp3 {

namespace {

using namespace ptx;

struct RowlaneArgs {
    const uint8_t *prog;      // quads of 304 bytes, one padded stream per warp of the bundle (rowlane_program.hpp)
    i32 stream_off[kRlMaxWarps];   // first quad of every stream
    i32 nslots, nnzA, lnz, unz;
    i64 batch;
    const double *Ax;
    double *Lw, *Uw;
    i32 *status;
};

// word fields (rowlane_program.hpp)
__device__ __forceinline__ unsigned f_slot64(unsigned w) { return w & 0xffc0u; }
__device__ __forceinline__ unsigned f_off64(unsigned w) { return (w >> 10) & 0x1fffc0u; }
__device__ __forceinline__ unsigned f_off8(unsigned w) { return (w >> 13) & 0x3fff8u; }
__device__ __forceinline__ bool f_valid(unsigned w) { return (int)w < 0; }
__device__ __forceinline__ unsigned word_of(const uint4 &v, int r) { return r == 0 ? v.x : r == 1 ? v.y : r == 2 ? v.z : v.w; }

// acc[slot(w)] -= l * m for this lane's two systems; branch-free: the load is unconditional (an empty lane word
// carries a readable slot), the store is predicated on the valid bit.  Multiply and subtract are not fused.
__device__ __forceinline__ void update_op(unsigned w, unsigned accb, const double2 &l, const double2 &m)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .f64 x0, x1, t0, t1;\n\t.reg .u32 a;\n\t"
        "setp.lt.s32 p, %0, 0;\n\t"
        "and.b32 a, %0, 0xffc0;\n\tadd.u32 a, a, %1;\n\t"
        "ld.shared.v2.f64 {x0, x1}, [a];\n\t"
        "mul.rn.f64 t0, %2, %4;\n\tmul.rn.f64 t1, %3, %5;\n\t"
        "sub.rn.f64 x0, x0, t0;\n\tsub.rn.f64 x1, x1, t1;\n\t"
        "@p st.shared.v2.f64 [a], {x0, x1};\n\t}"
        ::"r"(w), "r"(accb), "d"(l.x), "d"(l.y), "d"(m.x), "d"(m.y) : "memory");
}
// m = acc entry at shared address `addr` when `flag` is set (predicated, no branch)
__device__ __forceinline__ void load_multiplier(double2 &m, unsigned addr, unsigned flag)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\t@p ld.shared.v2.f64 {%0, %1}, [%2];\n\t}"
                 : "+d"(m.x), "+d"(m.y) : "r"(addr), "r"(flag) : "memory");
}

// One pass of the main loop = one program stage of NQ quads.  All global loads of a warp share ONE scoreboard slot
// (ptxas assigns every LDG to the same slot), so a wait for any load waits for all outstanding ones: the operands of a
// WHOLE stage are requested at once, one stage before they are used.  Two register sets alternate: at the top of a pass
// the set of this stage is complete (its loads are the only outstanding ones; one dummy use makes the wait happen
// there), then the loads of the next stage are issued into the other set, then the quads execute.
//
// W warps of a CTA share a bundle: every warp runs its own stream of the program (its own columns, accumulator and
// program ring).  A warp publishes the number of columns it has finished in shared memory at the top of a pass (no load
// is outstanding there, so the fence in front of the flag is cheap) and reads the other warps' counters before it
// requests the operands of a stage whose sources belong to them: ahead of time when they are already final, otherwise
// when the stage is reached (everything before it has been executed and published by then: no deadlock).
template <int W, int NQ>
// (resident CTAs per SM the selection rule of lu_kernels.cu::rowlane_variant counts on: 8 warps x one-quad stages: 2,
// 4 warps x two-quad stages: 3; 2 warps x one-quad stages: 10, every bundle of a 10,000-system batch resident)
__global__ void __launch_bounds__(32 * W, (W == 2 && NQ == 1) ? 10 : (W == 8 && NQ == 1) ? 2 : (W == 4 && NQ == 2) ? 3 : 1)
    lu_refactor_rowlane_kernel(const RowlaneArgs a)
{
    constexpr int S = 8, NR = kRlQuadRecords, SR = NQ * NR;
    constexpr unsigned QUAD_BYTES = kRlQuadWords * 4, STAGE_BYTES = NQ * QUAD_BYTES, RING_BYTES = kRlRingStages * STAGE_BYTES;
    static_assert(STAGE_BYTES % 16 == 0 && NR == 4, "geometry");
    extern __shared__ __align__(16) uint8_t smem_all[];
    const int lane = threadIdx.x & 31, g = lane >> 2, h = lane & 3, wid = threadIdx.x >> 5;
    const i64 b = blockIdx.x;                                            // one bundle per CTA
    // shared memory: [0, 16) progress counters (8 x 16 bits), [64, 64 + 8 * W * 4) failure codes, then per warp accumulator + ring
    const unsigned flags_s = (unsigned)__cvta_generic_to_shared(smem_all);
    int *fails = reinterpret_cast<int *>(smem_all + 64);
    uint8_t *smem_raw = smem_all + 64 + 32 * kRlMaxWarps + (size_t)wid * ((size_t)a.nslots * 64u + RING_BYTES);
    const i64 sys0 = b * S + 2 * h, sys1 = sys0 + 1;
    const char *Axs0 = reinterpret_cast<const char *>(a.Ax + (sys0 < a.batch ? sys0 : a.batch - 1) * a.nnzA);
    const char *Axs1 = reinterpret_cast<const char *>(a.Ax + (sys1 < a.batch ? sys1 : a.batch - 1) * a.nnzA);
    char *Lb = reinterpret_cast<char *>(a.Lw + (size_t)b * a.lnz * S) + h * 16;
    char *Ub = reinterpret_cast<char *>(a.Uw + (size_t)b * a.unz * S) + h * 16;
    const unsigned smem_s = (unsigned)__cvta_generic_to_shared(smem_raw);
    const unsigned accb = smem_s + h * 16;                               // accumulator entry e: accb + 64 e
    const unsigned ring_s = smem_s + (unsigned)a.nslots * 64u;           // program ring
    const unsigned lw_off = 48u + (unsigned)g * 16u;                     // this lane group's four lane words inside a quad
    const unsigned aw_off = 176u + (unsigned)g * 16u;                    // ... and its four address words
    for (int t = lane; t < a.nslots * S; t += 32) reinterpret_cast<double *>(smem_raw)[t] = 0.0;
    if (W > 1) {
        if (threadIdx.x < 4) reinterpret_cast<unsigned *>(smem_all)[threadIdx.x] = 0u;
        __syncthreads();
    }

    // program ring: stage s lives in ring slot s % 4.  Three stages are loaded up front; entering stage s requests
    // stage s + 3 (into the slot stage s - 1 has left) and waits until at most that one group is pending.
    const uint8_t *pnext = a.prog + (size_t)a.stream_off[wid] * QUAD_BYTES + lane * 16;
    unsigned pdst = 0;
    auto request_stage = [&]() {
#pragma unroll
        for (unsigned o = 0; o < STAGE_BYTES; o += 512)
            if (o + 512 <= STAGE_BYTES || o + lane * 16 < STAGE_BYTES) cp_async16(ring_s + pdst + o + lane * 16, pnext + o);
        pnext += STAGE_BYTES;
        pdst = (pdst + STAGE_BYTES == RING_BYTES) ? 0u : pdst + STAGE_BYTES;
        cp_async_commit();
    };
    request_stage(); request_stage(); request_stage();
    cp_async_wait<0>();
    __syncwarp();

    double2 Q0[SR], Q1[SR];
    const double2 zero2 = make_double2(0.0, 0.0);
    // request the operands of the quads of the stage at ring address sb
    auto request_operands = [&](unsigned sb, double2 (&Q)[SR]) {
#pragma unroll
        for (int qd = 0; qd < NQ; ++qd) {
            const unsigned qa = sb + qd * QUAD_BYTES;
            const unsigned kind = lds_u32(qa) & 7u;
            if (kind == (unsigned)kRlUpdate) {
                const uint4 aw = lds_u4(qa + aw_off);
#pragma unroll
                for (int r = 0; r < NR; ++r) {
                    const unsigned o = word_of(aw, r);
                    if (o != 0xffffffffu) Q[qd * NR + r] = ldg_d2(Lb + (size_t)o);
                }
            } else if (kind == (unsigned)kRlFin || kind == (unsigned)kRlLoad4) {
                const uint4 aw = lds_u4(qa + aw_off);
#pragma unroll
                for (int r = 0; r < NR; ++r) {
                    const unsigned o = word_of(aw, r);
                    if (o != 0xffffffffu) Q[qd * NR + r] = make_double2(ldg_nc_f64(Axs0 + (size_t)o), ldg_nc_f64(Axs1 + (size_t)o));
                }
            }
        }
    };
    // have the other warps finished the columns the stage at ring address sb reads?  (8 x 16-bit counters)
    auto sources_final = [&](unsigned sb) -> bool {
        if (!(lds_u32(sb) & kRlFlagCross)) return true;                // the stage reads no column of another warp
        const uint4 req = lds_u4(sb + 32);
        uint4 done;
        asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(done.x), "=r"(done.y), "=r"(done.z), "=r"(done.w) : "r"(flags_s) : "memory");
        return (__vcmpgeu2(done.x, req.x) & __vcmpgeu2(done.y, req.y) & __vcmpgeu2(done.z, req.z) & __vcmpgeu2(done.w, req.w)) == 0xffffffffu;
    };
    unsigned sb0 = ring_s, sb1 = ring_s + STAGE_BYTES;                   // ring addresses of this stage and the next
#pragma unroll
    for (int d = 0; d < SR; ++d) { Q0[d] = zero2; Q1[d] = zero2; }
    bool have0 = W == 1 || sources_final(sb0), have1 = false;            // operands of the stage already requested?
    if (have0) request_operands(sb0, Q0);

    double2 m = zero2, piv = make_double2(1.0, 1.0), rcp = piv;
    int fail0 = INT32_MAX, fail1 = INT32_MAX;
    unsigned sink = 0, cols_done = 0, cols_published = 0;
    auto publish = [&]() {
        if (W > 1 && cols_done != cols_published) {
            __threadfence_block();                 // this warp's L stores before the counter (no load is outstanding here)
            __syncwarp();
            if (lane == 0) asm volatile("st.volatile.shared.u16 [%0], %1;" ::"r"(flags_s + 2u * (unsigned)wid), "h"((unsigned short)cols_done) : "memory");
            cols_published = cols_done;
        }
    };
    auto pivot_prologue = [&](const uint4 &h0) {
        piv = lds_d2(accb + (h0.y & 0xffffu));
        rcp = make_double2(rcp_refined(piv.x), rcp_refined(piv.y));
        if (!(fabs(piv.x) > 0.0 && isfinite(piv.x))) fail0 = min(fail0, (int)h0.w);
        if (!(fabs(piv.y) > 0.0 && isfinite(piv.y))) fail1 = min(fail1, (int)h0.w);
    };
    auto store_l = [&](unsigned w, unsigned base) {
        if (f_valid(w)) {
            const unsigned t = accb + f_slot64(w);
            const double2 x = lds_d2(t);
            sts_d2(t, zero2);
            stg_d2(Lb + (size_t)(base + f_off64(w)), make_double2(div_shared(x.x, piv.x, rcp.x), div_shared(x.y, piv.y, rcp.y)));
        }
    };
    auto store_u = [&](unsigned w, unsigned base) {
        if (f_valid(w)) {
            const unsigned t = accb + f_slot64(w);
            const double2 x = lds_d2(t);
            sts_d2(t, zero2);
            stg_cs_d2(Ub + (size_t)(base + f_off64(w)), x);
        }
    };
    // one stage: operands of this stage in C (requested one pass ago when haveC), the next stage's go to Q.  false: END reached
    auto pass = [&](double2 (&C)[SR], double2 (&Q)[SR], bool &haveC, bool &haveQ) -> bool {
        request_stage();
        cp_async_wait<1>();
        __syncwarp();                        // also: L values stored by other lanes are visible to the loads below
        // every global load of the warp is on one scoreboard slot: this use waits for the loads of C, which are the only
        // ones outstanding, BEFORE the loads of the next stage are issued; the quads below then never wait for memory
        sink ^= (unsigned)__double2hiint(C[0].x);
        if (W > 1) {
            publish();
            if (!haveC) {
                // sources of other warps were not final a pass ago: everything before this stage is executed and
                // published, so waiting here cannot deadlock; then read the operands now
                unsigned spins = 0;
                while (!sources_final(sb0))
                    if (++spins > (1u << 26)) { fail0 = fail1 = -2; break; }      // a broken program must not hang the GPU
                __threadfence_block();
                request_operands(sb0, C);
                sink ^= (unsigned)__double2hiint(C[0].x);
            }
            haveQ = sources_final(sb1);
            if (haveQ) {
                if (lds_u32(sb1) & kRlFlagCross) __threadfence_block();       // other warps' L stores before our loads
                request_operands(sb1, Q);
            }
        } else {
            request_operands(sb1, Q);
        }
#pragma unroll
        for (int qd = 0; qd < NQ; ++qd) {
            const unsigned qa = sb0 + qd * QUAD_BYTES;
            const uint4 hq = lds_u4(qa), lq = lds_u4(qa + lw_off);
            const unsigned kind = hq.x & 7u;
            if (kind == (unsigned)kRlUpdate) {
                load_multiplier(m, accb + (hq.y & 0xffffu), hq.x & 0x100u);
                update_op(lq.x, accb, C[qd * NR + 0], m);
                load_multiplier(m, accb + (hq.y >> 16), hq.x & 0x200u);
                update_op(lq.y, accb, C[qd * NR + 1], m);
                load_multiplier(m, accb + (hq.z & 0xffffu), hq.x & 0x400u);
                update_op(lq.z, accb, C[qd * NR + 2], m);
                load_multiplier(m, accb + (hq.z >> 16), hq.x & 0x800u);
                update_op(lq.w, accb, C[qd * NR + 3], m);
                continue;
            }
            if (NQ <= 2 && kind == (unsigned)kRlFin) {
                // (stages of 1 - 2 quads: the several-warps geometries; with 3 quads per stage the six copies of this
                // path cost more in the instruction cache than they save: config 4, one warp per bundle, 61 -> 79 ms)
                // column boundary, once per column and on the critical path of the warps that wait for this column: all
                // shared-memory loads first, the U entries / clearing / the next column's A values while the reciprocal
                // of the pivot is refined, the divisions last
                const uint4 bs = lds_u4(qa + 16);
                const bool vL = (hq.x & kRlHasL) && f_valid(lq.x), vU = (hq.x & kRlHasU) && f_valid(lq.y), vA = (hq.x & kRlHasA) && f_valid(lq.z);
                const unsigned tL = accb + f_slot64(lq.x), tU = accb + f_slot64(lq.y);
                double2 xp = piv;
                if (hq.x & kRlFlagP) xp = lds_d2(accb + (hq.y & 0xffffu));
                const double2 xL = lds_d2(tL), xU = lds_d2(tU);          // (an empty word carries a readable slot)
                if (vU) { sts_d2(tU, zero2); stg_cs_d2(Ub + (size_t)(bs.y + f_off64(lq.y)), xU); }
                if (vL) sts_d2(tL, zero2);
                if (vA) sts_d2(accb + f_slot64(lq.z), C[qd * NR + 2]);    // after the clearing: the next column reuses the slots
                if (hq.x & kRlFlagP) {
                    piv = xp;
                    rcp = make_double2(rcp_refined(piv.x), rcp_refined(piv.y));
                    if (!(fabs(piv.x) > 0.0 && isfinite(piv.x))) fail0 = min(fail0, (int)hq.w);
                    if (!(fabs(piv.y) > 0.0 && isfinite(piv.y))) fail1 = min(fail1, (int)hq.w);
                }
                if (vL) stg_d2(Lb + (size_t)(bs.x + f_off64(lq.x)), make_double2(div_shared(xL.x, piv.x, rcp.x), div_shared(xL.y, piv.y, rcp.y)));
                ++cols_done;
                continue;
            }
            if (kind == (unsigned)kRlUpdLate) {
                // rare: a source column of this warp was finalised less than two stages ago, read the operands now
                const uint4 aw = lds_u4(qa + aw_off);
                __syncwarp();
                double2 l[NR];
#pragma unroll
                for (int r = 0; r < NR; ++r) {
                    const unsigned o = word_of(aw, r);
                    l[r] = o != 0xffffffffu ? ldg_d2(Lb + (size_t)o) : zero2;
                }
#pragma unroll
                for (int r = 0; r < NR; ++r) {
                    const unsigned ms = r == 0 ? (hq.y & 0xffffu) : r == 1 ? (hq.y >> 16) : r == 2 ? (hq.z & 0xffffu) : (hq.z >> 16);
                    load_multiplier(m, accb + ms, hq.x & (0x100u << r));
                    update_op(word_of(lq, r), accb, l[r], m);
                }
            } else if (kind != (unsigned)kRlEnd) {
                // STOREL4 / STOREU4 / LOAD4 (columns with more than 8 entries of a role; FIN with 3 quads per stage): one
                // shared loop over the records
                if (hq.x & kRlFlagP) pivot_prologue(hq);
                unsigned roles = hq.x >> 24;
#pragma unroll 1
                for (int r = 0; roles != 0u; ++r, roles >>= 2) {
                    const unsigned role = roles & 3u;
                    if (role == 0u) continue;
                    const unsigned w = lds_u32(qa + lw_off + 4 * r);
                    if (role == kRlRoleL) store_l(w, lds_u32(qa + 16 + 4 * r));
                    else if (role == kRlRoleU) store_u(w, lds_u32(qa + 16 + 4 * r));
                    else if (f_valid(w)) {
                        const double2 c = r == 0 ? C[qd * NR + 0] : r == 1 ? C[qd * NR + 1] : r == 2 ? C[qd * NR + 2] : C[qd * NR + 3];
                        sts_d2(accb + f_slot64(w), c);
                    }
                }
                if (kind == (unsigned)kRlFin) ++cols_done;
            } else {
                return false;
            }
        }
        sb0 = sb1;
        sb1 = (sb1 + STAGE_BYTES == ring_s + RING_BYTES) ? ring_s : sb1 + STAGE_BYTES;
        return true;
    };
#pragma unroll 1
    for (;;) {
        if (!pass(Q0, Q1, have0, have1)) break;
        if (!pass(Q1, Q0, have1, have0)) break;
    }
    cp_async_wait<0>();
    publish();
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) {
        fail0 = min(fail0, __shfl_xor_sync(0xffffffffu, fail0, o));
        fail1 = min(fail1, __shfl_xor_sync(0xffffffffu, fail1, o));
    }
    if (W > 1) {
        if (g == 0) { fails[(wid * 4 + h) * 2] = fail0; fails[(wid * 4 + h) * 2 + 1] = fail1; }
        __syncthreads();
        if (wid == 0 && g == 0)
            for (int v = 1; v < W; ++v) { fail0 = min(fail0, fails[(v * 4 + h) * 2]); fail1 = min(fail1, fails[(v * 4 + h) * 2 + 1]); }
    }
    if (wid == 0 && g == 0 && a.status != nullptr) {
        if (sys0 < a.batch) a.status[sys0] = (fail0 == INT32_MAX) ? 0 : fail0;
        if (sys1 < a.batch) a.status[sys1] = (fail1 == INT32_MAX) ? 0 : fail1;
        if (sink == 0x7ff7dead && a.nslots < 0) a.status[0] = (i32)sink;          // keeps the dummy uses alive (never true)
    }
}

template <int W, int NQ>
int launch_T(const RowlaneArgs &a, i64 grid, size_t smem, cudaStream_t st)
{
    CSP3_CUDA(cudaFuncSetAttribute(lu_refactor_rowlane_kernel<W, NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lu_refactor_rowlane_kernel<W, NQ><<<(unsigned)grid, 32 * W, smem, st>>>(a);
    CSP3_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace

int launch_refactor_rowlane(const DevSchedule &D, i64 batch, const double *Ax, double *Lw, double *Uw, i32 *status, cudaStream_t st)
{
    if (batch <= 0) return 0;
    const int v = rowlane_variant(D, batch);
    if (v < 0) { set_error("row-lane refactor program not available for this pattern"); return -1; }
    const DevSchedule::RlVariant &R = D.rl[v];
    RowlaneArgs a;
    a.prog = R.prog;
    for (int w = 0; w < kRlMaxWarps; ++w) a.stream_off[w] = R.stream_off[w];
    a.nslots = R.nslots; a.nnzA = D.nnzA; a.lnz = D.lnz; a.unz = D.unz;
    a.batch = batch; a.Ax = Ax; a.Lw = Lw; a.Uw = Uw; a.status = status;
    const i64 grid = (batch + 7) / 8;
    const size_t smem = R.smem;
    switch (R.warps * 8 + R.stage_quads) {
        case 1 * 8 + 3: return launch_T<1, 3>(a, grid, smem, st);
        case 1 * 8 + 2: return launch_T<1, 2>(a, grid, smem, st);
        case 1 * 8 + 1: return launch_T<1, 1>(a, grid, smem, st);
        case 2 * 8 + 1: return launch_T<2, 1>(a, grid, smem, st);
        case 2 * 8 + 2: return launch_T<2, 2>(a, grid, smem, st);
        case 4 * 8 + 1: return launch_T<4, 1>(a, grid, smem, st);
        case 8 * 8 + 1: return launch_T<8, 1>(a, grid, smem, st);
        case 8 * 8 + 2: return launch_T<8, 2>(a, grid, smem, st);
        case 4 * 8 + 2: return launch_T<4, 2>(a, grid, smem, st);
        case 4 * 8 + 3: return launch_T<4, 3>(a, grid, smem, st);
        case 8 * 8 + 3: return launch_T<8, 3>(a, grid, smem, st);
        default: break;
    }
    set_error("row-lane refactor: no kernel for %d warps per bundle and %d quads per stage", R.warps, R.stage_quads);
    return -1;
}

}  // namespace csp3
