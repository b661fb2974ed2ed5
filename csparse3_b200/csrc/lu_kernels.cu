// lu_kernels.cu -- batched fp64 LU refactorisation and triangular solves on one sparsity pattern (sm_100a).
//
// No reference counterpart exists; semantics are the frozen-pattern / frozen-pivot
// refactorisation and the cs_ipvec -> cs_lsolve -> cs_usolve -> cs_ipvec solve defined by
// oracle/csp3_oracle.c (orc_csc_lu_refactor, orc_csc_lu_solve).  Both kernels reproduce the oracle's
// floating-point operation order exactly (unfused multiply / subtract, IEEE division), so results are
// bit-identical to the sequential algorithm and independent of bundle width, grid shape and GPU count.
//
// Work decomposition: ONE WARP owns a bundle of S systems and executes the compiled program of the pattern
// (program.hpp) for them, record after record, with no block- or grid-level barrier.  Lanes are split S x E
// (E = 32/S lanes over the entries of a column), so all metadata is shared by S systems.  Power-grid factors
// have elimination trees that are hundreds of levels deep and one or two columns wide, so parallelism inside a
// system is scarce; throughput comes from thousands of independent bundles in flight.
//
// Data movement
//   program  : global (L2-resident, shared by all bundles) -> shared-memory ring via cp.async, read with LDS
//   A, b     : system-major input, pulled into L1 by prefetch directives compiled into the program
//   L re-use : the last `win` L entries live in a shared-memory ring (the elimination order is a postorder,
//              most re-reads are recent); older columns are prefetched into L1 two columns ahead
//   factors  : written once, streamed back once by the sweeps
//
// Factor layouts in HBM
//   system-major (API):      Lx[batch][lnz], Ux[batch][unz]                   (cs_lu column layout)
//   bundle-interleaved (ws): Lw[bundle][lnz][S], Uw[bundle][unz][S], z[bundle][n][S]
//     the workspace form of the fused refactor+solve path: one entry of all S systems of a bundle is one
//     contiguous 8*S-byte run, so a warp-wide access to a column is one contiguous segment.
#include "common.cuh"
#include "program.hpp"

#include <algorithm>
#include <cmath>

namespace csp3 {

namespace {

constexpr size_t kMaxSmem = 200 * 1024;
constexpr int kProgStages = 4;      // ring slots of the program stream

__device__ __forceinline__ void pf_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];\n" ::"l"(p)); }
__device__ __forceinline__ void cp_async16(unsigned smem_dst, const void *gsrc)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// Sequential reader of a compiled program: stage s of the global byte stream lives in ring slot s % 4.
// Invariant while a record starting in stage `cur` is read: stages <= cur + 1 have landed (records are shorter
// than a stage, and the look-ahead never leaves stage cur + 1).
struct ProgStream {
    const uint8_t *src;
    const uint8_t *ring;        // generic pointer to the shared-memory ring
    unsigned ring_s;            // same, as a shared-window address
    unsigned shift, mask, stage;
    int nstages, cur;

    __device__ __forceinline__ void issue(int s, int lane)
    {
        if (s < nstages) {
            const unsigned dst = ring_s + ((unsigned)(s % kProgStages) << shift);
            const uint8_t *from = src + ((size_t)s << shift);
            for (unsigned u = lane * 16; u < stage; u += 32 * 16) cp_async16(dst + u, from + u);
        }
        cp_async_commit();                                   // exactly one group per stage, possibly empty
    }
    __device__ __forceinline__ void start(const uint8_t *program, int bytes, int stage_bytes, uint8_t *ring_ptr, int lane)
    {
        src = program; ring = ring_ptr; ring_s = (unsigned)__cvta_generic_to_shared(ring_ptr);
        stage = (unsigned)stage_bytes; shift = 31 - __clz(stage_bytes); mask = kProgStages * stage - 1;
        nstages = bytes >> shift; cur = 0;
        issue(0, lane); issue(1, lane); issue(2, lane);
        cp_async_wait<1>();
        __syncwarp();
    }
    // call with the offset of the record about to be read
    __device__ __forceinline__ void touch(unsigned p, int lane)
    {
        while ((int)(p >> shift) > cur) {
            ++cur;
            issue(cur + 2, lane);                            // overwrites stage cur-2: nothing reads it any more
            cp_async_wait<1>();                              // stages <= cur + 1 have landed
            __syncwarp();
        }
    }
    template <class T>
    __device__ __forceinline__ T ld(unsigned p) const { return *reinterpret_cast<const T *>(ring + (p & mask)); }
};

struct RefactorArgs {
    const uint8_t *prog;
    i32 prog_bytes, prog_stage;
    i32 n, nnzA, lnz, unz, acc_doubles, win_entries;
    i64 batch;
    const double *Ax;
    double *Lx, *Ux;
    i32 *status;
};

// ---------------------------------------------------------------------------------------------------------
// refactorisation
// ---------------------------------------------------------------------------------------------------------
template <int S, bool IL>
__global__ void __launch_bounds__(32) lu_refactor_kernel(const RefactorArgs a)
{
    constexpr int E = 32 / S;
    constexpr int LS = IL ? S : 1;              // stride between consecutive entries of one system
    extern __shared__ double smem[];
    __shared__ int fail[32];
    const int lane = threadIdx.x, sys = lane / E, e = lane - sys * E;
    const i64 b = blockIdx.x;
    const i64 g_raw = b * S + sys;
    const bool valid = g_raw < a.batch;
    const i64 g = valid ? g_raw : a.batch - 1;
    const bool store = IL || valid;             // padded lanes of the last bundle own workspace rows of their own
    const double *Axg = a.Ax + g * a.nnzA;
    double *Lg = IL ? a.Lx + (size_t)b * a.lnz * S + sys : a.Lx + g * a.lnz;
    double *Ug = IL ? a.Ux + (size_t)b * a.unz * S + sys : a.Ux + g * a.unz;
    double *acc = smem + sys;                   // accumulator slot t of this system: acc[t * S]
    double *win = smem + a.acc_doubles + sys;   // ring of the most recent L entries: win[(p & wmask) * S]
    const int wmask = a.win_entries - 1;
    fail[lane] = INT32_MAX;
    ProgStream ps;
    ps.start(a.prog, a.prog_bytes, a.prog_stage, reinterpret_cast<uint8_t *>(smem + a.acc_doubles + (size_t)a.win_entries * S), lane);

    unsigned p = 0;
    for (int k = 0; k < a.n; ++k) {
        ps.touch(p, lane);
        const int2 ha = ps.ld<int2>(p), hb = ps.ld<int2>(p + 8);          // records are 8-byte aligned
        const uint2 h2 = ps.ld<uint2>(p + 16);
        const int up = ha.x, lp = ha.y;
        const int ucnt = hb.x & 0xffff, lcnt = (int)((unsigned)hb.x >> 16);
        const int a_cnt = hb.y & 0xffff, pair_cnt = (int)((unsigned)hb.y >> 16);
        const int pf_cnt = (int)(h2.x & 0xffff), mpf_cnt = (int)(h2.x >> 16);
        const unsigned pa = p + kRfHeader;                                   // A sources (i32)
        const unsigned po = pa + 4 * a_cnt;                                  // A accumulator slots (u16)
        const unsigned pp = (po + 2 * a_cnt + 7) & ~7u;                      // A prefetch directives (i32)
        const unsigned pm = (pp + 4 * pf_cnt + 7) & ~7u;                     // far-source prefetch directives (8 B)
        p = pm + 8 * mpf_cnt;                                                // first pair record
        const int len = ucnt + lcnt - 1;
        const int wlo = lp - a.win_entries;                                  // L entries >= wlo are in the ring

        // this column's A values (their lines were prefetched kPfCols columns ago)
        int ao = -1;
        double av = 0.0;
        if (e < a_cnt) { ao = ps.ld<uint16_t>(po + 2 * e); av = __ldg(Axg + ps.ld<int>(pa + 4 * e)); }
        // prefetch directives: A values of column k + kPfCols, far-back L columns of column k + kPfMissCols
        for (int t = e; t < pf_cnt; t += E) pf_l1(Axg + ps.ld<int>(pp + 4 * t));
        for (int m = 0; m < mpf_cnt; ++m) {
            const uint2 d = ps.ld<uint2>(pm + 8 * m);
            const int fl = (int)(d.y & 0xffff);
            if (IL) {                                                        // one 128-byte line = 16/S bundle entries
                constexpr int EPL = (16 / S > 0) ? 16 / S : 1;
                if (lane * EPL < fl) pf_l1(Lg - sys + (size_t)((int)d.x + lane * EPL) * S);
            } else if (e * 16 < fl) {
                pf_l1(Lg + (int)d.x + e * 16);
            }
        }
        for (int t = e; t < len; t += E) acc[t * S] = 0.0;
        __syncwarp();
        if (ao >= 0) acc[ao * S] = av;
        for (int t = e + E; t < a_cnt; t += E) acc[ps.ld<uint16_t>(po + 2 * t) * S] = __ldg(Axg + ps.ld<int>(pa + 4 * t));
        __syncwarp();

        // left-looking updates, one pair record per off-diagonal entry of U(:,k), in the stored (topological) order.
        // One-pair look-ahead: the next record's header, its first map entry and (for a far-back source column)
        // its first L values are requested before the dependent work of the current pair starts.
        int2 ph = make_int2(0, 0);
        int offn = 0;
        double lvn = 0.0;
        if (pair_cnt > 0) {
            ps.touch(p, lane);
            ph = ps.ld<int2>(p);
            if (e < (int)((unsigned)ph.y >> 16)) {
                offn = ps.ld<uint16_t>(p + 8 + 2 * e);
                if (ph.x < wlo) lvn = Lg[(size_t)(ph.x + e) * LS];
            }
        }
        for (int pi = 0; pi < pair_cnt; ++pi) {
            const int lstart = ph.x, moff = ph.y & 0xffff, llen = (int)((unsigned)ph.y >> 16);
            const unsigned pmap = p + 8;
            p = (pmap + 2 * llen + 7) & ~7u;
            const int off0 = offn;
            const double lv0 = lvn;
            if (pi + 1 < pair_cnt) {
                ps.touch(p, lane);
                ph = ps.ld<int2>(p);
                if (e < (int)((unsigned)ph.y >> 16)) {
                    offn = ps.ld<uint16_t>(p + 8 + 2 * e);
                    if (ph.x < wlo) lvn = Lg[(size_t)(ph.x + e) * LS];
                }
            }
            const double mult = acc[moff * S];
            if (lstart >= wlo) {                                             // whole source column is in the ring
                if (e < llen) acc[off0 * S] = __dsub_rn(acc[off0 * S], __dmul_rn(win[((lstart + e) & wmask) * S], mult));
                for (int t = e + E; t < llen; t += E) {
                    const int off = ps.ld<uint16_t>(pmap + 2 * t);
                    acc[off * S] = __dsub_rn(acc[off * S], __dmul_rn(win[((lstart + t) & wmask) * S], mult));
                }
            } else {
                if (e < llen) acc[off0 * S] = __dsub_rn(acc[off0 * S], __dmul_rn(lv0, mult));
                for (int t = e + E; t < llen; t += E) {
                    const int off = ps.ld<uint16_t>(pmap + 2 * t);
                    acc[off * S] = __dsub_rn(acc[off * S], __dmul_rn(Lg[(size_t)(lstart + t) * LS], mult));
                }
            }
            __syncwarp();
        }

        // finalize: U(:,k) as accumulated; L(:,k) = x / pivot, unit diagonal first; mirror L into the ring
        const double pivot = acc[(ucnt - 1) * S];
        if (store) {
            for (int t = e; t < ucnt; t += E) Ug[(size_t)(up + t) * LS] = acc[t * S];
            if (e == 0) Lg[(size_t)lp * LS] = 1.0;
        }
        for (int t = e; t < lcnt - 1; t += E) {
            const double v = acc[(ucnt + t) * S] / pivot;
            if (store) Lg[(size_t)(lp + 1 + t) * LS] = v;
            win[((lp + 1 + t) & wmask) * S] = v;
        }
        if (e == 0 && valid && !(fabs(pivot) > 0.0 && isfinite(pivot))) atomicMin(&fail[sys], k + 1);
        __syncwarp();
    }
    cp_async_wait<0>();
    if (a.status != nullptr && lane < S) {
        const i64 gs = b * S + lane;
        if (gs < a.batch) a.status[gs] = (fail[lane] == INT32_MAX) ? 0 : fail[lane];
    }
}

// ---------------------------------------------------------------------------------------------------------
// triangular solves
// ---------------------------------------------------------------------------------------------------------
struct SolveArgs {
    const uint8_t *lprog, *uprog;
    i32 lprog_bytes, lprog_stage, uprog_bytes, uprog_stage;
    i32 n, lnz, unz, nslots, ring_bytes, scratch_doubles;
    i64 batch;
    const double *Lx, *Ux, *b;
    double *x, *z;            // z: forward result, indexed like x; [bundle][n][S] (interleaved path only)
};

// One column-oriented sweep in exactly the order of cs_lsolve (columns ascending) / cs_usolve (descending).
// The live rows of y sit in host-assigned shared-memory slots; factor values are read straight from global
// memory behind a sequential L1 prefetcher (each column is one contiguous run).
template <int S, bool IL, bool UPPER>
__device__ __forceinline__ void sweep(const SolveArgs &a, ProgStream &ps, double *yc, const double *Fg,
                                      const double *rhs_base, int rhs_stride, double *out_base, int out_stride,
                                      bool out_ok, int lane, int sys, int e)
{
    constexpr int E = 32 / S;
    constexpr int LS = IL ? S : 1;
    constexpr int kAheadEntries = 96;                  // prefetch distance along the factor stream, in entries
    const int n = a.n;
    unsigned p = 0;
    for (int step = 0; step < n; ++step) {
        ps.touch(p, lane);
        const int2 ha = ps.ld<int2>(p), hb = ps.ld<int2>(p + 8);          // records are 8-byte aligned
        const int outpos = ps.ld<int>(p + 16);
        const int start = ha.x, rhs = ha.y;
        const int slot = (int)(short)(hb.x & 0xffff), len = (int)((unsigned)hb.x >> 16);
        const int nalloc = hb.y & 0xffff, pf_cnt = (int)((unsigned)hb.y >> 16);
        const unsigned psl = p + kSvHeader;                                  // target slots (u16)
        const unsigned pal = (psl + 2 * len + 7) & ~7u;                      // alloc rhs indices (i32)
        const unsigned pas = pal + 4 * nalloc;                               // alloc slots (u16)
        const unsigned ppf = (pas + 2 * nalloc + 7) & ~7u;                   // rhs prefetch directives (i32)
        p = (ppf + 4 * pf_cnt + 7) & ~7u;

        // operands that do not depend on y
        int sl0 = 0;
        double fv0 = 0.0;
        if (e < len) { sl0 = ps.ld<uint16_t>(psl + 2 * e); fv0 = Fg[(size_t)(start + e) * LS]; }
        int aslot = -1;
        double arhs = 0.0;
        if (e < nalloc) { aslot = ps.ld<uint16_t>(pas + 2 * e); arhs = rhs_base[(size_t)ps.ld<int>(pal + 4 * e) * rhs_stride]; }
        double d = 1.0;
        if (UPPER) d = Fg[(size_t)(start + len) * LS];
        // prefetch: right-hand sides of the rows allocated kPfCols columns from now, next part of the factor stream
        for (int t = e; t < pf_cnt; t += E) pf_l1(rhs_base + (size_t)ps.ld<int>(ppf + 4 * t) * rhs_stride);
        {
            const int ahead = UPPER ? start - kAheadEntries : start + len + kAheadEntries;
            const int limit = UPPER ? a.unz : a.lnz;
            if (ahead >= 0 && ahead < limit) {
                if (IL) { if (lane == 0) pf_l1(Fg - sys + (size_t)ahead * S); }
                else if (e == 0) pf_l1(Fg + ahead);
            }
        }
        const double yraw = (slot >= 0) ? yc[slot * S] : rhs_base[(size_t)rhs * rhs_stride];
        const double yj = UPPER ? yraw / d : yraw;
        if (e == 0 && out_ok) out_base[(size_t)outpos * out_stride] = yj;
        __syncwarp();                                               // y[j] is read before its slot is reused
        if (aslot >= 0) yc[aslot * S] = arhs;
        for (int t = e + E; t < nalloc; t += E)
            yc[ps.ld<uint16_t>(pas + 2 * t) * S] = rhs_base[(size_t)ps.ld<int>(pal + 4 * t) * rhs_stride];
        __syncwarp();
        if (e < len) yc[sl0 * S] = __dsub_rn(yc[sl0 * S], __dmul_rn(fv0, yj));
        for (int t = e + E; t < len; t += E) {
            const int sl = ps.ld<uint16_t>(psl + 2 * t);
            yc[sl * S] = __dsub_rn(yc[sl * S], __dmul_rn(Fg[(size_t)(start + t) * LS], yj));
        }
        __syncwarp();
    }
    cp_async_wait<0>();
    __syncwarp();
}

// Backward sweep, row-oriented: x_i = (y_i - sum_{j>i} U(i,j) x_j) / U(i,i), subtractions in descending j (the order
// in which cs_usolve's column sweep touches row i).  The E lanes of a system fetch the row's U values and form the
// unfused products in parallel, park them in shared memory, and lane 0 subtracts them in order.  Only the x_j that a
// later row still needs live on chip (host-assigned slots); U is read in place from its column-major array.
template <int S, bool IL>
__device__ __forceinline__ void usweep_rows(const SolveArgs &a, ProgStream &ps, double *xc, double *scratch,
                                            const double *Ug, double *park, int park_stride, double *xg,
                                            bool valid, int lane, int e)
{
    constexpr int E = 32 / S;
    constexpr int LS = IL ? S : 1;
    const int n = a.n;
    unsigned p = 0;
    // operands of the row about to be processed (first chunk), requested one row ahead
    ps.touch(p, lane);
    int2 ha = ps.ld<int2>(p), hb = ps.ld<int2>(p + 8);
    double un = 0.0, dn = Ug[(size_t)ha.x * LS], yn = park[(size_t)ha.y * park_stride];
    int sn = 0;
    if (e < (int)((unsigned)hb.x >> 16)) { const int2 en = ps.ld<int2>(p + kSvHeader + 8 * e); un = Ug[(size_t)en.x * LS]; sn = en.y & 0xffff; }
    for (int step = 0; step < n; ++step) {
        const int xpos = ha.y, slot_out = (int)(short)(hb.x & 0xffff), len = (int)((unsigned)hb.x >> 16);
        const int pf_cnt = hb.y & 0xffff;
        const int pf_x = ps.ld<int>(p + 16);
        const unsigned pe = p + kSvHeader;                                   // entries (8 B each)
        const unsigned ppf = pe + 8 * len;                                   // prefetch directives (i32)
        p = (ppf + 4 * pf_cnt + 7) & ~7u;
        const double u0 = un, d = dn, y0 = yn;
        const int s0 = sn;
        // prefetch directives for the row kPfCols steps ahead (its U entries, its diagonal, its parked y)
        for (int t = e; t < pf_cnt; t += E) pf_l1(Ug + (size_t)ps.ld<int>(ppf + 4 * t) * LS);
        if (pf_x >= 0 && e == 0) pf_l1(park + (size_t)pf_x * park_stride);
        // look-ahead: next row's header and first-chunk operands
        if (step + 1 < n) {
            ps.touch(p, lane);
            ha = ps.ld<int2>(p); hb = ps.ld<int2>(p + 8);
            dn = Ug[(size_t)ha.x * LS];
            yn = park[(size_t)ha.y * park_stride];
            if (e < (int)((unsigned)hb.x >> 16)) { const int2 en = ps.ld<int2>(p + kSvHeader + 8 * e); un = Ug[(size_t)en.x * LS]; sn = en.y & 0xffff; }
        }
        // products of this row, parked in order
        if (e < len) scratch[e * S] = __dmul_rn(u0, xc[s0 * S]);
        for (int t = e + E; t < len; t += E) {
            const int2 en = ps.ld<int2>(pe + 8 * t);
            scratch[t * S] = __dmul_rn(Ug[(size_t)en.x * LS], xc[(en.y & 0xffff) * S]);
        }
        __syncwarp();
        if (e == 0) {
            double s = y0;
            for (int t = 0; t < len; ++t) s = __dsub_rn(s, scratch[t * S]);
            const double xi = s / d;
            if (valid) xg[xpos] = xi;
            if (slot_out >= 0) xc[slot_out * S] = xi;
        }
        __syncwarp();
    }
    cp_async_wait<0>();
    __syncwarp();
}

template <int S, bool IL>
__global__ void __launch_bounds__(32) lu_solve_kernel(const SolveArgs a)
{
    constexpr int E = 32 / S;
    extern __shared__ double smem[];
    const int lane = threadIdx.x, sys = lane / E, e = lane - sys * E;
    const i64 bnd = blockIdx.x;
    const i64 g_raw = bnd * S + sys;
    const bool valid = g_raw < a.batch;
    const i64 g = valid ? g_raw : a.batch - 1;
    const int n = a.n;
    const double *Lg = IL ? a.Lx + (size_t)bnd * a.lnz * S + sys : a.Lx + g * a.lnz;
    const double *Ug = IL ? a.Ux + (size_t)bnd * a.unz * S + sys : a.Ux + g * a.unz;
    const double *bg = a.b + g * n;
    double *xg = a.x + g * n;
    // forward results are parked at their final x position: in the interleaved scratch z, or in x itself
    double *park = IL ? a.z + (size_t)bnd * n * S + sys : xg;
    const int park_stride = IL ? S : 1;
    double *yc = smem + sys;                                        // live row in slot s: yc[s * S]
    uint8_t *ring = reinterpret_cast<uint8_t *>(smem + (size_t)a.ring_bytes / 8);   // ring_bytes: byte offset of the ring
    ProgStream ps;
    // y = L \ (P b)
    ps.start(a.lprog, a.lprog_bytes, a.lprog_stage, ring, lane);
    sweep<S, IL, false>(a, ps, yc, Lg, bg, 1, park, park_stride, IL || valid, lane, sys, e);
    // x = Q (U \ y), row-oriented
    ps.start(a.uprog, a.uprog_bytes, a.uprog_stage, ring, lane);
    usweep_rows<S, IL>(a, ps, yc, smem + a.scratch_doubles + sys, Ug, park, park_stride, xg, valid, lane, e);
}

template <int S, bool IL>
int launch_refactor_T(const RefactorArgs &a, size_t smem, cudaStream_t st)
{
    CSP3_CUDA(cudaFuncSetAttribute(lu_refactor_kernel<S, IL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const i64 grid = (a.batch + S - 1) / S;
    lu_refactor_kernel<S, IL><<<(unsigned)grid, 32, smem, st>>>(a);
    CSP3_CUDA(cudaGetLastError());
    return 0;
}

template <int S, bool IL>
int launch_solve_T(const SolveArgs &a, size_t smem, cudaStream_t st)
{
    CSP3_CUDA(cudaFuncSetAttribute(lu_solve_kernel<S, IL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const i64 grid = (a.batch + S - 1) / S;
    lu_solve_kernel<S, IL><<<(unsigned)grid, 32, smem, st>>>(a);
    CSP3_CUDA(cudaGetLastError());
    return 0;
}

int pow2_floor(int v) { int p = 1; while (2 * p <= v) p *= 2; return p; }

int pick_width(int requested, i64 batch)
{
    if (requested == 1 || requested == 2 || requested == 4 || requested == 8 || requested == 16) return requested;
    return (batch >= 8 * kNumSMs) ? 4 : (batch >= 2 * kNumSMs ? 2 : 1);
}

}  // namespace

bool use_wide(const DevSchedule &D, i64 batch)
{
    (void)batch;
    return D.wide_ok && tuning().wide != 0 && tuning().ws_S == 0;
}

bool use_panel(const DevSchedule &D, i64 batch)
{
    (void)batch;
    return D.panel_ok && tuning().panel != 0 && tuning().ws_S == 0 && (!D.wide_ok || D.wide_S == 8);
}
// Which refactor kernel serves `batch` systems on the workspace path (measured on B200):
//  * the wide kernel is the throughput kernel: one warp per bundle, everything of a bundle in shared memory; best
//    when the batch fills the GPU (config 3: 6.6 ms for 1,250 bundles) but a bundle is one sequential walk (4.7 ms
//    whatever the batch), and patterns with long columns leave 3 bundles per SM (config 4);
//  * the row-lane kernel runs 1 / 2 / 4 / 8 warps per bundle on different columns (9 KB of shared memory per warp):
//    8 warps (stages of one quad, 120 registers: two CTAs per SM) up to 2 bundles per SM (config 3: 1.66 ms for one
//    bundle, 2.23 ms for 157), 4 warps up to one wave of 3 bundles per SM (2.4 - 2.8 ms), above
//    that the wide kernel -- or, for the long-column patterns, one warp per bundle (config 4: 60.9 ms against 80.9 ms).
//    (2 warps per bundle win by 10 % between 3 and 6 bundles per SM on an idle GPU, 4.8 against 5.3 ms at 625 bundles,
//    but lose when two such launches share the GPU, as the chunks of csp3_nr_solve_host do: not selected automatically.)
int rowlane_variant(const DevSchedule &D, i64 batch)
{
    const Tuning &t = tuning();
    if (!D.rl_enabled || t.rowlane == 0 || t.ws_S != 0) return -1;
    if (t.rowlane < 0 && (t.tmem != 0 || t.panel != 0)) return -1;      // an experimental kernel was asked for explicitly
    if (t.rl_warps > 0) return ensure_rowlane_variant(D, kRlVariants - 1) == 0 ? kRlVariants - 1 : -1;   // CSP3_RL_W / CSP3_RL_NQ: forced geometry
    const i64 bundles = (batch + 7) / 8;
    const bool long_columns = D.wrf_smem > (size_t)40 * 1024;
    // a geometry may be unavailable for a pattern (8 accumulators of a very long column exceed shared memory): next one
    if (bundles <= (i64)2 * kNumSMs && ensure_rowlane_variant(D, 3) == 0) return 3;
    if (bundles <= (i64)3 * kNumSMs && ensure_rowlane_variant(D, 2) == 0) return 2;
    if ((long_columns || t.rowlane > 0) && ensure_rowlane_variant(D, 0) == 0) return 0;
    return -1;
}

bool use_rowlane(const DevSchedule &D, i64 batch) { return rowlane_variant(D, batch) >= 0; }

// The sweeps of a small batch: row by row on 8 warps per bundle, in the regime where the refactorisation runs 8 warps per
// bundle too (up to 2 bundles per SM); otherwise the wide sweeps (one warp per bundle).
bool use_rowsweep(const DevSchedule &D, i64 batch)
{
    if (tuning().rowsweep == 0 || !D.rl_enabled || tuning().ws_S != 0 || tuning().tmem != 0 || tuning().panel != 0) return false;
    if ((batch + 7) / 8 > (i64)2 * kNumSMs && tuning().rowsweep < 2) return false;
    return ensure_rowsweep(D) == 0;
}
// (use_tmem: lu_wide.cu.  With the panel kernel selected the factors stay in 8-system bundles.)

int workspace_bundle_width(const DevSchedule &D, i64 batch)
{
    if (use_panel(D, batch) || use_rowlane(D, batch)) return 8;
    if (use_wide(D, batch)) return D.wide_S;
    const int S = tuning().ws_S;
    if (S == 2 || S == 4 || S == 8 || S == 16) return S;
    // The kernels are latency-bound per warp: use the narrowest bundle that still lets every bundle of the
    // batch be resident at once (32 one-warp CTAs per SM), wider only when the batch would not fit in one wave.
    for (int w = 2; w <= 8; w *= 2)
        if ((batch + w - 1) / w <= (i64)32 * kNumSMs) return w;
    return 16;
}

// growth[s] = max |L(i,j)|, i > j, of system s from the bundle-interleaved factor workspace: the element growth a frozen
// pivot sequence must be watched for.  One CTA per bundle; every step reads 2 KB contiguous.
__global__ void lu_growth_kernel(i64 batch, int lnz, int S, const uint8_t *__restrict__ ldiag, const double *__restrict__ Lw,
                                 double *__restrict__ growth)
{
    __shared__ double part[256];
    const i64 b = blockIdx.x;
    const int s = threadIdx.x % S, q = threadIdx.x / S, nq = blockDim.x / S;
    const double *base = Lw + (size_t)b * lnz * S + s;
    double m = 0.0;
    for (int e = q; e < lnz; e += nq)
        if (!ldiag[e]) m = fmax(m, fabs(base[(size_t)e * S]));
    part[threadIdx.x] = m;
    __syncthreads();
    if (q == 0) {
        for (int t = 1; t < nq; ++t) m = fmax(m, part[t * S + s]);          // NaN-free max: a NaN factor shows up in `status`
        const i64 g = b * S + s;
        if (g < batch) growth[g] = m;
    }
}

int launch_growth(const DevSchedule &D, i64 batch, const double *Lw, double *growth, cudaStream_t st)
{
    if (batch <= 0) return 0;
    const int S = use_tmem(D, batch) ? 32 : workspace_bundle_width(D, batch);
    lu_growth_kernel<<<(unsigned)((batch + S - 1) / S), 256, 0, st>>>(batch, D.lnz, S, D.d_ldiag, Lw, growth);
    CSP3_CUDA(cudaGetLastError());
    return 0;
}

int launch_refactor(const DevSchedule &D, i64 batch, const double *Ax, double *Lx, double *Ux, i32 *status,
                    bool interleaved, cudaStream_t st)
{
    if (batch <= 0) return 0;
    if (interleaved && use_panel(D, batch)) return launch_refactor_panel(D, batch, Ax, Lx, Ux, status, nullptr, st);
    if (interleaved && use_rowlane(D, batch)) return launch_refactor_rowlane(D, batch, Ax, Lx, Ux, status, st);
    if (interleaved && use_tmem(D, batch)) return launch_refactor_tmem(D, batch, Ax, Lx, Ux, status, st);
    if (interleaved && use_wide(D, batch)) return launch_refactor_wide(D, batch, Ax, Lx, Ux, status, st);
    RefactorArgs a;
    a.prog = D.rf_prog; a.prog_bytes = D.rf_prog_bytes; a.prog_stage = D.rf_prog_stage;
    a.n = D.n; a.nnzA = D.nnzA; a.lnz = D.lnz; a.unz = D.unz;
    a.batch = batch; a.Ax = Ax; a.Lx = Lx; a.Ux = Ux; a.status = status;
    int S = interleaved ? workspace_bundle_width(D, batch) : pick_width(tuning().rf_S, batch);
    const int len = D.max_col_len > 0 ? D.max_col_len : 1;
    const size_t ring = (size_t)kProgStages * D.rf_prog_stage;
    while (!interleaved && S > 1 && (size_t)len * S * 8 + ring > kMaxSmem / 2) S >>= 1;
    a.acc_doubles = (len * S + 1) & ~1;                               // keep the program ring 16-byte aligned
    const size_t acc_bytes = (size_t)a.acc_doubles * 8;
    if (acc_bytes + ring + (size_t)64 * S * 8 > kMaxSmem) {
        set_error("factor column too long for shared memory (%d entries, bundle width %d)", len, S);
        return -1;
    }
    // Recent-L ring: as large as possible while every bundle of the batch stays co-resident (one wave): a
    // second wave costs far more than a smaller ring, because the kernel is latency- not capacity-bound.
    int win;
    if (tuning().rf_win > 0) {
        win = pow2_floor(std::max(64, tuning().rf_win));
    } else {
        const i64 bundles = (batch + S - 1) / S;
        const i64 per_sm = std::max<i64>(1, std::min<i64>(32, (bundles + kNumSMs - 1) / kNumSMs));
        const size_t budget = (size_t)(227 * 1024) / (size_t)per_sm - 1024 - 256;      // per-CTA reservation + static
        win = 1024;
        while (win > 64 && acc_bytes + ring + (size_t)win * S * 8 > budget) win >>= 1;
    }
    while (win > 64 && acc_bytes + ring + (size_t)win * S * 8 > kMaxSmem) win >>= 1;
    a.win_entries = win;
    const size_t smem = acc_bytes + (size_t)win * S * 8 + ring;
#define CSP3_RF(SV) return interleaved ? launch_refactor_T<SV, true>(a, smem, st) : launch_refactor_T<SV, false>(a, smem, st)
    switch (S) {
        case 1: return launch_refactor_T<1, false>(a, smem, st);
        case 2: CSP3_RF(2);
        case 4: CSP3_RF(4);
        case 8: CSP3_RF(8);
        case 16: CSP3_RF(16);
    }
#undef CSP3_RF
    set_error("invalid refactor bundle width %d", S);
    return -1;
}

int launch_solve(const DevSchedule &D, i64 batch, const double *Lx, const double *Ux, const double *b,
                 double *x, double *z, bool interleaved, cudaStream_t st)
{
    if (batch <= 0) return 0;
    if (interleaved && use_rowsweep(D, batch)) return launch_solve_rows(D, batch, Lx, Ux, b, x, z, st);
    if (interleaved && use_wide(D, batch) && D.wide_solve_ok && tuning().wide_solve != 0) {
        const i64 padded = (batch + 31) / 32 * 32;
        return launch_solve_wide(D, batch, Lx, Ux, b, x, z, z + padded * D.n, st);
    }
    SolveArgs a;
    a.lprog = D.ls_prog; a.lprog_bytes = D.ls_prog_bytes; a.lprog_stage = D.ls_prog_stage;
    a.uprog = D.ur_prog; a.uprog_bytes = D.ur_prog_bytes; a.uprog_stage = D.ur_prog_stage;
    a.n = D.n; a.lnz = D.lnz; a.unz = D.unz;
    a.nslots = (std::max(D.ls_nslots, D.ur_nslots) + 2) & ~1;          // live rows of y (forward) / x (backward)
    const int scratch_rows = (D.ur_max_len + 2) & ~1;                   // products of one row of U
    a.batch = batch; a.Lx = Lx; a.Ux = Ux; a.b = b; a.x = x; a.z = z;
    const size_t ring = (size_t)kProgStages * std::max(D.ls_prog_stage, D.ur_prog_stage);
    int S = interleaved ? workspace_bundle_width(D, batch) : pick_width(tuning().sv_S, batch);
    while (!interleaved && S > 1 && (size_t)(a.nslots + scratch_rows) * S * 8 + ring > kMaxSmem / 2) S >>= 1;
    a.scratch_doubles = a.nslots * S;
    a.ring_bytes = (a.nslots + scratch_rows) * S * 8;
    const size_t smem = (size_t)a.ring_bytes + ring;
    if (smem > kMaxSmem) {
        set_error("solve working set too large for shared memory (%d live rows, bundle width %d, %zu bytes)", a.nslots, S, smem);
        return -1;
    }
#define CSP3_SV(SV) return interleaved ? launch_solve_T<SV, true>(a, smem, st) : launch_solve_T<SV, false>(a, smem, st)
    switch (S) {
        case 1: return launch_solve_T<1, false>(a, smem, st);
        case 2: CSP3_SV(2);
        case 4: CSP3_SV(4);
        case 8: CSP3_SV(8);
        case 16: CSP3_SV(16);
    }
#undef CSP3_SV
    set_error("invalid solve bundle width %d", S);
    return -1;
}

}  // namespace csp3
