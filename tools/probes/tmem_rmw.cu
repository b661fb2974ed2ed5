// tmem_rmw.cu -- probe: tensor memory (TMEM) as a dynamically addressed per-lane scratchpad for fp64 accumulators.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_rmw tmem_rmw.cu && ./tmem_rmw
// Every lane of a warp owns one TMEM lane; slot s of a lane = columns 2s, 2s+1 (one double).  The kernel runs the
// access pattern of a left-looking LU chunk: load G accumulator slots (dynamic columns), subtract l * m, store them
// back, with the slot indices read from a table -- and checks the result against the same recurrence on the host.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ double tm_ld(unsigned taddr)
{
    unsigned lo, hi;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(taddr));
    return __hiloint2double((int)hi, (int)lo);
}
__device__ __forceinline__ void tm_st(unsigned taddr, double v)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"((unsigned)__double2loint(v)), "r"((unsigned)__double2hiint(v)) : "memory");
}
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <int G>
__global__ void k_rmw(int nslots, int rounds, const int *__restrict__ table, const double *__restrict__ init, double *out, long long *cycles)
{
    __shared__ unsigned tbase;
    __shared__ int tab[4096];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&tbase)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    for (int i = threadIdx.x; i < rounds * G && i < 4096; i += blockDim.x) tab[i] = table[i];
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const unsigned base = tbase + ((unsigned)(warp * 32) << 16);           // this warp's 32 TMEM lanes
    const int sys = (blockIdx.x * (blockDim.x >> 5) + warp) * 32 + lane;
    for (int s = 0; s < nslots; ++s) tm_st(base + 2 * s, init[(size_t)sys * nslots + s]);
    tm_wait_st();
    const double l = 1.0 + 1e-3 * lane, m = 0.5;
    const long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
        double x[G];
#pragma unroll
        for (int g = 0; g < G; ++g) x[g] = tm_ld(base + 2 * tab[(r * G + g) & 4095]);
        tm_wait_ld();
#pragma unroll
        for (int g = 0; g < G; ++g) x[g] = __dsub_rn(x[g], __dmul_rn(l, m));
#pragma unroll
        for (int g = 0; g < G; ++g) tm_st(base + 2 * tab[(r * G + g) & 4095], x[g]);
        tm_wait_st();
    }
    const long long t1 = clock64();
    for (int s = 0; s < nslots; ++s) { double v = tm_ld(base + 2 * s); tm_wait_ld(); out[(size_t)sys * nslots + s] = v; }
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(256));
}

template <int G>
void run(int warps, int blocks)
{
    const int nslots = 128, rounds = 4096 / G;
    const int nsys = blocks * warps * 32;
    std::vector<int> table(rounds * G);
    for (int r = 0; r < rounds; ++r)
        for (int g = 0; g < G; ++g) table[r * G + g] = (r * 37 + g * 5) % nslots;          // distinct inside a round (G <= 25)
    std::vector<double> init((size_t)nsys * nslots), ref;
    for (size_t i = 0; i < init.size(); ++i) init[i] = (double)(i % 1000) * 0.25;
    ref = init;
    for (int s = 0; s < nsys; ++s) {
        const double l = 1.0 + 1e-3 * (s & 31), m = 0.5;
        for (int i = 0; i < rounds * G; ++i) ref[(size_t)s * nslots + table[i]] -= l * m;
    }
    int *dt; double *di, *dout; long long *dc;
    CK(cudaMalloc(&dt, table.size() * 4)); CK(cudaMalloc(&di, init.size() * 8)); CK(cudaMalloc(&dout, init.size() * 8)); CK(cudaMalloc(&dc, blocks * 8));
    CK(cudaMemcpy(dt, table.data(), table.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(di, init.data(), init.size() * 8, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k_rmw<G><<<blocks, warps * 32>>>(nslots, rounds, dt, di, dout, dc);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    k_rmw<G><<<blocks, warps * 32>>>(nslots, rounds, dt, di, dout, dc);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    std::vector<double> out(init.size()); std::vector<long long> cyc(blocks);
    CK(cudaMemcpy(out.data(), dout, out.size() * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(cyc.data(), dc, blocks * 8, cudaMemcpyDeviceToHost));
    size_t bad = 0;
    for (size_t i = 0; i < out.size(); ++i) bad += out[i] != ref[i];
    printf("G=%2d warps/CTA=%d CTAs=%4d: %s, %.1f cycles per read-modify-write (warp 0 of CTA 0: %lld cycles for %d ops), kernel %.3f ms\n", G, warps, blocks,
           bad ? "MISMATCH" : "bit-exact", (double)cyc[0] / (rounds * G), cyc[0], rounds * G, ms);
    cudaFree(dt); cudaFree(di); cudaFree(dout); cudaFree(dc);
}

int main()
{
    run<1>(1, 148); run<4>(1, 148); run<8>(1, 148); run<16>(1, 148);
    run<16>(2, 148); run<16>(4, 148); run<16>(2, 296); run<8>(4, 296);
    return 0;
}
