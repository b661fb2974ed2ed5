// Microbenchmark: does prefetch.global.L1 / L2 shorten a later dependent load?  One warp, cold lines each time.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void pf_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void pf_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__global__ void probe(const double *buf, long stride, int mode, int delay, long long *out, double *sink)
{
    long long tot = 0; double acc = 0;
    for (int it = 0; it < 64; ++it) {
        const double *p = buf + (long)it * stride + threadIdx.x;          // a fresh 256-byte region per iteration
        if (mode == 1) pf_l1(p);
        if (mode == 2) pf_l2(p);
        if (mode == 3) acc += *p;                                           // warm with a real load
        long long t0 = clock64();
        while (clock64() - t0 < delay) {}
        __syncwarp();
        long long t1 = clock64();
        double v = (mode == 4) ? __ldg(p) : *(volatile const double *)p;
        acc += v;
        long long t2 = clock64();
        if (acc == 12345.678) printf("x");
        tot += t2 - t1;
    }
    if (threadIdx.x == 0) { out[0] = tot / 64; sink[0] = acc; }
}
int main()
{
    const long stride = 1 << 20;                      // 8 MB apart
    double *buf; long long *out; double *sink;
    cudaMalloc(&buf, 64 * stride * 8 + 4096); cudaMemset(buf, 0, 64 * stride * 8 + 4096);
    cudaMalloc(&out, 8); cudaMalloc(&sink, 8);
    const char *names[] = {"no prefetch", "prefetch.L1", "prefetch.L2", "warm by load", "no prefetch (ldg)"};
    for (int mode = 0; mode < 5; ++mode)
        for (int delay : {0, 2000}) {
            // flush L2 with a big memset so every iteration starts cold
            cudaMemset(buf, 0, 64 * stride * 8);
            probe<<<1, 32>>>(buf, stride, mode, delay, out, sink);
            long long h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
            printf("%-20s delay %5d cycles -> load latency %lld cycles\n", names[mode], delay, h);
        }
    return 0;
}
