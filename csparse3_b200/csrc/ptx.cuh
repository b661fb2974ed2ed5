// ptx.cuh -- the inline-PTX memory helpers shared by the LU kernels that address shared memory by 32-bit shared-window
// addresses (lu_rowlane.cu, lu_panel.cu).  All of them are volatile: the kernels rely on program order between them.
#pragma once
#include <cuda_runtime.h>

namespace csp3 {
namespace ptx {

__device__ __forceinline__ double2 lds_d2(unsigned a)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ uint4 lds_u4(unsigned a)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ unsigned lds_u32(unsigned a)
{
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_d2(unsigned a, double2 v)
{
    asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(a), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ double2 ldg_d2(const void *p)
{
    double2 v;
    asm volatile("ld.global.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ldg_nc_f64(const void *p)
{
    double v;
    asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void stg_d2(void *p, double2 v) { asm volatile("st.global.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory"); }
__device__ __forceinline__ void stg_cs_d2(void *p, double2 v) { asm volatile("st.global.cs.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory"); }
__device__ __forceinline__ void cp_async16(unsigned dst, const void *src) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_async8(unsigned dst, const void *src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

}  // namespace ptx
}  // namespace csp3
