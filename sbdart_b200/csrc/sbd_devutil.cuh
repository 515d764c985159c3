// Device helpers shared by the register kernels (sbd_fast.cu, sbd_wide.cu).
#pragma once
#include <cuda_runtime.h>

namespace sbd {

#ifndef FULLMASK
#define FULLMASK 0xffffffffu
#endif

__device__ __forceinline__ double shfl_d(double v, int src, int width)
{
    return __shfl_sync(FULLMASK, v, src, width);
}

// value of lane j of a layer group of GW lanes (phase 1 of the register kernels); a group of 10
// lanes (NSTR = 20: three groups per warp) names the source lane in the warp, gbase = first lane
template <int GW>
__device__ __forceinline__ double group_get(double v, int j, int gbase)
{
    if constexpr ((GW & (GW - 1)) == 0) return shfl_d(v, j, GW);
    else return __shfl_sync(FULLMASK, v, gbase + j);
}

// ---- asynchronous global -> shared staging (LDGSTS) ------------------------
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_one() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }
// the warp copies `ndoubles` (even, both sides 16-byte aligned) doubles
__device__ __forceinline__ void warp_copy_async(double *dst, const double *src, int ndoubles, int lane)
{
    for (int i = lane; i < ndoubles / 2; i += 32) cp_async16(dst + 2 * i, src + 2 * i);
}

// ---- bulk asynchronous copies (TMA engine, cp.async.bulk) completing on an mbarrier ------
// One instruction by one lane moves a whole record (global -> shared) instead of one LDGSTS
// per 16 bytes and lane.  Sizes and both addresses must be multiples of 16 bytes.
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *smem, const void *gmem, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem), "r"(bytes),
                   "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
// waits for the phase with the given parity; gives up after ~1 s (returns false)
__device__ __forceinline__ bool mbar_wait(unsigned long long *bar, unsigned parity)
{
    const unsigned b = (unsigned)__cvta_generic_to_shared(bar);
#pragma unroll 1
    for (int spin = 0; spin < (1 << 26); spin++) {
        unsigned done;
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(b), "r"(parity) : "memory");
        if (done) return true;
    }
    return false;
}
// orders this thread's earlier generic-proxy accesses (ordinary loads / stores) before its
// later async-proxy operations (the bulk copies)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

// 1/x to ~1 ulp without the IEEE slow paths of the division operator:
// MUFU.RCP64H seed + two Newton steps (x must be normal and non-zero).
__device__ __forceinline__ double fast_rcp(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    y = fma(y, fma(-x, y, 1.0), y);
    y = fma(y, fma(-x, y, 1.0), y);
    return y;
}

// 1/sqrt(x) to ~1 ulp: MUFU.RSQ64H seed (2^-26) + one third-order correction
// (x must be normal and positive).
__device__ __forceinline__ double fast_rsqrt(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y * y, 1.0);
    return fma(fma(e, 0.375, 0.5), y * e, y);
}

// One-sided Jacobi: a sweep in which every pair's squared cosine stayed below this
// is the last one (tools/accuracy_probe.py: tightening it does not change the result).
#ifdef SBD_JACOBI_BIG
constexpr double kJacobiBig = SBD_JACOBI_BIG;
#else
constexpr double kJacobiBig = 1.0e-10;
#endif


}  // namespace sbd
