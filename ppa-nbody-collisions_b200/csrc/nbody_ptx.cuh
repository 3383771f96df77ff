// nbody_ptx.cuh -- device-side helpers shared by the kernel translation units: mbarrier + 1-D TMA bulk copy,
// approximate reciprocal square root, compensated addition.
#pragma once
#include <cuda_runtime.h>

namespace nb {
namespace {

// ------------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D TMA bulk copy (SASS: SYNCS.*, UBLKCP)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p)
{
    return static_cast<unsigned>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    const unsigned addr = smem_u32(bar);
    unsigned ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity), "r"(0x989680u)      // suspend-time hint (ns): sleep in hardware, do not spin
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ float rsqrt_approx(float x)
{
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// (hi, lo) += x with the rounding error of hi + x captured in lo (Knuth TwoSum; needs -fmad=false and
// no fast-math, which is how this file is compiled)
__device__ __forceinline__ void two_sum(float &hi, float &lo, const float x)
{
    const float s = hi + x;
    const float bb = s - hi;
    const float err = (hi - (s - bb)) + (x - bb);
    hi = s;
    lo += err;
}

}  // namespace
}  // namespace nb
