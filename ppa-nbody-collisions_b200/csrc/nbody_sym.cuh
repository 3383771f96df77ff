// nbody_sym.cuh -- device code shared by the two two-sided force kernels (nbody_sym.cu: one CTA per tile pair, TMA ring,
// cell-sorted order; nbody_symw.cu: one warp per work item, bodies' own order): the systolic sub-step loop, the
// fixed-point force sums and the candidate list.
#pragma once
#include "nbody_device.cuh"
#include "nbody_ptx.cuh"

namespace nb {
namespace {

__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void red_add64(long long *addr, long long v)
{
    asm volatile("red.global.add.u64 [%0], %1;" ::"l"(addr), "l"(v) : "memory");
}

// fixed-point image of a float sum: v * 2^k rounded to the nearest integer (the product is exact)
__device__ __forceinline__ long long to_fixed(float v, float fscale) { return __float2ll_rn(v * fscale); }

// UNROLL: sub-steps per loop iteration.  The CTA-level kernel unrolls all 32 (its 8 warps run in step and share the
// fetched instructions; the back-edge of a partially unrolled loop costs register moves).  The warp-level kernel must
// not: its 24 warps per SM are all in different places, and 47 KB of straight-line code per round thrash the
// instruction cache (ncu: 3 warps stalled on "no instruction" per issue, profiles/r02_force_16k_warp_level_unrolled_ncu.json).
template <bool TEST, int IPT, int UNROLL = 32>
__device__ __forceinline__ void sym_substeps(float2 &xs, float2 &ys, float2 &ms, float2 &gx, float2 &gy,
                                             const float (&nx)[IPT], const float (&ny)[IPT],
                                             const float (&nm)[IPT], const float (&thr)[IPT], const float2 s2,
                                             float2 (&tfx)[IPT], float2 (&tfy)[IPT], unsigned &mask, const int lane)
{
    const int src = (lane + 1) & 31;
#pragma unroll UNROLL
    for (int s = 0; s < 32; ++s) {
        bool flagged = false;
#pragma unroll
        for (int q = 0; q < IPT; ++q) {
            const float2 dx = __fadd2_rn(xs, make_float2(nx[q], nx[q]));
            const float2 dy = __fadd2_rn(ys, make_float2(ny[q], ny[q]));
            const float2 d2 = __ffma2_rn(dx, dx, __ffma2_rn(dy, dy, s2));
            const float2 inv = make_float2(rsqrt_approx(d2.x), rsqrt_approx(d2.y));
            float2 i3 = __fmul2_rn(__fmul2_rn(inv, inv), inv);
            if (TEST) {                           // a pair that passes the pre-test stays out of the sums
                const bool f0 = d2.x <= thr[q], f1 = d2.y <= thr[q];
                i3.x = f0 ? 0.f : i3.x;
                i3.y = f1 ? 0.f : i3.y;
                flagged |= f0 | f1;
            }
            const float2 sj = __fmul2_rn(i3, ms);
            const float2 si = __fmul2_rn(i3, make_float2(nm[q], nm[q]));
            tfx[q] = __ffma2_rn(dx, sj, tfx[q]);
            tfy[q] = __ffma2_rn(dy, sj, tfy[q]);
            gx = __ffma2_rn(dx, si, gx);
            gy = __ffma2_rn(dy, si, gy);
        }
        if (TEST) mask |= flagged ? (1u << s) : 0u;
        xs.x = __shfl_sync(0xffffffffu, xs.x, src);
        xs.y = __shfl_sync(0xffffffffu, xs.y, src);
        ys.x = __shfl_sync(0xffffffffu, ys.x, src);
        ys.y = __shfl_sync(0xffffffffu, ys.y, src);
        ms.x = __shfl_sync(0xffffffffu, ms.x, src);
        ms.y = __shfl_sync(0xffffffffu, ms.y, src);
        gx.x = __shfl_sync(0xffffffffu, gx.x, src);
        gx.y = __shfl_sync(0xffffffffu, gx.y, src);
        gy.x = __shfl_sync(0xffffffffu, gy.x, src);
        gy.y = __shfl_sync(0xffffffffu, gy.y, src);
    }
}

__device__ __forceinline__ void push_candidate(const DevState &st, const int rank, const int row, const int partner)
{
    if (st.xbuf) {                                // sharded: the pair travels to every rank (sym_chain_kernel threads it)
        const unsigned idx = atomicAdd(&x_header(st, rank)->count, 1u);
        if (idx < (unsigned)st.x_cap) {
            x_pairs(st, rank)[idx] = make_int2(row, partner);
        } else {
            st.ctr->overflow_flag = 1;
        }
        return;
    }
    const unsigned idx = atomicAdd(&st.ctr->cand_count, 1u);
    if (idx < (unsigned)st.cand_cap) {
        const int prev = atomicExch(&st.head[row], (int)idx);
        st.cand[idx] = make_int2(partner, prev);
    } else {
        st.ctr->overflow_flag = 1;
    }
}

}  // namespace
}  // namespace nb
