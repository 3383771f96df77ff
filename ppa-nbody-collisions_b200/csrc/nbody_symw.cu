// nbody_symw.cu -- the two-sided force kernel for SMALL n (one GPU, below kSymWarpMaxN bodies): one WARP per work item, no
// shared memory, no barrier; on the cell-sorted order (bounding boxes cull the collision pre-test) or on the bodies' own
// order (every round pre-tested).  Same compile flags as nbody_kernels.cu.
//
// The CTA-per-tile-pair kernel of nbody_sym.cu needs tens of tile pairs per CTA to hide what a tile pair costs to set
// up (row loads, the TMA ring, one barrier phase); at n = 16 384 there are 528 tile pairs for 444 CTAs.  Here the unit
// of work is what one warp does in one ROUND of that kernel: a group of 128 rows (4 per lane) against chunks of 64
// bodies (one pair per lane, handed round the lanes together with its accumulators: the same systolic loop,
// sym_substeps).  Work items are (group g, run of chunks) with chunk index >= 2 g -- the triangle of unordered pairs --
// taken from a queue by every warp on its own, so the 24 warps of an SM are always in different phases and hide each
// other's loads.  The two chunks that make up the group itself are evaluated one-sided (every ordered pair of the
// group is met there on its own; the self pair drops out in the exact path).
//
// Everything arrives by plain loads -- from the body store (float4 {x, y, m, r}) or from the planes of the sorted tiles --
// and leaves through RED.ADD.64 into the fixed-point force sums (nbody_sym.cuh): 4 per lane and round for the chunk's
// bodies, 8 per lane and item for the rows.  A round whose bounding boxes (64-body boxes of the sorted tiles) are apart
// runs the test-free loop; any other round -- every round on the bodies' own order -- carries the collision pre-test:
// pairs that pass it are left out of the packed sums and re-evaluated exactly afterwards, as in the large kernel.
#include <cstdlib>

#include "nbody_sym.cuh"

namespace nb {
namespace {

constexpr int kWThreads = 128;                 // 4 independent warps per CTA
constexpr int kWIpt = kWGroup / 32;            // rows per lane

// the flagged (lane, sub-step) pairs of one round, exactly (see sym_redo in nbody_sym.cu; here radii and indices are at hand)
__device__ __forceinline__ void symw_redo(const DevState &st, const int rank, const int rbase, const int cbase, const bool own,
                                          const float soft2, const unsigned mask, const float2 xs, const float2 ys,
                                          const float2 ms, const float2 rj, const int2 oj2, const int (&oi4)[kWIpt], float2 &gx, float2 &gy,
                                          const float (&nx)[kWIpt], const float (&ny)[kWIpt], const float (&nm)[kWIpt],
                                          const float (&ri)[kWIpt], const float (&thr)[kWIpt], float2 (&tfx)[kWIpt],
                                          float2 (&tfy)[kWIpt], const int lane, unsigned &n_redo)
{
    unsigned any = __reduce_or_sync(0xffffffffu, mask);
#pragma unroll 1
    while (any) {
        const int s = __ffs(any) - 1;
        any &= any - 1u;
        ++n_redo;
        const int p = (lane + s) & 31;
        const float xj[2] = {__shfl_sync(0xffffffffu, xs.x, p), __shfl_sync(0xffffffffu, xs.y, p)};
        const float yj[2] = {__shfl_sync(0xffffffffu, ys.x, p), __shfl_sync(0xffffffffu, ys.y, p)};
        const float mj[2] = {__shfl_sync(0xffffffffu, ms.x, p), __shfl_sync(0xffffffffu, ms.y, p)};
        const float rjj[2] = {__shfl_sync(0xffffffffu, rj.x, p), __shfl_sync(0xffffffffu, rj.y, p)};
        const int ojj[2] = {__shfl_sync(0xffffffffu, oj2.x, p), __shfl_sync(0xffffffffu, oj2.y, p)};   // original indices, < 0: padding
        float gxe[2] = {0.f, 0.f}, gye[2] = {0.f, 0.f};
        if ((mask >> s) & 1u) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int oj = ojj[e], js = cbase + 2 * p + e;
#pragma unroll
                for (int q = 0; q < kWIpt; ++q) {
                    const int oi = oi4[q], is = rbase + 32 * q + lane;
                    const float dx = xj[e] + nx[q], dy = yj[e] + ny[q];
                    const float d2 = fmaf(dx, dx, dy * dy);
                    const float d2s = soft2 > 0.f ? fmaf(dx, dx, fmaf(dy, dy, soft2)) : d2;
                    if (d2s <= thr[q]) {                          // else: the pair was part of the packed sums
                        const float rsum = ri[q] + rjj[e];
                        if (oi < 0 || oj < 0 || is == js) {
                            // padding, or the self pair: nothing
                        } else if (d2 <= rsum * rsum) {           // src/nbody.cu:126-134
                            push_candidate(st, rank, oi, oj);
                            if (!own) push_candidate(st, rank, oj, oi);
                        } else {
                            const float inv = rsqrt_approx(d2s);
                            const float i3 = (inv * inv) * inv;
                            const float sj = i3 * mj[e], si = i3 * nm[q];
                            tfx[q].x = fmaf(dx, sj, tfx[q].x);
                            tfy[q].x = fmaf(dy, sj, tfy[q].x);
                            gxe[e] = fmaf(dx, si, gxe[e]);
                            gye[e] = fmaf(dy, si, gye[e]);
                        }
                    }
                }
            }
        }
        const int from = (lane - s) & 31;
        gx.x += __shfl_sync(0xffffffffu, gxe[0], from);
        gx.y += __shfl_sync(0xffffffffu, gxe[1], from);
        gy.x += __shfl_sync(0xffffffffu, gye[0], from);
        gy.y += __shfl_sync(0xffffffffu, gye[1], from);
    }
}

template <int UNROLL>
__global__ void __launch_bounds__(kWThreads, 6) force_symw_kernel(const DevState st, const StepParams p)
{
    if (st.desc->sym != 2) return;
    const int lane = threadIdx.x & 31;
    const int n = st.desc->n;
    const float rmax = st.desc->rmax, fscale = st.desc->fscale;
    const float2 s2 = make_float2(p.soft2, p.soft2);
    const WGeom w = symw_geom(n, st.desc->sym_S);          // sym_S: chunks per work item on this path (plan)
    const float4 *__restrict__ pm = st.pm;
    const bool sorted = st.desc->sorted != 0;
    const float Rb = sqrtf((4.f * rmax * rmax + p.soft2) * 1.001f);     // no pre-test can pass beyond this separation
    unsigned n_rounds = 0, n_redo = 0, n_culled = 0;

    // Work distribution: ids from an atomic counter, the next one fetched while the current item is being worked on.  The
    // expensive own items have the lowest ids, so they start first and the queue evens out the rest.  (p.symw_queue = 0
    // deals the ids round robin instead, warp w taking w, w + W, ...: a measurement knob.)
    const unsigned total_warps = gridDim.x * (kWThreads / 32), my_warp = blockIdx.x * (kWThreads / 32) + (threadIdx.x >> 5);
    // several GPUs: rank r takes ids r, r + world, ... of the same order
    const bool queue = p.symw_queue != 0;
    const unsigned W = (unsigned)p.world, R = (unsigned)p.rank;
    unsigned next = my_warp * W + R;
    if (queue && lane == 0) next = atomicAdd(&st.res->sym_next, 1u) * W + R;
#pragma unroll 1
    for (;;) {
        const unsigned id = queue ? __shfl_sync(0xffffffffu, next, 0) : next;
        if (id >= (unsigned)w.ids) break;
        if (!queue) next += total_warps * W;
        else if (lane == 0) next = atomicAdd(&st.res->sym_next, 1u) * W + R;
        int g, c_lo, c_hi;
        if (!symw_decode(w, (int)id, g, c_lo, c_hi)) continue;
        const int rbase = kWGroup * g;
        float nx[kWIpt], ny[kWIpt], nm[kWIpt], ri[kWIpt], thr[kWIpt];
        int oi4[kWIpt];
        float2 tfx[kWIpt], tfy[kWIpt];
        // sorted order: rows and chunks are slots of the sorted tiles (5 planes of 512 + 64-body bounding boxes)
        const float *__restrict__ rt = st.jts + (size_t)(rbase / kTJ) * kSortedTileFloats;
#pragma unroll
        for (int q = 0; q < kWIpt; ++q) {
            const int i = rbase + 32 * q + lane;
            float4 b = make_float4(kDummyCoord, kDummyCoord, 0.f, 0.f);
            int o = -1;
            if (sorted) {
                const int t = i & (kTJ - 1);
                o = __float_as_int(rt[4 * kTJ + t]);
                if (o >= 0) b = make_float4(rt[t], rt[kTJ + t], rt[2 * kTJ + t], rt[3 * kTJ + t]);
            } else if (i < n) {
                o = i;
                b = pm[i];
            }
            oi4[q] = o;
            nx[q] = -b.x;
            ny[q] = -b.y;
            nm[q] = -b.z;
            ri[q] = b.w;
            const float rr = b.w + rmax;
            const float bound = p.soft2 > 0.f ? (rr * rr + p.soft2) * 1.000001f : rr * rr;
            thr[q] = o >= 0 ? bound : -1.0f;               // pads never flag
            tfx[q] = make_float2(0.f, 0.f);
            tfy[q] = make_float2(0.f, 0.f);
        }
        float4 rb = make_float4(0.f, 0.f, 0.f, 0.f);       // bounding box of the group's 128 rows
        if (sorted) {
            const float4 *bx = reinterpret_cast<const float4 *>(rt + 5 * kTJ) + 2 * ((rbase & (kTJ - 1)) / kWGroup);
            const float4 b0 = bx[0], b1 = bx[1];
            rb = make_float4(fminf(b0.x, b1.x), fminf(b0.y, b1.y), fmaxf(b0.z, b1.z), fmaxf(b0.w, b1.w));
        }
#pragma unroll 1
        for (int c = c_lo; c < c_hi; ++c) {
            const bool own = (c >> 1) == g;
            const int cbase = kWChunk * c, j0 = cbase + 2 * lane;
            float2 xs, ys, ms, rj;
            int2 oj2;
            bool may_hit = true;
            if (sorted) {
                const float *__restrict__ ct = st.jts + (size_t)(cbase / kTJ) * kSortedTileFloats;
                const int t = j0 & (kTJ - 1);
                xs = *reinterpret_cast<const float2 *>(ct + t);
                ys = *reinterpret_cast<const float2 *>(ct + kTJ + t);
                ms = *reinterpret_cast<const float2 *>(ct + 2 * kTJ + t);
                rj = *reinterpret_cast<const float2 *>(ct + 3 * kTJ + t);
                oj2 = *reinterpret_cast<const int2 *>(ct + 4 * kTJ + t);
                const float4 cb = reinterpret_cast<const float4 *>(ct + 5 * kTJ)[(cbase & (kTJ - 1)) / kWChunk];
                may_hit = __any_sync(0xffffffffu, !((cb.x - rb.z > Rb) | (rb.x - cb.z > Rb) | (cb.y - rb.w > Rb) | (rb.y - cb.w > Rb)));
            } else {
                const float4 pad = make_float4(kPadCoord, kPadCoord, 0.f, 0.f);
                const float4 b0 = j0 < n ? pm[j0] : pad, b1 = j0 + 1 < n ? pm[j0 + 1] : pad;
                xs = make_float2(b0.x, b1.x);
                ys = make_float2(b0.y, b1.y);
                ms = make_float2(b0.z, b1.z);
                rj = make_float2(b0.w, b1.w);
                oj2 = make_int2(j0 < n ? j0 : -1, j0 + 1 < n ? j0 + 1 : -1);
            }
            float2 gx = make_float2(0.f, 0.f), gy = make_float2(0.f, 0.f);
            unsigned mask = 0;
            if (may_hit) {
                sym_substeps<true, kWIpt, UNROLL>(xs, ys, ms, gx, gy, nx, ny, nm, thr, s2, tfx, tfy, mask, lane);
                if (__any_sync(0xffffffffu, mask != 0u))
                    symw_redo(st, p.rank, rbase, cbase, own, p.soft2, mask, xs, ys, ms, rj, oj2, oi4, gx, gy, nx, ny, nm, ri, thr, tfx, tfy, lane, n_redo);
            } else {
                sym_substeps<false, kWIpt, UNROLL>(xs, ys, ms, gx, gy, nx, ny, nm, thr, s2, tfx, tfy, mask, lane);
                ++n_culled;
            }
            ++n_rounds;
            if (!own) {                                    // the chunk's bodies: {gx0, gx1, gy0, gy1} of slots j0, j0 + 1
                long long *dst = st.facc + 2 * (size_t)j0;
                red_add64(dst, to_fixed(gx.x, fscale));
                red_add64(dst + 1, to_fixed(gy.x, fscale));
                red_add64(dst + 2, to_fixed(gx.y, fscale));
                red_add64(dst + 3, to_fixed(gy.y, fscale));
            }
        }
        long long *dst = st.facc + 2 * ((size_t)rbase + lane);
#pragma unroll
        for (int q = 0; q < kWIpt; ++q) {
            red_add64(dst + 64 * q, to_fixed(tfx[q].x + tfx[q].y, fscale));
            red_add64(dst + 64 * q + 1, to_fixed(tfy[q].x + tfy[q].y, fscale));
        }
    }
    if (p.count_stats && lane == 0) {
        atomicAdd(&st.ctr->fast_chunks, (unsigned long long)n_rounds * 2ull);
        atomicAdd(&st.ctr->exact_chunks, (unsigned long long)n_redo);
        atomicAdd(&st.ctr->culled_parts, (unsigned long long)n_culled);
    }
}

}  // namespace

// sub-steps unrolled per loop iteration: 4 by default (profiles/r02_small_n_sweep.md); NBODY_B200_SYMW_UNROLL for measurements
static int symw_unroll()
{
    static int u = 0;
    if (u == 0) {
        const char *e = getenv("NBODY_B200_SYMW_UNROLL");
        const int v = e ? atoi(e) : 4;
        u = (v == 2 || v == 8 || v == 32) ? v : 4;
    }
    return u;
}

cudaError_t launch_force_symw(const DevState &st, const StepParams &p, cudaStream_t s)
{
    switch (symw_unroll()) {
    case 2: force_symw_kernel<2><<<p.symw_grid, kWThreads, 0, s>>>(st, p); break;
    case 8: force_symw_kernel<8><<<p.symw_grid, kWThreads, 0, s>>>(st, p); break;
    case 32: force_symw_kernel<32><<<p.symw_grid, kWThreads, 0, s>>>(st, p); break;
    default: force_symw_kernel<4><<<p.symw_grid, kWThreads, 0, s>>>(st, p); break;
    }
    count_launch();
    return cudaGetLastError();
}

int force_symw_occupancy(int *regs)
{
    int occ = 0;
    cudaFuncAttributes fa = {};
    switch (symw_unroll()) {
    case 2: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, force_symw_kernel<2>, kWThreads, 0); cudaFuncGetAttributes(&fa, force_symw_kernel<2>); break;
    case 8: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, force_symw_kernel<8>, kWThreads, 0); cudaFuncGetAttributes(&fa, force_symw_kernel<8>); break;
    case 32: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, force_symw_kernel<32>, kWThreads, 0); cudaFuncGetAttributes(&fa, force_symw_kernel<32>); break;
    default: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, force_symw_kernel<4>, kWThreads, 0); cudaFuncGetAttributes(&fa, force_symw_kernel<4>); break;
    }
    if (regs) *regs = fa.numRegs;
    return occ;
}

}  // namespace nb
