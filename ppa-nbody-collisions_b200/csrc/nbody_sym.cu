// nbody_sym.cu -- the two-sided (pair-halving) force kernel and, for several GPUs, the threading of the exchanged
// candidate pairs.  Same compile flags as nbody_kernels.cu (-fmad=false: every fused multiply-add is explicit).
#include "nbody_device.cuh"
#include "nbody_ptx.cuh"
#include "nbody_sym.cuh"

namespace nb {
namespace {

// ------------------------------------------------------------------------------------------------
// Two-sided force kernel: every unordered pair is evaluated once and its force goes to both bodies (Newton's
// third law).  The collision predicate is symmetric bit for bit, so one evaluation also serves
// both rows' bookkeeping.
//
// Work: the triangle of tile pairs (I, J), J >= I, of the T 512-body tiles, cut into blocks of S x S tile pairs;
// CTAs take blocks from a queue (rank r of W takes every W-th).  A CTA holds the 512 bodies of tile I as rows
// (4 row groups x 2 warps x 32 lanes x 4 rows) and streams the J tiles through a TMA ring.  A J tile is 8 chunks
// of 64 bodies; in round r = 0..3 warp (k, h) works on chunk 4 h + (k + r) % 4.  Within a round the warp is a
// systolic ring: every lane owns one pair of j bodies plus their j-side accumulators and hands them to its
// neighbour after each of the 32 sub-steps (10 SHFL), while its 4 rows stay put.  12 packed f32x2 operations +
// 2 MUFU per (row, j pair) give four ordered interactions.
//
// Sums.  Forces meet in st.facc: per body two 64-bit FIXED-POINT sums (scale 2^k from the step plan).  Row sums
// of a tile pair are converted once and kept in a thread-private shared-memory accumulator until the CTA leaves
// the row of tile pairs; j-side sums of a tile pair are combined over the four warps that met the chunk (fixed
// order) and added at once.  Both go to L2 with RED.ADD.64.  Integer addition is associative: the total does not
// depend on which CTA or which GPU took which block, so results are reproducible bit for bit and there is no
// per-(body, block) partial array to write and re-read.
//
// Collisions.  Rounds whose bounding boxes may touch (all rounds on the bodies' own order) carry the pre-test
// d2 <= (r_i + r_max)^2: a pair that passes it is left OUT of the packed sums and its sub-step is flagged in a
// 32-bit mask per lane; after the round only the flagged (lane, sub-step) pairs are re-evaluated with the
// reference predicate (src/nbody.cu:126-134) -- a hit becomes a candidate of both rows and gives no force
// (:215-226), anything else gets its force added scalar-wise on both sides.
//
// Synchronisation: one mbarrier phase per tile pair (every warp arrives when its rounds are done); the wait
// sits one round into the NEXT tile pair, so warps may drift by a round without stalling.  What follows the wait
// is everything that needs all warps: the j-side combine of the previous tile pair (double-buffered) and the
// refill of its ring stage.  Thread 0 is the producer: it walks the queue, writes a descriptor per tile pair
// next to the TMA it issues, and the other threads just consume descriptors until the sentinel.
// ------------------------------------------------------------------------------------------------
// Blocks of the pair triangle in queue order: the Q (Q - 1) / 2 full-size blocks (R < C), row by row, then the Q
// half-size diagonal ones (a short tail).  sym_block_index is the inverse of sym_block_decode.
__host__ __device__ inline void sym_block_decode(int b, int Q, int &R, int &C)
{
    const int noff = Q * (Q - 1) / 2;
    if (b >= noff) {
        R = C = b - noff;
        return;
    }
    // rows 0 .. r - 1 hold r (2 Q - 1 - r) / 2 blocks: solve for r, then fix the rounding
    const double q2 = 2.0 * Q - 1.0;
    int r = (int)((q2 - sqrt(q2 * q2 - 8.0 * (double)b)) * 0.5);
    if (r < 0) r = 0;
    if (r > Q - 2) r = Q - 2;
    while (r > 0 && (long long)r * (2 * Q - 1 - r) / 2 > b) --r;
    while ((long long)(r + 1) * (2 * Q - 2 - r) / 2 <= b) ++r;
    R = r;
    C = r + 1 + (b - (int)((long long)r * (2 * Q - 1 - r) / 2));
}
__host__ __device__ inline int sym_block_index(int X, int Y, int Q)
{
    const int R = X < Y ? X : Y, C = X < Y ? Y : X;
    if (R == C) return Q * (Q - 1) / 2 + R;
    return R * (Q - 1) - R * (R - 1) / 2 + (C - R - 1);
}

constexpr int kSymThreads = 256;
constexpr int kSymStageFloats = 3 * kTJ + 64;            // x, y, m planes + 9 float4 boxes (padded to 128 B)
constexpr int kSymStageBytes = kSymStageFloats * 4;      // 6400
constexpr int kSymPlaneBytes = 3 * kTJ * 4;              // 6144
constexpr int kSymBoxBytes = 4 * (kTJ / kSubPart + 1) * 4;   // 144
// Geometry for IPT rows per lane: the 512 rows are GROUPS = 16 / IPT row groups of 32 IPT rows; the HSPLIT = 8 / GROUPS warps
// of a group share its rows and split the 8 chunks of a J tile: ROUNDS = GROUPS chunks each, one per round; in round r
// warp (k, h) works on chunk ROUNDS h + (k + r) % ROUNDS, so the 8 warps are always on 8 different chunks.
//   IPT = 4: 4 groups x 2 warps, 4 rounds, <= 80 registers, 3 CTAs per SM  (10 SHFL per 4 rows and sub-step)
//   IPT = 8: 2 groups x 4 warps, 2 rounds, <= 128 registers, 2 CTAs per SM (10 SHFL per 8 rows and sub-step)
// dynamic shared memory: ring | j-side round results [2][8 warps][ROUNDS][32 lanes] float4 | row accumulators
// [IPT rows][2][256 threads] int64
template <int IPT>
struct SymGeom {
    static constexpr int kGroups = 16 / IPT, kHsplit = 8 / kGroups, kRounds = kGroups;
    static constexpr int kGprivBytes = 2 * 8 * kRounds * 32 * 16;
    static constexpr int kAccBytes = IPT * 2 * kSymThreads * 8;
    static constexpr int kDynSmem = kStages * kSymStageBytes + kGprivBytes + kAccBytes;     // 74752 either way
    static constexpr int kMinBlocks = IPT == 4 ? 3 : 2;
};

constexpr unsigned kOwn = 1u, kFirst = 2u, kLast = 4u;   // descriptor flags; bits 4..6 r0, bits 8..10 r1

struct SymProducer {              // thread 0's walk over the work queue (shared memory, used by thread 0 only)
    int have;                     // inside an item
    int I, J, I1, J0, J1, diag, r0, r1;
};

// The flagged (lane, sub-step) pairs of one round, re-evaluated with the reference predicate.  The ring is home again:
// lane p holds j pair p and its accumulators; in sub-step s lane l met j pair (l + s) % 32.  Ig / Jg: the tiles in
// global memory (radius and original-index planes are not staged); `own`: rows and chunk come from the same tile --
// every ordered pair is met there on its own, so only the row side counts and the self pair is skipped.
template <int IPT>
__device__ __forceinline__ void sym_redo(const DevState &st, const float *__restrict__ Ig, const float *__restrict__ Jg,
                                         const int ibase, const int jbase, const int n, const bool sorted, const int c,
                                         const int k, const bool own, const float soft2, const int rank, const unsigned mask,
                                         const float2 xs, const float2 ys, const float2 ms, float2 &gx, float2 &gy,
                                         const float (&nx)[IPT], const float (&ny)[IPT], const float (&nm)[IPT],
                                         const float (&thr)[IPT], float2 (&tfx)[IPT], float2 (&tfy)[IPT],
                                         const int lane, unsigned &n_redo)
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
        float gxe[2] = {0.f, 0.f}, gye[2] = {0.f, 0.f};
        if ((mask >> s) & 1u) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int js = 64 * c + 2 * p + e;
                const float rj = Jg[3 * kTJ + js];
                const int oj = sorted ? __float_as_int(Jg[4 * kTJ + js]) : (jbase + js < n ? jbase + js : -1);
#pragma unroll
                for (int q = 0; q < IPT; ++q) {
                    const int rs = 32 * IPT * k + 32 * q + lane;
                    const float dx = xj[e] + nx[q], dy = yj[e] + ny[q];
                    const float d2 = fmaf(dx, dx, dy * dy);
                    const float d2s = soft2 > 0.f ? fmaf(dx, dx, fmaf(dy, dy, soft2)) : d2;
                    if (d2s <= thr[q]) {                          // else: the pair was part of the packed sums
                        const float ri = Ig[3 * kTJ + rs];
                        const int oi = sorted ? __float_as_int(Ig[4 * kTJ + rs]) : (ibase + rs < n ? ibase + rs : -1);
                        const float rsum = ri + rj;
                        if (oi < 0 || oj < 0 || (own && js == rs)) {
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
        // the j-side corrections go home: lane p receives what lane (p - s) % 32 found for j pair p
        const int from = (lane - s) & 31;
        gx.x += __shfl_sync(0xffffffffu, gxe[0], from);
        gx.y += __shfl_sync(0xffffffffu, gxe[1], from);
        gy.x += __shfl_sync(0xffffffffu, gye[0], from);
        gy.y += __shfl_sync(0xffffffffu, gye[1], from);
    }
}

template <int IPT>
__global__ void __launch_bounds__(kSymThreads, SymGeom<IPT>::kMinBlocks) force_sym_kernel(const DevState st, const StepParams p)
{
    using G = SymGeom<IPT>;
    constexpr int HSPLIT = G::kHsplit, ROUNDS = G::kRounds;
    extern __shared__ __align__(128) unsigned char sym_dyn[];
    __shared__ __align__(8) unsigned long long full_bar[kStages];
    __shared__ __align__(8) unsigned long long done_bar;
    __shared__ int4 s_desc[kStages];              // per ring stage: {I, J, flags, -}; I < 0: end of work
    __shared__ SymProducer s_prod;
    __shared__ float4 s_rb[kSymThreads / 32];     // per warp: bounding box of its 128 rows (sorted order)
    if (st.desc->sym != 1) return;
    float *ring = reinterpret_cast<float *>(sym_dyn);
    float4 *gpriv = reinterpret_cast<float4 *>(sym_dyn + kStages * kSymStageBytes);
    long long *acc_s = reinterpret_cast<long long *>(sym_dyn + kStages * kSymStageBytes + G::kGprivBytes);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int k = warp / HSPLIT, h = warp % HSPLIT;   // row group, share of the chunks
    const bool sorted = st.desc->sorted != 0;
    const int n = st.desc->n;
    const int tile_floats = sorted ? kSortedTileFloats : kTileFloats;
    const float *__restrict__ jsrc = sorted ? st.jts : st.jt;
    const float rmax = st.desc->rmax;
    const float fscale = st.desc->fscale;
    const float2 s2 = make_float2(p.soft2, p.soft2);
    const float Rb = sqrtf((4.f * rmax * rmax + p.soft2) * 1.001f);     // no pre-test can pass beyond this separation

    // ---- producer (thread 0) ----------------------------------------------------------------------------
    unsigned next_item = 0;                       // the item after the current one, fetched one item ahead
    const unsigned n_items = (unsigned)st.desc->sym_items;
    auto fetch = [&]() -> unsigned {
        const unsigned long long b = (unsigned long long)atomicAdd(&st.res->sym_next, 1u) * (unsigned)p.world + (unsigned)p.rank;
        return b < n_items ? (unsigned)b : 0xffffffffu;
    };
    auto produce = [&](const int stage) {         // thread 0: descriptor + TMA of the next tile pair into `stage`
        SymProducer &P = s_prod;
        if (!P.have) {
            const unsigned item = next_item;
            if (item == 0xffffffffu) {            // queue drained: sentinel (completes the stage's phase without data)
                s_desc[stage] = make_int4(-1, -1, 0, 0);
                mbar_arrive(&full_bar[stage]);
                return;
            }
            next_item = fetch();
            const int lgu = st.desc->sym_lgu, S = st.desc->sym_S, T = st.desc->n_jtiles;
            int R, C;
            sym_block_decode((int)(item >> lgu), st.desc->sym_Q, R, C);
            const int u = (int)(item & ((1u << lgu) - 1u)), rounds = ROUNDS >> lgu;
            P.diag = R == C;
            P.I = R * S;
            P.I1 = min(P.I + S, T);
            P.J0 = C * S;
            P.J1 = min(P.J0 + S, T);
            P.J = P.diag ? P.I : P.J0;
            P.r0 = u * rounds;
            P.r1 = P.r0 + rounds;
            P.have = 1;
        }
        const int I = P.I, J = P.J;
        const unsigned flags = (I == J ? kOwn : 0u) | (J == (P.diag ? I : P.J0) ? kFirst : 0u) | (J == P.J1 - 1 ? kLast : 0u) |
                               ((unsigned)P.r0 << 4) | ((unsigned)P.r1 << 8);
        s_desc[stage] = make_int4(I, J, (int)flags, 0);
        float *dst = ring + stage * kSymStageFloats;
        const float *src = jsrc + (size_t)J * tile_floats;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (sorted) {
            mbar_expect_tx(&full_bar[stage], kSymPlaneBytes + kSymBoxBytes);
            bulk_g2s(dst, src, kSymPlaneBytes, &full_bar[stage]);
            bulk_g2s(dst + 3 * kTJ, src + 5 * kTJ, kSymBoxBytes, &full_bar[stage]);
        } else {
            mbar_expect_tx(&full_bar[stage], kSymPlaneBytes);
            bulk_g2s(dst, src, kSymPlaneBytes, &full_bar[stage]);
        }
        if (++P.J == P.J1) {
            ++P.I;
            P.J = P.diag ? P.I : P.J0;
            if (P.I == P.I1) P.have = 0;
        }
    };

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) mbar_init(&full_bar[s], 1);
        mbar_init(&done_bar, kSymThreads / 32);   // one arrival per warp: every arrival wakes the warps asleep on the barrier
        fence_barrier_init();
        s_prod.have = 0;
        next_item = fetch();
#pragma unroll 1
        for (int s = 0; s < kStages; ++s) produce(s);
    }
#pragma unroll
    for (int q = 0; q < 2 * IPT; ++q) acc_s[q * kSymThreads + tid] = 0;
    __syncthreads();

    float nx[IPT], ny[IPT], nm[IPT];
    float2 tfx[IPT], tfy[IPT];
#pragma unroll
    for (int q = 0; q < IPT; ++q) {
        nx[q] = ny[q] = nm[q] = 0.f;
        tfx[q] = make_float2(0.f, 0.f);
        tfy[q] = make_float2(0.f, 0.f);
    }
    int prevJ = -1;                               // the previous tile pair: its j side is combined one round into this one
    unsigned prev_flags = kOwn;
    unsigned n_rounds = 0, n_redo = 0, n_culled = 0;

    // everything that needs all warps to be done with tile pair t - 1
    auto late = [&](const unsigned t) {
        if (t == 0) return;
        const unsigned tp = t - 1;
        mbar_wait(&done_bar, tp & 1u);
        if (tid == 0) produce((int)(tp % kStages));
        if (!(prev_flags & kOwn)) {
            // j side of tile pair t - 1: body pair `lane` of chunk `warp` of tile prevJ <- the GROUPS warps that met it
            const int pr0 = (prev_flags >> 4) & 7, pr1 = (prev_flags >> 8) & 7;
            const float4 *g = gpriv + (size_t)(tp & 1u) * (8 * ROUNDS * 32);
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int kk = 0; kk < G::kGroups; ++kk) {
                const int r = ((warp % ROUNDS) - kk + ROUNDS) % ROUNDS;
                if (r >= pr0 && r < pr1) {
                    const float4 v = g[((HSPLIT * kk + warp / ROUNDS) * ROUNDS + r) * 32 + lane];
                    a = make_float4(a.x + v.x, a.y + v.y, a.z + v.z, a.w + v.w);
                }
            }
            long long *dst = st.facc + 2 * ((size_t)prevJ * kTJ + 64 * warp + 2 * lane);      // {gx0, gx1, gy0, gy1}
            red_add64(dst, to_fixed(a.x, fscale));
            red_add64(dst + 1, to_fixed(a.z, fscale));
            red_add64(dst + 2, to_fixed(a.y, fscale));
            red_add64(dst + 3, to_fixed(a.w, fscale));
        }
    };

    unsigned t = 0;
#pragma unroll 1
    for (;; ++t) {
        const int stage = (int)(t % kStages);
        mbar_wait(&full_bar[stage], (t / kStages) & 1u);
        const int4 desc = s_desc[stage];
        const int I = desc.x, J = desc.y;
        const unsigned flags = (unsigned)desc.z;
        if (I < 0) break;
        const float *tl = ring + stage * kSymStageFloats;
        const float *__restrict__ Ig = jsrc + (size_t)I * tile_floats;
        const bool own = flags & kOwn;
        const int r0 = (flags >> 4) & 7, r1 = (flags >> 8) & 7;
        if (flags & kFirst) {                     // a new row of tile pairs: this warp's rows
#pragma unroll
            for (int q = 0; q < IPT; ++q) {
                const int rs = 32 * IPT * k + 32 * q + lane;
                nx[q] = -Ig[rs];
                ny[q] = -Ig[kTJ + rs];
                nm[q] = -Ig[2 * kTJ + rs];
            }
            if (sorted) {
                const float4 *bx = reinterpret_cast<const float4 *>(Ig + 5 * kTJ);
                constexpr int BOXES = 32 * IPT / kSubPart;            // 64-body boxes per row group
                float4 b = bx[BOXES * k];
#pragma unroll
                for (int e = 1; e < BOXES; ++e) {
                    const float4 o = bx[BOXES * k + e];
                    b = make_float4(fminf(b.x, o.x), fminf(b.y, o.y), fmaxf(b.z, o.z), fmaxf(b.w, o.w));
                }
                __syncwarp();
                if (lane == 0) s_rb[warp] = b;
                __syncwarp();
            }
        }
        float4 *gp = gpriv + ((size_t)(t & 1u) * 8 + warp) * (ROUNDS * 32) + lane;
#pragma unroll 1
        for (int r = r0; r < r1; ++r) {
            const int c = ROUNDS * h + ((k + r) % ROUNDS);
            bool may_hit = true;
            if (sorted) {
                const float4 cb = reinterpret_cast<const float4 *>(tl + 3 * kTJ)[c], rb = s_rb[warp];
                // (the vote only tells the compiler what is true anyway: the box test is the same in every lane)
                may_hit = __any_sync(0xffffffffu, !((cb.x - rb.z > Rb) | (rb.x - cb.z > Rb) | (cb.y - rb.w > Rb) | (rb.y - cb.w > Rb)));
            }
            float2 xs = *reinterpret_cast<const float2 *>(tl + 64 * c + 2 * lane);
            float2 ys = *reinterpret_cast<const float2 *>(tl + kTJ + 64 * c + 2 * lane);
            float2 ms = *reinterpret_cast<const float2 *>(tl + 2 * kTJ + 64 * c + 2 * lane);
            float2 gx = make_float2(0.f, 0.f), gy = make_float2(0.f, 0.f);
            ++n_rounds;
            if (may_hit) {
                float thr[IPT];
#pragma unroll
                for (int q = 0; q < IPT; ++q) {
                    const int rs = 32 * IPT * k + 32 * q + lane;
                    const float rr = Ig[3 * kTJ + rs] + rmax;
                    const float bound = p.soft2 > 0.f ? (rr * rr + p.soft2) * 1.000001f : rr * rr;
                    const bool real = sorted ? __float_as_int(Ig[4 * kTJ + rs]) >= 0 : I * kTJ + rs < n;
                    thr[q] = real ? bound : -1.0f;                    // pads never flag
                }
                unsigned mask = 0;
                sym_substeps<true, IPT>(xs, ys, ms, gx, gy, nx, ny, nm, thr, s2, tfx, tfy, mask, lane);
                if (__any_sync(0xffffffffu, mask != 0u))
                    sym_redo<IPT>(st, Ig, jsrc + (size_t)J * tile_floats, I * kTJ, J * kTJ, n, sorted, c, k, own, p.soft2, p.rank,
                             mask, xs, ys, ms, gx, gy, nx, ny, nm, thr, tfx, tfy, lane, n_redo);
            } else {
                const float thr[IPT] = {};
                unsigned mask = 0;
                sym_substeps<false, IPT>(xs, ys, ms, gx, gy, nx, ny, nm, thr, s2, tfx, tfy, mask, lane);
                ++n_culled;
            }
            if (r == r0) late(t);                 // one round of slack for the slowest warp of the previous tile pair
            gp[r * 32] = make_float4(gx.x, gx.y, gy.x, gy.y);
        }
        // bank this tile pair's row sums: exact from here on
#pragma unroll
        for (int q = 0; q < IPT; ++q) {
            acc_s[(2 * q) * kSymThreads + tid] += to_fixed(tfx[q].x + tfx[q].y, fscale);
            acc_s[(2 * q + 1) * kSymThreads + tid] += to_fixed(tfy[q].x + tfy[q].y, fscale);
            tfx[q] = make_float2(0.f, 0.f);
            tfy[q] = make_float2(0.f, 0.f);
        }
        if (flags & kLast) {                      // leaving this row of tile pairs: the rows' sums go to the global accumulators
            long long *dst = st.facc + 2 * ((size_t)I * kTJ + 32 * IPT * k + lane);
#pragma unroll
            for (int q = 0; q < IPT; ++q) {
                red_add64(dst + 64 * q, acc_s[(2 * q) * kSymThreads + tid]);
                red_add64(dst + 64 * q + 1, acc_s[(2 * q + 1) * kSymThreads + tid]);
                acc_s[(2 * q) * kSymThreads + tid] = 0;
                acc_s[(2 * q + 1) * kSymThreads + tid] = 0;
            }
        }
        prevJ = J;
        prev_flags = flags;
        __syncwarp();                             // the lanes' round results of tile pair t are in place ...
        if (lane == 0) mbar_arrive(&done_bar);    // ... and published by one release per warp
    }
    late(t);                                      // j side of the last tile pair
    if (p.count_stats && lane == 0) {
        atomicAdd(&st.ctr->fast_chunks, (unsigned long long)n_rounds * 2ull);   // a round = 2 sub-chunks of 32 bodies
        atomicAdd(&st.ctr->exact_chunks, (unsigned long long)n_redo);           // sub-steps re-evaluated exactly
        atomicAdd(&st.ctr->culled_parts, (unsigned long long)n_culled);
    }
}

// Sharded two-sided kernel, after the allgather of xbuf: thread every rank's candidate pairs whose row this rank
// finishes into the rows' chains
__global__ void __launch_bounds__(256) sym_chain_kernel(const DevState st, const StepParams p)
{
    const StepDesc &d = *st.desc;
    if (!d.sym) return;
    unsigned mine = 0;
    for (int r = 0; r < p.world; ++r) {
        const unsigned found = x_header(st, r)->count;
        if (found > (unsigned)st.x_cap && blockIdx.x == 0 && threadIdx.x == 0) st.ctr->overflow_flag = 1;   // every rank reports it
        const unsigned cnt = min(found, (unsigned)st.x_cap);
        const int2 *src = x_pairs(st, r);
        for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < cnt; e += gridDim.x * blockDim.x) {
            const int2 pr = src[e];
            const int slot = d.sorted ? st.sinv[pr.x] : pr.x;
            if (slot < d.row_lo || slot >= d.row_hi) continue;
            const unsigned idx = (unsigned)r * (unsigned)st.x_cap + e;      // cand holds world * x_cap entries
            const int prev = atomicExch(&st.head[pr.x], (int)idx);
            st.cand[idx] = make_int2(pr.y, prev);
            ++mine;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&st.ctr->cand_count, mine);
}

}  // namespace

cudaError_t launch_force_sym(const DevState &st, const StepParams &p, cudaStream_t s)
{
    if (p.sym_rows == 8)
        force_sym_kernel<8><<<p.sym_grid, kSymThreads, SymGeom<8>::kDynSmem, s>>>(st, p);
    else
        force_sym_kernel<4><<<p.sym_grid, kSymThreads, SymGeom<4>::kDynSmem, s>>>(st, p);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_sym_chain(const DevState &st, const StepParams &p, cudaStream_t s)
{
    sym_chain_kernel<<<296, 256, 0, s>>>(st, p);
    count_launch();
    return cudaGetLastError();
}

template <int IPT>
static int sym_occupancy(int *regs)
{
    int occ = 0;
    cudaFuncAttributes fa = {};
    cudaFuncSetAttribute(force_sym_kernel<IPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SymGeom<IPT>::kDynSmem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, force_sym_kernel<IPT>, kSymThreads, SymGeom<IPT>::kDynSmem);
    cudaFuncGetAttributes(&fa, force_sym_kernel<IPT>);
    if (regs) *regs = fa.numRegs;
    return occ;
}

int force_sym_occupancy(int rows, int *regs)
{
    return rows == 8 ? sym_occupancy<8>(regs) : sym_occupancy<4>(regs);
}

int force_sym_rounds(int rows) { return rows == 8 ? SymGeom<8>::kRounds : SymGeom<4>::kRounds; }

void sym_block_host(int b, int Q, int *R, int *C)
{
    sym_block_decode(b, Q, *R, *C);
}

int sym_block_index_host(int X, int Y, int Q)
{
    return sym_block_index(X, Y, Q);
}

}  // namespace nb
