// nbody_sym.cu -- the two-sided (pair-halving) force kernel of the cell-sorted order and, for several GPUs, the
// reduction of its partial forces and the threading of the exchanged candidate pairs.  Same compile flags as
// nbody_kernels.cu (-fmad=false: every fused multiply-add is explicit).
#include "nbody_device.cuh"
#include "nbody_ptx.cuh"

namespace nb {
namespace {

// ------------------------------------------------------------------------------------------------
// two-sided force kernel (cell-sorted order only): every unordered pair is evaluated once and its force goes
// to both bodies (Newton's third law).  The collision predicate is symmetric bit for bit, so
// one evaluation also serves both rows' bookkeeping.
//
// Work: the triangle of tile pairs (I, J), J >= I, of the T sorted tiles, cut into blocks of S x S tile pairs
// (S = ceil(T / 256)); CTAs take blocks from a queue.  A CTA holds the 512 bodies of tile I as rows (4 warp
// pairs x 32 lanes x 4 rows) and streams the J tiles through the TMA ring.  A J tile is 8 chunks of 64 bodies;
// in round r = 0..3 warp (k, h) works on chunk 4 h + (k + r) % 4, so all 8 warps are on different chunks and
// the chunk's j-side sums in shared memory have one writer at a time (block barrier between rounds).
// Within a round the warp is a systolic ring: every lane owns one pair of j bodies plus their j-side
// accumulators and hands them to its neighbour after each of the 32 sub-steps (10 SHFL), while its 4 rows stay
// put.  12 packed f32x2 operations + 2 MUFU per (row, j pair) give four ordered interactions.
// Results go to part[Y][slot] (Y = super-tile of the other side); each entry has exactly one writer block and
// a fixed summation order, so the forces do not depend on which CTA took which block.
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
    int r = 0, rem = b;
    while (rem >= Q - 1 - r) {
        rem -= Q - 1 - r;
        ++r;
    }
    R = r;
    C = r + 1 + rem;
}
__host__ __device__ inline int sym_block_index(int X, int Y, int Q)
{
    const int R = X < Y ? X : Y, C = X < Y ? Y : X;
    if (R == C) return Q * (Q - 1) / 2 + R;
    return R * (Q - 1) - R * (R - 1) / 2 + (C - R - 1);
}

#ifndef NB_SYM_UNROLL
#define NB_SYM_UNROLL 32
#endif
constexpr int kSymUnroll = NB_SYM_UNROLL;   // sub-steps per iteration of the ring loop
constexpr int kSymThreads = 256;
constexpr int kSymDynSmem = kStages * kSortedTileFloats * 4;

template <bool TEST, int IPT>
__device__ __forceinline__ void sym_substeps(float2 &xs, float2 &ys, float2 &ms, float2 &gx, float2 &gy,
                                             const float (&nx)[IPT], const float (&ny)[IPT],
                                             const float (&nm)[IPT], const float (&thr)[IPT], const float2 s2,
                                             float2 (&tfx)[IPT], float2 (&tfy)[IPT], bool &cand, const int lane)
{
    const int src = (lane + 1) & 31;
#pragma unroll kSymUnroll
    for (int s = 0; s < 32; ++s) {
#pragma unroll
        for (int q = 0; q < IPT; ++q) {
            const float2 dx = __fadd2_rn(xs, make_float2(nx[q], nx[q]));
            const float2 dy = __fadd2_rn(ys, make_float2(ny[q], ny[q]));
            const float2 d2 = __ffma2_rn(dx, dx, __ffma2_rn(dy, dy, s2));
            if (TEST) {
                cand |= (d2.x <= thr[q]);
                cand |= (d2.y <= thr[q]);
            }
            const float2 inv = make_float2(rsqrt_approx(d2.x), rsqrt_approx(d2.y));
            const float2 i3 = __fmul2_rn(__fmul2_rn(inv, inv), inv);
            const float2 sj = __fmul2_rn(i3, ms);
            const float2 si = __fmul2_rn(i3, make_float2(nm[q], nm[q]));
            tfx[q] = __ffma2_rn(dx, sj, tfx[q]);
            tfy[q] = __ffma2_rn(dy, sj, tfy[q]);
            gx = __ffma2_rn(dx, si, gx);
            gy = __ffma2_rn(dy, si, gy);
        }
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

// One round redone with the reference predicate (src/nbody.cu:126-134): pairs that hit give no force to either
// body (:215-226) and become candidates of both rows; `own_tile`: rows and chunk come from the same tile, every
// ordered pair is met there on its own, so only the row side counts and the self pair is skipped.
template <int IPT>
__device__ __noinline__ void sym_exact_round(const DevState &st, const float *tl, const float *rows, const int c,
                                             const int k, const bool own_tile, const float soft2, const int rank,
                                             float4 (*acc_s)[kSymThreads], float4 *gacc)
{
    const int lane = threadIdx.x & 31;
    const int src = (lane + 1) & 31;
    float xi[IPT], yi[IPT], mi[IPT], ri[IPT];
    float2 tfx[IPT], tfy[IPT];
    int oi[IPT];
#pragma unroll
    for (int q = 0; q < IPT; ++q) {
        const int rs = 32 * IPT * k + 32 * q + lane;
        xi[q] = rows[rs];
        yi[q] = rows[kTJ + rs];
        mi[q] = rows[2 * kTJ + rs];
        ri[q] = rows[3 * kTJ + rs];
        oi[q] = __float_as_int(rows[4 * kTJ + rs]);
        tfx[q] = make_float2(0.f, 0.f);
        tfy[q] = make_float2(0.f, 0.f);
    }
    float2 xs = *reinterpret_cast<const float2 *>(tl + 64 * c + 2 * lane);
    float2 ys = *reinterpret_cast<const float2 *>(tl + kTJ + 64 * c + 2 * lane);
    float2 ms = *reinterpret_cast<const float2 *>(tl + 2 * kTJ + 64 * c + 2 * lane);
    float2 gx = make_float2(0.f, 0.f), gy = make_float2(0.f, 0.f);
#pragma unroll 1
    for (int s = 0; s < 32; ++s) {
        const int jp = (lane + s) & 31;           // the lane this j pair started on
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int js = 64 * c + 2 * jp + e;
            const float xj = e ? xs.y : xs.x, yj = e ? ys.y : ys.x, mj = e ? ms.y : ms.x;
            const float rj = tl[3 * kTJ + js];
            const int oj = __float_as_int(tl[4 * kTJ + js]);
            float gxe = 0.f, gye = 0.f;
#pragma unroll
            for (int q = 0; q < IPT; ++q) {
                const bool valid = (oi[q] >= 0) & (oj >= 0) & !(own_tile & (js == 32 * IPT * k + 32 * q + lane));
                const float dx = xj - xi[q], dy = yj - yi[q];
                const float d2 = fmaf(dx, dx, dy * dy);
                const float rs = ri[q] + rj;
                const bool hit = d2 <= rs * rs;
                if (valid && hit) {
                    push_candidate(st, rank, oi[q], oj);
                    if (!own_tile) push_candidate(st, rank, oj, oi[q]);
                } else if (valid) {
                    const float inv = rsqrt_approx(soft2 > 0.f ? fmaf(dx, dx, fmaf(dy, dy, soft2)) : d2);
                    const float i3 = (inv * inv) * inv;
                    const float sj = i3 * mj, si = i3 * -mi[q];
                    if (e) {
                        tfx[q].y = fmaf(dx, sj, tfx[q].y);
                        tfy[q].y = fmaf(dy, sj, tfy[q].y);
                    } else {
                        tfx[q].x = fmaf(dx, sj, tfx[q].x);
                        tfy[q].x = fmaf(dy, sj, tfy[q].x);
                    }
                    gxe = fmaf(dx, si, gxe);
                    gye = fmaf(dy, si, gye);
                }
            }
            if (e) {
                gx.y += gxe;
                gy.y += gye;
            } else {
                gx.x += gxe;
                gy.x += gye;
            }
        }
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
#pragma unroll
    for (int q = 0; q < IPT; ++q) {
        float4 a = acc_s[q][threadIdx.x];
        two_sum(a.x, a.y, tfx[q].x + tfx[q].y);
        two_sum(a.z, a.w, tfy[q].x + tfy[q].y);
        acc_s[q][threadIdx.x] = a;
    }
    if (!own_tile) {
        float4 ga = gacc[32 * c + lane];
        ga.x += gx.x;
        ga.y += gx.y;
        ga.z += gy.x;
        ga.w += gy.y;
        gacc[32 * c + lane] = ga;
    }
}

template <int IPT, int MINB>
__global__ void __launch_bounds__(kSymThreads, MINB) force_sym_kernel(const DevState st, const StepParams p)
{
    extern __shared__ __align__(128) float tiles_dyn[];
    __shared__ __align__(8) unsigned long long full_bar[kStages];
    __shared__ float4 acc_s[IPT][kSymThreads];        // per thread and row {fx_hi, fx_lo, fy_hi, fy_lo}
    __shared__ float4 gacc[2][kTJ / 2];                   // per j pair {gx0, gx1, gy0, gy1}, double-buffered over tile pairs
    __shared__ int s_rc[2];
    if (!st.desc->sym) return;
    float(*tiles)[kSortedTileFloats] = reinterpret_cast<float(*)[kSortedTileFloats]>(tiles_dyn);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // 512 rows = GROUPS row groups of 32 IPT rows; the HSPLIT warps of a group share its rows and split the 8 chunks
    // of a J tile: ROUNDS chunks each, one per round
    constexpr int GROUPS = kTJ / (32 * IPT), HSPLIT = 8 / GROUPS, ROUNDS = 8 / HSPLIT;
    static_assert(GROUPS * HSPLIT == 8 && ROUNDS == GROUPS, "8 warps, all on different chunks in every round");
    const int k = warp / HSPLIT, h = warp % HSPLIT;
    const int T = st.desc->n_jtiles, S = st.desc->sym_S, Q = st.desc->sym_Q;
    const int nblk = st.desc->sym_blocks;
    const float rmax = st.desc->rmax;
    const float2 s2 = make_float2(p.soft2, p.soft2);
    const float Rb = sqrtf((4.f * rmax * rmax + p.soft2) * 1.001f);     // no pre-test can pass beyond this separation
    const size_t stride = st.part_stride;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) mbar_init(&full_bar[s], 1);
        fence_barrier_init();
    }
    gacc[0][tid] = make_float4(0.f, 0.f, 0.f, 0.f);
    gacc[1][tid] = make_float4(0.f, 0.f, 0.f, 0.f);

    auto issue = [&](int stage, int tile) {       // one thread
        mbar_expect_tx(&full_bar[stage], (unsigned)kSortedTileFloats * 4u);
        bulk_g2s(tiles[stage], st.jts + (size_t)tile * kSortedTileFloats, (unsigned)kSortedTileFloats * 4u, &full_bar[stage]);
    };

    unsigned it = 0;                              // tile pairs this CTA has consumed: ring position and parity
    unsigned n_exact = 0, n_culled = 0;
    for (;;) {
        __syncthreads();
        if (tid == 0) {
            // several GPUs: rank r takes blocks r, r + world, ... of the same order
            const long long b = (long long)atomicAdd(&st.res->sym_next, 1u) * p.world + p.rank;
            int R = -1, C = -1;
            if (b < nblk) sym_block_decode((int)b, Q, R, C);
            s_rc[0] = R;
            s_rc[1] = C;
        }
        __syncthreads();
        const int R = s_rc[0], C = s_rc[1];
        if (R < 0) break;
        const bool diag = R == C;
        const int I0 = R * S, I1 = min(I0 + S, T), J0 = C * S, J1 = min(J0 + S, T);

        int pI = I0, pJ = diag ? I0 : J0;         // producer cursor (every thread keeps a copy)
        bool pmore = true;
        auto padvance = [&]() {
            if (++pJ == J1) {
                ++pI;
                pJ = diag ? pI : J0;
                if (pI == I1) pmore = false;
            }
        };
#pragma unroll 1
        for (int s = 0; s < kStages && pmore; ++s) {
            if (tid == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue((int)((it + s) % kStages), pJ);
            }
            padvance();
        }

#pragma unroll 1
        for (int I = I0; I < I1; ++I) {
            const float *rows = st.jts + (size_t)I * kSortedTileFloats;
            float nx[IPT], ny[IPT], nm[IPT], thr[IPT];
#pragma unroll
            for (int q = 0; q < IPT; ++q) {
                const int rs = 32 * IPT * k + 32 * q + lane;
                nx[q] = -rows[rs];
                ny[q] = -rows[kTJ + rs];
                nm[q] = -rows[2 * kTJ + rs];
                const float rr = rows[3 * kTJ + rs] + rmax;
                const float bound = p.soft2 > 0.f ? (rr * rr + p.soft2) * 1.000001f : rr * rr;
                thr[q] = __float_as_int(rows[4 * kTJ + rs]) >= 0 ? bound : -1.0f;      // pads never flag
                acc_s[q][tid] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            // row-side sums of the rounds since the last fold (registers); banked in the compensated shared
            // accumulators before every pre-tested round and at the end of every tile pair
            float2 tfx[IPT], tfy[IPT];
#pragma unroll
            for (int q = 0; q < IPT; ++q) {
                tfx[q] = make_float2(0.f, 0.f);
                tfy[q] = make_float2(0.f, 0.f);
            }
            auto fold_rows = [&]() {
#pragma unroll
                for (int q = 0; q < IPT; ++q) {
                    float4 a = acc_s[q][tid];
                    two_sum(a.x, a.y, tfx[q].x + tfx[q].y);
                    two_sum(a.z, a.w, tfy[q].x + tfy[q].y);
                    acc_s[q][tid] = a;
                    tfx[q] = make_float2(0.f, 0.f);
                    tfy[q] = make_float2(0.f, 0.f);
                }
            };
            float4 rb;                            // bounding box of this warp's 128 rows
            {
                const float4 *bx = reinterpret_cast<const float4 *>(rows + 5 * kTJ);
                constexpr int BOXES = 32 * IPT / kSubPart;        // 64-body boxes per row group
                rb = bx[BOXES * k];
#pragma unroll
                for (int e = 1; e < BOXES; ++e) {
                    const float4 b = bx[BOXES * k + e];
                    rb = make_float4(fminf(rb.x, b.x), fminf(rb.y, b.y), fmaxf(rb.z, b.z), fmaxf(rb.w, b.w));
                }
            }
#pragma unroll 1
            for (int J = diag ? I : J0; J < J1; ++J, ++it) {
                const int stage = (int)(it % kStages);
                mbar_wait(&full_bar[stage], (it / kStages) & 1u);
                const float *tl = tiles[stage];
                const bool own = J == I;
                const int buf = (int)(it & 1u);
#pragma unroll 1
                for (int r = 0; r < ROUNDS; ++r) {
                    const int c = ROUNDS * h + ((k + r) % ROUNDS);
                    const float4 cb = reinterpret_cast<const float4 *>(tl + 5 * kTJ)[c];
                    const bool may_hit = !((cb.x - rb.z > Rb) | (rb.x - cb.z > Rb) | (cb.y - rb.w > Rb) | (rb.y - cb.w > Rb));
                    float2 xs = *reinterpret_cast<const float2 *>(tl + 64 * c + 2 * lane);
                    float2 ys = *reinterpret_cast<const float2 *>(tl + kTJ + 64 * c + 2 * lane);
                    float2 ms = *reinterpret_cast<const float2 *>(tl + 2 * kTJ + 64 * c + 2 * lane);
                    float2 gx = make_float2(0.f, 0.f), gy = make_float2(0.f, 0.f);
                    bool cand = false;
                    if (may_hit) {
                        // the round's row sums must be separable (they are dropped if the pre-test fires):
                        // bank what earlier rounds left in the registers first
                        fold_rows();
                        sym_substeps<true, IPT>(xs, ys, ms, gx, gy, nx, ny, nm, thr, s2, tfx, tfy, cand, lane);
                    } else {
                        sym_substeps<false, IPT>(xs, ys, ms, gx, gy, nx, ny, nm, thr, s2, tfx, tfy, cand, lane);
                        ++n_culled;
                    }
                    if (may_hit && __any_sync(0xffffffffu, cand)) {
                        // rare: a possible hit somewhere in the round; its sums are dropped and the round redone
                        sym_exact_round<IPT>(st, tl, rows, c, k, own, p.soft2, p.rank, acc_s, gacc[buf]);
                        n_exact += 2;
#pragma unroll
                        for (int q = 0; q < IPT; ++q) {
                            tfx[q] = make_float2(0.f, 0.f);
                            tfy[q] = make_float2(0.f, 0.f);
                        }
                    } else if (!own) {
                        float4 ga = gacc[buf][32 * c + lane];
                        ga.x += gx.x;
                        ga.y += gx.y;
                        ga.z += gy.x;
                        ga.w += gy.y;
                        gacc[buf][32 * c + lane] = ga;
                    }
                    __syncthreads();
                }
                fold_rows();
                // every warp is done with the stage: refill it with the tile pair kStages ahead
                if (tid == 0 && pmore) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    issue(stage, pJ);
                }
                if (pmore) padvance();
                if (!own) {
                    // j side of this tile pair: bodies 2 tid, 2 tid + 1 of tile J <- the rows of tile I.  The first
                    // tile of the block's rows writes, the others add (same CTA, program order).
                    const float4 ga = gacc[buf][tid];
                    gacc[buf][tid] = make_float4(0.f, 0.f, 0.f, 0.f);
                    float4 *dst = reinterpret_cast<float4 *>(st.part + (size_t)R * stride + (size_t)J * kTJ) + tid;
                    float4 v = make_float4(ga.x, ga.z, ga.y, ga.w);       // {gx0, gy0, gx1, gy1}
                    if (I != I0) {
                        const float4 o = *dst;
                        v = make_float4(o.x + v.x, o.y + v.y, o.z + v.z, o.w + v.w);
                    }
                    *dst = v;
                }
            }
            // i side of the finished row of tile pairs: the two warps that share these rows, in fixed order
            __syncthreads();
            if (h == 0) {
                float2 *dst = st.part + (size_t)C * stride + (size_t)I * kTJ + 32 * IPT * k + lane;
#pragma unroll
                for (int q = 0; q < IPT; ++q) {
                    float4 a = acc_s[q][tid];
#pragma unroll
                    for (int e = 1; e < HSPLIT; ++e) {
                        const float4 b = acc_s[q][tid + 32 * e];
                        two_sum(a.x, a.y, b.x);
                        two_sum(a.z, a.w, b.z);
                        a.y += b.y;
                        a.w += b.w;
                    }
                    float2 v = make_float2(a.x + a.y, a.z + a.w);
                    if (diag && I != I0) {        // the diagonal block's rows already hold j-side sums of earlier tiles
                        const float2 o = dst[32 * q];
                        v = make_float2(o.x + v.x, o.y + v.y);
                    }
                    dst[32 * q] = v;
                }
            }
            __syncthreads();
        }
    }
    if (p.count_stats && lane == 0) {
        atomicAdd(&st.ctr->fast_chunks, (unsigned long long)it * (2ull * ROUNDS));   // rounds x 2 sub-chunks of 32 bodies
        atomicAdd(&st.ctr->exact_chunks, (unsigned long long)n_exact);
        atomicAdd(&st.ctr->culled_parts, (unsigned long long)n_culled);
    }
}

// Sharded two-sided kernel, before the allgather: this rank's partial force on every slot = the sum, in
// super-tile order, of the part[][] entries its own blocks wrote.
__global__ void __launch_bounds__(256) sym_reduce_kernel(const DevState st, const StepParams p)
{
    const StepDesc &d = *st.desc;
    if (!d.sym) return;
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= d.n) return;
    const int X = slot / kTJ / d.sym_S, Q = d.sym_Q, W = p.world, me = p.rank;
    float fx = 0.f, fy = 0.f, lx = 0.f, ly = 0.f;
    const float2 *src = st.part + slot;
    auto add = [&](int y) {
        const float2 v = __ldcs(src + (size_t)y * st.part_stride);
        two_sum(fx, lx, v.x);
        two_sum(fy, ly, v.y);
    };
    // super-tiles before X: block (y, X) has index y (Q - 1) - y (y - 1) / 2 + X - y - 1, stepping by Q - 2 - y
    int b = X - 1;
    for (int y = 0; y < X; ++y) {
        if (b % W == me) add(y);
        b += Q - 2 - y;
    }
    if (sym_block_index(X, X, Q) % W == me) add(X);
    // super-tiles after X: block (X, y), consecutive indices, so every W-th one is this rank's
    if (X + 1 < Q) {
        const int b0 = sym_block_index(X, X + 1, Q);
        int first = (me - b0 % W + W) % W;
        for (int y = X + 1 + first; y < Q; y += W) add(y);
    }
    x_force(st, p.rank)[slot] = make_float2(fx + lx, fy + ly);
}

// ... and after it: thread every rank's candidate pairs whose row this rank finishes into the rows' chains
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
            const int slot = st.sinv[pr.x];
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
        force_sym_kernel<8, 2><<<p.sym_grid, kSymThreads, kSymDynSmem, s>>>(st, p);
    else
        force_sym_kernel<4, 3><<<p.sym_grid, kSymThreads, kSymDynSmem, s>>>(st, p);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_sym_reduce(const DevState &st, const StepParams &p, cudaStream_t s)
{
    sym_reduce_kernel<<<(st.cap + 255) / 256, 256, 0, s>>>(st, p);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_sym_chain(const DevState &st, const StepParams &p, cudaStream_t s)
{
    sym_chain_kernel<<<296, 256, 0, s>>>(st, p);
    count_launch();
    return cudaGetLastError();
}

int force_sym_occupancy(int rows, int *regs)
{
    int occ = 0;
    cudaFuncAttributes fa = {};
    if (rows == 8) {
        cudaFuncSetAttribute(force_sym_kernel<8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSymDynSmem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, force_sym_kernel<8, 2>, kSymThreads, kSymDynSmem);
        cudaFuncGetAttributes(&fa, force_sym_kernel<8, 2>);
    } else {
        cudaFuncSetAttribute(force_sym_kernel<4, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSymDynSmem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, force_sym_kernel<4, 3>, kSymThreads, kSymDynSmem);
        cudaFuncGetAttributes(&fa, force_sym_kernel<4, 3>);
    }
    if (regs) *regs = fa.numRegs;
    return occ;
}

void sym_block_host(int b, int Q, int *R, int *C)
{
    sym_block_decode(b, Q, *R, *C);
}

int sym_block_index_host(int X, int Y, int Q)
{
    return sym_block_index(X, Y, Q);
}

}  // namespace nb
