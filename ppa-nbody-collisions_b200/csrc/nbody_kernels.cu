// nbody_kernels.cu -- hand-written sm_100a kernels of the ppa-nbody-collisions time step.
//
// One step = force (ComputeForces' O(n^2) pair loop, src/nbody.cu:182-242; the one-sided kernel lives here, the
//                   two-sided one of the cell-sorted order in nbody_sym.cu)
//          -> finish (collision bookkeeping :215-226,245-246, velocity + walls :250-264, MoveBodies :277-292)
//          -> [allgather of the post-step rows when sharded]
//          -> [count when sharded] + scatter (the host compaction of :488-510 as a stable device compaction)
//             + plan of the next step in the last CTA.
//
// Compiled with -fmad=false: every fused multiply-add below is written explicitly and sits exactly where
// the reference's PTX has one (SURVEY.md 8a "arithmetic contract"); the force sum itself uses rsqrt and a
// different summation order and is therefore tolerance-checked, not bit-exact.
#include "nbody_device.cuh"
#include "nbody_ptx.cuh"

namespace nb {
namespace {

// ------------------------------------------------------------------------------------------------
// plan: the coverage descriptor of the next step (src/nbody.cu:473, :194, :142-143) + sharding
// ------------------------------------------------------------------------------------------------
__host__ __device__ inline void plan_fill(StepDesc &d, const StepParams &p, int n, float rmax, unsigned step, float mmax,
                                          float rmin)
{
    const int T = kGroup;
    d.n = n;
    if (p.coverage == NB_COVERAGE_REFERENCE) {
        d.blocks = n < T ? 1 : n / T;                         // src/nbody.cu:473 (floor)
        d.limit_last = n % (T + 1);                           // src/nbody.cu:194
        const int threads = d.blocks * T;
        d.n_active = n < threads ? n : threads;               // src/nbody.cu:142-143
    } else {
        d.blocks = (n + T - 1) / T;
        d.limit_last = d.blocks > 0 ? n - T * (d.blocks - 1) : 0;
        d.n_active = n;
    }
    d.limit_first = d.blocks <= 1 ? d.limit_last : T;
    d.window_len = d.blocks > 0 ? T * (d.blocks - 1) + d.limit_last : 0;
    d.excl_len = n - d.window_len;
    const int world = p.world > 1 ? p.world : 1;
    const int rank = p.world > 1 ? p.rank : 0;
    const int IB = p.iblock > 0 ? p.iblock : kIBlock;      // rows per i-block = rows per force CTA
    const int iblocks_total = (n + IB - 1) / IB;
    const int per_rank = (iblocks_total + world - 1) / world;
    d.rows_per_rank = per_rank * IB;
    long long lo = (long long)rank * d.rows_per_rank;
    d.row_lo = lo < n ? (int)lo : n;
    long long hi = lo + d.rows_per_rank;
    d.row_hi = hi < n ? (int)hi : n;
    if (d.row_hi < d.row_lo) d.row_hi = d.row_lo;
    d.row_act_hi = d.row_hi < d.n_active ? d.row_hi : d.n_active;
    if (d.row_act_hi < d.row_lo) d.row_act_hi = d.row_lo;
    d.n_iblocks = (d.row_act_hi - d.row_lo + IB - 1) / IB;
    d.n_jtiles = (n + kTJ - 1) / kTJ;
    d.force_exact = n < 2 * T ? 1 : 0;
    // Unit size.  Splitting every j-tile into 2, 4 or 8 parts keeps the static partition over the force grid
    // balanced when there are few tiles per CTA, and a part is also the granularity at which a row that saw a
    // possible hit is redone, so smaller parts pay off while such rows are frequent.  Thresholds (whole tiles
    // per CTA) calibrated on B200 with the shipped surface density, profiles/r01_unit_size_sweep.log.
    const long long whole = (long long)d.n_iblocks * d.n_jtiles;
    const long long grid = p.force_grid > 0 ? p.force_grid : 1;
    d.lg_parts = whole < 16 * grid ? 3 : (whole < 100 * grid ? 2 : (whole < 400 * grid ? 1 : 0));
    if (p.lg_parts_override >= 0) d.lg_parts = p.lg_parts_override;
    d.units = whole << d.lg_parts;
    d.sorted = (p.sort_min_n > 0 && p.coverage == NB_COVERAGE_FULL && n >= p.sort_min_n && n >= 2 * kTJ) ? 1 : 0;
    d.sym = 0;
    d.sym_S = 1;
    d.sym_Q = 0;
    d.sym_blocks = 0;
    d.sym_lgu = 0;
    d.sym_items = 0;
    d.fscale = 1.f;
    d.finv = 1.0;
    // Two-sided kernel: always on the cell-sorted order; on the bodies' own order (every round pre-tested) from sym_min_n
    // bodies on, one GPU only.  It needs a fixed-point scale for its force sums; without one the one-sided kernel runs.
    // Two-sided kernels (all-pairs coverage, a fixed-point scale for the force sums must exist; else the one-sided kernel
    // runs).  One GPU, sym_min_n <= n < kSymWarpMaxN: a warp per work item (nbody_symw.cu), on the sorted order if the step
    // has one.  Otherwise, on the sorted order: a CTA per tile pair (nbody_sym.cu).
    // (several GPUs: only on the sorted order -- the plain step graph holds no exchange of force sums and candidates)
    const bool small = p.sym_min_n > 0 && n >= p.sym_min_n && n < p.symw_max_n && (p.world <= 1 || d.sorted);
    if (p.sym && p.coverage == NB_COVERAGE_FULL && small && p.sym_small == 2 &&
        sym_scale(n, mmax, rmin, p.field_w > p.field_h ? p.field_w : p.field_h, &d.fscale, &d.finv)) {
        d.sym = 2;
        d.sym_S = p.symw_run > 0 ? p.symw_run : symw_run(n, 4 * (p.symw_grid > 0 ? p.symw_grid : 1));
        d.sym_items = symw_geom(n, d.sym_S).ids;
    } else if (p.sym && p.coverage == NB_COVERAGE_FULL && (d.sorted || (small && p.sym_small == 1 && n >= 2 * kTJ)) &&
        sym_scale(n, mmax, rmin, p.field_w > p.field_h ? p.field_w : p.field_h, &d.fscale, &d.finv)) {
        // the triangle of tile pairs in blocks of S x S; S depends on the tile count alone, so that one GPU and several
        // cut the work -- and round the partial sums -- alike: their results agree bit for bit
        d.sym = 1;
        int S = d.n_jtiles / 1024;
        d.sym_S = S < 1 ? 1 : (S > kSymSMax ? kSymSMax : S);
        d.sym_Q = (d.n_jtiles + d.sym_S - 1) / d.sym_S;
        d.sym_blocks = d.sym_Q * (d.sym_Q + 1) / 2;
        // few tile pairs: split each into 2 or 4 items of 2 or 1 rounds, so that the queue still balances the grid
        const int grid = p.sym_grid > 0 ? p.sym_grid : 1;
        const int max_lgu = p.sym_rows == 8 ? 1 : 2;      // a tile pair is 2 (8 rows per lane) or 4 rounds
        if (d.sym_S == 1)
            while (d.sym_lgu < max_lgu && ((long long)d.sym_blocks << d.sym_lgu) < 16LL * grid) ++d.sym_lgu;
        d.sym_items = d.sym_blocks << d.sym_lgu;
    }
    d.rmax = rmax;
    d.step = step;
}

__global__ void plan_kernel(DevState st, StepParams p, int n)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    StepDesc d;
    plan_fill(d, p, n, __uint_as_float(st.res->rmax_bits), 0u, __uint_as_float(st.res->mmax_bits),
              __uint_as_float(0x7f800000u - st.res->rmin_inv));
    d.n_prev = n;
    *st.desc = d;
    st.res->rmax_bits = 0u;
    st.res->mmax_bits = 0u;
    st.res->rmin_inv = 0u;
    st.res->ticket = 0u;
    st.res->sym_next = 0u;
    if (st.xbuf) x_header(st, p.rank)->count = 0u;
    Counters c = {};
    *st.ctr = c;
    *st.host_n = n;
    __threadfence_system();
}

// owner(u): the force CTA whose unit range [c U / G, (c+1) U / G) holds unit u
__device__ __forceinline__ int unit_owner(long long u, long long U, int G)
{
    return (int)(((u + 1) * (long long)G - 1) / U);
}

// visit-order key of candidate j for row i: tile k ascending, then off with s = (t + off) % limit
// (src/nbody.cu:182-207)
__device__ __forceinline__ int visit_key(const StepDesc &d, int i, int j)
{
    const int gbase = i & ~(kGroup - 1), t = i & (kGroup - 1);
    int q = j - gbase;
    if (q < 0) q += d.n;
    const int k = q >> 7, s = q & (kGroup - 1);
    const int limit = (k == d.blocks - 1) ? d.limit_last : kGroup;
    int off = s - (t % limit);
    if (off < 0) off += limit;
    return k * kGroup + off;
}

// ------------------------------------------------------------------------------------------------
// force kernel
// ------------------------------------------------------------------------------------------------
struct Window {            // the j ranges a group never visits: [a0,b0) u [a1,b1)
    int a0, b0, a1, b1;
};

template <bool PACKED, bool TEST = true>
__device__ __forceinline__ void pair2(const float2 xs, const float2 ys, const float2 ms, const float2 nxi,
                                      const float2 nyi, const float thr, const float2 soft2, float2 &fx, float2 &fy, bool &cand)
{
    if (PACKED) {
        const float2 dx = __fadd2_rn(xs, nxi);
        const float2 dy = __fadd2_rn(ys, nyi);
        const float2 d2 = __ffma2_rn(dx, dx, __ffma2_rn(dy, dy, soft2));     // soft2 = 0: fma(dy, dy, 0) == dy * dy
        if (TEST) {
            cand |= (d2.x <= thr);
            cand |= (d2.y <= thr);
        }
        const float2 inv = make_float2(rsqrt_approx(d2.x), rsqrt_approx(d2.y));
        const float2 s = __fmul2_rn(__fmul2_rn(inv, inv), __fmul2_rn(inv, ms));
        fx = __ffma2_rn(dx, s, fx);
        fy = __ffma2_rn(dy, s, fy);
    } else {
        const float dx0 = xs.x + nxi.x, dy0 = ys.x + nyi.x;
        const float dx1 = xs.y + nxi.y, dy1 = ys.y + nyi.y;
        const float d20 = fmaf(dx0, dx0, fmaf(dy0, dy0, soft2.x)), d21 = fmaf(dx1, dx1, fmaf(dy1, dy1, soft2.y));
        if (TEST) {
            cand |= (d20 <= thr);
            cand |= (d21 <= thr);
        }
        const float i0 = rsqrt_approx(d20), i1 = rsqrt_approx(d21);
        const float s0 = (i0 * i0) * (i0 * ms.x), s1 = (i1 * i1) * (i1 * ms.y);
        fx.x = fmaf(dx0, s0, fx.x);
        fy.x = fmaf(dy0, s0, fy.x);
        fx.y = fmaf(dx1, s1, fx.y);
        fy.y = fmaf(dy1, s1, fy.y);
    }
}

// Exact evaluation of one 32-body sub-chunk for one row per lane: the reference predicate
// (src/nbody.cu:126-134), its bookkeeping split (hit pairs are excluded from the force sum, :215-226) and
// candidate emission.  Hits are collected in a per-lane bit mask first (no warp-level operation inside the
// j loop) and pushed afterwards: slots of the global candidate list are reserved per warp with one
// atomicAdd (warp-aggregated), then threaded into the row's chain.
__device__ __forceinline__ void exact_chunk(const DevState &st, const float *px, const int j0, const float xi,
                                            const float yi, const float ri, const bool active, const int row,
                                            const int excl, const Window &w, float2 &ax, float2 &ay, const int lane,
                                            const bool sorted, const float soft2)
{
    unsigned hits = 0;
#pragma unroll 2
    for (int jj = 0; jj < kSC; ++jj) {
        const float xj = px[jj], yj = px[kTJ + jj], mj = px[2 * kTJ + jj], rj = px[3 * kTJ + jj];
        const float dx = xj - xi, dy = yj - yi;
        const float d2 = fmaf(dx, dx, dy * dy);
        const float rs = ri + rj;
        const float rs2 = rs * rs;
        const bool hit = d2 <= rs2;
        // sorted j stream: the body's original index rides in the fifth plane (-1 for padding)
        const int j = sorted ? __float_as_int(px[4 * kTJ + jj]) : j0 + jj;
        const bool in_excl = ((j >= w.a0) & (j < w.b0)) | ((j >= w.a1) & (j < w.b1));
        const bool valid = active & (j >= 0) & (j != excl) & !in_excl;
        const float inv = rsqrt_approx(soft2 > 0.f ? fmaf(dx, dx, fmaf(dy, dy, soft2)) : d2);   // predicate above: unsoftened
        const float s = (inv * inv) * (inv * mj);
        if (valid && !hit) {                      // even j -> .x, odd j -> .y: the fast path's lane assignment
            if (jj & 1) {
                ax.y = fmaf(dx, s, ax.y);
                ay.y = fmaf(dy, s, ay.y);
            } else {
                ax.x = fmaf(dx, s, ax.x);
                ay.x = fmaf(dy, s, ay.x);
            }
        }
        hits |= (valid && hit ? 1u : 0u) << jj;
    }
    unsigned pending = __ballot_sync(0xffffffffu, hits != 0u);
    while (pending) {
        const bool push = hits != 0u;
        const int jj = push ? __ffs(hits) - 1 : 0;
        hits &= hits - 1u;
        const int leader = __ffs(pending) - 1;
        unsigned base = 0;
        if (lane == leader) base = atomicAdd(&st.ctr->cand_count, (unsigned)__popc(pending));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (push) {
            const unsigned idx = base + __popc(pending & ((1u << lane) - 1u));
            if (idx < (unsigned)st.cand_cap) {
                const int prev = atomicExch(&st.head[row], (int)idx);
                st.cand[idx] = make_int2(sorted ? __float_as_int(px[4 * kTJ + jj]) : j0 + jj, prev);
            } else {
                st.ctr->overflow_flag = 1;
            }
        }
        pending = __ballot_sync(0xffffffffu, hits != 0u);
    }
}

// One CTA = WARPS warps x 32 lanes x IPT rows per lane = one 512-row i-block (kIBlock).  The CTA owns a
// contiguous run of work units (i-block, j-tile, part) of the step's static partition and walks it in
// SEGMENTS: the parts of one j-tile that fall into its run (one whole tile when n is large).  A segment is one
// ring stage: its bodies arrive by 1-D TMA bulk copies completing on an mbarrier, the last warp to finish a
// stage refills it with the segment kStages ahead.  Inside a segment every part is a checkpoint: its sums are
// kept only if the row's collision pre-test stayed clear over the part, otherwise the part is redone exactly.
template <bool PACKED, int WARPS, int IPT, bool SORTED>
__device__ __forceinline__ void force_body(const DevState &st, const StepParams &p, float *tiles_dyn,
                                           unsigned long long *full_bar, unsigned *done_cnt,
                                           float4 (*acc_s)[WARPS * 32])
{
    constexpr int IBLOCK = WARPS * 32 * IPT;      // rows per CTA = rows per i-block (p.iblock)
    static_assert(IPT << kMaxLgParts <= 32, "one redo bit per (part, row) of a segment");
    constexpr int THREADS = WARPS * 32;
    constexpr int WROWS = 32 * IPT;               // rows per warp (a divisor of the 128-row visit-order group)
    float(*tiles)[kSortedTileFloats] = reinterpret_cast<float(*)[kSortedTileFloats]>(tiles_dyn);   // kStages of the larger layout
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long U = st.desc->units;
    const int c = blockIdx.x;
    const int G = (long long)gridDim.x < U ? (int)gridDim.x : (int)U;   // CTAs that own at least one unit
    if (c >= G) return;
    const long long u0 = (long long)c * U / G, u1 = (long long)(c + 1) * U / G;
    if (u0 >= u1) return;
    const int lgP = st.desc->lg_parts;            // a unit (part) is kTJ >> lgP bodies of a j-tile
    const int P = 1 << lgP;
    const int TP = st.desc->n_jtiles << lgP;      // units per i-block
    const int n = st.desc->n;
    const int row_lo = st.desc->row_lo, row_act_hi = st.desc->row_act_hi;
    const int excl_len = st.desc->excl_len, limit_first = st.desc->limit_first;
    const bool fexact = st.desc->force_exact != 0;
    const float rmax = st.desc->rmax;
    const float2 s2 = make_float2(p.soft2, p.soft2);
    const int jw = kTJ >> lgP;                    // bodies per part
    constexpr bool sorted = SORTED;               // cell-sorted order: 5 planes + bounding boxes per tile, rows are slots
    const float *jsrc = sorted ? st.jts : st.jt;
    const int tile_floats = sorted ? kSortedTileFloats : kTileFloats;
    const int nsc = jw / kSC;                     // 32-body sub-chunks per part

    int ib = (int)(u0 / TP);
    int v = (int)(u0 - (long long)ib * TP);       // unit index inside the i-block: tile << lgP | part
    int left = (int)(u1 - u0);                    // units of the run not yet consumed
    // parts of the segment that starts at unit v with `rem` units of the run left: up to the end of the tile
    auto seg_parts = [&](int vv, int rem) { const int to_tile_end = P - (vv & (P - 1)); return rem < to_tile_end ? rem : to_tile_end; };
    // producer cursor: the segment kStages ahead of the consumer (every warp keeps its own, identical, copy)
    int pv = v, pleft = left;

    auto issue = [&](int stage, int vv, int parts) {   // one thread at a time
        const int tile = vv >> lgP, part = vv & (P - 1);
        const float *src = jsrc + (size_t)tile * tile_floats + part * jw;
        float *dst = tiles[stage] + part * jw;
        const unsigned plane_bytes = (unsigned)(parts * jw) * 4u;
        const unsigned box_bytes = (unsigned)(parts * jw / kSubPart) * 16u;
        if (parts == P) {                              // the whole tile is contiguous in either layout
            mbar_expect_tx(&full_bar[stage], (unsigned)tile_floats * 4u);
            bulk_g2s(tiles[stage], jsrc + (size_t)tile * tile_floats, (unsigned)tile_floats * 4u, &full_bar[stage]);
        } else if (!sorted) {
            mbar_expect_tx(&full_bar[stage], 4u * plane_bytes);
#pragma unroll
            for (int pl = 0; pl < 4; ++pl) bulk_g2s(dst + pl * kTJ, src + pl * kTJ, plane_bytes, &full_bar[stage]);
        } else {
            mbar_expect_tx(&full_bar[stage], 5u * plane_bytes + box_bytes);
#pragma unroll
            for (int pl = 0; pl < 5; ++pl) bulk_g2s(dst + pl * kTJ, src + pl * kTJ, plane_bytes, &full_bar[stage]);
            const int box0 = 5 * kTJ + 4 * (part * jw / kSubPart);
            bulk_g2s(tiles[stage] + box0, jsrc + (size_t)tile * tile_floats + box0, box_bytes, &full_bar[stage]);
        }
    };
    auto advance = [&](int &vv, int &rem) {            // step a cursor over one segment
        const int parts = seg_parts(vv, rem);
        rem -= parts;
        vv += parts;
        if (vv == TP) vv = 0;
    };

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            done_cnt[s] = 0;
        }
        fence_barrier_init();
    }
    __syncthreads();
#pragma unroll 1
    for (int k = 0; k < kStages; ++k) {               // prologue: fill the ring
        if (pleft > 0 && threadIdx.x == 0) issue(k, pv, seg_parts(pv, pleft));
        if (pleft > 0) advance(pv, pleft);
    }

    float2 nxi[IPT], nyi[IPT], fx[IPT], fy[IPT];
    float thr[IPT];
    Window w = {0, 0, 0, 0};
    int wbase = 0, gbase = 0;
    bool warp_active = false;
    unsigned n_fast = 0, n_exact = 0, n_culled = 0;

    for (int it = 0; left > 0; ++it) {
        if (it == 0 || v == 0) {
            // (re)load this warp's rows: lane l holds rows wbase + 32 q + l; gbase = their 128-row group
            wbase = row_lo + ib * IBLOCK + warp * WROWS;
            gbase = wbase & ~(kGroup - 1);
            warp_active = wbase < row_act_hi;
#pragma unroll
            for (int q = 0; q < IPT; ++q) {
                const int i = wbase + 32 * q + lane;
                const bool act = i < row_act_hi;
                float4 b = make_float4(kDummyCoord, kDummyCoord, 0.f, 0.f);
                if (act) {
                    if (sorted) {                 // rows are slots of the sorted order: the body sits in the sorted tiles
                        const float *t = st.jts + (size_t)(i / kTJ) * kSortedTileFloats + (i & (kTJ - 1));
                        b = make_float4(t[0], t[kTJ], 0.f, t[3 * kTJ]);
                    } else {
                        b = st.pm[i];
                    }
                }
                nxi[q] = make_float2(-b.x, -b.x);
                nyi[q] = make_float2(-b.y, -b.y);
                const float rs = b.w + rmax;
                // softened distances are larger by eps^2: so is the pre-test bound (with a margin for its rounding)
                const float bound = p.soft2 > 0.f ? (rs * rs + p.soft2) * 1.000001f : rs * rs;
                thr[q] = act ? bound : -1.0f;
                fx[q] = make_float2(0.f, 0.f);
                fy[q] = make_float2(0.f, 0.f);
                acc_s[q][threadIdx.x] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (excl_len == 0) {
                w = Window{0, 0, 0, 0};
            } else if (gbase >= excl_len) {
                w = Window{gbase - excl_len, gbase, 0, 0};
            } else {
                w = Window{0, gbase, n + gbase - excl_len, n};
            }
        }
        const int parts = seg_parts(v, left);
        const int stage = it % kStages;
        mbar_wait(&full_bar[stage], (it / kStages) & 1);

        if (warp_active) {
            const int tile = v >> lgP, part0 = v & (P - 1);
            // Fast pass: per part, plain running sums and ONE pre-test flag per row, no warp-level
            // synchronisation inside.  A part's sums are folded into the segment sums only for rows whose flag
            // stayed clear; the others (a self pair, a real neighbour, a window edge, tiny n) get a redo bit.
            unsigned cmask = 0;                   // bit pp * IPT + q: row q of this lane must redo part pp
#pragma unroll 1
            for (int pp = 0; pp < parts; ++pp) {
                const float *tl = tiles[stage] + (part0 + pp) * jw;
                const int jt0 = tile * kTJ + (part0 + pp) * jw;
                const bool pspecial = fexact | ((jt0 < w.b0) & (jt0 + jw > w.a0)) | ((jt0 < w.b1) & (jt0 + jw > w.a1));
                float2 tfx[IPT], tfy[IPT];
                bool cand[IPT];
#pragma unroll
                for (int q = 0; q < IPT; ++q) {
                    tfx[q] = make_float2(0.f, 0.f);
                    tfy[q] = make_float2(0.f, 0.f);
                    cand[q] = pspecial;
                }
                // Sorted stream: the part's bodies are close together.  If no row of this warp lies inside their
                // bounding box inflated by the row's pre-test radius, no pair of the part can pass the pre-test
                // (|dx| or |dy| alone already exceeds the radius) and the loop runs without it.
                bool may_hit = true;
                if (sorted && !pspecial) {
                    const float4 *boxes = reinterpret_cast<const float4 *>(tiles[stage] + 5 * kTJ);
                    float4 bb;
                    if (lgP == 0) {
                        bb = boxes[kTJ / kSubPart];                     // the tile's own box
                    } else {
                        boxes += (part0 + pp) * (jw / kSubPart);
                        bb = boxes[0];
                        for (int k = 1; k < jw / kSubPart; ++k) {
                            const float4 o = boxes[k];
                            bb = make_float4(fminf(bb.x, o.x), fminf(bb.y, o.y), fmaxf(bb.z, o.z), fmaxf(bb.w, o.w));
                        }
                    }
                    bool inside = false;
#pragma unroll
                    for (int q = 0; q < IPT; ++q) {
                        // pre-test radius sqrt(thr) with a margin that covers the approximate rsqrt; thr < 0
                        // (inactive row) gives NaN: never inside
                        const float R = thr[q] * rsqrt_approx(thr[q]) * 1.0002f;
                        const float x = -nxi[q].x, y = -nyi[q].x;
                        inside |= (x >= bb.x - R) & (x <= bb.z + R) & (y >= bb.y - R) & (y <= bb.w + R);
                    }
                    may_hit = __any_sync(0xffffffffu, inside);
                }
                if (may_hit) {
#pragma unroll 8
                    for (int k4 = 0; k4 < jw / 4; ++k4) {
                        const float4 X = *reinterpret_cast<const float4 *>(tl + 4 * k4);
                        const float4 Y = *reinterpret_cast<const float4 *>(tl + kTJ + 4 * k4);
                        const float4 M = *reinterpret_cast<const float4 *>(tl + 2 * kTJ + 4 * k4);
#pragma unroll
                        for (int q = 0; q < IPT; ++q) {
                            pair2<PACKED, true>(make_float2(X.x, X.y), make_float2(Y.x, Y.y), make_float2(M.x, M.y), nxi[q],
                                                nyi[q], thr[q], s2, tfx[q], tfy[q], cand[q]);
                            pair2<PACKED, true>(make_float2(X.z, X.w), make_float2(Y.z, Y.w), make_float2(M.z, M.w), nxi[q],
                                                nyi[q], thr[q], s2, tfx[q], tfy[q], cand[q]);
                        }
                    }
                } else {
#pragma unroll 8
                    for (int k4 = 0; k4 < jw / 4; ++k4) {
                        const float4 X = *reinterpret_cast<const float4 *>(tl + 4 * k4);
                        const float4 Y = *reinterpret_cast<const float4 *>(tl + kTJ + 4 * k4);
                        const float4 M = *reinterpret_cast<const float4 *>(tl + 2 * kTJ + 4 * k4);
#pragma unroll
                        for (int q = 0; q < IPT; ++q) {
                            pair2<PACKED, false>(make_float2(X.x, X.y), make_float2(Y.x, Y.y), make_float2(M.x, M.y), nxi[q],
                                                 nyi[q], thr[q], s2, tfx[q], tfy[q], cand[q]);
                            pair2<PACKED, false>(make_float2(X.z, X.w), make_float2(Y.z, Y.w), make_float2(M.z, M.w), nxi[q],
                                                 nyi[q], thr[q], s2, tfx[q], tfy[q], cand[q]);
                        }
                    }
                    ++n_culled;
                }
                unsigned bits = 0;
#pragma unroll
                for (int q = 0; q < IPT; ++q) {
                    if (!cand[q]) {
                        fx[q] = __fadd2_rn(fx[q], tfx[q]);
                        fy[q] = __fadd2_rn(fy[q], tfy[q]);
                    }
                    bits |= (cand[q] ? 1u : 0u) << q;
                }
                cmask |= bits << (pp * IPT);
            }
            n_fast += parts * nsc;
            unsigned redo = __reduce_or_sync(0xffffffffu, cmask);
            while (redo) {
                // rare: row q of some lanes may hold a hit in part pp.  Those lanes redo the part sub-chunk by
                // sub-chunk: fast sums where the pre-test stays clear, the exact predicate (with window / self
                // exclusion and candidate emission) where it does not.
                const int bit = __ffs(redo) - 1;
                redo &= redo - 1;
                const int pp = bit / IPT, q = bit - pp * IPT;
                const bool mine = (cmask >> bit) & 1u;
                const float *tl = tiles[stage] + (part0 + pp) * jw;
                const int jt0 = tile * kTJ + (part0 + pp) * jw;
                float2 nx = nxi[0], ny = nyi[0], ax = fx[0], ay = fy[0];
                float th = thr[0];
#pragma unroll
                for (int k = 1; k < IPT; ++k)
                    if (q == k) {
                        nx = nxi[k];
                        ny = nyi[k];
                        ax = fx[k];
                        ay = fy[k];
                        th = thr[k];
                    }
                const int slot = wbase + 32 * q + lane;
                const int t = slot - gbase;
                const bool act = mine && slot < row_act_hi;
                // the row's ORIGINAL index: what candidates, the self test and the window refer to
                const int row = (sorted && act) ? st.sidx[0][slot] : slot;
                int excl = row;
                if (limit_first != kGroup) excl = limit_first > 0 ? gbase + (t % limit_first) : -1;
                const float xi = -nx.x, yi = -ny.x;
                const float ri = act ? st.pm[row].w : 0.f;
#pragma unroll 1
                for (int sc = 0; sc < nsc; ++sc) {
                    const int j0 = jt0 + sc * kSC;
                    const float *px = tl + sc * kSC;
                    float2 tx = make_float2(0.f, 0.f), ty = make_float2(0.f, 0.f);
                    bool cc = fexact | ((j0 < w.b0) & (j0 + kSC > w.a0)) | ((j0 < w.b1) & (j0 + kSC > w.a1));
                    if (mine) {
#pragma unroll 1
                        for (int k4 = 0; k4 < kSC / 4; ++k4) {
                            const float4 X = *reinterpret_cast<const float4 *>(px + 4 * k4);
                            const float4 Y = *reinterpret_cast<const float4 *>(px + kTJ + 4 * k4);
                            const float4 M = *reinterpret_cast<const float4 *>(px + 2 * kTJ + 4 * k4);
                            pair2<PACKED>(make_float2(X.x, X.y), make_float2(Y.x, Y.y), make_float2(M.x, M.y), nx, ny, th, s2, tx, ty, cc);
                            pair2<PACKED>(make_float2(X.z, X.w), make_float2(Y.z, Y.w), make_float2(M.z, M.w), nx, ny, th, s2, tx, ty, cc);
                        }
                    }
                    const bool need = mine && cc;
                    if (__any_sync(0xffffffffu, need)) {
                        ++n_exact;
                        exact_chunk(st, px, j0, xi, yi, ri, act && need, row, excl, w, ax, ay, lane, sorted, p.soft2);
                    }
                    if (mine && !cc) {
                        ax = __fadd2_rn(ax, tx);
                        ay = __fadd2_rn(ay, ty);
                    }
                }
#pragma unroll
                for (int k = 0; k < IPT; ++k)
                    if (q == k) {
                        fx[k] = ax;
                        fy[k] = ay;
                    }
            }
        }
        // release the stage; the last warp to get here refills it with the segment kStages ahead (no warp ever
        // waits for another one: the only blocking point is the full barrier above)
        __syncwarp();
        if (lane == 0) {
            __threadfence_block();
            if (atomicAdd(&done_cnt[stage], 1u) == WARPS - 1) {
                done_cnt[stage] = 0;
                if (pleft > 0) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    issue(stage, pv, seg_parts(pv, pleft));
                }
            }
        }
        if (pleft > 0) advance(pv, pleft);

        if (warp_active) {                        // fold this segment's sums into the compensated accumulators
#pragma unroll
            for (int q = 0; q < IPT; ++q) {
                float4 a = acc_s[q][threadIdx.x];
                two_sum(a.x, a.y, fx[q].x + fx[q].y);
                two_sum(a.z, a.w, fy[q].x + fy[q].y);
                acc_s[q][threadIdx.x] = a;
                fx[q] = make_float2(0.f, 0.f);
                fy[q] = make_float2(0.f, 0.f);
            }
        }
        left -= parts;
        v += parts;
        if (v == TP || left == 0) {               // leaving this i-block: flush the partial sums of this run
            if (warp_active) {
                float2 *slab = st.fpart + (size_t)(c + ib) * IBLOCK + warp * WROWS + lane;
#pragma unroll
                for (int q = 0; q < IPT; ++q) {
                    const float4 a = acc_s[q][threadIdx.x];
                    slab[32 * q] = make_float2(a.x + a.y, a.z + a.w);
                }
            }
            if (v == TP) {
                v = 0;
                ++ib;
            }
        }
    }
    if (p.count_stats && lane == 0) {
        atomicAdd(&st.ctr->fast_chunks, (unsigned long long)n_fast);
        atomicAdd(&st.ctr->exact_chunks, (unsigned long long)n_exact);
        atomicAdd(&st.ctr->culled_parts, (unsigned long long)n_culled);
    }
}

// The step's descriptor decides (on the device) whether this step runs on the cell-sorted order; both bodies are
// instantiated so that neither pays for the other's code in its hot loop.
template <bool PACKED, int WARPS, int IPT, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) force_kernel(const DevState st, const StepParams p)
{
    extern __shared__ __align__(128) float tiles_dyn[];
    __shared__ __align__(8) unsigned long long full_bar[kStages];
    __shared__ unsigned done_cnt[kStages];        // warps that finished the segment in each stage (refill trigger)
    // second summation level: per row {fx_hi, fx_lo, fy_hi, fy_lo}, updated once per segment with a
    // compensated (TwoSum) add, so the rounding error of a row's force stays at the level of one short
    // float sum instead of growing with n as a single running float sum does
    __shared__ float4 acc_s[IPT][WARPS * 32];
    if (st.desc->sym) return;                     // this step belongs to force_sym_kernel
    if (st.desc->sorted)
        force_body<PACKED, WARPS, IPT, true>(st, p, tiles_dyn, full_bar, done_cnt, acc_s);
    else
        force_body<PACKED, WARPS, IPT, false>(st, p, tiles_dyn, full_bar, done_cnt, acc_s);
}

// ------------------------------------------------------------------------------------------------
// finish: per-row epilogue of ComputeForces + MoveBodies
// ------------------------------------------------------------------------------------------------
// `slot` is the thread's position in this rank's range [row_lo, row_hi): the body's index itself, or -- when
// the step ran on the cell-sorted order -- its slot in that order (partial sums and post rows are by slot,
// everything else by the body's original index).
__device__ __forceinline__ bool finish_row(const DevState &st, const StepParams &p, const StepDesc &d, const int slot)
{
    if (slot >= d.row_hi) return false;
    const int local = slot - d.row_lo;
    const int row = d.sorted ? st.sidx[0][slot] : slot;
    float4 *out_pm = post_pm(st, p.world > 1 ? p.rank : 0);
    float2 *out_vel = post_vel(st, p.world > 1 ? p.rank : 0);
    int *out_abs = post_abs(st, p.world > 1 ? p.rank : 0);
    const float4 b = st.pm[row];
    float2 v = st.vel[row];
    if (slot >= d.row_act_hi) {                   // frozen tail: no thread in either reference kernel
        if (p.merge) out_abs[local] = row;
        out_pm[local] = b;
        out_vel[local] = v;
        if (b.z == 0.f && p.world <= 1 && !p.merge) atomicAdd(&st.tile_count[row / kCompactTile], 1);
        return b.z != 0.f;
    }
    // force = sum of the segment partials in CTA order
    const int IB = p.iblock > 0 ? p.iblock : kIBlock;
    const int ib = local / IB, within = local % IB;
    const long long U = d.units;
    const int TP = d.n_jtiles << d.lg_parts;
    const int G = (long long)p.force_grid < U ? p.force_grid : (int)U;
    const int c_first = unit_owner((long long)ib * TP, U, G);
    const int c_last = unit_owner((long long)(ib + 1) * TP - 1, U, G);
    float fx = 0.f, fy = 0.f;
    if (d.sym) {                                  // two-sided kernel: one exact fixed-point sum per body (all ranks' after the
        longlong2 *acc = reinterpret_cast<longlong2 *>(st.facc) + slot;        // all-reduce when sharded)
        const longlong2 a = *acc;
        *acc = make_longlong2(0, 0);              // ready for the next step
        fx = (float)((double)a.x * d.finv);
        fy = (float)((double)a.y * d.finv);
    } else {
        for (int c = c_first; c <= c_last; ++c) {
            const float2 part = st.fpart[(size_t)(c + ib) * IB + within];
            fx += part.x;
            fy += part.y;
        }
    }
    // collision bookkeeping in the reference's visit order (src/nbody.cu:215-226)
    float umass = b.z, uradius = b.w;
    bool deleted = false;
    int lowest = row;                             // conserving merge: lowest index among the row and its partners
    const int h = st.head[row];
    if (h >= 0) {
        st.head[row] = -1;
        int last_key = -1;
        while (true) {
            int best_key = 0x7fffffff, best_j = -1;
            for (int e = h; e >= 0;) {
                const int2 ce = st.cand[e];
                const int key = visit_key(d, row, ce.x);
                if (key > last_key && key < best_key) {
                    best_key = key;
                    best_j = ce.x;
                }
                e = ce.y;
            }
            if (best_j < 0) break;
            const float4 o = st.pm[best_j];
            int kind;
            if (p.merge) {                        // opt-in: only the pointer and the event, merge_*_kernel does the rest
                lowest = best_j < lowest ? best_j : lowest;
                kind = row < best_j ? NB_EV_ABSORB : NB_EV_KILLED;
            } else if (b.z >= o.z) {              // :215-221
                umass += o.z;
                uradius = fmaf(p.growth, o.w, uradius);
                kind = NB_EV_ABSORB;
            } else {                              // :222-226
                deleted = true;
                kind = NB_EV_KILLED;
            }
            if (st.ev_cap > 0) {
                const unsigned slot = atomicAdd(&st.ctr->ev_count, 1u);
                if (slot < (unsigned)st.ev_cap) {
                    st.ev[slot] = EventRec{(int)d.step, row, best_j, ((unsigned)best_key << 1) | (unsigned)kind};
                } else {
                    st.ctr->ev_dropped = 1;
                }
            }
            last_key = best_key;
        }
    }
    if (p.merge) out_abs[local] = lowest;
    // velocity + walls (:250-264), position (:288)
    const float ax = fx * p.grav, ay = fy * p.grav;
    const float dvx = p.dt * ax, dvy = p.dt * ay;
    const float W = (float)p.field_w, H = (float)p.field_h;
    const float nW = (float)(-p.field_w), nH = (float)(-p.field_h);
    const float tpx = dvx + b.x, tpy = dvy + b.y;
    if (tpx > W - b.w || tpx < b.w + nW) v.x = -v.x;
    if (tpy > H - b.w || tpy < b.w + nH) v.y = -v.y;
    v.x = v.x + dvx;
    v.y = v.y + dvy;
    float4 o;
    o.x = fmaf(p.dt, v.x, b.x);
    o.y = fmaf(p.dt, v.y, b.y);
    o.z = deleted ? 0.f : umass;                  // :245
    o.w = uradius;                                // :246
    out_pm[local] = o;
    out_vel[local] = v;
    // single GPU: the compaction needs the number of removed bodies per tile of 1024 bodies (by body index).  They
    // are few, so one atomic each is cheaper than a counting pass (tile_count is zero on entry: upload and the
    // previous scatter clear it).  Sharded runs and the opt-in merge count afterwards instead (count_kernel).
    if (!(o.z != 0.f) && p.world <= 1 && !p.merge) atomicAdd(&st.tile_count[row / kCompactTile], 1);
    return o.z != 0.f;
}

__global__ void __launch_bounds__(256) finish_kernel(const DevState st, const StepParams p)
{
    const StepDesc &d = *st.desc;
    const int row = d.row_lo + blockIdx.x * blockDim.x + threadIdx.x;
    finish_row(st, p, d, row);
}

// post rows are stored by slot (= index, or position in the sorted order), rank chunk by rank chunk
__device__ __forceinline__ int post_slot(const DevState &st, int i) { return st.desc->sorted ? st.sinv[i] : i; }
__device__ __forceinline__ float4 load_post_pm(const DevState &st, int rpr, int i)
{
    const int s = post_slot(st, i), rk = s / rpr, loc = s - rk * rpr;
    return post_pm(st, rk)[loc];
}
__device__ __forceinline__ float2 load_post_vel(const DevState &st, int rpr, int i)
{
    const int s = post_slot(st, i), rk = s / rpr, loc = s - rk * rpr;
    return post_vel(st, rk)[loc];
}

// ------------------------------------------------------------------------------------------------
// opt-in conserving lowest-index merge (NB_FLAG_MERGE_CONSERVING; not reference behaviour).  finish_kernel left, per row,
// the lowest index among the body and its hit partners next to the post-force state; after the allgather every rank
// holds all of them and runs the same two kernels over all bodies (deterministic: chains are walked in index order).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) merge_link_kernel(const DevState st)
{
    const int n = st.desc->n, rpr = st.desc->rows_per_rank;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    auto absorber = [&](int b) {
        const int s = post_slot(st, b), rk = s / rpr;
        return post_abs(st, rk)[s - rk * rpr];
    };
    int r = i;
    for (int a = absorber(r); a != r; a = absorber(r)) r = a;          // pointers only ever go down: terminates
    if (r != i) st.mnext[i] = atomicExch(&st.mhead[r], i);             // thread i onto its root's chain
}

__global__ void __launch_bounds__(256) merge_apply_kernel(const DevState st, const StepParams p)
{
    const int n = st.desc->n, rpr = st.desc->rows_per_rank;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int h = st.mhead[i];
    if (h < 0) return;                             // not a root with members (members are zeroed by their root)
    st.mhead[i] = -1;
    auto pm_of = [&](int b) {
        const int s = post_slot(st, b), rk = s / rpr;
        return post_pm(st, rk) + (s - rk * rpr);
    };
    auto vel_of = [&](int b) {
        const int s = post_slot(st, b), rk = s / rpr;
        return post_vel(st, rk) + (s - rk * rpr);
    };
    float4 me = *pm_of(i);
    const float2 v = *vel_of(i);
    float M = me.z, Px = me.z * v.x, Py = me.z * v.y;
    int last = -1;
    while (true) {                                 // members in ascending index order (the chain is unordered)
        int k = 0x7fffffff;
        for (int e = h; e >= 0; e = st.mnext[e])
            if (e > last && e < k) k = e;
        if (k == 0x7fffffff) break;
        float4 *ok = pm_of(k);
        const float4 o = *ok;
        const float2 ov = *vel_of(k);
        M += o.z;
        Px = fmaf(o.z, ov.x, Px);
        Py = fmaf(o.z, ov.y, Py);
        me.w = fmaf(p.growth, o.w, me.w);
        ok->z = 0.f;                               // removed by the compaction
        last = k;
    }
    me.z = M;
    *pm_of(i) = me;
    *vel_of(i) = make_float2(__fdiv_rn(Px, M), __fdiv_rn(Py, M));
}

// ------------------------------------------------------------------------------------------------
// compaction: count, then scatter (+ plan of the next step in the last CTA)
// ------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(kCompactThreads) count_kernel(const DevState st, const StepParams p)
{
    __shared__ int s_cnt[kCompactThreads / 32];
    if (p.world <= 1 && !p.merge) return;          // finish_kernel already counted the removed bodies
    const int n = st.desc->n, rpr = st.desc->rows_per_rank;
    const int base = blockIdx.x * kCompactTile;
    if (base >= n) return;
    int cnt = 0;
#pragma unroll
    for (int r = 0; r < kCompactTile / kCompactThreads; ++r) {
        const int i = base + r * kCompactThreads + threadIdx.x;
        if (i < n) cnt += load_post_pm(st, rpr, i).z != 0.f ? 0 : 1;       // removed bodies
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
#pragma unroll
        for (int k = 0; k < kCompactThreads / 32; ++k) tot += s_cnt[k];
        st.tile_count[blockIdx.x] = tot;
    }
}

__device__ __forceinline__ void store_body(const DevState &st, int o, const float4 b, const float2 v)
{
    st.pm[o] = b;
    st.vel[o] = v;
    float *t = st.jt + (size_t)(o / kTJ) * kTileFloats + (o & (kTJ - 1));
    t[0] = b.x;
    t[kTJ] = b.y;
    t[2 * kTJ] = b.z;
    t[3 * kTJ] = b.w;
}
__device__ __forceinline__ void store_pad(const DevState &st, int o)
{
    float *t = st.jt + (size_t)(o / kTJ) * kTileFloats + (o & (kTJ - 1));
    t[0] = kPadCoord;
    t[kTJ] = kPadCoord;
    t[2 * kTJ] = 0.f;
    t[3 * kTJ] = 0.f;
}

__device__ __forceinline__ int block_sum(int v, int *s_buf)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_buf[threadIdx.x >> 5] = v;
    __syncthreads();
    int tot = 0;
#pragma unroll
    for (int k = 0; k < kCompactThreads / 32; ++k) tot += s_buf[k];
    return tot;
}

__global__ void __launch_bounds__(kCompactThreads) scatter_kernel(const DevState st, const StepParams p)
{
    __shared__ int s_buf[kCompactThreads / 32];
    __shared__ int s_warp[kCompactThreads / 32];
    __shared__ bool s_last;
    const int n = st.desc->n, rpr = st.desc->rows_per_rank;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ct = blockIdx.x, base = ct * kCompactTile;
    if (base < n) {
        int part = 0;
        for (int t = threadIdx.x; t < ct; t += kCompactThreads) part += kCompactTile - st.tile_count[t];   // tiles < ct are full
        int run = block_sum(part, s_buf);
        float rmx = 0.f, mmx = 0.f, rmn = __int_as_float(0x7f800000);
#pragma unroll 1
        for (int r = 0; r < kCompactTile / kCompactThreads; ++r) {
            const int i = base + r * kCompactThreads + threadIdx.x;
            float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
            float2 v = make_float2(0.f, 0.f);
            if (i < n) {
                b = load_post_pm(st, rpr, i);
                v = load_post_vel(st, rpr, i);
            }
            const bool keep = (i < n) && (b.z != 0.f);            // src/nbody.cu:490
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            __syncthreads();
            if (lane == 0) s_warp[warp] = __popc(m);
            __syncthreads();
            int before = 0, total = 0;
#pragma unroll
            for (int k = 0; k < kCompactThreads / 32; ++k) {
                const int cnt = s_warp[k];
                before += k < warp ? cnt : 0;
                total += cnt;
            }
            const int dst_index = run + before + __popc(m & ((1u << lane) - 1u));
            if (st.remap && i < n) st.remap[i] = keep ? dst_index : -1;
            if (keep) {
                store_body(st, dst_index, b, v);
                rmx = fmaxf(rmx, b.w);
                mmx = fmaxf(mmx, b.z);
                rmn = fminf(rmn, b.w);
            }
            run += total;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            rmx = fmaxf(rmx, __shfl_xor_sync(0xffffffffu, rmx, o));
            mmx = fmaxf(mmx, __shfl_xor_sync(0xffffffffu, mmx, o));
            rmn = fminf(rmn, __shfl_xor_sync(0xffffffffu, rmn, o));
        }
        if (lane == 0 && rmx > 0.f) atomicMax(&st.res->rmax_bits, __float_as_uint(rmx));
        if (lane == 0 && mmx > 0.f) atomicMax(&st.res->mmax_bits, __float_as_uint(mmx));
        if (lane == 0 && rmn >= 0.f) atomicMax(&st.res->rmin_inv, 0x7f800000u - __float_as_uint(rmn));   // min: +0 .. +inf order like uints
    }
    // last CTA to finish plans the next step
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&st.res->ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const int tiles = (n + kCompactTile - 1) / kCompactTile;
    int part = 0;
    for (int t = threadIdx.x; t < tiles; t += kCompactThreads) part += __ldcg(&st.tile_count[t]);
    const int n_new = n - block_sum(part, s_buf);                      // tile_count holds the removed bodies
    for (int t = threadIdx.x; t < tiles; t += kCompactThreads) st.tile_count[t] = 0;   // finish_kernel adds into them
    const int pad_end = (n_new + kTJ - 1) / kTJ * kTJ;
    for (int o = n_new + threadIdx.x; o < pad_end; o += kCompactThreads) store_pad(st, o);
    if (threadIdx.x == 0) {
        const StepDesc old = *st.desc;
        Counters *c = st.ctr;
        const long long per_row = old.window_len > 0 ? old.window_len - 1 : 0;
        c->pairs += (unsigned long long)((long long)(old.row_act_hi - old.row_lo) * per_row);
        c->candidates += c->cand_count;
        c->cand_count = 0;
        c->steps += 1;
        StepDesc d;
        plan_fill(d, p, n_new, __uint_as_float(__ldcg(&st.res->rmax_bits)), old.step + 1,
                  __uint_as_float(__ldcg(&st.res->mmax_bits)), __uint_as_float(0x7f800000u - __ldcg(&st.res->rmin_inv)));
        d.n_prev = old.n;
        *st.desc = d;
        st.res->rmax_bits = 0u;
        st.res->mmax_bits = 0u;
        st.res->rmin_inv = 0u;
        st.res->ticket = 0u;
        st.res->sym_next = 0u;
        if (st.xbuf) x_header(st, p.rank)->count = 0u;
        if (p.sort_min_n > 0) {                   // the host picks the next steps' graph by this (pinned, mapped);
            *st.host_n = n_new;                   // once it has switched to the lean graph it never looks again
            __threadfence_system();
        }
    }
}

// ------------------------------------------------------------------------------------------------
// ingest / export: the reference's BodiesData block (src/nbody.cu:66-77) <-> the SoA store
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ingest_kernel(const DevState st, const float2 *__restrict__ pos_in,
                                                     const float2 *__restrict__ vel_in, const float *__restrict__ mass_in,
                                                     const float *__restrict__ rad_in, const int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int pad_end = (n + kTJ - 1) / kTJ * kTJ;
    float r = 0.f, mx = 0.f, rn = __int_as_float(0x7f800000);
    if (i < n) {
        const float2 pos = pos_in[i];
        const float2 v = vel_in[i];
        const float m = mass_in[i];
        r = rad_in[i];
        mx = m;
        rn = r;
        store_body(st, i, make_float4(pos.x, pos.y, m, r), v);
    } else if (i < pad_end) {
        store_pad(st, i);
    }
    if (i < st.cap) st.head[i] = -1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        r = fmaxf(r, __shfl_xor_sync(0xffffffffu, r, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        rn = fminf(rn, __shfl_xor_sync(0xffffffffu, rn, o));
    }
    if ((threadIdx.x & 31) == 0 && r > 0.f) atomicMax(&st.res->rmax_bits, __float_as_uint(r));
    if ((threadIdx.x & 31) == 0 && mx > 0.f) atomicMax(&st.res->mmax_bits, __float_as_uint(mx));
    if ((threadIdx.x & 31) == 0 && rn >= 0.f) atomicMax(&st.res->rmin_inv, 0x7f800000u - __float_as_uint(rn));
}

__global__ void __launch_bounds__(256) export_kernel(const DevState st, float *__restrict__ block, const int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 b = st.pm[i];
    reinterpret_cast<float2 *>(block)[i] = make_float2(b.x, b.y);
    reinterpret_cast<float2 *>(block + 2 * (size_t)n)[i] = st.vel[i];
    block[4 * (size_t)n + i] = b.z;
    block[5 * (size_t)n + i] = b.w;
}

// ------------------------------------------------------------------------------------------------
// render: filled discs into an 8-bit image (src/nbody.cu:294-348), bounds-checked
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) render_kernel(const DevState st, const int n, unsigned char *img, const int w,
                                                     const int h, const int field_w, const int field_h)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 b = st.pm[i];
    // pixel radius (float) and centre, :310,:318-319 (field_w / field_h are half extents)
    const float prf = (b.w * (float)w) / (float)field_w;
    const int dfw = field_w << 1, dfh = field_h << 1;
    const int cx = (int)(((b.x + (float)field_w) / (float)dfw) * (float)w);
    const int cy = (int)(((b.y + (float)field_h) / (float)dfh) * (float)h);
    // bounding box, :323-326 (note the reference's asymmetric >= / > and its float -> int truncation)
    const int y_min = (float)cy - prf < 0.f ? 0 : (int)((float)cy - prf);
    const int y_max = (float)cy + prf >= (float)h ? h : (int)((float)cy + prf);
    const int x_min = (float)cx - prf < 0.f ? 0 : (int)((float)cx - prf);
    const int x_max = (float)cx + prf > (float)w ? w : (int)((float)cx + prf);
    const int r2 = (int)(prf * prf);
    for (int y = y_min; y < y_max; ++y) {
        for (int x = x_min; x < x_max; ++x) {
            const int x_sq = (x - cx) * (x - cx), y_sq = (y - cy) * (y - cy);
            if (x_sq + y_sq <= r2 && x >= 0 && x < w && y >= 0 && y < h) img[(size_t)w * y + x] = 0;   // :344
        }
    }
}

// force-kernel variants (selected by nb_params.flags, see NB_FLAG_VARIANT): occupancy vs rows per lane.
// Measured on B200 at n = 131072 (profiles/r01_variants.md): 0 is the fastest.
//   0: packed f32x2, 8 warps x 2 rows/lane, <=  80 registers (3 CTAs = 24 warps per SM)   [default]
//   1: packed,       8 warps x 2 rows/lane, <=  64 registers (4 CTAs = 32 warps per SM)
//   2: packed,       4 warps x 4 rows/lane, <= 128 registers (4 CTAs = 16 warps per SM)
//   3: packed,       8 warps x 2 rows/lane, <= 128 registers (2 CTAs = 16 warps per SM)
//   4: scalar FP32,  4 warps x 4 rows/lane (A/B reference for the packed path)
//   5: packed,       8 warps x 4 rows/lane (1024-row i-blocks), <= 128 registers (2 CTAs = 16 warps per SM)
constexpr int kForceDynSmem = kStages * kSortedTileFloats * 4;

#define NB_FORCE_VARIANTS(X)     \
    X(0, true, 8, 2, 3)          \
    X(1, true, 8, 2, 4)          \
    X(2, true, 4, 4, 4)          \
    X(3, true, 8, 2, 2)          \
    X(4, false, 4, 4, 4)         \
    X(5, true, 8, 4, 2)

}  // namespace

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
long long &launch_counter()
{
    static thread_local long long n = 0;
    return n;
}

cudaError_t launch_plan(const DevState &st, const StepParams &p, int n, cudaStream_t s)
{
    plan_kernel<<<1, 32, 0, s>>>(st, p, n);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_force(const DevState &st, const StepParams &p, int variant, cudaStream_t s)
{
    switch (variant) {
#define X(ID, PK, W, I, MB)                                                                    \
    case ID:                                                                                   \
        force_kernel<PK, W, I, MB><<<p.force_grid, W * 32, kForceDynSmem, s>>>(st, p);         \
        count_launch();                                                                        \
        break;
        NB_FORCE_VARIANTS(X)
#undef X
    default:
        return cudaErrorInvalidValue;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess || !p.sym) return e;
    // a context that can run the two-sided kernels: whichever kernel the step descriptor does not name returns at once
    // (a kernel the context's capacity can never reach is not launched at all: n <= n_max)
    // CTA-level kernel: sorted steps outside the warp-level kernel's range [sym_min_n, symw_max_n)
    const bool cta_level = p.sym_small == 1 || (p.sort_min_n > 0 && (p.sym_small != 2 || st.cap >= p.symw_max_n || p.sort_min_n < p.sym_min_n));
    const bool warp_level = p.sym_small == 2 && p.sym_min_n > 0 && st.cap >= p.sym_min_n;
    if (cta_level) e = launch_force_sym(st, p, s);
    if (e == cudaSuccess && warp_level) e = launch_force_symw(st, p, s);
    return e;
}

cudaError_t launch_finish(const DevState &st, const StepParams &p, cudaStream_t s)
{
    const int grid = (st.shard_cap + 255) / 256;
    finish_kernel<<<grid, 256, 0, s>>>(st, p);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_compact(const DevState &st, const StepParams &p, bool recount, cudaStream_t s)
{
    const int grid = (st.cap + kCompactTile - 1) / kCompactTile;
    if (p.world > 1 || recount) {            // rows finished on other GPUs, or removed after finish (merge): count now
        count_kernel<<<grid, kCompactThreads, 0, s>>>(st, p);
        count_launch();
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    scatter_kernel<<<grid, kCompactThreads, 0, s>>>(st, p);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_merge(const DevState &st, const StepParams &p, cudaStream_t s)
{
    const int grid = (st.cap + 255) / 256;
    merge_link_kernel<<<grid, 256, 0, s>>>(st);
    count_launch();
    merge_apply_kernel<<<grid, 256, 0, s>>>(st, p);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_ingest(const DevState &st, const float *pos, const float *vel, const float *mass, const float *rad, int n,
                          cudaStream_t s)
{
    const int span = st.cap + kTJ;
    ingest_kernel<<<(span + 255) / 256, 256, 0, s>>>(st, reinterpret_cast<const float2 *>(pos), reinterpret_cast<const float2 *>(vel),
                                                     mass, rad, n);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_export(const DevState &st, float *block, int n, cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    export_kernel<<<(n + 255) / 256, 256, 0, s>>>(st, block, n);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_render(const DevState &st, int n, unsigned char *img, int w, int h, int field_w, int field_h,
                          cudaStream_t s)
{
    if (n <= 0) return cudaSuccess;
    render_kernel<<<(n + 127) / 128, 128, 0, s>>>(st, n, img, w, h, field_w, field_h);
    count_launch();
    return cudaGetLastError();
}

int force_occupancy(int variant, int *regs, int *threads, int *iblock)
{
    int occ = 0;
    cudaFuncAttributes fa = {};
    int thr = 0, ibl = 0;
    switch (variant) {
#define X(ID, PK, W, I, MB)                                                                        \
    case ID:                                                                                       \
        thr = W * 32;                                                                              \
        ibl = W * 32 * I;                                                                          \
        cudaFuncSetAttribute(force_kernel<PK, W, I, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, kForceDynSmem); \
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, force_kernel<PK, W, I, MB>, thr, kForceDynSmem);   \
        cudaFuncGetAttributes(&fa, force_kernel<PK, W, I, MB>);                                    \
        break;
        NB_FORCE_VARIANTS(X)
#undef X
    default:
        return 0;
    }
    if (regs) *regs = fa.numRegs;
    if (threads) *threads = thr;
    if (iblock) *iblock = ibl;
    return occ;
}

size_t fpart_slabs(int force_grid, int shard_cap, int iblock)
{
    return (size_t)force_grid + (size_t)(shard_cap + iblock - 1) / iblock + 1;
}

void plan_host(StepDesc *d, const StepParams *p, int n)
{
    plan_fill(*d, *p, n, 0.f, 0u, 1.0f, 1.0f);     // unit mass and radius: the fixed-point scale always exists
    d->n_prev = n;
}

}  // namespace nb
