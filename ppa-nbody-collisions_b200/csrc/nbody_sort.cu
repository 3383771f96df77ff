// nbody_sort.cu -- the cell-sorted shadow copy of the j stream (full coverage, large n).
//
// The force kernel's per-pair collision pre-test costs one issue slot in thirteen.  It can be skipped for a
// (warp rows, j part) pair when no row of the warp lies inside the part's bounding box inflated by the row's
// pre-test radius -- which only happens often if the bodies of a part are close together.  So, after every
// compaction, the bodies are sorted by a 16-bit Morton cell key (stable LSD radix sort, two 8-bit passes:
// deterministic, ties keep index order) and a second set of j-tiles is built in that order:
//   jts[tile] = { x[512] y[512] m[512] r[512] orig[512] bbox[8] }       (bbox: float4 per 64 bodies)
// A step that runs on this order takes BOTH sides from it: a warp's rows are 64 consecutive slots (so they are
// close together as well, and their own self pairs sit in one part), partial force sums and post-step rows are
// stored by slot (sinv maps a body to its slot for the finish / compaction kernels).  Everything that the
// reference defines by body index -- visit order, candidates, events, survivors' order -- keeps using the
// original index, which rides along in the fifth plane.
#include "nbody_device.cuh"

namespace nb {
namespace {

constexpr int kSortThreads = 256;
constexpr int kSortMinPerBlock = 2048;          // elements per radix block: at least 8 warps x 256 contiguous elements,
constexpr int kSortMaxBlocks = 256;             // more when that keeps the histogram table (256 x blocks) short to scan

inline int sort_per_block(int cap)
{
    const int per = (cap + kSortMaxBlocks - 1) / kSortMaxBlocks;
    const int rounded = (per + 255) / 256 * 256;
    return rounded > kSortMinPerBlock ? rounded : kSortMinPerBlock;
}

__device__ __forceinline__ unsigned spread8(unsigned v)      // abcdefgh -> 0a0b0c0d0e0f0g0h
{
    v = (v | (v << 4)) & 0x0f0fu;
    v = (v | (v << 2)) & 0x3333u;
    v = (v | (v << 1)) & 0x5555u;
    return v;
}

__global__ void __launch_bounds__(kSortThreads) keys_kernel(const DevState st, const StepParams p)
{
    if (!st.desc->sorted) return;
    const int n = st.desc->n;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 b = st.pm[i];
    // 256 x 256 cells over the field [-W, W] x [-H, H]; bodies outside are clamped to the border cells
    const float sx = 128.0f / (float)p.field_w, sy = 128.0f / (float)p.field_h;
    int cx = (int)((b.x + (float)p.field_w) * sx), cy = (int)((b.y + (float)p.field_h) * sy);
    cx = cx < 0 ? 0 : (cx > 255 ? 255 : cx);
    cy = cy < 0 ? 0 : (cy > 255 ? 255 : cy);
    st.skey[0][i] = spread8((unsigned)cx) | (spread8((unsigned)cy) << 1);
    st.sidx[0][i] = i;
}

// histogram of one 8-bit digit per radix block: hist[bin * nblocks + block]
__global__ void __launch_bounds__(kSortThreads) radix_hist_kernel(const DevState st, const int src, const int shift,
                                                                  const int per_block)
{
    __shared__ unsigned s_hist[256];
    if (!st.desc->sorted) return;
    const int n = st.desc->n;
    const int nblocks = gridDim.x;
    s_hist[threadIdx.x] = 0;
    __syncthreads();
    const int base = blockIdx.x * per_block;
    for (int k = threadIdx.x; k < per_block; k += kSortThreads) {
        const int i = base + k;
        if (i < n) atomicAdd(&s_hist[(st.skey[src][i] >> shift) & 0xffu], 1u);
    }
    __syncthreads();
    st.shist[(size_t)threadIdx.x * nblocks + blockIdx.x] = s_hist[threadIdx.x];
}

// exclusive scan of the 256 * nblocks table, in place: one block walks it in coalesced chunks of 1024 entries
// (warp shuffles + one shared array of warp sums per chunk) and carries the running total
__global__ void __launch_bounds__(1024) radix_scan_kernel(const DevState st, const int nblocks)
{
    __shared__ unsigned s_warp[32];
    __shared__ unsigned s_carry;
    if (!st.desc->sorted) return;
    const int m = 256 * nblocks;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < m; base += 1024) {
        const int k = base + threadIdx.x;
        const unsigned v = k < m ? st.shist[k] : 0u;
        unsigned inc = v;                                        // inclusive scan inside the warp
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        if (warp == 0) {                                         // exclusive scan of the 32 warp sums
            const unsigned w = s_warp[lane];
            unsigned winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned t = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += t;
            }
            s_warp[lane] = winc - w;
        }
        __syncthreads();
        const unsigned carry = s_carry;
        if (k < m) st.shist[k] = carry + s_warp[warp] + inc - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + s_warp[31] + inc;
        __syncthreads();
    }
}

// stable scatter of one digit: every warp walks its contiguous share of the block's elements in order
__global__ void __launch_bounds__(kSortThreads) radix_scatter_kernel(const DevState st, const int src, const int shift,
                                                                     const int per_block)
{
    __shared__ unsigned s_cnt[kSortThreads / 32][256];        // per warp: digit counts, then running offsets
    if (!st.desc->sorted) return;
    const int n = st.desc->n;
    const int nblocks = gridDim.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int dst = src ^ 1;
    for (int k = threadIdx.x; k < (kSortThreads / 32) * 256; k += kSortThreads) (&s_cnt[0][0])[k] = 0;
    __syncthreads();
    const int per_warp = per_block / (kSortThreads / 32);
    const int wbase = blockIdx.x * per_block + warp * per_warp;
    for (int k = lane; k < per_warp; k += 32) {
        const int i = wbase + k;
        if (i < n) atomicAdd(&s_cnt[warp][(st.skey[src][i] >> shift) & 0xffu], 1u);
    }
    __syncthreads();
    {   // thread b: turn the per-warp counts of bin b into output offsets (global base + earlier warps)
        const int b = threadIdx.x;
        unsigned run = st.shist[(size_t)b * nblocks + blockIdx.x];
#pragma unroll
        for (int w = 0; w < kSortThreads / 32; ++w) {
            const unsigned c = s_cnt[w][b];
            s_cnt[w][b] = run;
            run += c;
        }
    }
    __syncthreads();
    for (int k = lane; k < per_warp; k += 32) {               // warp-uniform trip count
        const int i = wbase + k;
        const bool valid = i < n;
        const unsigned key = valid ? st.skey[src][i] : 0u;
        const unsigned digit = valid ? (key >> shift) & 0xffu : 0x100u + lane;      // invalid lanes match nobody
        const unsigned peers = __match_any_sync(0xffffffffu, digit);
        const int rank = __popc(peers & ((1u << lane) - 1u));
        unsigned pos = 0;
        if (valid) pos = s_cnt[warp][digit] + rank;
        __syncwarp();
        if (valid && rank == __popc(peers) - 1) s_cnt[warp][digit] += (unsigned)__popc(peers);   // last peer advances
        __syncwarp();
        if (valid) {
            st.skey[dst][pos] = key;
            st.sidx[dst][pos] = st.sidx[src][i];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Carrying the order over a compaction (all steps but every kResortEvery-th): bodies move a small fraction of a cell per
// step, so the previous step's order is as good as a fresh sort for culling (bounding boxes are rebuilt from the current
// positions either way; correctness never depends on the order).  The survivors keep their relative places: a stable
// compaction of the slot list through the body compaction's index map -- 2 small kernels instead of the 7 of a sort.
// ------------------------------------------------------------------------------------------------
constexpr int kCarryTile = 1024;

__global__ void __launch_bounds__(256) carry_count_kernel(const DevState st)
{
    __shared__ int s_cnt[8];
    if (!st.desc->sorted) return;
    const int n_prev = st.desc->n_prev;
    const int base = blockIdx.x * kCarryTile;
    if (base >= n_prev) return;
    int cnt = 0;
#pragma unroll
    for (int r = 0; r < kCarryTile / 256; ++r) {
        const int s = base + r * 256 + threadIdx.x;
        if (s < n_prev) cnt += st.remap[st.sidx[0][s]] >= 0 ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) tot += s_cnt[k];
        st.carry_count[blockIdx.x] = tot;
    }
}

__global__ void __launch_bounds__(256) carry_scatter_kernel(const DevState st)
{
    __shared__ int s_buf[8];
    __shared__ int s_warp[8];
    if (!st.desc->sorted) return;
    const int n_prev = st.desc->n_prev;
    const int base = blockIdx.x * kCarryTile;
    if (base >= n_prev) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int part = 0;
    for (int t = threadIdx.x; t < (int)blockIdx.x; t += 256) part += st.carry_count[t];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) s_buf[warp] = part;
    __syncthreads();
    int run = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) run += s_buf[k];
#pragma unroll 1
    for (int r = 0; r < kCarryTile / 256; ++r) {
        const int s = base + r * 256 + threadIdx.x;
        const int now = s < n_prev ? st.remap[st.sidx[0][s]] : -1;
        const bool keep = now >= 0;
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        __syncthreads();
        if (lane == 0) s_warp[warp] = __popc(m);
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int c = s_warp[k];
            before += k < warp ? c : 0;
            total += c;
        }
        if (keep) st.sidx[1][run + before + __popc(m & ((1u << lane) - 1u))] = now;
        run += total;
    }
}

// sorted j-tiles + bounding boxes; slot s of the order holds body src[s] (src = the radix sort's result sidx[0], or the
// carried-over list sidx[1], which is copied into sidx[0] on the way).  One block per tile.
__global__ void __launch_bounds__(kTJ) gather_kernel(const DevState st, const int resort)
{
    __shared__ float4 s_box[kTJ / 32];
    if (!st.desc->sorted) return;
    const int n = st.desc->n;
    const int s = blockIdx.x * kTJ + threadIdx.x;
    if (blockIdx.x * kTJ >= n) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4 b = make_float4(kPadCoord, kPadCoord, 0.f, 0.f);
    int orig = -1;
    if (s < n) {
        orig = st.sidx[resort ? 0 : 1][s];
        if (!resort) st.sidx[0][s] = orig;
        b = st.pm[orig];
        st.sinv[orig] = s;
    }
    // the two-sided kernel's force sums start every step from zero (on several GPUs finish only clears its own rows)
    if (st.facc) reinterpret_cast<longlong2 *>(st.facc)[s] = make_longlong2(0, 0);
    float *tile = st.jts + (size_t)blockIdx.x * kSortedTileFloats;
    float *t = tile + threadIdx.x;
    t[0] = b.x;
    t[kTJ] = b.y;
    t[2 * kTJ] = b.z;
    t[3 * kTJ] = b.w;
    t[4 * kTJ] = __int_as_float(orig);
    // bounding box per 64 slots and of the whole tile; pads do not count (an empty box overlaps nothing)
    const float inf = __int_as_float(0x7f800000);
    float x0 = s < n ? b.x : inf, y0 = s < n ? b.y : inf, x1 = s < n ? b.x : -inf, y1 = s < n ? b.y : -inf;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        x0 = fminf(x0, __shfl_xor_sync(0xffffffffu, x0, o));
        y0 = fminf(y0, __shfl_xor_sync(0xffffffffu, y0, o));
        x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, o));
        y1 = fmaxf(y1, __shfl_xor_sync(0xffffffffu, y1, o));
    }
    if (lane == 0) s_box[warp] = make_float4(x0, y0, x1, y1);
    __syncthreads();
    float4 *boxes = reinterpret_cast<float4 *>(tile + 5 * kTJ);
    if (threadIdx.x < kTJ / kSubPart) {
        const float4 a = s_box[2 * threadIdx.x], c = s_box[2 * threadIdx.x + 1];
        boxes[threadIdx.x] = make_float4(fminf(a.x, c.x), fminf(a.y, c.y), fmaxf(a.z, c.z), fmaxf(a.w, c.w));
    } else if (threadIdx.x == 32) {
        float4 u = s_box[0];
        for (int k = 1; k < kTJ / 32; ++k) {
            const float4 o = s_box[k];
            u = make_float4(fminf(u.x, o.x), fminf(u.y, o.y), fmaxf(u.z, o.z), fmaxf(u.w, o.w));
        }
        boxes[kTJ / kSubPart] = u;
    }
}

}  // namespace

cudaError_t launch_sort(const DevState &st, const StepParams &p, cudaStream_t s)
{
    if (p.resort) {
        const int per_block = sort_per_block(st.cap);
        const int nblocks = (st.cap + per_block - 1) / per_block;
        keys_kernel<<<(st.cap + kSortThreads - 1) / kSortThreads, kSortThreads, 0, s>>>(st, p);
        count_launch();
        for (int pass = 0; pass < 2; ++pass) {
            radix_hist_kernel<<<nblocks, kSortThreads, 0, s>>>(st, pass, 8 * pass, per_block);
            count_launch();
            radix_scan_kernel<<<1, 1024, 0, s>>>(st, nblocks);
            count_launch();
            radix_scatter_kernel<<<nblocks, kSortThreads, 0, s>>>(st, pass, 8 * pass, per_block);
            count_launch();
        }
    } else {
        const int tiles = (st.cap + kCarryTile - 1) / kCarryTile;
        carry_count_kernel<<<tiles, 256, 0, s>>>(st);
        count_launch();
        carry_scatter_kernel<<<tiles, 256, 0, s>>>(st);
        count_launch();
    }
    gather_kernel<<<(st.cap + kTJ - 1) / kTJ, kTJ, 0, s>>>(st, p.resort);
    count_launch();
    return cudaGetLastError();
}

size_t sort_hist_entries(int cap)
{
    const int per_block = sort_per_block(cap);
    return (size_t)256 * ((cap + per_block - 1) / per_block);
}

}  // namespace nb
