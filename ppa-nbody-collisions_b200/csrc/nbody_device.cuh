// nbody_device.cuh -- device-side data layout shared by the kernels and the C ABI.
//
// HBM layout (all arrays owned by the context, allocated once for n_max bodies):
//   pm   [cap]          float4 {x, y, m, r}   current compacted bodies (the i side, 16 B coalesced)
//   vel  [cap]          float2 {vx, vy}
//   jt   [tiles][4][512] float                 the same bodies as planar 8 KB j-tiles {x[512] y[512] m[512] r[512]}:
//                                              one 1-D TMA bulk copy per tile; the last tile is padded with
//                                              benign bodies (far away, m = 0, r = 0)
//   post [world][shard_cap*24 B]               uncompacted post-step rows, one chunk per rank:
//                                              float4 pm[shard_cap] then float2 vel[shard_cap]  (allgather payload;
//                                              + int absorber[shard_cap] with the opt-in conserving merge)
//   fpart[G + iblocks][512] float2             partial force sums, one slab per (force CTA, i-block) segment,
//                                              summed in CTA order by the finish kernel (deterministic)
//   head [cap] int, cand[cand_cap] int2{j,next} collision candidates: one global list filled through
//                                              warp-aggregated atomics, threaded into per-row chains
//   facc [tiles*512][2] int64                  two-sided force kernel: the force on the body in `slot` as 64-bit FIXED-POINT
//                                              sums (x, y), scale 2^k chosen per step from a bound on |F| (plan).  Every
//                                              CTA adds its partial sums with RED.ADD.64: integer addition is associative,
//                                              so the total does not depend on which CTA (or which GPU) took which block
//                                              -- deterministic without one partial per (body, block) in HBM
//   tile_count[cap/1024] int                   removed bodies per compaction tile of 1024 bodies
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "nbody_b200.h"

namespace nb {

constexpr int kGroup = 128;                  // rows per visit-order group == reference THREADS_PER_BLOCK (src/nbody.cu:36)
constexpr int kIBlock = 512;                 // default rows per i-block == rows per force CTA (warps x 32 lanes x rows per lane)
constexpr int kTJ = 512;                     // j bodies per shared-memory tile
constexpr int kMaxLgParts = 3;               // smallest work unit = kTJ >> 3 = 64 bodies
constexpr int kTileFloats = 4 * kTJ;         // x, y, m, r planes
constexpr int kTileBytes = kTileFloats * 4;
constexpr int kSubPart = 64;                 // bodies per bounding box of the sorted j stream (= the smallest part)
constexpr int kSortedTileFloats = 5 * kTJ + 4 * (kTJ / kSubPart + 1);   // x, y, m, r, orig planes + 8 float4 boxes + their union
constexpr int kSC = 32;                      // j bodies per sub-chunk (granularity of the collision pre-test)
constexpr int kStages = 4;                   // TMA ring depth
constexpr int kCompactThreads = 256;
constexpr int kCompactTile = 1024;           // rows per compaction tile (4 rounds of 256)
constexpr int kResortEvery = 32;             // the cell-sorted order is rebuilt from scratch every so many steps and carried
                                             // over (survivors keep their places) in between: bodies move little per step
constexpr int kSymSMax = 8;                  // two-sided force kernel: a block of the pair triangle is S x S tile pairs, S <= 8
constexpr int kSymMinNDefault = 12288;       // smallest n the warp-level two-sided kernel takes (one GPU); below it the one-sided
                                             // kernel is faster (profiles/r02_small_n_sweep.md)
constexpr int kSymWarpMaxN = 196608;         // from here on the CTA-level kernel of nbody_sym.cu takes over (the warp-level one is
                                             // 21 % / 11 % faster at n = 65 536 / 131 072, 2 % / 4 % slower at 262 144 / 1M)
constexpr int kWGroup = 128;                 // warp-level two-sided kernel (nbody_symw.cu): rows per group, bodies per chunk
constexpr int kWChunk = 64;
constexpr float kPadCoord = 1.0e18f;         // padding j bodies sit here: d2 ~ 2e36, finite, contributes exactly 0
constexpr float kDummyCoord = -1.0e18f;      // inactive i lanes sit here

struct StepDesc {                 // rewritten on the device at the end of every step (plan)
    int n;                        // live bodies
    int n_prev;                   // live bodies of the step before (slots of the order being carried over)
    int blocks;                   // B: j tiles of the reference's visit order        (src/nbody.cu:473)
    int limit_last;               // slots of tile B-1                                (src/nbody.cu:194)
    int limit_first;              // slots of tile 0 (128 unless B == 1)
    int n_active;                 // bodies that own a thread                         (src/nbody.cu:142-143)
    int window_len;               // L = 128 (B-1) + limit_last: length of every group's cyclic j window
    int excl_len;                 // n - L: bodies cyclically preceding a group's start that it never visits
    int row_lo, row_hi;           // this rank's rows [row_lo, row_hi) of [0, n)
    int row_act_hi;               // min(row_hi, n_active), >= row_lo
    int rows_per_rank;            // rows per rank this step (multiple of 512)
    int n_iblocks;                // 512-row i-blocks holding this rank's active rows
    int n_jtiles;                 // T = ceil(n / kTJ)
    int force_exact;              // 1: every sub-chunk takes the exact path (n < 256)
    int lg_parts;                 // a work unit is kTJ >> lg_parts bodies of one j-tile (0 .. kMaxLgParts)
    int sorted;                   // 1: the force kernel streams the cell-sorted j-tiles (jts) this step
    int sym;                      // this step's force kernel: 0 one-sided; 1 two-sided, a CTA per tile pair (nbody_sym.cu: on the
                                  //    sorted order when `sorted`); 2 two-sided, a warp per work item (nbody_symw.cu: small n,
                                  //    bodies' own order, one GPU)
    int sym_S, sym_Q;             //    1: tiles per super-tile, super-tiles (Q = ceil(T / S));  2: sym_S = chunks per work item
    int sym_blocks;               //    Q (Q + 1) / 2 blocks of the pair triangle
    int sym_lgu;                  //    a tile pair is split into 1 << sym_lgu work items of 4 >> sym_lgu rounds (S == 1 only)
    int sym_items;                //    sym_blocks << sym_lgu items in the work queue
    float fscale;                 //    fixed-point scale of facc (a power of two)
    double finv;                  //    1 / fscale
    long long units;              // U = n_iblocks * T << lg_parts  (work units of this rank)
    float rmax;                   // max radius over live bodies
    unsigned step;                // steps since upload
};

struct StepResult {               // accumulated by the scatter kernel, consumed by plan
    unsigned rmax_bits;
    unsigned ticket;
    unsigned sym_next;            // work queue of the two-sided force kernel: next item
    unsigned mmax_bits;           // max mass / min radius over the surviving bodies (float bits; positive floats order like
    unsigned rmin_inv;            // unsigned ints; the minimum is kept as a maximum of 0x7f800000 - bits so that zero is
                                  // its neutral start): the bound on |F| behind the fixed-point scale
};

struct Counters {
    unsigned long long pairs;
    unsigned long long candidates;
    unsigned long long exact_chunks;
    unsigned long long fast_chunks;
    unsigned long long culled_parts;   // parts that ran without the pre-test (sorted stream)
    unsigned long long steps;
    unsigned cand_count;          // entries pushed to the candidate list this step
    int overflow_flag;            // sticky
    unsigned ev_count;            // event records logged since the last nb_events
    int ev_dropped;               // sticky
};

struct EventRec {                 // device-side event record (sorted and unpacked on the host)
    int step, i, j;
    unsigned key_kind;            // visit-order key << 1 | kind
};

struct StepParams {
    float dt, growth, grav;
    float soft2;                  // squared Plummer softening length (0 = off: bit-identical to the unsoftened arithmetic)
    int field_w, field_h;
    int coverage;
    int rank, world;
    int force_grid;
    int count_stats;              // 1: the force kernel counts fast/exact sub-chunks
    int iblock;                   // rows per i-block of the force-kernel variant in use (512 or 1024)
    int merge;                    // 0: the reference's absorb rule; 1: conserving lowest-index merge (opt-in)
    int resort;                   // this graph's end-of-step order rebuild: 1 full radix sort, 0 carry the previous order over
    int lg_parts_override;        // >= 0: fixed unit size (tuning experiments); -1: cost model
    int sort_min_n;               // > 0: full-coverage steps with n >= sort_min_n use the cell-sorted j stream
    int sym;                      // 1: steps on the cell-sorted order evaluate each unordered pair once (two-sided kernel)
    int sym_grid;                 //    its grid (resident CTAs x SMs)
    int sym_min_n;                //    smallest n that runs it on the bodies' own order (one GPU); sorted steps always do
    int sym_small;                //    which kernel takes those steps: 2 the warp-level one (default), 1 the CTA-level one
    int symw_grid;                //    grid of the warp-level kernel (CTAs of 128 threads)
    int symw_max_n;               //    the warp-level kernel takes steps with sym_min_n <= n < symw_max_n
    int symw_queue;               //    1: its warps take work items from an atomic counter instead of round robin (measurements)
    int symw_run;                 //    > 0: chunks per work item, instead of the plan's choice (measurements)
    int sym_rows;                 //    rows per lane: 4 (default), 8 (NB_FLAG_SYM_ROWS8)
};

struct DevState {
    float4 *pm;
    float2 *vel;
    float *jt;
    float *jts;                   // cell-sorted j-tiles (kSortedTileFloats each), rebuilt every step when desc->sorted
    unsigned *skey[2];            // radix sort ping-pong: Morton cell keys
    int *sidx[2];                 //                       body indices (sidx[0][slot] = body after the sort)
    int *sinv;                    // body -> slot
    int *remap;                   // compaction: body index before -> after (-1: removed); carries the sorted order over
    int *carry_count;             // survivors per 1024 slots of the previous order
    int *host_n;                  // device pointer to a pinned host int: the live body count after every step
    unsigned *shist;              // 256 x radix blocks
    unsigned char *post;          // world chunks of shard_cap * 24 B
    float2 *fpart;
    long long *facc;              // two-sided kernel: [slots][2] fixed-point force sums
    size_t slots;                 //                   tiles * 512
    // two-sided kernel on several GPUs: every rank evaluates its share of the triangle's blocks.  The forces meet in
    // one all-reduce (integer sum) of `facc`; the collision candidates a rank found travel in one allgather of
    // `xbuf`: per rank {XHeader; int2 {row, partner}[x_cap]}
    unsigned char *xbuf;
    size_t x_stride;              // bytes per rank region
    int x_cap;                    // candidate pairs a rank can contribute per step
    int *head;
    int2 *cand;
    EventRec *ev;
    int *tile_count;
    int *mhead, *mnext;           // conserving merge: per-root chains of absorbed bodies (by body index)
    StepDesc *desc;
    StepResult *res;
    Counters *ctr;
    int cap;                      // n_max
    int shard_cap;                // rows per rank chunk in `post` (multiple of 512)
    int post_row_bytes;           // 24, or 28 with the conserving merge (absorber plane)
    int cand_cap;
    int ev_cap;
};

struct XHeader {
    unsigned count;               // candidate pairs this rank found this step
    unsigned pad[3];
};
__host__ __device__ inline XHeader *x_header(const DevState &st, int rank)
{
    return reinterpret_cast<XHeader *>(st.xbuf + (size_t)rank * st.x_stride);
}
__host__ __device__ inline int2 *x_pairs(const DevState &st, int rank)
{
    return reinterpret_cast<int2 *>(st.xbuf + (size_t)rank * st.x_stride + sizeof(XHeader));
}

// a rank's chunk of `post`: float4 pm[shard_cap], float2 vel[shard_cap] and -- conserving merge only (post_row_bytes
// = 28) -- int absorber[shard_cap]: the lowest index among the row's body and its hit partners
__host__ __device__ inline float4 *post_pm(const DevState &st, int rank)
{
    return reinterpret_cast<float4 *>(st.post + (size_t)rank * st.shard_cap * st.post_row_bytes);
}
__host__ __device__ inline float2 *post_vel(const DevState &st, int rank)
{
    return reinterpret_cast<float2 *>(st.post + (size_t)rank * st.shard_cap * st.post_row_bytes + (size_t)st.shard_cap * 16);
}
__host__ __device__ inline int *post_abs(const DevState &st, int rank)
{
    return reinterpret_cast<int *>(st.post + (size_t)rank * st.shard_cap * st.post_row_bytes + (size_t)st.shard_cap * 24);
}

// Kernels this library itself has launched (or recorded into a graph being captured) on the calling thread: every
// launch site bumps it, the C ABI turns it into nb_stats.kernel_launches (a count, not a formula).
long long &launch_counter();
inline void count_launch(int k = 1) { launch_counter() += k; }


// ---- warp-level two-sided kernel: its work queue (shared with the plan) ------------------------------------------
// Work item ids.  First the 2 G "own" items (group g against one of the two chunks that make it up: they cost several
// times a normal item because every row meets itself there, so they are handed out first), then the triangle proper:
// group g against the chunk slots s0(g) = floor((2 g + 2) / run) .. S - 1; pairing g with G - 1 - g makes (almost)
// equal-length rows of L ids each, so an id decodes with one division; ids that fall off the end of a row pair are void.
struct WGeom {
    int G, C, S, run, L, ids;
};
__host__ __device__ inline WGeom symw_geom(int n, int run)
{
    WGeom w;
    w.G = (n + kWGroup - 1) / kWGroup;
    w.C = (n + kWChunk - 1) / kWChunk;
    w.run = run;
    w.S = (w.C + run - 1) / run;
    w.L = 2 * w.S - (2 * w.G + 2) / run + 2;      // an upper bound of every row pair's length (floors, empty last rows)
    if (w.L < 1) w.L = 1;
    w.ids = 2 * w.G + ((w.G + 1) / 2) * w.L;
    return w;
}
// id -> group g and chunk range [c_lo, c_hi); false: a void id
__host__ __device__ __forceinline__ bool symw_decode(const WGeom &w, int id, int &g, int &c_lo, int &c_hi)
{
    if (id < 2 * w.G) {                        // an own item: one chunk
        g = id >> 1;
        c_lo = 2 * g + (id & 1);
        c_hi = min(c_lo + 1, w.C);
        return c_lo < w.C;
    }
    id -= 2 * w.G;
    const int pi = id / w.L;
    int off = id - pi * w.L, s;
    const int s0a = (2 * pi + 2) / w.run, na = max(w.S - s0a, 0);
    const int gb = w.G - 1 - pi;
    if (off < na) {
        g = pi;
        s = s0a + off;
    } else {
        off -= na;
        if (gb == pi) return false;
        const int s0b = (2 * gb + 2) / w.run;
        if (off >= w.S - s0b) return false;       // (also when that row is empty: S - s0b <= 0)
        g = gb;
        s = s0b + off;
    }
    c_lo = max(s * w.run, 2 * g + 2);
    c_hi = min((s + 1) * w.run, w.C);
    return c_lo < c_hi;
}

// chunks per work item for n bodies on `warps` resident warps: the longest run that still leaves >= 12 items per warp
__host__ __device__ inline int symw_run(int n, int warps)
{
    int run = 8;
    while (run > 1 && symw_geom(n, run).ids < 12 * warps) run >>= 1;
    return run;
}

// kernels (nbody_kernels.cu)
cudaError_t launch_plan(const DevState &st, const StepParams &p, int n, cudaStream_t s);
cudaError_t launch_force(const DevState &st, const StepParams &p, int variant, cudaStream_t s);
cudaError_t launch_finish(const DevState &st, const StepParams &p, cudaStream_t s);
cudaError_t launch_force_sym(const DevState &st, const StepParams &p, cudaStream_t s);    // nbody_sym.cu
cudaError_t launch_force_symw(const DevState &st, const StepParams &p, cudaStream_t s);   // nbody_symw.cu
int force_symw_occupancy(int *regs);
cudaError_t launch_sym_chain(const DevState &st, const StepParams &p, cudaStream_t s);    // sharded two-sided kernel: after the allgather of xbuf
cudaError_t launch_compact(const DevState &st, const StepParams &p, bool recount, cudaStream_t s);
cudaError_t launch_merge(const DevState &st, const StepParams &p, cudaStream_t s);
cudaError_t launch_sort(const DevState &st, const StepParams &p, cudaStream_t s);      // nbody_sort.cu
size_t sort_hist_entries(int cap);
// the four arrays of a BodiesData block (src/nbody.cu:66-77), wherever they lie on the device
cudaError_t launch_ingest(const DevState &st, const float *pos, const float *vel, const float *mass, const float *rad, int n,
                          cudaStream_t s);
cudaError_t launch_export(const DevState &st, float *block, int n, cudaStream_t s);
cudaError_t launch_render(const DevState &st, int n, unsigned char *img, int w, int h, int field_w, int field_h,
                          cudaStream_t s);
int force_occupancy(int variant, int *regs, int *threads, int *iblock);   // resident CTAs per SM of the force kernel
int force_sym_occupancy(int rows, int *regs);                             // same for the two-sided kernel (4 or 8 rows per lane)
constexpr int kForceVariants = 6;
size_t fpart_slabs(int force_grid, int shard_cap, int iblock);   // slabs of `iblock` float2 needed
void plan_host(StepDesc *d, const StepParams *p, int n);   // the device plan, run on the host (tests, sharding)
// Fixed-point scale of the two-sided kernel's force sums for n bodies with masses <= mmax and radii >= rmin in a field
// of half-width `field`: 2^k with n * mmax / (2 rmin)^2 * 2^k < 2^62.  False when that leaves too little resolution.
__host__ __device__ inline bool sym_scale(int n, float mmax, float rmin, int field, float *fscale, double *finv)
{
    if (n <= 0 || !(mmax > 0.f) || !(rmin > 0.f)) return false;
    // a pair that is not a hit is at least r_i + r_j >= 2 rmin apart, so no body ever sees more than this
    const double bound = (double)n * (double)mmax / (4.0 * (double)rmin * (double)rmin);
    if (!(bound > 0.0) || !(bound < 1.0e300)) return false;
    const int k = 61 - ilogb(bound);              // bound < 2^(ilogb + 1)  =>  bound * 2^k < 2^62
    if (k < -120 || k > 120) return false;
    // one unit is 2^-k; against a typical force n mbar / field^2 that is about 2^-62 (field / (2 rmin))^2 relative:
    // demand 30 bits below it (float32 carries 24)
    const double ratio = (double)field / (2.0 * (double)rmin);
    if (ratio * ratio > 2147483648.0) return false;
    *fscale = ldexpf(1.0f, k);
    *finv = ldexp(1.0, -k);
    return true;
}
void sym_block_host(int b, int Q, int *R, int *C);         // two-sided kernel: block b of the queue order -> super-tiles (R, C)
int sym_block_index_host(int X, int Y, int Q);             //                   and back

}  // namespace nb
