// nbody_api.cu -- the C ABI of include/nbody_b200.h over the kernels in nbody_kernels.cu.
//
// Replaces the body of the reference's main loop (src/nbody.cu:460-545): where the reference
// re-allocates, uploads, launches, downloads and compacts on the host every step, a context here
// owns all device memory once, keeps n on the device and replays one CUDA graph per step.
// There is no CPU fallback: without a usable sm_100 device nb_create fails.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>   // types only: the library itself is bound at run time (see nccl_api below)

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <new>
#include <vector>

#include "nbody_device.cuh"

using namespace nb;

// the two-sided (pair-halving) force kernel is used wherever it applies unless NB_FLAG_ONE_SIDED is given
static constexpr bool kPairHalvingDefault = true;

struct nb_ctx {
    nb_params par;
    DevState st;
    StepParams sp;
    int device;
    int sm_count;
    int force_regs;
    int sym_regs;
    int variant;
    int force_threads;
    cudaStream_t stream;
    cudaGraphExec_t graph[3];          // [0] plain step, [1] step + the cell-sorted order carried over, [2] step + full re-sort
    bool graph_ready[3];
    StepParams sp_plain;               // sp with the sorted order switched off (what graph[0] bakes in)
    StepParams sp_resort;              // sp with a full radix sort at the end of the step (graph[2]); sp itself carries the order over
    unsigned long long host_step;      // steps enqueued since nb_upload: decides carry / re-sort, the same on every rank
    volatile int *host_n;              // pinned + mapped: live body count, written by the device every step
    cudaEvent_t ring[16];              // bounds how far the host runs ahead when it picks a graph per step
    unsigned long long ring_pos;
    cudaEvent_t ev0, ev1;
    std::vector<cudaEvent_t> fev;     // per-step force-kernel brackets of nb_step_timed
    ncclComm_t comm;
    bool comm_ready;
    void *dev_block;                   // staging for upload/download: the BodiesData block, 24 * cap bytes
    unsigned char *dev_img;
    size_t dev_img_bytes;
    long long launches;                // kernels of this library executed on behalf of this context (direct + graph nodes)
    long long graph_nodes[3];          // kernel nodes of graph[k]
    char err[512];
};

// counts the kernels launched directly (not captured) while it is alive
struct LaunchScope {
    nb_ctx *c;
    long long c0;
    explicit LaunchScope(nb_ctx *ctx) : c(ctx), c0(launch_counter()) {}
    ~LaunchScope() { c->launches += launch_counter() - c0; }
};

static char g_create_err[512] = "";

// NCCL is bound lazily with dlopen so that (a) single-GPU users need no NCCL at all and (b) inside a
// process that already loaded an NCCL (e.g. torch's bundled one) we share that copy instead of
// bringing a second one.
struct NcclApi {
    void *lib;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    const char *(*GetErrorString)(ncclResult_t);
};
static NcclApi *nccl_api()
{
    static NcclApi api = {};
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (h) {
            api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
            api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
            api.AllGather = (decltype(api.AllGather))dlsym(h, "ncclAllGather");
            api.AllReduce = (decltype(api.AllReduce))dlsym(h, "ncclAllReduce");
            api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
            api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
            if (api.GetUniqueId && api.CommInitRank && api.AllGather && api.AllReduce && api.CommDestroy && api.GetErrorString) api.lib = h;
        }
    }
    return api.lib ? &api : nullptr;
}

static void set_err(nb_ctx *c, const char *fmt, ...)
{
    char *dst = c ? c->err : g_create_err;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(dst, 512, fmt, ap);
    va_end(ap);
}

#define NB_CUDA(c, call)                                                                              \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            set_err(c, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__);   \
            return NB_ERR_CUDA;                                                                       \
        }                                                                                             \
    } while (0)

#define NB_NCCL(c, call)                                                                              \
    do {                                                                                              \
        ncclResult_t r_ = (call);                                                                     \
        if (r_ != ncclSuccess) {                                                                      \
            set_err(c, "%s failed: %s (%s:%d)", #call, nccl_api()->GetErrorString(r_), __FILE__, __LINE__); \
            return NB_ERR_COMM;                                                                       \
        }                                                                                             \
    } while (0)

extern "C" {

int nb_version(void) { return NB_VERSION; }

const char *nb_last_error(const nb_ctx *ctx) { return ctx ? ctx->err : g_create_err; }

static void free_all(nb_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    for (int k = 0; k < 3; ++k)
        if (c->graph_ready[k]) cudaGraphExecDestroy(c->graph[k]);
    for (cudaEvent_t e : c->ring)
        if (e) cudaEventDestroy(e);
    if (c->host_n) cudaFreeHost((void *)c->host_n);
    if (c->comm_ready) nccl_api()->CommDestroy(c->comm);
    for (cudaEvent_t e : c->fev) cudaEventDestroy(e);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->stream) cudaStreamDestroy(c->stream);
    void *ptrs[] = {c->st.remap, c->st.carry_count, c->st.mhead, c->st.mnext, c->st.sinv, c->st.jts, c->st.skey[0], c->st.skey[1], c->st.sidx[0], c->st.sidx[1], c->st.shist,
                    c->st.pm,   c->st.vel, c->st.jt,         c->st.post, c->st.fpart, c->st.facc, c->st.xbuf, c->st.head, c->st.cand,
                    c->st.ev,   c->st.tile_count, c->st.desc, c->st.res,  c->st.ctr,   c->dev_block, c->dev_img};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    delete c;
}

int nb_create(nb_ctx **out, const nb_params *params)
{
    if (!out || !params) {
        set_err(nullptr, "nb_create: null argument");
        return NB_ERR_INVALID;
    }
    *out = nullptr;
    if (params->n_max <= 0 || params->n_max > (1 << 26) || params->field_w <= 0 || params->field_h <= 0 ||
        (params->coverage != NB_COVERAGE_REFERENCE && params->coverage != NB_COVERAGE_FULL) ||
        (params->world > 1 && (params->rank < 0 || params->rank >= params->world))) {
        set_err(nullptr, "nb_create: invalid parameters");
        return NB_ERR_INVALID;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        set_err(nullptr, "nb_create: no CUDA device (this library has no CPU fallback)");
        return NB_ERR_CUDA;
    }
    if (params->device < 0 || params->device >= ndev) {
        set_err(nullptr, "nb_create: device %d out of range (%d devices)", params->device, ndev);
        return NB_ERR_INVALID;
    }
    cudaDeviceProp prop;
    NB_CUDA(nullptr, cudaGetDeviceProperties(&prop, params->device));
    if (prop.major != 10) {
        set_err(nullptr, "nb_create: device %d is sm_%d%d; this library is built for sm_100a only", params->device,
                prop.major, prop.minor);
        return NB_ERR_CUDA;
    }
    nb_ctx *c = new (std::nothrow) nb_ctx();
    if (!c) return NB_ERR_INVALID;
    c->par = *params;
    c->device = params->device;
    c->sm_count = prop.multiProcessorCount;
    c->err[0] = 0;
    {
        const cudaError_t e = cudaSetDevice(c->device);
        if (e != cudaSuccess) {
            set_err(nullptr, "nb_create: cudaSetDevice(%d): %s", c->device, cudaGetErrorString(e));
            delete c;
            return NB_ERR_CUDA;
        }
    }

    const int world = params->world > 1 ? params->world : 1;
    DevState &st = c->st;
    st.cap = params->n_max;
    st.cand_cap = params->candidate_capacity > 0 ? params->candidate_capacity
                                                 : (int)std::min<long long>(std::max<long long>(4LL * st.cap, 65536), 1LL << 30);
    st.ev_cap = params->event_capacity > 0 ? params->event_capacity : 0;

    c->variant = (params->flags >> NB_FLAG_VARIANT_SHIFT) & 0xf;
    if (params->flags & NB_FLAG_SCALAR_FORCE) c->variant = 4;
    if (c->variant >= kForceVariants) {
        set_err(nullptr, "nb_create: unknown force-kernel variant %d", c->variant);
        free_all(c);
        return NB_ERR_INVALID;
    }
    int iblock = kIBlock;
    const int occ = force_occupancy(c->variant, &c->force_regs, &c->force_threads, &iblock);
    const int iblocks_total = (st.cap + iblock - 1) / iblock;
    st.shard_cap = (iblocks_total + world - 1) / world * iblock;
    if (occ <= 0) {
        set_err(nullptr, "nb_create: force kernel does not fit on an SM");
        free_all(c);
        return NB_ERR_CUDA;
    }
    StepParams &sp = c->sp;
    sp.dt = params->dt;
    sp.growth = params->growth;
    sp.grav = params->grav != 0.f ? params->grav : NB_GRAV_CONSTANT;
    sp.soft2 = params->softening > 0.f ? params->softening * params->softening : 0.f;
    sp.field_w = params->field_w;
    sp.field_h = params->field_h;
    sp.coverage = params->coverage;
    sp.rank = world > 1 ? params->rank : 0;
    sp.world = world;
    sp.force_grid = c->sm_count * occ;
    sp.count_stats = 1;
    sp.iblock = iblock;
    sp.merge = (params->flags & NB_FLAG_MERGE_CONSERVING) ? 1 : 0;
    // the cell-sorted order pays for itself from about 6e4 bodies on (profiles/r01_sort_threshold.log); it needs
    // all-pairs coverage (the reference's excluded windows are defined by body index) and is sized in only if the
    // capacity can ever reach the threshold
    sp.sort_min_n = 0;
    if (params->coverage == NB_COVERAGE_FULL && !(params->flags & NB_FLAG_NO_SORT)) {
        const int min_n = params->sort_min_n > 0 ? params->sort_min_n : NB_SORT_MIN_N_DEFAULT;
        if (st.cap >= min_n) sp.sort_min_n = min_n;
    }
    // two-sided force kernel: all-pairs coverage; on the cell-sorted order, and on one GPU also on the bodies' own order
    sp.sym = 0;
    sp.sym_grid = 0;
    sp.sym_min_n = 0;
    sp.sym_rows = (params->flags & NB_FLAG_SYM_ROWS8) ? 8 : 4;
    if (const char *e = getenv("NBODY_B200_SYM_ROWS")) sp.sym_rows = atoi(e) == 8 ? 8 : 4;                // tuning only
    if (params->coverage == NB_COVERAGE_FULL && !(params->flags & NB_FLAG_ONE_SIDED) &&
        ((params->flags & NB_FLAG_PAIR_HALVING) || kPairHalvingDefault) && (sp.sort_min_n > 0 || world == 1)) {
        const int socc = force_sym_occupancy(sp.sym_rows, &c->sym_regs);
        if (socc > 0) {
            sp.sym = 1;
            sp.sym_grid = c->sm_count * socc;
            sp.sym_min_n = kSymMinNDefault;
            sp.sym_small = 2;
            sp.symw_queue = 1;
            sp.symw_max_n = kSymWarpMaxN;
            if (const char *e = getenv("NBODY_B200_SYM_MIN_N")) sp.sym_min_n = atoi(e);           // tuning only
            if (const char *e = getenv("NBODY_B200_SYMW_MAX_N")) sp.symw_max_n = atoi(e);
            if (const char *e = getenv("NBODY_B200_SYM_SMALL")) sp.sym_small = atoi(e) == 1 ? 1 : 2;
            if (const char *e = getenv("NBODY_B200_SYMW_QUEUE")) sp.symw_queue = atoi(e) != 0;
            if (const char *e = getenv("NBODY_B200_SYMW_RUN")) sp.symw_run = atoi(e) > 0 ? atoi(e) : 0;
            int wregs = 0;
            sp.symw_grid = c->sm_count * std::max(1, force_symw_occupancy(&wregs));
        }
    }
    sp.lg_parts_override = -1;
    if (const char *e = getenv("NBODY_B200_LG_PARTS")) sp.lg_parts_override = atoi(e) < 0 ? -1 : (atoi(e) > kMaxLgParts ? kMaxLgParts : atoi(e));   // tuning only

    const size_t tiles = (size_t)(st.cap + kTJ - 1) / kTJ + 1;
    const size_t ctiles = (size_t)(st.cap + kCompactTile - 1) / kCompactTile;
#define NB_ALLOC(ptr, bytes)                                                                    \
    do {                                                                                        \
        cudaError_t e_ = cudaMalloc((void **)&(ptr), (bytes));                                  \
        if (e_ != cudaSuccess) {                                                                \
            set_err(nullptr, "cudaMalloc(%s, %zu) failed: %s", #ptr, (size_t)(bytes), cudaGetErrorString(e_)); \
            free_all(c);                                                                        \
            return NB_ERR_CUDA;                                                                 \
        }                                                                                       \
    } while (0)
    NB_ALLOC(st.pm, sizeof(float4) * (size_t)st.cap);
    NB_ALLOC(st.vel, sizeof(float2) * (size_t)st.cap);
    NB_ALLOC(st.jt, (size_t)kTileBytes * tiles);
    st.post_row_bytes = (params->flags & NB_FLAG_MERGE_CONSERVING) ? 28 : 24;
    NB_ALLOC(st.post, (size_t)world * st.shard_cap * st.post_row_bytes);
    if (sp.sort_min_n > 0) {
        NB_ALLOC(st.jts, sizeof(float) * kSortedTileFloats * tiles);
        for (int k = 0; k < 2; ++k) {
            NB_ALLOC(st.skey[k], sizeof(unsigned) * (size_t)st.cap);
            NB_ALLOC(st.sidx[k], sizeof(int) * (size_t)st.cap);
        }
        NB_ALLOC(st.shist, sizeof(unsigned) * sort_hist_entries(st.cap));
        NB_ALLOC(st.sinv, sizeof(int) * (size_t)st.cap);
        NB_ALLOC(st.remap, sizeof(int) * (size_t)st.cap);
        NB_ALLOC(st.carry_count, sizeof(int) * ((size_t)st.cap / 1024 + 1));
    }
    if (sp.sym) {
        st.slots = tiles * kTJ;
        NB_ALLOC(st.facc, sizeof(long long) * 2 * st.slots);
        if (world > 1) {
            // candidate pairs one rank may contribute per step: its share of the pairs, with room for crowded starts
            st.x_cap = params->candidate_capacity > 0 ? params->candidate_capacity : std::max(131072, st.cap / 4);
            st.x_stride = (sizeof(XHeader) + (size_t)st.x_cap * sizeof(int2) + 255) / 256 * 256;
            NB_ALLOC(st.xbuf, st.x_stride * (size_t)world);
            st.cand_cap = (int)std::min<long long>(std::max<long long>(st.cand_cap, (long long)world * st.x_cap), 1LL << 30);
        }
    }
    NB_ALLOC(st.fpart, sizeof(float2) * iblock * fpart_slabs(sp.force_grid, st.shard_cap, iblock));
    NB_ALLOC(st.head, sizeof(int) * (size_t)st.cap);
    NB_ALLOC(st.cand, sizeof(int2) * (size_t)st.cand_cap);
    if (st.ev_cap > 0) NB_ALLOC(st.ev, sizeof(EventRec) * (size_t)st.ev_cap);
    NB_ALLOC(st.tile_count, sizeof(int) * ctiles);
    if (sp.merge) {
        NB_ALLOC(st.mhead, sizeof(int) * (size_t)st.cap);
        NB_ALLOC(st.mnext, sizeof(int) * (size_t)st.cap);
    }
    NB_ALLOC(st.desc, sizeof(StepDesc));
    NB_ALLOC(st.res, sizeof(StepResult));
    NB_ALLOC(st.ctr, sizeof(Counters));
    NB_ALLOC(c->dev_block, (size_t)24 * (st.cap + 4 * world));      // room for the padded arrays of a sharded upload
#undef NB_ALLOC
    sp.resort = 0;
    c->sp_resort = sp;
    c->sp_resort.resort = 1;
    c->sp_plain = sp;
    c->sp_plain.sort_min_n = 0;
    if (world > 1) c->sp_plain.sym = 0;            // sharded: the two-sided kernel runs on the sorted order only
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaHostAlloc((void **)&c->host_n, sizeof(int), cudaHostAllocMapped);
    if (e == cudaSuccess) {
        *c->host_n = 0;
        e = cudaHostGetDevicePointer((void **)&st.host_n, (void *)c->host_n, 0);
    }
    for (int k = 0; k < 16 && e == cudaSuccess; ++k) e = cudaEventCreateWithFlags(&c->ring[k], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev1);
    if (e == cudaSuccess) e = cudaMemsetAsync(st.res, 0, sizeof(StepResult), c->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(st.head, 0xff, sizeof(int) * (size_t)st.cap, c->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(st.tile_count, 0, sizeof(int) * ctiles, c->stream);
    if (e == cudaSuccess && sp.merge) e = cudaMemsetAsync(st.mhead, 0xff, sizeof(int) * (size_t)st.cap, c->stream);
    if (e == cudaSuccess && st.facc) e = cudaMemsetAsync(st.facc, 0, sizeof(long long) * 2 * st.slots, c->stream);
    if (e == cudaSuccess) e = launch_plan(st, sp, 0, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) {
        set_err(nullptr, "nb_create: %s", cudaGetErrorString(e));
        free_all(c);
        return NB_ERR_CUDA;
    }
    *out = c;
    return NB_OK;
}

void nb_destroy(nb_ctx *ctx) { free_all(ctx); }

int nb_upload(nb_ctx *c, const void *bodies, int n)
{
    if (!c || (!bodies && n > 0) || n < 0) return NB_ERR_INVALID;
    if (n > c->st.cap) {
        set_err(c, "nb_upload: n = %d exceeds n_max = %d", n, c->st.cap);
        return NB_ERR_CAPACITY;
    }
    NB_CUDA(c, cudaSetDevice(c->device));
    NB_CUDA(c, cudaStreamSynchronize(c->stream));
    LaunchScope scope(c);
    c->ring_pos = 0;
    const float *host = static_cast<const float *>(bodies);
    float *dev = static_cast<float *>(c->dev_block);
    const float *d_pos = dev, *d_vel = dev + 2 * (size_t)n, *d_mass = dev + 4 * (size_t)n, *d_rad = dev + 5 * (size_t)n;
    if (c->sp.world > 1 && c->comm_ready && n >= 4 * c->sp.world) {
        // sharded upload: every rank holds the same host block, so each copies only its 1 / world of every array over
        // PCIe and the rest arrives from the peers over NVLink (four in-place allgathers into arrays padded to world
        // equal chunks)
        const int W = c->sp.world, r = c->sp.rank;
        const size_t chunk = (((size_t)n + W - 1) / W + 3) / 4 * 4, padded = chunk * W;
        const size_t lo = std::min<size_t>((size_t)r * chunk, (size_t)n), hi = std::min<size_t>(lo + chunk, (size_t)n);
        float *p_pos = dev, *p_vel = dev + 2 * padded, *p_mass = dev + 4 * padded, *p_rad = dev + 5 * padded;
        if (hi > lo) {
            NB_CUDA(c, cudaMemcpyAsync(p_pos + 2 * lo, host + 2 * lo, 8 * (hi - lo), cudaMemcpyHostToDevice, c->stream));
            NB_CUDA(c, cudaMemcpyAsync(p_vel + 2 * lo, host + 2 * (size_t)n + 2 * lo, 8 * (hi - lo), cudaMemcpyHostToDevice, c->stream));
            NB_CUDA(c, cudaMemcpyAsync(p_mass + lo, host + 4 * (size_t)n + lo, 4 * (hi - lo), cudaMemcpyHostToDevice, c->stream));
            NB_CUDA(c, cudaMemcpyAsync(p_rad + lo, host + 5 * (size_t)n + lo, 4 * (hi - lo), cudaMemcpyHostToDevice, c->stream));
        }
        NB_NCCL(c, nccl_api()->AllGather(p_pos + 2 * r * chunk, p_pos, 8 * chunk, ncclChar, c->comm, c->stream));
        NB_NCCL(c, nccl_api()->AllGather(p_vel + 2 * r * chunk, p_vel, 8 * chunk, ncclChar, c->comm, c->stream));
        NB_NCCL(c, nccl_api()->AllGather(p_mass + r * chunk, p_mass, 4 * chunk, ncclChar, c->comm, c->stream));
        NB_NCCL(c, nccl_api()->AllGather(p_rad + r * chunk, p_rad, 4 * chunk, ncclChar, c->comm, c->stream));
        d_pos = p_pos;
        d_vel = p_vel;
        d_mass = p_mass;
        d_rad = p_rad;
    } else if (n > 0) {
        NB_CUDA(c, cudaMemcpyAsync(c->dev_block, bodies, (size_t)24 * n, cudaMemcpyHostToDevice, c->stream));
    }
    NB_CUDA(c, cudaMemsetAsync(c->st.res, 0, sizeof(StepResult), c->stream));
    NB_CUDA(c, cudaMemsetAsync(c->st.tile_count, 0, sizeof(int) * ((size_t)(c->st.cap + kCompactTile - 1) / kCompactTile), c->stream));
    if (c->st.facc) NB_CUDA(c, cudaMemsetAsync(c->st.facc, 0, sizeof(long long) * 2 * c->st.slots, c->stream));
    NB_CUDA(c, launch_ingest(c->st, d_pos, d_vel, d_mass, d_rad, n, c->stream));
    NB_CUDA(c, launch_plan(c->st, c->sp, n, c->stream));
    if (c->sp.sort_min_n > 0) NB_CUDA(c, launch_sort(c->st, c->sp_resort, c->stream));
    c->host_step = 0;
    // the caller may reuse `bodies` as soon as we return
    NB_CUDA(c, cudaStreamSynchronize(c->stream));
    return NB_OK;
}

static int check_flags(nb_ctx *c, const Counters &ctr)
{
    if (ctr.overflow_flag) {
        set_err(c, "collision candidate list overflowed (capacity %d): results after that step are invalid", c->st.cand_cap);
        return NB_ERR_CANDIDATE_OVERFLOW;
    }
    return NB_OK;
}

static int fetch_state(nb_ctx *c, StepDesc *d, Counters *ctr)
{
    NB_CUDA(c, cudaSetDevice(c->device));
    if (d) NB_CUDA(c, cudaMemcpyAsync(d, c->st.desc, sizeof(StepDesc), cudaMemcpyDeviceToHost, c->stream));
    if (ctr) NB_CUDA(c, cudaMemcpyAsync(ctr, c->st.ctr, sizeof(Counters), cudaMemcpyDeviceToHost, c->stream));
    NB_CUDA(c, cudaStreamSynchronize(c->stream));
    return NB_OK;
}

int nb_sync(nb_ctx *c)
{
    if (!c) return NB_ERR_INVALID;
    Counters ctr;
    int rc = fetch_state(c, nullptr, &ctr);
    if (rc != NB_OK) return rc;
    return check_flags(c, ctr);
}

int nb_num_bodies(nb_ctx *c, int *n)
{
    if (!c || !n) return NB_ERR_INVALID;
    StepDesc d;
    Counters ctr;
    int rc = fetch_state(c, &d, &ctr);
    if (rc != NB_OK) return rc;
    *n = d.n;
    return check_flags(c, ctr);
}

int nb_download(nb_ctx *c, void *bodies, int capacity_n, int *n_out)
{
    if (!c || !n_out) return NB_ERR_INVALID;
    StepDesc d;
    Counters ctr;
    int rc = fetch_state(c, &d, &ctr);
    if (rc != NB_OK) return rc;
    *n_out = d.n;
    if (d.n > capacity_n || (!bodies && d.n > 0)) {
        set_err(c, "nb_download: host buffer holds %d bodies, %d are live", capacity_n, d.n);
        return NB_ERR_CAPACITY;
    }
    LaunchScope scope(c);
    if (d.n > 0) {
        NB_CUDA(c, launch_export(c->st, (float *)c->dev_block, d.n, c->stream));
        NB_CUDA(c, cudaMemcpyAsync(bodies, c->dev_block, (size_t)24 * d.n, cudaMemcpyDeviceToHost, c->stream));
        NB_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    return check_flags(c, ctr);
}

// one step's launches on the context's stream; f0/f1 (optional) bracket the force kernel, marks (optional,
// 6 events) separate force | finish | allgather | compaction | sort
static int enqueue_step(nb_ctx *c, const StepParams &sp, cudaEvent_t f0, cudaEvent_t f1, cudaEvent_t *marks = nullptr)
{
    if (f0) NB_CUDA(c, cudaEventRecord(f0, c->stream));
    if (marks) NB_CUDA(c, cudaEventRecord(marks[0], c->stream));
    NB_CUDA(c, launch_force(c->st, sp, c->variant, c->stream));
    if (f1) NB_CUDA(c, cudaEventRecord(f1, c->stream));
    if (marks) NB_CUDA(c, cudaEventRecord(marks[1], c->stream));
    if (sp.sym && c->sp.world > 1) {
        // two-sided kernel on several GPUs: every rank holds a part of every body's force (fixed-point sums: one exact
        // integer all-reduce brings them together, the same bits on every rank and as on one GPU) and of the candidates
        NB_NCCL(c, nccl_api()->AllReduce(c->st.facc, c->st.facc, 2 * c->st.slots, ncclInt64, ncclSum, c->comm, c->stream));
        NB_NCCL(c, nccl_api()->AllGather(c->st.xbuf + (size_t)c->sp.rank * c->st.x_stride, c->st.xbuf, c->st.x_stride, ncclChar, c->comm, c->stream));
        NB_CUDA(c, launch_sym_chain(c->st, sp, c->stream));
    }
    NB_CUDA(c, launch_finish(c->st, sp, c->stream));
    if (marks) NB_CUDA(c, cudaEventRecord(marks[2], c->stream));
    if (c->sp.world > 1) {
        const size_t chunk = (size_t)c->st.shard_cap * c->st.post_row_bytes;
        NB_NCCL(c, nccl_api()->AllGather(c->st.post + (size_t)c->sp.rank * chunk, c->st.post, chunk, ncclChar, c->comm, c->stream));
    }
    if (sp.merge) NB_CUDA(c, launch_merge(c->st, sp, c->stream));
    if (marks) NB_CUDA(c, cudaEventRecord(marks[3], c->stream));
    NB_CUDA(c, launch_compact(c->st, sp, sp.merge != 0, c->stream));
    if (marks) NB_CUDA(c, cudaEventRecord(marks[4], c->stream));
    if (sp.sort_min_n > 0) NB_CUDA(c, launch_sort(c->st, sp, c->stream));   // shadow order of the next step
    if (marks) NB_CUDA(c, cudaEventRecord(marks[5], c->stream));
    return NB_OK;
}

static const StepParams &graph_params(const nb_ctx *c, int which);

static int ensure_graph(nb_ctx *c, int which)
{
    if (c->graph_ready[which]) return NB_OK;
    cudaGraph_t g = nullptr;
    NB_CUDA(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    const long long c0 = launch_counter();
    int rc = enqueue_step(c, graph_params(c, which), nullptr, nullptr);
    c->graph_nodes[which] = launch_counter() - c0;      // recorded, not executed: they count once per replay
    launch_counter() = c0;
    cudaError_t e = cudaStreamEndCapture(c->stream, &g);
    if (rc != NB_OK) {
        if (g) cudaGraphDestroy(g);
        return rc;
    }
    NB_CUDA(c, e);
    e = cudaGraphInstantiate(&c->graph[which], g, 0);
    cudaGraphDestroy(g);
    NB_CUDA(c, e);
    c->graph_ready[which] = true;
    return NB_OK;
}

// The step parameters of graph `which`: 0 plain, 1 sorted order carried over at the end of the step, 2 full re-sort.
static const StepParams &graph_params(const nb_ctx *c, int which)
{
    return which == 2 ? c->sp_resort : (which ? c->sp : c->sp_plain);
}

// Which graph the next step uses.  Sorted or plain: by the live body count the device mirrors to pinned host memory (n
// only shrinks, so a stale -- too large -- count can only pick the richer graph, whose extra kernels then exit at once).
// Carry or re-sort: by the host's own step counter (every kResortEvery-th step ends with a full radix sort), which is
// the same on every rank.
static int pick_graph(nb_ctx *c, bool sorted_graph)
{
    if (!sorted_graph) return 0;
    return (c->host_step + 1) % kResortEvery == 0 ? 2 : 1;
}

// n_steps steps through the CUDA graphs (or plain launches with NB_FLAG_NO_GRAPH).  To keep the body count reasonably
// fresh a single-GPU host stays at most 16 steps ahead of its device.
static int run_steps(nb_ctx *c, int n_steps)
{
    int rc;
    LaunchScope scope(c);
    const bool sortable = c->sp.sort_min_n > 0;
    const bool entry_sorted = sortable && *c->host_n >= c->sp.sort_min_n;
    for (int s = 0; s < n_steps; ++s) {
        bool sorted_graph = false;
        if (sortable && c->sp.world > 1) {
            // sharded: the graphs hold different collectives (the sorted ones also exchange the two-sided kernel's sums and
            // candidates), so the choice must be the same on every rank and cannot depend on how far each host has read
            // ahead of its device.  It is made once per call from the body count at its start -- every sharded nb_step
            // and nb_upload ends with a stream synchronisation, and the replicas hold the same bodies, so every rank
            // reads the same number.  Should the count cross the threshold inside the call, the rich graph's sort and
            // two-sided kernels exit at once (the step descriptor, identical on every rank, does not name them).
            sorted_graph = entry_sorted;
        } else if (sortable) {
            cudaEvent_t &slot = c->ring[c->ring_pos % 16];
            if (c->ring_pos >= 16) NB_CUDA(c, cudaEventSynchronize(slot));
            sorted_graph = *c->host_n >= c->sp.sort_min_n;
        }
        const int which = pick_graph(c, sorted_graph);
        if (c->par.flags & NB_FLAG_NO_GRAPH) {
            if ((rc = enqueue_step(c, graph_params(c, which), nullptr, nullptr)) != NB_OK) return rc;
        } else {
            if ((rc = ensure_graph(c, which)) != NB_OK) return rc;
            NB_CUDA(c, cudaGraphLaunch(c->graph[which], c->stream));
            c->launches += c->graph_nodes[which];
        }
        ++c->host_step;
        if (sortable && c->sp.world <= 1) {
            NB_CUDA(c, cudaEventRecord(c->ring[c->ring_pos % 16], c->stream));
            ++c->ring_pos;
        }
    }
    return NB_OK;
}

static int step_precheck(nb_ctx *c, int n_steps)
{
    if (!c || n_steps < 0) return NB_ERR_INVALID;
    if (c->sp.world > 1 && !c->comm_ready) {
        set_err(c, "nb_step: world = %d but nb_comm_init has not been called", c->sp.world);
        return NB_ERR_COMM;
    }
    NB_CUDA(c, cudaSetDevice(c->device));
    return NB_OK;
}

int nb_step(nb_ctx *c, int n_steps)
{
    int rc = step_precheck(c, n_steps);
    if (rc != NB_OK) return rc;
    if ((rc = run_steps(c, n_steps)) != NB_OK) return rc;
    if (c->sp.world > 1) NB_CUDA(c, cudaStreamSynchronize(c->stream));
    return NB_OK;
}

int nb_step_timed(nb_ctx *c, int n_steps, float *ms_total, float *ms_force)
{
    int rc = step_precheck(c, n_steps);
    if (rc != NB_OK) return rc;
    if (ms_total) *ms_total = 0.f;
    if (ms_force) *ms_force = 0.f;
    if (!ms_force) {            // whole region only: keep the graph path
        NB_CUDA(c, cudaEventRecord(c->ev0, c->stream));
        if ((rc = run_steps(c, n_steps)) != NB_OK) return rc;
        NB_CUDA(c, cudaEventRecord(c->ev1, c->stream));
        NB_CUDA(c, cudaEventSynchronize(c->ev1));
        if (ms_total) NB_CUDA(c, cudaEventElapsedTime(ms_total, c->ev0, c->ev1));
        return NB_OK;
    }
    while ((int)c->fev.size() < 2 * n_steps) {
        cudaEvent_t e;
        NB_CUDA(c, cudaEventCreate(&e));
        c->fev.push_back(e);
    }
    NB_CUDA(c, cudaEventRecord(c->ev0, c->stream));
    LaunchScope scope(c);
    for (int s = 0; s < n_steps; ++s) {
        if ((rc = enqueue_step(c, graph_params(c, pick_graph(c, c->sp.sort_min_n > 0)), c->fev[2 * s], c->fev[2 * s + 1])) != NB_OK) return rc;
        ++c->host_step;
    }
    NB_CUDA(c, cudaEventRecord(c->ev1, c->stream));
    NB_CUDA(c, cudaEventSynchronize(c->ev1));
    if (ms_total) NB_CUDA(c, cudaEventElapsedTime(ms_total, c->ev0, c->ev1));
    float sum = 0.f;
    for (int s = 0; s < n_steps; ++s) {
        float ms = 0.f;
        NB_CUDA(c, cudaEventElapsedTime(&ms, c->fev[2 * s], c->fev[2 * s + 1]));
        sum += ms;
    }
    *ms_force = sum;
    return NB_OK;
}

int nb_step_profile(nb_ctx *c, int n_steps, float ms[5])
{
    int rc = step_precheck(c, n_steps);
    if (rc != NB_OK) return rc;
    if (!ms) return NB_ERR_INVALID;
    while ((int)c->fev.size() < 6) {
        cudaEvent_t e;
        NB_CUDA(c, cudaEventCreate(&e));
        c->fev.push_back(e);
    }
    for (int k = 0; k < 5; ++k) ms[k] = 0.f;
    LaunchScope scope(c);
    for (int s = 0; s < n_steps; ++s) {
        if ((rc = enqueue_step(c, graph_params(c, pick_graph(c, c->sp.sort_min_n > 0)), nullptr, nullptr, c->fev.data())) != NB_OK) return rc;
        ++c->host_step;
        NB_CUDA(c, cudaEventSynchronize(c->fev[5]));
        for (int k = 0; k < 5; ++k) {
            float t = 0.f;
            NB_CUDA(c, cudaEventElapsedTime(&t, c->fev[k], c->fev[k + 1]));
            ms[k] += t;
        }
    }
    return NB_OK;
}

int nb_get_stats(nb_ctx *c, nb_stats *out)
{
    if (!c || !out) return NB_ERR_INVALID;
    StepDesc d;
    Counters ctr;
    int rc = fetch_state(c, &d, &ctr);
    if (rc != NB_OK) return rc;
    memset(out, 0, sizeof(*out));
    out->steps = (int64_t)ctr.steps;
    out->pairs = (int64_t)ctr.pairs;
    out->candidates = (int64_t)ctr.candidates;
    out->exact_chunks = (int64_t)ctr.exact_chunks;
    out->fast_chunks = (int64_t)ctr.fast_chunks;
    out->culled_parts = (int64_t)ctr.culled_parts;
    out->n = d.n;
    out->overflow = ctr.overflow_flag;
    out->events_dropped = ctr.ev_dropped;
    out->sm_count = c->sm_count;
    out->force_grid = c->sp.force_grid;
    out->force_regs = c->force_regs;
    out->force_threads = c->force_threads;
    out->force_variant = c->variant;
    out->pair_halving = d.sym != 0 ? 1 : 0;
    out->sym_regs = c->sp.sym ? c->sym_regs : 0;
    out->row_lo = d.row_lo;
    out->row_hi = d.row_hi;
    out->kernel_launches = c->launches;
    out->force_partials = d.sym ? 2 : 1;       // one 16-byte fixed-point pair, or one float2 slab (two at a CTA boundary)
    return NB_OK;
}

int nb_events(nb_ctx *c, nb_event *buf, int capacity, int *count)
{
    if (!c || !count || capacity < 0 || (!buf && capacity > 0)) return NB_ERR_INVALID;
    *count = 0;
    if (c->st.ev_cap <= 0) {
        set_err(c, "nb_events: the context was created with event_capacity = 0");
        return NB_ERR_INVALID;
    }
    Counters ctr;
    int rc = fetch_state(c, nullptr, &ctr);
    if (rc != NB_OK) return rc;
    const unsigned have = std::min<unsigned>(ctr.ev_count, (unsigned)c->st.ev_cap);
    std::vector<EventRec> rec(have);
    if (have > 0)
        NB_CUDA(c, cudaMemcpy(rec.data(), c->st.ev, sizeof(EventRec) * have, cudaMemcpyDeviceToHost));
    // reset the log
    unsigned zero = 0;
    NB_CUDA(c, cudaMemcpy(&c->st.ctr->ev_count, &zero, sizeof(unsigned), cudaMemcpyHostToDevice));
    std::sort(rec.begin(), rec.end(), [](const EventRec &a, const EventRec &b) {
        if (a.step != b.step) return a.step < b.step;
        if (a.i != b.i) return a.i < b.i;
        return (a.key_kind >> 1) < (b.key_kind >> 1);
    });
    *count = (int)have;
    if ((int)have > capacity) {
        set_err(c, "nb_events: %u records, buffer holds %d", have, capacity);
        return NB_ERR_CAPACITY;
    }
    for (unsigned k = 0; k < have; ++k) {
        buf[k].step = rec[k].step;
        buf[k].i = rec[k].i;
        buf[k].j = rec[k].j;
        buf[k].kind = (int32_t)(rec[k].key_kind & 1u);
    }
    if (ctr.ev_dropped) {
        set_err(c, "nb_events: the event log overflowed (capacity %d); records were dropped", c->st.ev_cap);
        return NB_ERR_EVENT_OVERFLOW;
    }
    return NB_OK;
}

int nb_comm_unique_id(void *id_out)
{
    if (!id_out) return NB_ERR_INVALID;
    static_assert(sizeof(ncclUniqueId) == NB_UNIQUE_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId id;
    if (!nccl_api()) {
        set_err(nullptr, "libnccl.so.2 could not be loaded");
        return NB_ERR_COMM;
    }
    if (nccl_api()->GetUniqueId(&id) != ncclSuccess) {
        set_err(nullptr, "ncclGetUniqueId failed");
        return NB_ERR_COMM;
    }
    memcpy(id_out, &id, sizeof(id));
    return NB_OK;
}

int nb_comm_init(nb_ctx *c, const void *id_bytes)
{
    if (!c || !id_bytes) return NB_ERR_INVALID;
    if (c->sp.world <= 1) return NB_OK;
    if (c->comm_ready) return NB_OK;
    if (!nccl_api()) {
        set_err(c, "libnccl.so.2 could not be loaded");
        return NB_ERR_COMM;
    }
    NB_CUDA(c, cudaSetDevice(c->device));
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof(id));
    NB_NCCL(c, nccl_api()->CommInitRank(&c->comm, c->sp.world, id, c->sp.rank));
    c->comm_ready = true;
    return NB_OK;
}

int nb_render(nb_ctx *c, uint8_t *image, int w, int h) { return nb_render_grid(c, image, w, h, 0x7fffffff); }

int nb_render_grid(nb_ctx *c, uint8_t *image, int w, int h, int grid_threads)
{
    if (!c || !image || w <= 0 || h <= 0 || grid_threads < 0) return NB_ERR_INVALID;
    StepDesc d;
    int rc = fetch_state(c, &d, nullptr);
    if (rc != NB_OK) return rc;
    const size_t bytes = (size_t)w * h;
    if (bytes > c->dev_img_bytes) {
        if (c->dev_img) cudaFree(c->dev_img);
        c->dev_img = nullptr;
        c->dev_img_bytes = 0;
        NB_CUDA(c, cudaMalloc((void **)&c->dev_img, bytes));
        c->dev_img_bytes = bytes;
    }
    LaunchScope scope(c);
    NB_CUDA(c, cudaMemsetAsync(c->dev_img, 254, bytes, c->stream));      // src/nbody.cu:534
    NB_CUDA(c, launch_render(c->st, std::min(d.n, grid_threads), c->dev_img, w, h, c->sp.field_w, c->sp.field_h, c->stream));
    NB_CUDA(c, cudaMemcpyAsync(image, c->dev_img, bytes, cudaMemcpyDeviceToHost, c->stream));
    NB_CUDA(c, cudaStreamSynchronize(c->stream));
    return NB_OK;
}

int nb_plan_host(const nb_params *params, int n, int force_grid, nb_plan *out)
{
    if (!params || !out || n < 0) return NB_ERR_INVALID;
    StepParams sp = {};
    sp.coverage = params->coverage;
    sp.world = params->world > 1 ? params->world : 1;
    sp.rank = params->world > 1 ? params->rank : 0;
    sp.force_grid = force_grid;
    sp.lg_parts_override = -1;
    // the same rules as nb_create (without the device-dependent parts: occupancy, memory budget)
    sp.merge = (params->flags & NB_FLAG_MERGE_CONSERVING) ? 1 : 0;
    sp.sort_min_n = 0;
    if (params->coverage == NB_COVERAGE_FULL && !(params->flags & NB_FLAG_NO_SORT)) {
        const int min_n = params->sort_min_n > 0 ? params->sort_min_n : NB_SORT_MIN_N_DEFAULT;
        if (params->n_max >= min_n) sp.sort_min_n = min_n;
    }
    sp.sym = (params->coverage == NB_COVERAGE_FULL && !(params->flags & NB_FLAG_ONE_SIDED) &&
              ((params->flags & NB_FLAG_PAIR_HALVING) || kPairHalvingDefault) && (sp.sort_min_n > 0 || sp.world == 1)) ? 1 : 0;
    sp.sym_rows = (params->flags & NB_FLAG_SYM_ROWS8) ? 8 : 4;
    sp.sym_grid = (sp.sym_rows == 8 ? 2 : 3) * 148; // the queue granularity rule (sym_lgu) is quoted for a B200
    sp.sym_min_n = kSymMinNDefault;
    sp.sym_small = 2;
    sp.symw_max_n = kSymWarpMaxN;
    sp.symw_grid = 6 * 148;
    sp.field_w = sp.field_h = 1;
    {
        int variant = (params->flags >> NB_FLAG_VARIANT_SHIFT) & 0xf;
        if (params->flags & NB_FLAG_SCALAR_FORCE) variant = 4;
        sp.iblock = variant == 5 ? 1024 : kIBlock;
    }
    StepDesc d;
    plan_host(&d, &sp, n);
    memset(out, 0, sizeof(*out));
    out->n = d.n;
    out->blocks = d.blocks;
    out->limit_last = d.limit_last;
    out->limit_first = d.limit_first;
    out->n_active = d.n_active;
    out->window_len = d.window_len;
    out->row_lo = d.row_lo;
    out->row_hi = d.row_hi;
    out->row_act_hi = d.row_act_hi;
    out->rows_per_rank = d.rows_per_rank;
    out->n_iblocks = d.n_iblocks;
    out->n_jtiles = d.n_jtiles;
    out->units = d.units;
    out->sorted = d.sorted;
    out->two_sided = d.sym != 0 ? 1 : 0;
    out->sym_S = d.sym_S;
    out->sym_Q = d.sym_Q;
    out->sym_blocks = d.sym_blocks;
    out->sym_lgu = d.sym_lgu;
    return NB_OK;
}

int nb_plan_warp_items(int n, int run, int *ids)
{
    if (n <= 0 || run <= 0 || !ids) return NB_ERR_INVALID;
    *ids = symw_geom(n, run).ids;
    return NB_OK;
}

int nb_plan_warp_item(int n, int run, int id, int *group, int *chunk_lo, int *chunk_hi)
{
    if (n <= 0 || run <= 0 || !group || !chunk_lo || !chunk_hi) return NB_ERR_INVALID;
    const WGeom w = symw_geom(n, run);
    if (id < 0 || id >= w.ids) return NB_ERR_INVALID;
    if (!symw_decode(w, id, *group, *chunk_lo, *chunk_hi)) {
        *group = -1;                               // a void id: no work
        *chunk_lo = *chunk_hi = 0;
    }
    return NB_OK;
}

int nb_plan_force_scale(int n, float m_max, float r_min, int field, int *log2_scale)
{
    if (!log2_scale) return NB_ERR_INVALID;
    float fscale = 0.f;
    double finv = 0.0;
    if (!sym_scale(n, m_max, r_min, field, &fscale, &finv)) return NB_ERR_INVALID;
    int e = 0;
    frexpf(fscale, &e);
    *log2_scale = e - 1;                           // fscale is a power of two: 0.5 * 2^e
    return NB_OK;
}

int nb_plan_block(int Q, int b, int *R, int *C)
{
    if (!R || !C || Q <= 0 || b < 0 || b >= Q * (Q + 1) / 2) return NB_ERR_INVALID;
    sym_block_host(b, Q, R, C);
    return NB_OK;
}

int nb_plan_block_index(int Q, int X, int Y)
{
    if (Q <= 0 || X < 0 || Y < 0 || X >= Q || Y >= Q) return NB_ERR_INVALID;
    return sym_block_index_host(X, Y, Q);
}

}  // extern "C"
