/*
 * nbody_b200.h -- C ABI of the B200-native ppa-nbody-collisions time step.
 *
 * The reference (Aidan900/ppa-nbody-collisions) has no plugin/FFI layer: its
 * only boundary is the body of main()'s loop in src/nbody.cu:460-545, the two
 * kernel signatures it launches (ComputeForces src/nbody.cu:139-140,
 * MoveBodies src/nbody.cu:277-278) and the contiguous BodiesData block they
 * share (src/nbody.cu:47-124).  This header is that boundary turned into a
 * C-callable library: plain pointers and sizes, no C++ or torch types, every
 * entry point returns 0 or a negative NB_ERR_* code and never throws.
 *
 * One context owns one GPU (one process per GPU when sharded).  The context
 * owns all device memory; the caller owns every host buffer.  A context is not
 * re-entrant: one host thread at a time.
 *
 * There is NO CPU fallback: nb_create() fails with NB_ERR_CUDA when no sm_100
 * device is usable.
 */
#ifndef NBODY_B200_H
#define NBODY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NB_VERSION 103   /* 103: nb_render_grid, nb_plan_warp_item(s), nb_plan_force_scale; 102: nb_stats.kernel_launches, nb_probe_fp32; 101: nb_stats / nb_plan grew (two-sided kernel) */

/* error codes */
#define NB_OK 0
#define NB_ERR_INVALID (-1)           /* bad argument                                   */
#define NB_ERR_CUDA (-2)              /* CUDA runtime / driver error (see nb_last_error) */
#define NB_ERR_CAPACITY (-3)          /* n exceeds n_max, or host buffer too small      */
#define NB_ERR_CANDIDATE_OVERFLOW (-4)/* collision candidate list was too small         */
#define NB_ERR_COMM (-5)              /* NCCL error or communicator not initialised     */
#define NB_ERR_IO (-6)                /* file could not be read / written               */
#define NB_ERR_EVENT_OVERFLOW (-7)    /* event log was too small (events were dropped)  */

/* coverage: which ordered pairs (i, j) a step evaluates */
#define NB_COVERAGE_REFERENCE 0       /* exactly the pairs src/nbody.cu:182-207 visits (floor(n/128)
                                         blocks, last j-tile truncated to n % 129, frozen tail)      */
#define NB_COVERAGE_FULL 1            /* true all-pairs: every i against every j != i               */

/* event kinds (the reference emits none; they are derived from src/nbody.cu:215-226) */
#define NB_EV_ABSORB 0                /* thread i saw a hit with m_i >= m_j: i gains m_j, growth*r_j */
#define NB_EV_KILLED 1                /* thread i saw a hit with m_i <  m_j: i's mass becomes 0      */

#define NB_GRAV_CONSTANT 6.67408e-11f /* GRAV_CONSTANT, src/nbody.cu:37 */

typedef struct nb_ctx nb_ctx;

/* Replaces the scalar arguments main() passes to both kernels (src/nbody.cu:481-483). */
typedef struct nb_params {
    int   n_max;              /* capacity in bodies (>= every n ever uploaded)                      */
    float dt;                 /* ConfigData::timestep          include/nbodyConfig.h:8              */
    float growth;             /* ConfigData::growthRate        include/nbodyConfig.h:13             */
    int   field_w;            /* ConfigData::fieldWidth        include/nbodyConfig.h:16             */
    int   field_h;            /* ConfigData::fieldHeight       include/nbodyConfig.h:17             */
    float grav;               /* 0 -> NB_GRAV_CONSTANT                                              */
    int   coverage;           /* NB_COVERAGE_*                                                      */
    int   device;             /* CUDA device ordinal                                                */
    int   candidate_capacity; /* entries of the per-step collision candidate list (ordered hit
                                 pairs), 0 -> max(4 * n_max, 64Ki).  With the two-sided kernel on
                                 several GPUs it is also the number of pairs ONE rank can contribute
                                 to the per-step exchange, 0 -> max(128Ki, n_max / 4); exceeding
                                 either is reported as NB_ERR_CANDIDATE_OVERFLOW                     */
    int   event_capacity;     /* 0 = no event log; else records kept between nb_events() calls      */
    int   rank;               /* this context's shard, 0 <= rank < world                            */
    int   world;              /* number of shards (GPUs); <= 1 means single GPU                     */
    int   flags;              /* NB_FLAG_*                                                          */
    int   sort_min_n;         /* smallest n that uses the sorted order, 0 -> NB_SORT_MIN_N_DEFAULT   */
    float softening;          /* opt-in Plummer softening length eps: forces use |r|^2 + eps^2 (the collision
                                 test stays unsoftened).  0 = off = the reference's arithmetic                */
} nb_params;

#define NB_FLAG_NO_GRAPH 1    /* launch kernels one by one instead of replaying a CUDA graph        */
#define NB_FLAG_SCALAR_FORCE 2/* use the scalar-FP32 force kernel instead of the packed f32x2 one   */
#define NB_FLAG_NO_SORT 4     /* never use the cell-sorted shadow order.  By default, with NB_COVERAGE_FULL and
                                 n >= sort_min_n, every step runs on a Morton-cell-sorted copy of the bodies (sorted every 32
                                 steps, carried over the compaction in between) so that the collision pre-test is skipped
                                 where bounding boxes are apart; results keep the bodies' own order, events and survivors
                                 are unchanged                                                                       */
#define NB_FLAG_MERGE_CONSERVING 16 /* opt-in physics beyond parity (every force path, any number of GPUs): instead of the reference's "heavier
                                 absorbs, nothing is conserved" rule (src/nbody.cu:215-226), every body points at the
                                 lowest index among itself and its hit partners; following the pointers ends at a root,
                                 which takes mass, momentum (at the post-force velocities) and growth * radius of
                                 everything that ends at it, in ascending index order, keeps its position and moves on
                                 with P / M; the others are removed.  Events: kind = ABSORB when i < j, else KILLED   */
#define NB_FLAG_ONE_SIDED 32   /* never use the two-sided force kernels: the ONLY switch of that path.  By default every step
                                 with NB_COVERAGE_FULL and n >= 12288 (a warp per work item below 196608 bodies, a CTA
                                 per tile pair from there on) evaluates every unordered pair once
                                 and applies the force to both bodies (Newton's third law): 12 instead of 2 x 9 packed
                                 operations per pair of interactions.  The collision predicate is symmetric bit for bit
                                 (src/nbody.cu:126-134), so events, survivors, masses and radii are unchanged; the force
                                 sums are exact 64-bit fixed-point integers, hence deterministic and the same on any
                                 number of GPUs                                                                      */
#define NB_FLAG_PAIR_HALVING 64 /* accepted and ignored: the two-sided kernels are the default (see NB_FLAG_ONE_SIDED)    */
#define NB_FLAG_SYM_ROWS8 128  /* tuning variant of the CTA-level two-sided kernel: 8 rows per lane at 2 CTAs per SM (half the
                                 shuffles per evaluation).  Same results as the default (4 rows, 3 CTAs); measured 1 %
                                 slower at n = 1 048 576 and 20 % slower at n = 70 000 (profiles/r02_rows4_vs_rows8.jsonl):
                                 kept for tuning, never selected by default                                          */
#define NB_SORT_MIN_N_DEFAULT 12288
#define NB_FLAG_VARIANT_SHIFT 8   /* bits 8..11: force-kernel variant (occupancy / rows-per-lane trade-off,
                                     see nbody_kernels.cu); 0 = default                                */
#define NB_FLAG_VARIANT(v) ((v) << NB_FLAG_VARIANT_SHIFT)

typedef struct nb_event {
    int32_t step;             /* step index (0-based, counted since nb_upload)                      */
    int32_t i;                /* pre-step index of the body whose row saw the hit                   */
    int32_t j;                /* pre-step index of the other body                                   */
    int32_t kind;             /* NB_EV_*                                                            */
} nb_event;

typedef struct nb_stats {
    int64_t steps;            /* steps executed since nb_upload                                     */
    int64_t pairs;            /* ordered pairs evaluated by this context's shard since nb_upload    */
    int64_t candidates;       /* collision pair-events since nb_upload (this shard)                 */
    int64_t exact_chunks;     /* 32-body j sub-chunks that took the exact path (this shard)         */
    int64_t fast_chunks;      /* sub-chunks that took the packed fast path (this shard)             */
    int32_t n;                /* live bodies now                                                    */
    int32_t overflow;         /* 1 if the candidate list ever overflowed                            */
    int32_t events_dropped;   /* 1 if the event log overflowed                                      */
    int32_t sm_count;         /* multiprocessors of the device                                      */
    int32_t force_grid;       /* CTAs of the persistent force kernel                                */
    int32_t force_regs;       /* registers per thread of the force kernel                           */
    int32_t row_lo, row_hi;   /* this shard's rows [row_lo, row_hi) of the next step                */
    int32_t force_threads;    /* threads per CTA of the force kernel                                */
    int32_t force_variant;    /* force-kernel variant in use                                        */
    int64_t culled_parts;     /* j parts that ran without the collision pre-test (sorted stream)    */
    int32_t pair_halving;     /* 1: the next step will run the two-sided force kernel               */
    int32_t sym_regs;         /* registers per thread of the two-sided force kernel (0: not in use) */
    int64_t kernel_launches;  /* kernels of this library executed for this context since nb_create: direct launches
                                 plus the kernel nodes of every graph replay (counted at the launch sites, NCCL's
                                 own kernels not included)                                                         */
    int32_t force_partials;   /* 8-byte partial force sums per body the finish kernel adds up in the next step     */
    int32_t reserved;
} nb_stats;

/* ---- lifecycle ---------------------------------------------------------- */
int  nb_create(nb_ctx **ctx, const nb_params *params);
void nb_destroy(nb_ctx *ctx);
const char *nb_last_error(const nb_ctx *ctx);      /* ctx may be NULL: last nb_create error */
int  nb_version(void);

/* ---- body store --------------------------------------------------------- */
/*
 * bodies: the reference's BodiesData block (src/nbody.cu:66-77), 24*n bytes:
 *   Vec2f Positions[n]; Vec2f Velocities[n]; float Masses[n]; float Radii[n];
 * nb_upload replaces BodiesData::uploadToDevice (src/nbody.cu:88-96) and
 * resets the step counter, statistics and event log.  nb_download replaces the
 * D2H copy + host compaction of src/nbody.cu:486-510: it returns the already
 * compacted survivors in the same layout (re-laid-out for the current n).
 * Every rank of a sharded run passes the same full block; after nb_comm_init each rank copies only its 1 / world of
 * every array over PCIe and the rest arrives from the peers over NVLink.
 */
int nb_upload(nb_ctx *ctx, const void *bodies, int n);
int nb_download(nb_ctx *ctx, void *bodies, int capacity_n, int *n);
int nb_num_bodies(nb_ctx *ctx, int *n);            /* synchronises */

/* ---- time step ---------------------------------------------------------- */
/*
 * n_steps iterations of ComputeForces + MoveBodies + compaction
 * (src/nbody.cu:481-510).  Asynchronous on the context's stream when world<=1;
 * errors (candidate overflow, CUDA faults) surface at the next synchronising
 * call.  nb_step_timed also returns device times measured with CUDA events on
 * the context's stream: whole region and the force kernel alone (sum).
 */
int nb_step(nb_ctx *ctx, int n_steps);
int nb_step_timed(nb_ctx *ctx, int n_steps, float *ms_total, float *ms_force);
/* Per-kernel device times (ms, summed over n_steps, CUDA events on the context's stream, no graph):
 * ms[0] force, ms[1] finish (bookkeeping + integrate), ms[2] allgather (0 on one GPU), ms[3] compaction,
 * ms[4] rebuild of the cell-sorted order (0 when it is not in use). */
int nb_step_profile(nb_ctx *ctx, int n_steps, float ms[5]);
int nb_sync(nb_ctx *ctx);
int nb_get_stats(nb_ctx *ctx, nb_stats *out);      /* synchronises */

/*
 * Derived collision event list: all records logged since the
 * last nb_events/nb_upload, sorted by (step, i, visit order).  Needs
 * event_capacity > 0.  Sharded runs return this rank's rows only.
 */
int nb_events(nb_ctx *ctx, nb_event *buf, int capacity, int *count);

/* ---- multi-GPU (one process per GPU) ------------------------------------ */
/*
 * Rank 0 calls nb_comm_unique_id (128 bytes, an ncclUniqueId), the host
 * distributes it (threads of one process: shared memory, as `nbody --gpus N` does;
 * processes: e.g. a torch.distributed / MPI broadcast), then every rank calls
 * nb_comm_init.  After that nb_step exchanges the shard results over NVLink:
 * one ncclAllGather of the post-step rows per step, plus -- two-sided steps --
 * one integer ncclAllReduce of the fixed-point force sums and one
 * ncclAllGather of the candidate pairs.
 */
#define NB_UNIQUE_ID_BYTES 128
int nb_comm_unique_id(void *id_out);
int nb_comm_init(nb_ctx *ctx, const void *id);

/*
 * The step plan (coverage descriptor + shard bounds) for n live bodies, computed on the host by
 * the same code the device runs at the end of every step.  No GPU needed: used by the sharding
 * tests and by hosts that want to know which rows a rank owns.
 */
typedef struct nb_plan {
    int32_t n, blocks, limit_last, limit_first, n_active, window_len;
    int32_t row_lo, row_hi, row_act_hi, rows_per_rank, n_iblocks, n_jtiles;
    int64_t units;
    int32_t sorted;           /* 1: the step runs on the cell-sorted order                                   */
    int32_t two_sided;        /* 1: the step runs the two-sided force kernel: the triangle of tile pairs in    */
    int32_t sym_S, sym_Q;     /*    blocks of sym_S x sym_S tile pairs, sym_Q super-tiles per side             */
    int32_t sym_blocks;       /*    sym_Q (sym_Q + 1) / 2 blocks; rank r of W takes work items r, r + W, ...   */
    int32_t sym_lgu;          /*    a block is 1 << sym_lgu work items (> 0 only when sym_S == 1: few tile pairs) */
} nb_plan;
int nb_plan_host(const nb_params *params, int n, int force_grid, nb_plan *out);
/* Block b (queue order) of the two-sided kernel's pair triangle with Q super-tiles per side -> (R, C), R <= C;
   nb_plan_block_index is its inverse (any order of the two super-tiles).  Host only. */
int nb_plan_block(int Q, int b, int *R, int *C);
int nb_plan_block_index(int Q, int X, int Y);
/* The warp-level two-sided kernel's work items for n bodies with `run` chunks per item: how many ids there are, and which
   128-row group meets which 64-body chunks [chunk_lo, chunk_hi) under id (group = -1: a void id).  Rank r of W takes ids
   r, r + W, ...  Host only. */
int nb_plan_warp_items(int n, int run, int *ids);
int nb_plan_warp_item(int n, int run, int id, int *group, int *chunk_lo, int *chunk_hi);
/* The fixed-point scale 2^k of the two-sided kernels' force sums for n bodies with masses <= m_max and radii >= r_min in a
   field of half-width `field` (n * m_max / (2 r_min)^2 * 2^k < 2^62); NB_ERR_INVALID when no scale leaves 30 bits below a
   typical force -- the step plan then falls back to the one-sided kernel.  Host only. */
int nb_plan_force_scale(int n, float m_max, float r_min, int field, int *log2_scale);

/* ---- measurement ---------------------------------------------------------- */
/*
 * FP32 peak of `device` measured now: a stream of independent packed fma.rn.f32x2, 32 warps per SM, no memory
 * traffic (about 5 ms).  The denominator of the force kernels' roofline in bench.py.
 */
int nb_probe_fp32(int device, double *tflops);

/* ---- render (src/nbody.cu:294-348, 350-371) ------------------------------ */
/* Rasterise the current bodies into a w*h 8-bit image (background 254, body 0). */
int nb_render(nb_ctx *ctx, uint8_t *image, int w, int h);
/*
 * The same with the reference's launch grid: its loop draws with the grid of the step it has just done,
 * floor(n_before / 128) blocks of 128 threads (src/nbody.cu:473,535), so only bodies below grid_threads =
 * 128 * max(1, n_before / 128) are drawn -- up to 127 tail bodies are missing from its pictures.  The drop-in driver
 * uses this; threads beyond the live bodies (where the reference reads outside its body store) draw nothing here.
 */
int nb_render_grid(nb_ctx *ctx, uint8_t *image, int w, int h, int grid_threads);
/* Write a binary P5 file exactly as saveImageToDisk does (header "P5\n<w> <h>\n255\n"). */
int nb_write_pgm(const char *path, const uint8_t *image, int w, int h);

/* ---- driver surface: config, RNG, initial conditions (host only, no GPU) -- */
/* ConfigData, include/nbodyConfig.h:4-19 (imagePath as a fixed buffer). */
typedef struct nb_config {
    int   particleCount;
    int   totalIterations;
    int   save_Image_Every_Xth_Iteration;
    float timestep;
    float minRandBodyMass;
    float maxRandBodyMass;
    float minRadius;
    float maxRadius;
    float growthRate;
    int   imgWidth;
    int   imgHeight;
    int   fieldWidth;
    int   fieldHeight;
    char  imagePath[1024];
} nb_config;

/*
 * parseConfigFile (include/nbodyConfig.h:22-227): same keys, same echo lines
 * (written to echo_fd, e.g. 1 for stdout; -1 for none), same "Invalid
 * variable" handling.  Where the reference calls exit(1) this returns
 * NB_ERR_IO / NB_ERR_INVALID after writing the same message.  Fields whose key
 * is missing are left untouched (the reference leaves them uninitialised).
 */
int nb_config_parse(const char *path, nb_config *cfg, int echo_fd);

/* jbutil::randgen (include/jbutil.h:514-562) */
typedef struct nb_rng { uint64_t u, v, w; } nb_rng;
void     nb_rng_seed(nb_rng *g, uint64_t seed);
uint64_t nb_rng_ival64(nb_rng *g);
double   nb_rng_fval(nb_rng *g);
double   nb_rng_fval_range(nb_rng *g, double a, double b);

/* scenarios (NB_SCENARIO_SQUARE is the reference's src/nbody.cu:401-416) */
#define NB_SCENARIO_SQUARE 0      /* uniform square +-field, v = 0 (reference)                       */
#define NB_SCENARIO_DISC 1        /* uniform disc of radius `extent`, v = 0                          */
#define NB_SCENARIO_TWO_GALAXY 2  /* two spinning discs on an encounter course                       */
typedef struct nb_scenario {
    int      kind;            /* NB_SCENARIO_*                                                      */
    int      n;
    uint64_t seed;            /* 1024 in the reference (src/nbody.cu:403)                           */
    int      field_w, field_h;
    float    min_mass, max_mass, min_radius, max_radius;
    double   extent;          /* disc radius (DISC, TWO_GALAXY); ignored for SQUARE                 */
} nb_scenario;
/* Fills a BodiesData block (6*n floats). */
int nb_generate(const nb_scenario *sc, void *bodies);

#ifdef __cplusplus
}
#endif
#endif /* NBODY_B200_H */
