/*
 * nbody_oracle.c -- CPU restatement of the ppa-nbody-collisions time step.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (the
 * ppa-nbody-collisions_b200 package, its C-ABI library or the `nbody` driver)
 * may import, link or execute this file.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs use it, as the checker
 * and as the CPU baseline.
 *
 * Parity pin: the reference ships no tests and no golden vectors (SURVEY.md
 * section 4).  This restatement is pinned two ways:
 *   (1) tests/golden/host_*.json  -- RNG stream, initial bodies and config echo
 *       produced by compiling the reference's own headers here
 *       (tools/make_golden_host.py);
 *   (2) tests/golden/gpuref_*.json -- per-step state hashes of the UNMODIFIED
 *       reference kernels (oracle/_ref, built from /root/reference/src/nbody.cu
 *       by oracle/Makefile) run on a B200 (tools/make_golden_gpuref.py).
 * The restatement follows the reference's PTX arithmetic contract (all .rn):
 * fused multiply-add exactly where nvcc contracted, separate roundings
 * elsewhere.  Build with -ffp-contract=off (see oracle/Makefile).
 *
 * Every function cites the reference lines it follows
 * (paths relative to /root/reference).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_THREADS_PER_BLOCK 128          /* src/nbody.cu:36 */
#define ORC_GRAV_CONSTANT 6.67408e-11f     /* src/nbody.cu:37 */

enum { ORC_COVERAGE_REFERENCE = 0, ORC_COVERAGE_FULL = 1 };
enum { ORC_EV_ABSORB = 0, ORC_EV_KILLED = 1 };

typedef struct {
    float dt;        /* ConfigData::timestep        include/nbodyConfig.h:8  */
    float growth;    /* ConfigData::growthRate      include/nbodyConfig.h:13 */
    int field_w;     /* ConfigData::fieldWidth      include/nbodyConfig.h:16 */
    int field_h;     /* ConfigData::fieldHeight     include/nbodyConfig.h:17 */
    int coverage;    /* ORC_COVERAGE_*                                       */
    int threads;     /* OpenMP threads, <=0 -> all                           */
    float softening; /* opt-in Plummer softening length (NOT reference behaviour; 0 = off, the parity mode) */
    int merge;       /* 0: the reference's rule (src/nbody.cu:215-226).  1: opt-in conserving lowest-index merge */
} orc_params;

typedef struct {
    int i;           /* pre-step index of the body whose thread saw the hit */
    int j;           /* pre-step index of the other body                    */
    int kind;        /* ORC_EV_ABSORB: i absorbs j; ORC_EV_KILLED: i dies   */
} orc_event;

typedef struct {
    int n;           /* live bodies at the start of the step                */
    int blocks;      /* B  = numBlocks                 src/nbody.cu:473     */
    int limit_last;  /* innerLoopLimit of tile B-1     src/nbody.cu:194     */
    int n_active;    /* bodies that own a thread       src/nbody.cu:142-143 */
} orc_cov;

/* ---------------------------------------------------------------------- */
/* RNG: jbutil::randgen, include/jbutil.h:514-562 (Numerical Recipes Ran). */
/* ---------------------------------------------------------------------- */
typedef struct { uint64_t u, v, w; } orc_rng;

static inline void rng_advance(orc_rng *g)            /* jbutil.h:537-544 */
{
    g->u = g->u * 2862933555777941757ULL + 7046029254386353087ULL;
    g->v ^= g->v >> 17;
    g->v ^= g->v << 31;
    g->v ^= g->v >> 8;
    g->w = 4294957665ULL * (g->w & 0xffffffffULL) + (g->w >> 32);
}

uint64_t orc_rng_ival64(orc_rng *g)                   /* jbutil.h:546-553 */
{
    rng_advance(g);
    uint64_t x = g->u ^ (g->u << 21);
    x ^= x >> 35;
    x ^= x << 4;
    return (x + g->v) ^ g->w;
}

void orc_rng_seed(orc_rng *g, uint64_t s)             /* jbutil.h:525-535 */
{
    g->v = 4101842887655102017ULL;
    g->w = 1;
    g->u = s ^ g->v;
    orc_rng_ival64(g);
    g->v = g->u;
    orc_rng_ival64(g);
    g->w = g->v;
    orc_rng_ival64(g);
}

double orc_rng_fval(orc_rng *g)                       /* jbutil.h:554-557 */
{
    return 5.42101086242752217E-20 * (double)orc_rng_ival64(g);
}

double orc_rng_fval_range(orc_rng *g, double a, double b) /* jbutil.h:558-561 */
{
    return orc_rng_fval(g) * (b - a) + a;
}

/*
 * Initial conditions, src/nbody.cu:401-416: seed, then per body four draws in
 * the order x, y, m, r; float = double - int; velocities zero.  `block` is the
 * reference's BodiesData layout (src/nbody.cu:66-77):
 *   [x0 y0 x1 y1 ...][vx0 vy0 ...][m0 m1 ...][r0 r1 ...]   (24*n bytes)
 */
void orc_init_square(float *block, int n, uint64_t seed, int field_w, int field_h,
                     float min_mass, float max_mass, float min_radius, float max_radius)
{
    orc_rng g;
    orc_rng_seed(&g, seed);
    float *pos = block, *vel = block + 2 * (size_t)n;
    float *mass = block + 4 * (size_t)n, *rad = block + 5 * (size_t)n;
    int dw = field_w << 1, dh = field_h << 1;          /* src/nbody.cu:388,390 */
    for (int b = 0; b < n; ++b) {
        float x = (float)(orc_rng_fval_range(&g, 0, dw) - field_w);
        float y = (float)(orc_rng_fval_range(&g, 0, dh) - field_h);
        float m = (float)orc_rng_fval_range(&g, min_mass, max_mass);
        float r = (float)orc_rng_fval_range(&g, min_radius, max_radius);
        pos[2 * b] = x; pos[2 * b + 1] = y;
        vel[2 * b] = 0.f; vel[2 * b + 1] = 0.f;
        mass[b] = m; rad[b] = r;
    }
}

/*
 * Synthetic scenarios of BASELINE.json configs[1..4] (the reference itself only generates the
 * square above, src/nbody.cu:401-416).  Restated here so that the checkers and bench.py's reference arm can
 * build their inputs WITHOUT loading the product library; tests/test_oracle_golden.py checks bit equality
 * with the product's nb_generate.  Per body four draws in the reference's order (u1, u2, m, r), positions
 * from (radius^2 fraction, angle) in double, stored as float; solid-body spin omega about the disc centre.
 */
void orc_init_disc_part(float *block, int n, int first, int count, uint64_t seed, double cx, double cy, double R,
                        double bulk_vx, double bulk_vy, double omega,
                        float min_mass, float max_mass, float min_radius, float max_radius)
{
    float *pos = block, *vel = block + 2 * (size_t)n;
    float *mass = block + 4 * (size_t)n, *rad = block + 5 * (size_t)n;
    const double two_pi = 6.283185307179586476925286766559;
    orc_rng g;
    orc_rng_seed(&g, seed);
    for (int k = 0; k < count; ++k) {
        const int b = first + k;
        const double u1 = orc_rng_fval(&g), u2 = orc_rng_fval(&g);
        const double rr = R * sqrt(u1), th = two_pi * u2;
        const double x = rr * cos(th), y = rr * sin(th);
        pos[2 * b] = (float)(cx + x);
        pos[2 * b + 1] = (float)(cy + y);
        vel[2 * b] = (float)(bulk_vx - omega * y);
        vel[2 * b + 1] = (float)(bulk_vy + omega * x);
        mass[b] = (float)orc_rng_fval_range(&g, min_mass, max_mass);
        rad[b] = (float)orc_rng_fval_range(&g, min_radius, max_radius);
    }
}

/* ---------------------------------------------------------------------- */
/* Coverage: which threads exist and how long the last j-tile is.          */
/* ---------------------------------------------------------------------- */
void orc_coverage(int n, int mode, orc_cov *c)
{
    const int T = ORC_THREADS_PER_BLOCK;
    c->n = n;
    if (mode == ORC_COVERAGE_REFERENCE) {
        c->blocks = n < T ? 1 : n / T;                 /* src/nbody.cu:473 (floor) */
        c->limit_last = n % (T + 1);                   /* src/nbody.cu:194         */
        int threads = c->blocks * T;
        c->n_active = n < threads ? n : threads;       /* src/nbody.cu:142-143     */
    } else {
        /* True all-pairs: ceil(n/128) tiles, the last holding the remainder;
         * same slot/visit-order formulas, every body owns a thread. */
        c->blocks = (n + T - 1) / T;
        c->limit_last = n - T * (c->blocks - 1);
        if (c->blocks == 0) c->limit_last = 0;
        c->n_active = n;
    }
}

typedef struct {
    float vx, vy;    /* velocities[i] after ComputeForces   src/nbody.cu:264   */
    float px, py;    /* positions[i] after MoveBodies        src/nbody.cu:288   */
    float m, r;      /* updatedMasses/Radii[i]               src/nbody.cu:245-246 */
    int hits;        /* collision pair-events seen by thread i                  */
    int absorber;    /* merge mode 1: min(i, lowest index among i's hit partners)  */
    int asserts;     /* device asserts that would have fired src/nbody.cu:235, vec2f.h:51 */
    long long visited; /* pairs evaluated by this thread                        */
} orc_row;

typedef struct {
    orc_event *ev; size_t n, cap;
} ev_buf;

static void ev_push(ev_buf *b, int i, int j, int kind)
{
    if (b->n == b->cap) {
        b->cap = b->cap ? b->cap * 2 : 1024;
        b->ev = (orc_event *)realloc(b->ev, b->cap * sizeof(orc_event));
    }
    b->ev[b->n].i = i; b->ev[b->n].j = j; b->ev[b->n].kind = kind;
    b->n++;
}

/*
 * One ComputeForces thread + its MoveBodies update, src/nbody.cu:139-292.
 * Arithmetic contract (nvcc 12.9 PTX of the unmodified file, all .rn):
 *   d2     = fma(dx, dx, dy*dy)                       :129-131
 *   rs2    = (r_i + r_j) * (r_i + r_j)                :133
 *   radius = fma(growth, r_j, radius)                 :219
 *   d      = sqrt.rn(d2); d3 = d * (d * d)            :232,239 / vec2f.h:91-93
 *   inv    = rcp.rn(d3)  (== 1.0f / d3)               vec2f.h:49-53
 *   f     += fma(inv, dir * m_j, f)                   :239
 *   a = f * G; dv = dt * a (separate roundings)       :250-252
 *   v' = (+-)v + dv                                   :256-264
 *   p' = fma(dt, v', p)                               :288
 */
static void eval_row(const float *pos, const float *vel, const float *mass, const float *rad,
                     const orc_cov *cov, const orc_params *par, int i, orc_row *out, ev_buf *evb)
{
    const int T = ORC_THREADS_PER_BLOCK;
    const int n = cov->n, B = cov->blocks;
    const int b = i / T, t = i % T;
    const float xi = pos[2 * i], yi = pos[2 * i + 1];
    const float mi = mass[i], ri = rad[i];
    float umass = mi, uradius = ri;                    /* :174-175 */
    float fx = 0.f, fy = 0.f;                          /* :153     */
    int skip = 1, deleted = 0, hits = 0, asserts = 0;  /* :178,180 */
    long long visited = 0;
    int amin = i;

    for (int k = 0; k < B; ++k) {                      /* :182 */
        const int g = (int)(((long long)i + (long long)T * k) % n);   /* :186 */
        const int limit = (k == B - 1) ? cov->limit_last : T;         /* :194 */
        /* slot s of this tile holds body (128b + s + 128k) % n (thread s's load, :186-189) */
        const int base = (int)(((long long)T * b + (long long)T * k) % n);
        int snext = limit > 0 ? t % limit : 0;         /* s = (t + off) % limit, :207, kept incrementally */
        for (int off = 0; off < limit; ++off) {        /* :196 */
            const int s = snext;
            if (++snext == limit) snext = 0;
            if (skip && g == i) { skip = 0; continue; }   /* :200-204 */
            int j = base + s;                          /* base < n and s < 128 <= n (or s < n) */
            if (j >= n) j -= n;
            ++visited;
            const float xj = pos[2 * j], yj = pos[2 * j + 1], mj = mass[j], rj = rad[j];
            /* areParticlesColliding, :126-134 */
            const float dx = xj - xi, dy = yj - yi;
            const float dy2 = dy * dy;
            const float d2 = fmaf(dx, dx, dy2);
            const float rs = ri + rj;
            const float rs2 = rs * rs;
            const int intersect = d2 <= rs2;
            if (intersect && par->merge) {
                /* conserving lowest-index merge (NOT reference behaviour): only remember the lowest partner */
                if (j < amin) amin = j;
                ++hits;
                if (evb) ev_push(evb, i, j, i < j ? ORC_EV_ABSORB : ORC_EV_KILLED);
                continue;
            }
            if (intersect && (mi >= mj)) {             /* :215-221 */
                umass += mj;
                uradius = fmaf(par->growth, rj, uradius);
                ++hits;
                if (evb) ev_push(evb, i, j, ORC_EV_ABSORB);
                continue;
            } else if (intersect && (mi < mj)) {       /* :222-226 */
                deleted = 1;
                ++hits;
                if (evb) ev_push(evb, i, j, ORC_EV_KILLED);
                continue;
            }
            if (par->softening > 0.f) {
                /* opt-in softened force (no reference counterpart): |r|^2 + eps^2 with the roundings of the
                 * CUDA fast path, then the reference's formula; the collision predicate above is unsoftened */
                const float e2 = par->softening * par->softening;
                const float d2s = fmaf(dx, dx, fmaf(dy, dy, e2));
                const float ds = sqrtf(d2s);
                const float invs = 1.0f / (ds * (ds * ds));
                fx = fmaf(invs, dx * mj, fx);
                fy = fmaf(invs, dy * mj, fy);
                continue;
            }
            const float d = sqrtf(d2);                 /* :232 */
            if (!(d != 0)) ++asserts;                  /* :235 */
            const float tx = dx * mj, ty = dy * mj;    /* :239 */
            const float dd = d * d;
            const float d3 = d * dd;
            if (!(d3 != 0.0)) ++asserts;               /* vec2f.h:51 */
            const float inv = 1.0f / d3;               /* vec2f.h:52 */
            fx = fmaf(inv, tx, fx);
            fy = fmaf(inv, ty, fy);
        }
    }
    out->m = deleted ? 0.f : umass;                    /* :245 */
    out->r = uradius;                                  /* :246 */
    const float ax = fx * ORC_GRAV_CONSTANT, ay = fy * ORC_GRAV_CONSTANT;   /* :250 */
    const float dvx = par->dt * ax, dvy = par->dt * ay;                     /* :252 */
    float vx = vel[2 * i], vy = vel[2 * i + 1];
    const float W = (float)par->field_w, H = (float)par->field_h;
    const float nW = (float)(-par->field_w), nH = (float)(-par->field_h);
    const float tpx = dvx + xi, tpy = dvy + yi;
    if (tpx > W - ri || tpx < ri + nW) vx = -vx;       /* :256-258 */
    if (tpy > H - ri || tpy < ri + nH) vy = -vy;       /* :259-261 */
    out->vx = vx + dvx;                                /* :264 */
    out->vy = vy + dvy;
    out->px = fmaf(par->dt, out->vx, xi);              /* :288 */
    out->py = fmaf(par->dt, out->vy, yi);
    out->hits = hits;
    out->absorber = amin;
    out->asserts = asserts;
    out->visited = visited;
}

static int cmp_event(const void *a, const void *b)
{
    const orc_event *x = (const orc_event *)a, *y = (const orc_event *)b;
    return (x->i > y->i) - (x->i < y->i);
}

/*
 * Post-step values of selected rows (no commit, no compaction).  Each row
 * depends only on pre-step state (src/nbody.cu:210-264), so checking a sample
 * of rows is an exact test of those rows at any N.  out is nrows x 6 floats:
 * vx, vy, px, py, m, r.  Rows >= n_active are returned unchanged (frozen tail).
 */
void orc_rows(const float *block, int n, const orc_params *par, const int *rows, int nrows,
              float *out, int *hits, long long *visited)
{
    orc_cov cov;
    orc_coverage(n, par->coverage, &cov);
    const float *pos = block, *vel = block + 2 * (size_t)n;
    const float *mass = block + 4 * (size_t)n, *rad = block + 5 * (size_t)n;
#ifdef _OPENMP
    int nt = par->threads > 0 ? par->threads : omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 16) num_threads(nt)
#endif
    for (int q = 0; q < nrows; ++q) {
        int i = rows[q];
        orc_row r;
        if (i < cov.n_active) {
            eval_row(pos, vel, mass, rad, &cov, par, i, &r, NULL);
        } else {
            r.vx = vel[2 * i]; r.vy = vel[2 * i + 1];
            r.px = pos[2 * i]; r.py = pos[2 * i + 1];
            r.m = mass[i]; r.r = rad[i]; r.hits = 0; r.visited = 0;
        }
        out[6 * q + 0] = r.vx; out[6 * q + 1] = r.vy;
        out[6 * q + 2] = r.px; out[6 * q + 3] = r.py;
        out[6 * q + 4] = r.m;  out[6 * q + 5] = r.r;
        if (hits) hits[q] = r.hits;
        if (visited) visited[q] = r.visited;
    }
}

/*
 * Float64 yardstick for the force sum (NOT the reference's arithmetic): the same visited pairs and
 * the same float32 collision predicate decide which pairs contribute (src/nbody.cu:210-226), but
 * direction, distance and the sum itself are evaluated in double.  Used by the tests to show that the
 * CUDA path's force error is no larger than the reference's own sequential-float32 summation error
 * (which grows like sqrt(n) * 2^-24).  out is nrows x 2 doubles: dv = dt * G * force per row, i.e. the
 * velocity change of src/nbody.cu:250-252.
 */
void orc_rows_dv_f64(const float *block, int n, const orc_params *par, const int *rows, int nrows, double *out)
{
    orc_cov cov;
    orc_coverage(n, par->coverage, &cov);
    const float *pos = block, *mass = block + 4 * (size_t)n, *rad = block + 5 * (size_t)n;
    const int T = ORC_THREADS_PER_BLOCK;
#ifdef _OPENMP
    int nt = par->threads > 0 ? par->threads : omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 16) num_threads(nt)
#endif
    for (int q = 0; q < nrows; ++q) {
        const int i = rows[q];
        double fx = 0.0, fy = 0.0;
        if (i < cov.n_active) {
            const int b = i / T, t = i % T;
            const float xi = pos[2 * i], yi = pos[2 * i + 1], ri = rad[i];
            int skip = 1;
            for (int k = 0; k < cov.blocks; ++k) {
                const int g = (int)(((long long)i + (long long)T * k) % n);
                const int limit = (k == cov.blocks - 1) ? cov.limit_last : T;
                const int base = (int)(((long long)T * b + (long long)T * k) % n);
                int snext = limit > 0 ? t % limit : 0;
                for (int off = 0; off < limit; ++off) {
                    const int s = snext;
                    if (++snext == limit) snext = 0;
                    if (skip && g == i) { skip = 0; continue; }
                    int j = base + s;
                    if (j >= n) j -= n;
                    const float dxf = pos[2 * j] - xi, dyf = pos[2 * j + 1] - yi;
                    const float d2f = fmaf(dxf, dxf, dyf * dyf);
                    const float rs = ri + rad[j];
                    if (d2f <= rs * rs) continue;                       /* hit: no force from j */
                    const double dx = (double)pos[2 * j] - (double)xi, dy = (double)pos[2 * j + 1] - (double)yi;
                    const double d2 = dx * dx + dy * dy + (double)par->softening * (double)par->softening;
                    const double inv = 1.0 / (d2 * sqrt(d2));
                    fx += dx * (double)mass[j] * inv;
                    fy += dy * (double)mass[j] * inv;
                }
            }
        }
        out[2 * q] = (double)par->dt * (fx * (double)ORC_GRAV_CONSTANT);
        out[2 * q + 1] = (double)par->dt * (fy * (double)ORC_GRAV_CONSTANT);
    }
}

/*
 * One full step in place: ComputeForces + MoveBodies for every thread that
 * exists, then the host compaction of src/nbody.cu:488-510 (stable, keeps
 * bodies with mass != 0.f).  `block` holds n bodies in the BodiesData layout
 * on entry and the survivors (re-laid-out for the new n) on exit.
 * Events (optional) come out sorted by (i, visit order).  Returns the new n.
 * stats[0] = pairs evaluated, stats[1] = events, stats[2] = asserts that would
 * have fired in the reference.
 */
int orc_step(float *block, int n, const orc_params *par,
             orc_event *events, long long ev_cap, long long *stats)
{
    orc_cov cov;
    orc_coverage(n, par->coverage, &cov);
    float *pos = block, *vel = block + 2 * (size_t)n;
    float *mass = block + 4 * (size_t)n, *rad = block + 5 * (size_t)n;
    orc_row *rows = (orc_row *)malloc((size_t)(cov.n_active > 0 ? cov.n_active : 1) * sizeof(orc_row));
    int nt = 1;
#ifdef _OPENMP
    nt = par->threads > 0 ? par->threads : omp_get_max_threads();
#endif
    ev_buf *bufs = (ev_buf *)calloc((size_t)nt, sizeof(ev_buf));
    long long visited = 0, asserts = 0;
#ifdef _OPENMP
#pragma omp parallel num_threads(nt) reduction(+ : visited, asserts)
#endif
    {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        ev_buf *evb = events ? &bufs[tid] : NULL;
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 32)
#endif
        for (int i = 0; i < cov.n_active; ++i) {
            eval_row(pos, vel, mass, rad, &cov, par, i, &rows[i], evb);
            visited += rows[i].visited;
            asserts += rows[i].asserts;
        }
    }
    long long nev = 0;
    if (events) {
        /* Each row was walked by one thread, so per-row order is already the
         * visit order; a stable sort by i restores the global order. */
        size_t total = 0;
        for (int t = 0; t < nt; ++t) total += bufs[t].n;
        orc_event *all = (orc_event *)malloc((total ? total : 1) * sizeof(orc_event));
        /* stable: tag with sequence, sort by (i, seq) */
        size_t p = 0;
        for (int t = 0; t < nt; ++t) {
            memcpy(all + p, bufs[t].ev, bufs[t].n * sizeof(orc_event));
            p += bufs[t].n;
        }
        /* rows are contiguous inside one thread's buffer, so a merge sort keyed
         * on i alone is stable as long as the sort is stable; use mergesort by
         * hand to avoid qsort's instability. */
        orc_event *tmp = (orc_event *)malloc((total ? total : 1) * sizeof(orc_event));
        for (size_t width = 1; width < total; width *= 2) {
            for (size_t lo = 0; lo < total; lo += 2 * width) {
                size_t mid = lo + width < total ? lo + width : total;
                size_t hi = lo + 2 * width < total ? lo + 2 * width : total;
                size_t a = lo, c = mid, o = lo;
                while (a < mid && c < hi) tmp[o++] = (cmp_event(&all[c], &all[a]) < 0) ? all[c++] : all[a++];
                while (a < mid) tmp[o++] = all[a++];
                while (c < hi) tmp[o++] = all[c++];
            }
            orc_event *sw = all; all = tmp; tmp = sw;
        }
        nev = (long long)total;
        long long ncopy = nev < ev_cap ? nev : ev_cap;
        memcpy(events, all, (size_t)ncopy * sizeof(orc_event));
        free(all); free(tmp);
    } else {
        for (int i = 0; i < cov.n_active; ++i) nev += rows[i].hits;
    }
    for (int t = 0; t < nt; ++t) free(bufs[t].ev);
    free(bufs);

    /* commit: velocities (:264), positions/masses/radii (:288-290) for threads that exist */
    for (int i = 0; i < cov.n_active; ++i) {
        vel[2 * i] = rows[i].vx; vel[2 * i + 1] = rows[i].vy;
        pos[2 * i] = rows[i].px; pos[2 * i + 1] = rows[i].py;
        mass[i] = rows[i].m; rad[i] = rows[i].r;
    }
    if (par->merge) {
        /*
         * Conserving lowest-index merge, opt-in (the north star's wording; the reference neither conserves mass
         * nor transfers momentum).  Every body points at the lowest index among itself and its hit
         * partners; following the pointers ends at a root; a root takes mass, momentum (at the post-force
         * velocities) and growth * radius of everything that ends at it, in ascending index order, keeps its own
         * position and moves on with P / M; everything else is removed.  float32, one rounding per step as written.
         */
        int *root = (int *)malloc((size_t)(n > 0 ? n : 1) * sizeof(int));
        float *M = (float *)malloc((size_t)(n > 0 ? n : 1) * sizeof(float));
        float *Px = (float *)malloc((size_t)(n > 0 ? n : 1) * sizeof(float));
        float *Py = (float *)malloc((size_t)(n > 0 ? n : 1) * sizeof(float));
        int *members = (int *)calloc((size_t)(n > 0 ? n : 1), sizeof(int));
        for (int i = 0; i < n; ++i) {
            int r = i < cov.n_active ? rows[i].absorber : i;
            while (r != (r < cov.n_active ? rows[r].absorber : r)) r = rows[r].absorber;
            root[i] = r;
            M[i] = mass[i];
            Px[i] = mass[i] * vel[2 * i];
            Py[i] = mass[i] * vel[2 * i + 1];
        }
        for (int k = 0; k < n; ++k) {
            const int r = root[k];
            if (r == k) continue;
            ++members[r];
            M[r] += mass[k];
            Px[r] = fmaf(mass[k], vel[2 * k], Px[r]);
            Py[r] = fmaf(mass[k], vel[2 * k + 1], Py[r]);
            rad[r] = fmaf(par->growth, rad[k], rad[r]);
        }
        for (int k = 0; k < n; ++k) {
            if (root[k] != k) { mass[k] = 0.f; continue; }
            if (members[k] > 0) {                      /* a root that absorbed something */
                vel[2 * k] = Px[k] / M[k];
                vel[2 * k + 1] = Py[k] / M[k];
                mass[k] = M[k];
            }
        }
        free(root); free(M); free(Px); free(Py); free(members);
    }
    free(rows);

    /* host compaction, :488-510 */
    int new_n = 0;
    for (int i = 0; i < n; ++i) if (mass[i] != 0.f) ++new_n;
    float *nb = (float *)malloc((size_t)(new_n > 0 ? new_n : 1) * 6 * sizeof(float));
    float *npos = nb, *nvel = nb + 2 * (size_t)new_n, *nmass = nb + 4 * (size_t)new_n, *nrad = nb + 5 * (size_t)new_n;
    int w = 0;
    for (int i = 0; i < n; ++i) {
        if (mass[i] != 0.f) {
            npos[2 * w] = pos[2 * i]; npos[2 * w + 1] = pos[2 * i + 1];
            nvel[2 * w] = vel[2 * i]; nvel[2 * w + 1] = vel[2 * i + 1];
            nmass[w] = mass[i]; nrad[w] = rad[i];
            ++w;
        }
    }
    memcpy(block, nb, (size_t)new_n * 6 * sizeof(float));
    free(nb);
    if (stats) { stats[0] = visited; stats[1] = nev; stats[2] = asserts; }
    return new_n;
}

/*
 * generateImage, src/nbody.cu:294-348, for the bodies that exist (i < n; the reference launches a stale grid
 * without a bound check): filled discs of value 0 into an image pre-filled with 254
 * (:534).  Float expressions are kept in the reference's order; -ffp-contract=off matches its PTX (no
 * contraction is possible in these expressions anyway).
 */
void orc_render(const float *block, int n, int drawn, unsigned char *img, int width, int height, int field_w, int field_h)
{
    const float *pos = block, *rad = block + 5 * (size_t)n;
    memset(img, 254, (size_t)width * height);
    const int dfw = field_w << 1, dfh = field_h << 1;                       /* :314-315 */
    /* `drawn`: threads of the launch grid, 128 * floor(n_before_the_step / 128) in the reference's loop (:473,535) */
    for (int i = 0; i < n && i < drawn; ++i) {
        const float pr = (rad[i] * width) / field_w;                          /* :310 */
        const int cx = (int)(((pos[2 * i] + field_w) / dfw) * width);         /* :318 */
        const int cy = (int)(((pos[2 * i + 1] + field_h) / dfh) * height);    /* :319 */
        const int y_min = cy - pr < 0 ? 0 : (int)(cy - pr);                   /* :323-326 */
        const int y_max = cy + pr >= height ? height : (int)(cy + pr);
        const int x_min = cx - pr < 0 ? 0 : (int)(cx - pr);
        const int x_max = cx + pr > width ? width : (int)(cx + pr);
        const int r2 = (int)(pr * pr);
        for (int y = y_min; y < y_max; ++y)
            for (int x = x_min; x < x_max; ++x)
                if ((x - cx) * (x - cx) + (y - cy) * (y - cy) <= r2 && x >= 0 && x < width && y >= 0 && y < height)
                    img[(size_t)width * y + x] = 0;                           /* :344 */
    }
}

/* FNV-1a-64 over a byte range; used for compact golden fixtures. */
uint64_t orc_fnv1a64(const void *data, size_t nbytes, uint64_t h)
{
    const unsigned char *p = (const unsigned char *)data;
    if (h == 0) h = 0xcbf29ce484222325ULL;
    for (size_t k = 0; k < nbytes; ++k) { h ^= p[k]; h *= 0x100000001b3ULL; }
    return h;
}

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
