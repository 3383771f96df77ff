/*
 * gpu_ref_harness.cu -- drives the UNMODIFIED reference kernels.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/nbody_oracle.c header).  This file
 * contains no reference source: it #includes the reference translation unit
 * where it lies (path given by -DREF_NBODY_CU, normally
 * /root/reference/src/nbody.cu) with its main() renamed, and then calls the
 * reference's own ComputeForces / MoveBodies kernels and BodiesData store
 * through a small C interface.  oracle/Makefile builds it into
 * oracle/_ref/libnbody_gpuref.so (git-ignored; travels to the GPU box).
 *
 * gpuref_step() follows the body of the reference's main loop,
 * src/nbody.cu:461-545, without the image path (:513-522, :529-539).
 */
#define main nbody_reference_main
#include REF_NBODY_CU
#undef main

#include <cstdio>
#include <cstring>
#include <unistd.h>

namespace {
BodiesData g_bodies;
bool g_open = false;
int g_n = 0;
cudaStream_t g_calc = nullptr;

size_t ref_smem_bytes()
{
    /* src/nbody.cu:451 */
    return THREADS_PER_BLOCK * ((2 * (sizeof(Vec2f) + sizeof(float) + sizeof(float))) + 2 * sizeof(Vec2f));
}
int ref_blocks(int n)
{
    return n < THREADS_PER_BLOCK ? 1 : n / THREADS_PER_BLOCK;   /* src/nbody.cu:473 */
}
}  // namespace

extern "C" {

int gpuref_device_count()
{
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess) { cudaGetLastError(); return 0; }
    return c;
}

/* Take a copy of `block` (BodiesData layout, 6*n floats). */
int gpuref_open(const float *block, int n)
{
    if (g_open) return -1;
    if (n <= 0) return -2;
    g_bodies = BodiesData();
    g_bodies.alloc(n);
    memcpy(g_bodies.contiguousData, block, g_bodies.size);
    g_n = n;
    if (cudaStreamCreate(&g_calc) != cudaSuccess) return -3;
    g_open = true;
    return 0;
}

int gpuref_n() { return g_open ? g_n : -1; }

int gpuref_read(float *block_out)
{
    if (!g_open) return -1;
    memcpy(block_out, g_bodies.contiguousData, g_bodies.size);
    return g_n;
}

/*
 * One iteration of the reference loop.  Returns the new body count (>= 0) or a
 * negative CUDA error.  kernel_ms (optional) receives the device time of
 * ComputeForces + MoveBodies measured with events on the calculation stream.
 */
int gpuref_step(float dt, float growth, int field_w, int field_h, float *kernel_ms)
{
    if (!g_open) return -1;
    const int n = g_n;
    float *d_um = nullptr, *d_ur = nullptr;
    Vec2f *d_uv = nullptr;                         /* never dereferenced, src/nbody.cu:441 */
    cudaMalloc((void **)&d_um, n * sizeof(float));          /* :463 */
    cudaMalloc((void **)&d_ur, n * sizeof(float));          /* :464 */
    const int blocks = ref_blocks(n);                        /* :473 */
    g_bodies.uploadToDevice(g_calc);                         /* :476 */
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, g_calc);
    ComputeForces<<<blocks, THREADS_PER_BLOCK, ref_smem_bytes(), g_calc>>>(
        g_bodies.d_contiguousData, d_um, d_uv, d_ur, n, dt, field_w, field_h, blocks, growth);   /* :481 */
    MoveBodies<<<blocks, THREADS_PER_BLOCK, 0, g_calc>>>(
        g_bodies.d_contiguousData, d_um, d_uv, d_ur, g_bodies.numBodies, dt);                    /* :483 */
    cudaEventRecord(e1, g_calc);
    cudaMemcpyAsync(g_bodies.contiguousData, g_bodies.d_contiguousData, g_bodies.size,
                    cudaMemcpyDeviceToHost, g_calc);                                             /* :486 */
    cudaError_t err = cudaStreamSynchronize(g_calc);
    if (err == cudaSuccess) err = cudaGetLastError();
    if (kernel_ms) { *kernel_ms = 0.f; cudaEventElapsedTime(kernel_ms, e0, e1); }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d_um); cudaFree(d_ur);                                                              /* :541-542 */
    if (err != cudaSuccess) {
        fprintf(stderr, "gpuref_step: CUDA error %s\n", cudaGetErrorString(err));
        return -100 - (int)err;
    }
    int new_n = 0;                                                                               /* :488-494 */
    for (int i = 0; i < n; ++i)
        if (g_bodies.Masses[i] != 0.f) ++new_n;
    if (new_n == 0) {             /* the reference would malloc(0) here; stop cleanly instead */
        g_bodies.freeData();
        g_n = 0;
        g_open = false;
        cudaStreamDestroy(g_calc);
        return 0;
    }
    BodiesData fresh;
    fresh.alloc(new_n);                                                                          /* :496 */
    int w = 0;
    for (int i = 0; i < n; ++i) {                                                                /* :499-510 */
        if (g_bodies.Masses[i] != 0.f) {
            fresh.Positions[w] = g_bodies.Positions[i];
            fresh.Velocities[w] = g_bodies.Velocities[i];
            fresh.Masses[w] = g_bodies.Masses[i];
            fresh.Radii[w] = g_bodies.Radii[i];
            ++w;
        }
    }
    g_bodies.freeData();                                                                         /* :525 */
    g_bodies = fresh;                                                                            /* :526 */
    g_n = new_n;
    return new_n;
}

/*
 * The reference's image of the CURRENT bodies, launched as its main loop does (src/nbody.cu:529-539): background
 * 254, then generateImage<<<blocks, 128, 128 * (sizeof(Vec2f) + sizeof(float))>>> with `blocks` taken from
 * `grid_n` -- the loop re-uses the step's grid, floor(n_before_the_step / 128) blocks, so bodies at and beyond
 * 128 * blocks are not drawn.  Launches whose grid has threads beyond the live bodies read outside the body store in
 * the reference (positions[i], radii[i] for i >= numBodies): those are refused here (-4) instead of pinned.
 */
int gpuref_render(int width, int height, int field_w, int field_h, int grid_n, unsigned char *img_out)
{
    if (!g_open) return -1;
    if (width <= 0 || height <= 0 || !img_out) return -2;
    const int blocks = ref_blocks(grid_n);
    if (blocks * THREADS_PER_BLOCK > g_n) return -4;
    const size_t image_size = (size_t)width * height;
    char *d_img = nullptr;
    cudaStream_t image_stream;
    cudaStreamCreate(&image_stream);                                                             /* :457 */
    g_bodies.uploadToDevice();                                                                   /* :532 */
    cudaMalloc((void **)&d_img, image_size);                                                     /* :533 */
    cudaMemsetAsync(d_img, 254, image_size, image_stream);                                       /* :534 */
    generateImage<<<blocks, THREADS_PER_BLOCK, THREADS_PER_BLOCK * (sizeof(Vec2f) + sizeof(float)), image_stream>>>(
        g_bodies.d_contiguousData, g_bodies.numBodies, d_img, width, height, field_w, field_h);  /* :535-536 */
    cudaMemcpyAsync(img_out, d_img, image_size, cudaMemcpyDeviceToHost, image_stream);           /* :537 */
    cudaError_t err = cudaStreamSynchronize(image_stream);
    if (err == cudaSuccess) err = cudaGetLastError();
    cudaFree(d_img);
    cudaStreamDestroy(image_stream);
    /* the step path uploads again only if the device copy is gone (BodiesData::uploadToDevice, :88-96) */
    return err == cudaSuccess ? 0 : -100 - (int)err;
}

void gpuref_close()
{
    if (!g_open) return;
    g_bodies.freeData();
    cudaStreamDestroy(g_calc);
    g_open = false;
    g_n = 0;
}

/*
 * Kernel-only timing of the unmodified ComputeForces + MoveBodies on a state
 * that stays resident (no per-step malloc / copies): the best case for the
 * reference.  Runs `warmup` + `reps` launches; state is not compacted, so every
 * launch does the same amount of pair work.  Returns 0 or a negative error.
 */
int gpuref_time_kernels(const float *block, int n, float dt, float growth, int field_w, int field_h,
                        int warmup, int reps, float *ms_total)
{
    void *d_block = nullptr;
    float *d_um = nullptr, *d_ur = nullptr;
    Vec2f *d_uv = nullptr;
    const size_t bytes = (size_t)n * 24;
    cudaStream_t s;
    cudaStreamCreate(&s);
    cudaMalloc(&d_block, bytes);
    cudaMalloc((void **)&d_um, n * sizeof(float));
    cudaMalloc((void **)&d_ur, n * sizeof(float));
    cudaMemcpyAsync(d_block, block, bytes, cudaMemcpyHostToDevice, s);
    const int blocks = ref_blocks(n);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int it = 0; it < warmup + reps; ++it) {
        if (it == warmup) cudaEventRecord(e0, s);
        /* positions are restored each launch so collisions/forces stay identical */
        ComputeForces<<<blocks, THREADS_PER_BLOCK, ref_smem_bytes(), s>>>(
            d_block, d_um, d_uv, d_ur, n, dt, field_w, field_h, blocks, growth);
        MoveBodies<<<blocks, THREADS_PER_BLOCK, 0, s>>>(d_block, d_um, d_uv, d_ur, n, 0.0f);
        /* dt = 0 in MoveBodies keeps positions fixed; masses of dead bodies become 0,
         * which only removes their pull, not the pair tests. */
    }
    cudaEventRecord(e1, s);
    cudaError_t err = cudaStreamSynchronize(s);
    if (err == cudaSuccess) err = cudaGetLastError();
    if (ms_total) { *ms_total = 0.f; cudaEventElapsedTime(ms_total, e0, e1); }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d_block); cudaFree(d_um); cudaFree(d_ur);
    cudaStreamDestroy(s);
    return err == cudaSuccess ? 0 : -100 - (int)err;
}

/*
 * Run the reference program itself (its own main(), renamed) inside `dir`,
 * which must hold nbodyConfig.txt and the image directory it names.  The
 * reference ends with cudaDeviceReset(): call this from a dedicated process.
 */
int gpuref_main_in(const char *dir)
{
    if (chdir(dir) != 0) return -1;
    char arg0[] = "nbodyCuda";
    char *argv[] = {arg0, nullptr};
    return nbody_reference_main(1, argv);
}

}  // extern "C"
