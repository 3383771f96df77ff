// sym_probe.cu -- would evaluating each unordered pair once (Newton's third law) beat the one-sided loop?
//
//   build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -o build/sym_probe tools/sym_probe.cu
//   run  : build/sym_probe            (one JSON line per variant)
//
// one-sided (what force_kernel does): IPT rows per lane in registers, bodies j broadcast from shared memory,
//     9 packed f32x2 ops + 2 MUFU + pre-test per (row, j pair); one ordered interaction per evaluation.
// two-sided, systolic: IPT rows per lane AND one j pair per lane; the j pair and its own force accumulators travel
//     round the warp with SHFL (10 per sub-step), 32 sub-steps per 64-body chunk; 12 packed ops + 2 MUFU + pre-test
//     per (row, j pair), TWO ordered interactions per evaluation.
// two-sided, shared: same, but the j pair is read from shared memory at a rotating index and its accumulators are
//     read-modify-written in a per-warp private shared array (no SHFL).
// two-sided, hybrid (not measured yet, prepared for the next round): positions and masses are re-read from a doubled
//     copy of the chunk in shared memory at a rotating offset (3 LDS.64, immediate offsets once unrolled), only the
//     four accumulator registers travel by SHFL: 7 instead of 10 issue slots per sub-step.
// "ginter_per_s" counts ORDERED interactions in every variant.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ float rsqrt_ftz(float x)
{
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

constexpr int TJ = 512;
enum Mode { ONE_SIDED = 0, SYM_SHFL = 1, SYM_SHFL_NOTEST = 2, SYM_LDS = 3, ONE_SIDED_NOTEST = 4, SYM_HYB = 5 };

template <int IPT, int MODE, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_sym(float *out, int reps)
{
    __shared__ __align__(16) float sx[TJ], sy[TJ], sm[TJ];
    __shared__ __align__(16) float4 sg[MODE == SYM_LDS ? THREADS / 32 : 1][MODE == SYM_LDS ? TJ / 2 : 1];
    // SYM_HYB: per warp, the chunk's 32 (x, y, m) pairs twice in a row, so that lane + s never wraps
    __shared__ __align__(16) float2 dup[MODE == SYM_HYB ? THREADS / 32 : 1][3][MODE == SYM_HYB ? 64 : 1];
    for (int k = threadIdx.x; k < TJ; k += blockDim.x) {
        sx[k] = 1000.f + 37.f * k; sy[k] = -500.f + 11.f * k; sm[k] = 1e10f + k;
    }
    if (MODE == SYM_LDS)
        for (int k = threadIdx.x; k < THREADS / 32 * (TJ / 2); k += blockDim.x) (&sg[0][0])[k] = make_float4(0, 0, 0, 0);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float nx[IPT], ny[IPT], thr[IPT], nm[IPT];
    float2 fx[IPT], fy[IPT];
    bool cand[IPT];
#pragma unroll
    for (int q = 0; q < IPT; ++q) {
        nx[q] = 3.f * threadIdx.x + q; ny[q] = -7.f * threadIdx.x - q; thr[q] = 1.0f + q; nm[q] = -1e9f - threadIdx.x;
        fx[q] = make_float2(0, 0); fy[q] = make_float2(0, 0); cand[q] = false;
    }
    float gsum = 0.f;
    for (int rep = 0; rep < reps; ++rep) {
        if (MODE == ONE_SIDED || MODE == ONE_SIDED_NOTEST) {
#pragma unroll 8
            for (int j = 0; j < TJ; j += 4) {
                const float4 X = *reinterpret_cast<const float4 *>(&sx[j]);
                const float4 Y = *reinterpret_cast<const float4 *>(&sy[j]);
                const float4 M = *reinterpret_cast<const float4 *>(&sm[j]);
                const float2 xs[2] = {make_float2(X.x, X.y), make_float2(X.z, X.w)};
                const float2 ys[2] = {make_float2(Y.x, Y.y), make_float2(Y.z, Y.w)};
                const float2 ms[2] = {make_float2(M.x, M.y), make_float2(M.z, M.w)};
#pragma unroll
                for (int q = 0; q < IPT; ++q)
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const float2 dx = __fadd2_rn(xs[u], make_float2(nx[q], nx[q]));
                        const float2 dy = __fadd2_rn(ys[u], make_float2(ny[q], ny[q]));
                        const float2 d2 = __ffma2_rn(dx, dx, __fmul2_rn(dy, dy));
                        if (MODE == ONE_SIDED) cand[q] |= (d2.x <= thr[q]) | (d2.y <= thr[q]);
                        const float2 inv = make_float2(rsqrt_ftz(d2.x), rsqrt_ftz(d2.y));
                        const float2 s = __fmul2_rn(__fmul2_rn(inv, inv), __fmul2_rn(inv, ms[u]));
                        fx[q] = __ffma2_rn(dx, s, fx[q]);
                        fy[q] = __ffma2_rn(dy, s, fy[q]);
                    }
            }
        } else {
#pragma unroll 1
            for (int c = 0; c < TJ / 64; ++c) {
                const int cc = (c + warp) & (TJ / 64 - 1);
                float2 xs, ys, ms, gx = make_float2(0, 0), gy = make_float2(0, 0);
                if (MODE != SYM_LDS) {
                    xs = *reinterpret_cast<const float2 *>(&sx[cc * 64 + 2 * lane]);
                    ys = *reinterpret_cast<const float2 *>(&sy[cc * 64 + 2 * lane]);
                    ms = *reinterpret_cast<const float2 *>(&sm[cc * 64 + 2 * lane]);
                }
                if (MODE == SYM_HYB) {
                    __syncwarp();
                    dup[warp][0][lane] = xs; dup[warp][0][lane + 32] = xs;
                    dup[warp][1][lane] = ys; dup[warp][1][lane + 32] = ys;
                    dup[warp][2][lane] = ms; dup[warp][2][lane + 32] = ms;
                    __syncwarp();
                }
#pragma unroll(MODE == SYM_HYB ? 32 : 4)
                for (int s = 0; s < 32; ++s) {
                    int slot = 0;
                    if (MODE == SYM_HYB) {
                        xs = dup[warp][0][lane + s];
                        ys = dup[warp][1][lane + s];
                        ms = dup[warp][2][lane + s];
                    }
                    if (MODE == SYM_LDS) {
                        slot = cc * 32 + ((lane + s) & 31);
                        xs = *reinterpret_cast<const float2 *>(&sx[2 * slot]);
                        ys = *reinterpret_cast<const float2 *>(&sy[2 * slot]);
                        ms = *reinterpret_cast<const float2 *>(&sm[2 * slot]);
                        const float4 g = sg[warp][slot];
                        gx = make_float2(g.x, g.y);
                        gy = make_float2(g.z, g.w);
                    }
#pragma unroll
                    for (int q = 0; q < IPT; ++q) {
                        const float2 dx = __fadd2_rn(xs, make_float2(nx[q], nx[q]));
                        const float2 dy = __fadd2_rn(ys, make_float2(ny[q], ny[q]));
                        const float2 d2 = __ffma2_rn(dx, dx, __fmul2_rn(dy, dy));
                        if (MODE != SYM_SHFL_NOTEST) cand[q] |= (d2.x <= thr[q]) | (d2.y <= thr[q]);
                        const float2 inv = make_float2(rsqrt_ftz(d2.x), rsqrt_ftz(d2.y));
                        const float2 i3 = __fmul2_rn(__fmul2_rn(inv, inv), inv);
                        const float2 sj = __fmul2_rn(i3, ms);
                        const float2 si = __fmul2_rn(i3, make_float2(nm[q], nm[q]));
                        fx[q] = __ffma2_rn(dx, sj, fx[q]);
                        fy[q] = __ffma2_rn(dy, sj, fy[q]);
                        gx = __ffma2_rn(dx, si, gx);
                        gy = __ffma2_rn(dy, si, gy);
                    }
                    if (MODE == SYM_LDS) {
                        sg[warp][slot] = make_float4(gx.x, gx.y, gy.x, gy.y);
                    } else if (MODE == SYM_HYB) {
                        const int src = (lane + 1) & 31;
                        gx.x = __shfl_sync(0xffffffffu, gx.x, src); gx.y = __shfl_sync(0xffffffffu, gx.y, src);
                        gy.x = __shfl_sync(0xffffffffu, gy.x, src); gy.y = __shfl_sync(0xffffffffu, gy.y, src);
                    } else {
                        const int src = (lane + 1) & 31;
                        xs.x = __shfl_sync(0xffffffffu, xs.x, src); xs.y = __shfl_sync(0xffffffffu, xs.y, src);
                        ys.x = __shfl_sync(0xffffffffu, ys.x, src); ys.y = __shfl_sync(0xffffffffu, ys.y, src);
                        ms.x = __shfl_sync(0xffffffffu, ms.x, src); ms.y = __shfl_sync(0xffffffffu, ms.y, src);
                        gx.x = __shfl_sync(0xffffffffu, gx.x, src); gx.y = __shfl_sync(0xffffffffu, gx.y, src);
                        gy.x = __shfl_sync(0xffffffffu, gy.x, src); gy.y = __shfl_sync(0xffffffffu, gy.y, src);
                    }
                }
                gsum += gx.x + gx.y + gy.x + gy.y;
            }
        }
    }
    float s = gsum;
#pragma unroll
    for (int q = 0; q < IPT; ++q) s += fx[q].x + fx[q].y + fy[q].x + fy[q].y + (cand[q] ? 1.f : 0.f);
    if (MODE == SYM_LDS) s += sg[warp][lane].x;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float time_ms(F launch, int reps)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

template <int IPT, int MODE, int THREADS, int MINB>
static void run(const char *name, float *out, int sms)
{
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_sym<IPT, MODE, THREADS, MINB>, THREADS, 0));
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, k_sym<IPT, MODE, THREADS, MINB>));
    const int cps = occ < MINB ? occ : MINB;
    const int blocks = sms * cps, reps = 100;
    const double ordered = (double)blocks * THREADS * IPT * TJ * reps * (MODE == ONE_SIDED || MODE == ONE_SIDED_NOTEST ? 1 : 2);
    const double nameplate = (double)sms * 128 * 2 * 1.965e9;
    float ms = time_ms([&] { k_sym<IPT, MODE, THREADS, MINB><<<blocks, THREADS>>>(out, reps); }, 5);
    printf("{\"probe\": \"%s\", \"ipt\": %d, \"threads\": %d, \"ctas_per_sm\": %d, \"warps_per_sm\": %d, \"regs\": %d, \"ms\": %.4f, "
           "\"ginter_per_s\": %.1f, \"frac_20flop_nameplate\": %.3f}\n",
           name, IPT, THREADS, cps, cps * THREADS / 32, fa.numRegs, ms, ordered / ms * 1e-6, ordered * 20 / ms * 1e3 / nameplate);
}

int main()
{
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    float *out;
    CK(cudaMalloc(&out, sizeof(float) * 1024 * 1024 * 8));
#define SW(MODE, NAME)                                  \
    run<2, MODE, 256, 3>(NAME, out, sms);               \
    run<2, MODE, 256, 2>(NAME, out, sms);               \
    run<4, MODE, 256, 2>(NAME, out, sms);               \
    run<4, MODE, 256, 3>(NAME, out, sms);               \
    run<4, MODE, 128, 4>(NAME, out, sms);
    SW(ONE_SIDED, "one_sided")
    SW(ONE_SIDED_NOTEST, "one_sided_notest")
    SW(SYM_SHFL, "sym_shfl")
    SW(SYM_SHFL_NOTEST, "sym_shfl_notest")
    SW(SYM_LDS, "sym_lds")
    SW(SYM_HYB, "sym_hybrid")
    run<8, SYM_SHFL, 256, 2>("sym_shfl", out, sms);          // 8 rows per lane: 1.25 SHFL per evaluation, 16 warps per SM
    run<8, SYM_SHFL_NOTEST, 256, 2>("sym_shfl_notest", out, sms);
    run<8, SYM_HYB, 256, 2>("sym_hybrid", out, sms);
    CK(cudaFree(out));
    return 0;
}
