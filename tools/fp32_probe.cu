// fp32_probe.cu -- B200 FP32-pipe microbenchmarks used to size the force kernel.
//
//   build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -o build/fp32_probe tools/fp32_probe.cu
//   run  : build/fp32_probe            (prints one JSON line per probe)
//
// Probes:
//   ffma        : dependent-chain-free scalar FFMA stream (measured FP32 peak, 2 flop per lane-op)
//   ffma2       : same with packed fma.rn.f32x2 (does one issue slot carry two lane-ops?)
//   pair_scalar : the force inner loop, scalar form (11 issue slots per interaction)
//   pair_packed : the force inner loop, two j per packed op (9 packed FP32 ops + 2 MUFU + 2 FSETP per 2 interactions)
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

template <int ILP>
__global__ void k_ffma(float *out, int iters, float a, float b)
{
    float acc[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) acc[k] = threadIdx.x * 1e-3f + k;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) acc[k] = fmaf(acc[k], a, b);
    }
    float s = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += acc[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void k_ffma2(float *out, int iters, float a, float b)
{
    float2 acc[ILP];
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
#pragma unroll
    for (int k = 0; k < ILP; ++k) acc[k] = make_float2(threadIdx.x * 1e-3f + k, threadIdx.x * 2e-3f + k);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) acc[k] = __ffma2_rn(acc[k], a2, b2);
    }
    float s = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += acc[k].x + acc[k].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

constexpr int TJ = 256;
constexpr int IPT = 4;

// scalar inner loop: per interaction FADD,FADD,FMUL,FFMA,FSETP,MUFU,FMUL,FMUL,FMUL,FFMA,FFMA
__global__ void __launch_bounds__(128) k_pair_scalar(float *out, int reps)
{
    __shared__ __align__(16) float sx[TJ], sy[TJ], sm[TJ];
    for (int k = threadIdx.x; k < TJ; k += blockDim.x) {
        sx[k] = 1000.f + 37.f * k; sy[k] = -500.f + 11.f * k; sm[k] = 1e10f + k;
    }
    __syncthreads();
    float xi[IPT], yi[IPT], thr[IPT], fx[IPT], fy[IPT];
    bool cand = false;
#pragma unroll
    for (int q = 0; q < IPT; ++q) {
        xi[q] = -3.f * threadIdx.x - q; yi[q] = 7.f * threadIdx.x + q; thr[q] = 1.0f + q;
        fx[q] = 0; fy[q] = 0;
    }
    for (int rep = 0; rep < reps; ++rep) {
#pragma unroll 2
        for (int j = 0; j < TJ; j += 4) {
            const float4 X = *reinterpret_cast<const float4 *>(&sx[j]);
            const float4 Y = *reinterpret_cast<const float4 *>(&sy[j]);
            const float4 M = *reinterpret_cast<const float4 *>(&sm[j]);
            const float xs[4] = {X.x, X.y, X.z, X.w}, ys[4] = {Y.x, Y.y, Y.z, Y.w}, ms[4] = {M.x, M.y, M.z, M.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
#pragma unroll
                for (int q = 0; q < IPT; ++q) {
                    const float dx = xs[u] - xi[q], dy = ys[u] - yi[q];
                    const float d2 = fmaf(dx, dx, dy * dy);
                    cand |= (d2 <= thr[q]);
                    const float inv = rsqrt_ftz(d2);
                    const float s = (inv * inv) * (inv * ms[u]);
                    fx[q] = fmaf(dx, s, fx[q]);
                    fy[q] = fmaf(dy, s, fy[q]);
                }
            }
        }
    }
    float s = cand ? 1.f : 0.f;
#pragma unroll
    for (int q = 0; q < IPT; ++q) s += fx[q] + fy[q];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// packed inner loop: two j per packed op.
__global__ void __launch_bounds__(128) k_pair_packed(float *out, int reps)
{
    __shared__ __align__(16) float sx[TJ], sy[TJ], sm[TJ];
    for (int k = threadIdx.x; k < TJ; k += blockDim.x) {
        sx[k] = 1000.f + 37.f * k; sy[k] = -500.f + 11.f * k; sm[k] = 1e10f + k;
    }
    __syncthreads();
    float2 nxi[IPT], nyi[IPT], fx[IPT], fy[IPT];
    float thr[IPT];
    bool cand = false;
#pragma unroll
    for (int q = 0; q < IPT; ++q) {
        const float x = -3.f * threadIdx.x - q, y = 7.f * threadIdx.x + q;
        nxi[q] = make_float2(-x, -x); nyi[q] = make_float2(-y, -y); thr[q] = 1.0f + q;
        fx[q] = make_float2(0, 0); fy[q] = make_float2(0, 0);
    }
    for (int rep = 0; rep < reps; ++rep) {
#pragma unroll 2
        for (int j = 0; j < TJ; j += 4) {
            const float4 X = *reinterpret_cast<const float4 *>(&sx[j]);
            const float4 Y = *reinterpret_cast<const float4 *>(&sy[j]);
            const float4 M = *reinterpret_cast<const float4 *>(&sm[j]);
            const float2 xs[2] = {make_float2(X.x, X.y), make_float2(X.z, X.w)};
            const float2 ys[2] = {make_float2(Y.x, Y.y), make_float2(Y.z, Y.w)};
            const float2 ms[2] = {make_float2(M.x, M.y), make_float2(M.z, M.w)};
#pragma unroll
            for (int u = 0; u < 2; ++u) {
#pragma unroll
                for (int q = 0; q < IPT; ++q) {
                    const float2 dx = __fadd2_rn(xs[u], nxi[q]);
                    const float2 dy = __fadd2_rn(ys[u], nyi[q]);
                    const float2 d2 = __ffma2_rn(dx, dx, __fmul2_rn(dy, dy));
                    cand |= (d2.x <= thr[q]);
                    cand |= (d2.y <= thr[q]);
                    const float2 inv = make_float2(rsqrt_ftz(d2.x), rsqrt_ftz(d2.y));
                    const float2 s = __fmul2_rn(__fmul2_rn(inv, inv), __fmul2_rn(inv, ms[u]));
                    fx[q] = __ffma2_rn(dx, s, fx[q]);
                    fy[q] = __ffma2_rn(dy, s, fy[q]);
                }
            }
        }
    }
    float s = cand ? 1.f : 0.f;
#pragma unroll
    for (int q = 0; q < IPT; ++q) s += fx[q].x + fx[q].y + fy[q].x + fy[q].y;
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

int main(int argc, char **argv)
{
    const bool only_pair = argc > 1 && argv[1][0] == 'p';
    const int only_cps = argc > 2 ? atoi(argv[2]) : 0;
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int sms = p.multiProcessorCount;
    printf("{\"probe\": \"device\", \"name\": \"%s\", \"sms\": %d, \"clock_khz\": %d, \"cc\": \"%d.%d\"}\n",
           p.name, sms, clk_khz, p.major, p.minor);
    float *out;
    CK(cudaMalloc(&out, sizeof(float) * 1024 * 1024 * 4));
    const double nameplate = (double)sms * 128 * 2 * 1.965e9;

    for (int wpsm : {8, 16, 32}) {          // warps per SM
        if (only_pair) break;
        const int blocks = sms * wpsm / 4, threads = 128, iters = 20000;
        {
            float ms = time_ms([&] { k_ffma<8><<<blocks, threads>>>(out, iters, 1.0001f, 0.5f); }, 5);
            double flops = 2.0 * 8 * iters * (double)blocks * threads;
            printf("{\"probe\": \"ffma\", \"warps_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.2f, \"frac_nameplate\": %.3f}\n",
                   wpsm, ms, flops / ms * 1e-9, flops / ms * 1e3 / nameplate);
        }
        {
            float ms = time_ms([&] { k_ffma2<8><<<blocks, threads>>>(out, iters, 1.0001f, 0.5f); }, 5);
            double flops = 2.0 * 2 * 8 * iters * (double)blocks * threads;
            printf("{\"probe\": \"ffma2\", \"warps_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.2f, \"frac_nameplate\": %.3f}\n",
                   wpsm, ms, flops / ms * 1e-9, flops / ms * 1e3 / nameplate);
        }
    }
    for (int cps : {1, 2, 3, 4, 6, 8}) {    // CTAs (4 warps) per SM
        if (only_cps && cps != only_cps) continue;
        const int blocks = sms * cps, reps = only_pair ? 50 : 400;
        const double inter = (double)blocks * 128 * IPT * TJ * reps;
        {
            float ms = time_ms([&] { k_pair_scalar<<<blocks, 128>>>(out, reps); }, 5);
            printf("{\"probe\": \"pair_scalar\", \"ctas_per_sm\": %d, \"ms\": %.4f, \"ginter_per_s\": %.1f, \"frac_20flop_nameplate\": %.3f}\n",
                   cps, ms, inter / ms * 1e-6, inter * 20 / ms * 1e3 / nameplate);
        }
        {
            float ms = time_ms([&] { k_pair_packed<<<blocks, 128>>>(out, reps); }, 5);
            printf("{\"probe\": \"pair_packed\", \"ctas_per_sm\": %d, \"ms\": %.4f, \"ginter_per_s\": %.1f, \"frac_20flop_nameplate\": %.3f}\n",
                   cps, ms, inter / ms * 1e-6, inter * 20 / ms * 1e3 / nameplate);
        }
    }
    CK(cudaFree(out));
    return 0;
}
