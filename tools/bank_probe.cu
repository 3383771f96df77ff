// bank_probe.cu -- does a packed FFMA2 with three distinct register-pair operands cost more than one with two?
//   build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/bank_probe tools/bank_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, int iters, float seed)
{
    float2 acc[8], x[8], y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        acc[i] = make_float2(threadIdx.x * 1e-3f + i, seed + i);
        x[i] = make_float2(1.0f + 1e-6f * (threadIdx.x + i), 1.0f - 1e-6f * i);
        y[i] = make_float2(1e-3f * i + seed, -1e-3f * i);
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) acc[i] = __ffma2_rn(x[i], y[i], acc[i]);              // 3 distinct pairs
            if (MODE == 1) acc[i] = __ffma2_rn(x[i], x[i], acc[i]);              // 2 distinct pairs
            if (MODE == 2) acc[i] = __ffma2_rn(x[i], y[i & 1], acc[i]);          // y shared by neighbours (reuse cache)
            if (MODE == 3) { acc[i] = __fmul2_rn(x[i], acc[i]); }                // FMUL2, 2 distinct
            if (MODE == 4) { acc[i].x = fmaf(x[i].x, y[i].x, acc[i].x); acc[i].y = fmaf(x[i].y, y[i].y, acc[i].y); }   // scalar, 3 distinct
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
int run(const char *name, float *out, int sms)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int blocks = sms * 4, iters = 20000;
    k<MODE><<<blocks, 256>>>(out, iters, 0.5f);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        CK(cudaEventRecord(e0));
        k<MODE><<<blocks, 256>>>(out, iters, 0.5f);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    const double lane_ops = (double)blocks * 256 * iters * 8 * 2;
    printf("{\"probe\": \"%s\", \"ms\": %.4f, \"tflops\": %.2f, \"frac_nameplate\": %.3f}\n", name, best, lane_ops * 2 / best * 1e-9,
           lane_ops * 2 / best * 1e3 / ((double)sms * 128 * 2 * 1.965e9));
    return 0;
}

int main()
{
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    float *out;
    CK(cudaMalloc(&out, 4 << 20));
    run<0>("ffma2_3_distinct_pairs", out, p.multiProcessorCount);
    run<1>("ffma2_2_distinct_pairs", out, p.multiProcessorCount);
    run<2>("ffma2_shared_operand", out, p.multiProcessorCount);
    run<3>("fmul2_2_distinct_pairs", out, p.multiProcessorCount);
    run<4>("ffma_scalar_3_distinct", out, p.multiProcessorCount);
    return 0;
}
