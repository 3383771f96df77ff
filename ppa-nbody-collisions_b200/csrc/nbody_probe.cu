// nbody_probe.cu -- the FP32 peak the force kernels' roofline is quoted against, measured on the device at hand:
// a stream of independent packed fma.rn.f32x2 (SASS FFMA2), 32 warps per SM, no memory traffic.  bench.py calls it
// in the same run as the timed steps so that `roofline.peak` is a measurement of this GPU at its current clocks, not a
// nameplate product (MEASURED_PEAKS.json has no FP32 entry).
#include <cuda_runtime.h>

#include "nbody_b200.h"

namespace {

constexpr int kIlp = 8;

__global__ void __launch_bounds__(256) ffma2_probe_kernel(float *out, const int iters, const float a, const float b)
{
    float2 acc[kIlp];
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
#pragma unroll
    for (int k = 0; k < kIlp; ++k) acc[k] = make_float2(threadIdx.x * 1e-3f + k, threadIdx.x * 2e-3f + k);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < kIlp; ++k) acc[k] = __ffma2_rn(acc[k], a2, b2);
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < kIlp; ++k) s += acc[k].x + acc[k].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace

extern "C" int nb_probe_fp32(int device, double *tflops)
{
    if (!tflops) return NB_ERR_INVALID;
    *tflops = 0.0;
    cudaDeviceProp prop;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        cudaGetLastError();
        return NB_ERR_CUDA;
    }
    const int grid = prop.multiProcessorCount * 4, threads = 256, iters = 1 << 15;
    float *out = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaStream_t s = nullptr;
    int rc = NB_ERR_CUDA;
    float best = 0.f;
    if (cudaMalloc(&out, sizeof(float) * grid * threads) != cudaSuccess) goto done;
    if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) goto done;
    if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) goto done;
    for (int rep = 0; rep < 6; ++rep) {          // first repetition warms up
        cudaEventRecord(e0, s);
        ffma2_probe_kernel<<<grid, threads, 0, s>>>(out, iters, 0.999f, 0.001f);
        cudaEventRecord(e1, s);
        if (cudaEventSynchronize(e1) != cudaSuccess) goto done;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms > 0.f && (best == 0.f || ms < best)) best = ms;
    }
    if (best > 0.f) {
        // per thread and iteration: kIlp packed FMAs = 2 lanes x 2 flop each
        *tflops = (double)grid * threads * iters * kIlp * 4.0 / (best * 1e-3) / 1e12;
        rc = NB_OK;
    }
done:
    if (rc != NB_OK) cudaGetLastError();
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (s) cudaStreamDestroy(s);
    if (out) cudaFree(out);
    return rc;
}
