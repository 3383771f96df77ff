// pair_probe.cu -- schedule experiments for the force inner loop (packed f32x2 FMA pipe + MUFU.RSQ).
//
//   build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -o build/pair_probe tools/pair_probe.cu
//   run  : build/pair_probe            (one JSON line per variant x occupancy)
//
// Every variant evaluates the same interactions (IPT rows per lane against a 256-body shared tile,
// `reps` times); they differ in how the source orders the two halves of an interaction:
//   A: dx, dy, d2, rsqrt, pre-test          B: inv^3 m, two accumulating FMAs
//   plain    : A and B of one pair back to back (what the first force kernel did)
//   pipe     : A of group g+1 before B of group g (group = 4 bodies x IPT rows)
//   nomufu   : rsqrt replaced by an FMA-pipe op (ceiling of the FMA pipe with this dependency pattern)
//   notest   : no pre-test (cost of the FMNMX/FSETP stream)
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

constexpr int TJ = 256;
enum Mode { PLAIN = 0, PIPE = 1, NOMUFU = 2, NOTEST = 3, PIPE_NOTEST = 4, PIPE_SPREAD = 5, SUBCHUNK = 6, SUBCHUNK1 = 7, HYB_ACC = 8, HYB_ACC_S = 9, HYB_ALLB = 10, SCALAR = 11, HYB_D2 = 12 };

template <int IPT, int MODE>
struct Stage {
    float2 dx[IPT][2], dy[IPT][2], inv[IPT][2], m[2];
    __device__ __forceinline__ void a(const float *sx, const float *sy, const float *sm, int j, const float (&nx)[IPT],
                                      const float (&ny)[IPT], const float (&thr)[IPT], bool (&cand)[IPT])
    {
        const float4 X = *reinterpret_cast<const float4 *>(&sx[j]);
        const float4 Y = *reinterpret_cast<const float4 *>(&sy[j]);
        const float4 M = *reinterpret_cast<const float4 *>(&sm[j]);
        const float2 xs[2] = {make_float2(X.x, X.y), make_float2(X.z, X.w)};
        const float2 ys[2] = {make_float2(Y.x, Y.y), make_float2(Y.z, Y.w)};
        m[0] = make_float2(M.x, M.y);
        m[1] = make_float2(M.z, M.w);
#pragma unroll
        for (int q = 0; q < IPT; ++q)
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                float2 d2;
                if (MODE == SCALAR) {
                    dx[q][u] = make_float2(xs[u].x + nx[q], xs[u].y + nx[q]);
                    dy[q][u] = make_float2(ys[u].x + ny[q], ys[u].y + ny[q]);
                    d2 = make_float2(fmaf(dx[q][u].x, dx[q][u].x, dy[q][u].x * dy[q][u].x), fmaf(dx[q][u].y, dx[q][u].y, dy[q][u].y * dy[q][u].y));
                } else if (MODE == HYB_D2) {
                    dx[q][u] = __fadd2_rn(xs[u], make_float2(nx[q], nx[q]));
                    dy[q][u] = __fadd2_rn(ys[u], make_float2(ny[q], ny[q]));
                    const float2 t = __fmul2_rn(dy[q][u], dy[q][u]);
                    d2 = make_float2(fmaf(dx[q][u].x, dx[q][u].x, t.x), fmaf(dx[q][u].y, dx[q][u].y, t.y));
                } else {
                    dx[q][u] = __fadd2_rn(xs[u], make_float2(nx[q], nx[q]));
                    dy[q][u] = __fadd2_rn(ys[u], make_float2(ny[q], ny[q]));
                    d2 = __ffma2_rn(dx[q][u], dx[q][u], __fmul2_rn(dy[q][u], dy[q][u]));
                }
                if (MODE != NOTEST && MODE != PIPE_NOTEST) {
                    cand[q] |= (d2.x <= thr[q]);
                    cand[q] |= (d2.y <= thr[q]);
                }
                if (MODE == NOMUFU)
                    inv[q][u] = __fmul2_rn(d2, make_float2(1e-9f, 1e-9f));
                else
                    inv[q][u] = make_float2(rsqrt_ftz(d2.x), rsqrt_ftz(d2.y));
            }
    }
    __device__ __forceinline__ void b(float2 (&fx)[IPT], float2 (&fy)[IPT]) const
    {
#pragma unroll
        for (int q = 0; q < IPT; ++q)
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (MODE == HYB_ACC) {             // packed products, scalar accumulates
                    const float2 s = __fmul2_rn(__fmul2_rn(inv[q][u], inv[q][u]), __fmul2_rn(inv[q][u], m[u]));
                    fx[q].x = fmaf(dx[q][u].x, s.x, fx[q].x); fx[q].y = fmaf(dx[q][u].y, s.y, fx[q].y);
                    fy[q].x = fmaf(dy[q][u].x, s.x, fy[q].x); fy[q].y = fmaf(dy[q][u].y, s.y, fy[q].y);
                } else if (MODE == HYB_ACC_S) {    // packed inv^2, scalar inv*m, s and accumulates
                    const float2 i2 = __fmul2_rn(inv[q][u], inv[q][u]);
                    const float s0 = i2.x * (inv[q][u].x * m[u].x), s1 = i2.y * (inv[q][u].y * m[u].y);
                    fx[q].x = fmaf(dx[q][u].x, s0, fx[q].x); fx[q].y = fmaf(dx[q][u].y, s1, fx[q].y);
                    fy[q].x = fmaf(dy[q][u].x, s0, fy[q].x); fy[q].y = fmaf(dy[q][u].y, s1, fy[q].y);
                } else if (MODE == HYB_ALLB || MODE == SCALAR) {   // stage B fully scalar
                    const float s0 = (inv[q][u].x * inv[q][u].x) * (inv[q][u].x * m[u].x);
                    const float s1 = (inv[q][u].y * inv[q][u].y) * (inv[q][u].y * m[u].y);
                    fx[q].x = fmaf(dx[q][u].x, s0, fx[q].x); fx[q].y = fmaf(dx[q][u].y, s1, fx[q].y);
                    fy[q].x = fmaf(dy[q][u].x, s0, fy[q].x); fy[q].y = fmaf(dy[q][u].y, s1, fy[q].y);
                } else {
                    const float2 s = __fmul2_rn(__fmul2_rn(inv[q][u], inv[q][u]), __fmul2_rn(inv[q][u], m[u]));
                    fx[q] = __ffma2_rn(dx[q][u], s, fx[q]);
                    fy[q] = __ffma2_rn(dy[q][u], s, fy[q]);
                }
            }
    }
};

template <int IPT, int MODE, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_pair(float *out, int reps)
{
    __shared__ __align__(16) float sx[TJ], sy[TJ], sm[TJ];
    for (int k = threadIdx.x; k < TJ; k += blockDim.x) {
        sx[k] = 1000.f + 37.f * k; sy[k] = -500.f + 11.f * k; sm[k] = 1e10f + k;
    }
    __syncthreads();
    float nx[IPT], ny[IPT], thr[IPT];
    float2 fx[IPT], fy[IPT];
    bool cand[IPT];
#pragma unroll
    for (int q = 0; q < IPT; ++q) {
        nx[q] = 3.f * threadIdx.x + q; ny[q] = -7.f * threadIdx.x - q; thr[q] = 1.0f + q;
        fx[q] = make_float2(0, 0); fy[q] = make_float2(0, 0); cand[q] = false;
    }
    for (int rep = 0; rep < reps; ++rep) {
        if (MODE == SUBCHUNK || MODE == SUBCHUNK1) {
            // the force kernel's structure: fresh sums per 32-body sub-chunk, folded when the pre-test is clear
            unsigned cmask = 0;
            const unsigned smask = (unsigned)(reps >> 20);     // runtime zero
#pragma unroll(MODE == SUBCHUNK ? 2 : 1)
            for (int sc = 0; sc < TJ / 32; ++sc) {
                float2 tfx[IPT], tfy[IPT];
                bool c2[IPT];
#pragma unroll
                for (int q = 0; q < IPT; ++q) { tfx[q] = make_float2(0, 0); tfy[q] = make_float2(0, 0); c2[q] = (smask >> sc) & 1u; }
#pragma unroll
                for (int j = sc * 32; j < sc * 32 + 32; j += 4) {
                    Stage<IPT, MODE> st;
                    st.a(sx, sy, sm, j, nx, ny, thr, c2);
                    st.b(tfx, tfy);
                }
#pragma unroll
                for (int q = 0; q < IPT; ++q) {
                    if (!c2[q]) { fx[q] = __fadd2_rn(fx[q], tfx[q]); fy[q] = __fadd2_rn(fy[q], tfy[q]); }
                    cmask |= (c2[q] ? 1u : 0u) << (sc * IPT + q);
                }
            }
            if (__reduce_or_sync(0xffffffffu, cmask)) cand[0] = true;
        } else if (MODE == PIPE || MODE == PIPE_NOTEST) {
            Stage<IPT, MODE> cur, nxt;
            cur.a(sx, sy, sm, 0, nx, ny, thr, cand);
#pragma unroll 8
            for (int j = 0; j < TJ; j += 4) {
                nxt.a(sx, sy, sm, (j + 4) & (TJ - 1), nx, ny, thr, cand);
                cur.b(fx, fy);
                cur = nxt;
            }
        } else {
#pragma unroll 8
            for (int j = 0; j < TJ; j += 4) {
                Stage<IPT, MODE> st;
                st.a(sx, sy, sm, j, nx, ny, thr, cand);
                st.b(fx, fy);
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < IPT; ++q) s += fx[q].x + fx[q].y + fy[q].x + fy[q].y + (cand[q] ? 1.f : 0.f);
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
    int occ = 0, regs = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_pair<IPT, MODE, THREADS, MINB>, THREADS, 0));
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, k_pair<IPT, MODE, THREADS, MINB>));
    regs = fa.numRegs;
    const int cps = occ < MINB ? occ : MINB;      // run exactly MINB CTAs per SM (or what fits)
    const int blocks = sms * cps, reps = 200;
    const double inter = (double)blocks * THREADS * IPT * TJ * reps;
    const double nameplate = (double)sms * 128 * 2 * 1.965e9;
    float ms = time_ms([&] { k_pair<IPT, MODE, THREADS, MINB><<<blocks, THREADS>>>(out, reps); }, 5);
    printf("{\"probe\": \"%s\", \"ipt\": %d, \"threads\": %d, \"ctas_per_sm\": %d, \"warps_per_sm\": %d, \"regs\": %d, \"ms\": %.4f, "
           "\"ginter_per_s\": %.1f, \"frac_20flop_nameplate\": %.3f}\n",
           name, IPT, THREADS, cps, cps * THREADS / 32, regs, ms, inter / ms * 1e-6, inter * 20 / ms * 1e3 / nameplate);
}

int main()
{
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    float *out;
    CK(cudaMalloc(&out, sizeof(float) * 1024 * 1024 * 8));
#define SWEEP(IPT, MODE, NAME)                                  \
    run<IPT, MODE, 128, 4>(NAME, out, sms);                     \
    run<IPT, MODE, 128, 6>(NAME, out, sms);                     \
    run<IPT, MODE, 128, 8>(NAME, out, sms);                     \
    run<IPT, MODE, 256, 2>(NAME, out, sms);                     \
    run<IPT, MODE, 256, 3>(NAME, out, sms);                     \
    run<IPT, MODE, 256, 4>(NAME, out, sms);
#define SW(MODE, NAME) run<2, MODE, 256, 3>(NAME, out, sms); run<2, MODE, 256, 2>(NAME, out, sms); run<4, MODE, 128, 4>(NAME, out, sms);
    SW(PLAIN, "plain")
    SW(HYB_ACC, "hyb_acc")
    SW(HYB_ACC_S, "hyb_acc_s")
    SW(HYB_ALLB, "hyb_allb")
    SW(HYB_D2, "hyb_d2")
    SW(SCALAR, "scalar")
    CK(cudaFree(out));
    return 0;
}
