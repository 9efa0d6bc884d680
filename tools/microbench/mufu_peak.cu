// Measured ceiling of the SM special-function (XU / MUFU) pipe on this GPU: the denominator of
// the decoder's compute roofline.  Independent ex2/lg2 chains, many resident warps, CUDA-event
// timed.  Variants: pure MUFU; the decoder's 1 ex2 : 2 lg2 mix; MUFU with ~6 FMAs per MUFU (the
// decoder's instruction mix); the same plus shared-memory traffic (MUFU and LDS/STS share the
// MIO queue).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_peak mufu_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2(float x) { float y; asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, int iters, float seed)
{
    __shared__ float sm[256 * 8];
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = seed + 0.001f * (threadIdx.x + i);
    float *col = sm + threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) { a[i] = ex2(a[i]); }                              // 1 MUFU
            else if (MODE == 1) { a[i] = lg2(ex2(a[i])) - lg2(a[i] + 2.f); }  // 3 MUFU (1 ex2 : 2 lg2)
            else {
                float e = ex2(a[i]);
                float p = fmaf(e, a[i], 1.f), q = fmaf(e, p, a[i]);
                p = fmaf(q, e, p); q = fmaf(p, e, q); p = fmaf(q, e, p); q = fmaf(p, e, q);
                p = fmaf(q, e, p); q = fmaf(p, e, q); p = fmaf(q, e, p); q = fmaf(p, e, q);
                p = fmaf(q, e, p); q = fmaf(p, e, q); p = fmaf(q, e, p); q = fmaf(p, e, q);
                p = fmaf(q, e, p); q = fmaf(p, e, q); p = fmaf(q, e, p); q = fmaf(p, e, q);
                a[i] = lg2(p) - lg2(q);                                       // 3 MUFU + 18 FMA-pipe ops
                if (MODE == 3) { col[i * 256] = a[i]; }
            }
        }
        if (MODE == 3) {
            asm volatile("" ::: "memory");
#pragma unroll
            for (int i = 0; i < 8; i++) a[i] += col[((i + 1) & 7) * 256];     // 1 STS + 1 LDS per 3 MUFU
            asm volatile("" ::: "memory");
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name, int mufu_per_inner, int ctas_per_sm)
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float *out; cudaMalloc(&out, sizeof(float) * sms * 8 * 256);
    const int iters = 20000, grid = sms * ctas_per_sm;
    k<MODE><<<grid, 256>>>(out, 100, 0.5f);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    k<MODE><<<grid, 256>>>(out, iters, 0.5f);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double mufu = (double)grid * 256 * iters * 8 * mufu_per_inner;
    const double per_clk_sm = mufu / (ms * 1e-3) / sms / (khz * 1e3);
    printf("%-34s warps/SM=%2d  %.3f ms  %.3e MUFU/s  %.2f MUFU lanes/clk/SM (at %d MHz max clock)\n", name,
           ctas_per_sm * 8, ms, mufu / (ms * 1e-3), per_clk_sm, khz / 1000);
    cudaFree(out);
}

int main()
{
    for (int c : {1, 2, 4, 8}) {
        run<0>("pure ex2", 1, c);
        run<1>("1 ex2 : 2 lg2", 3, c);
        run<2>("3 MUFU + 18 FMA", 3, c);
        run<3>("3 MUFU + 18 FMA + LDS/STS", 3, c);
    }
    return 0;
}
