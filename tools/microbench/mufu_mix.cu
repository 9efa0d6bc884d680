// How well do MUFU (XU pipe) and FFMA (FMA pipe) overlap?  Per inner step: 1 MUFU + NF
// independent FFMAs, 8 independent streams per thread.  Reports cycles per inner step per SMSP.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int NF, bool USE_MUFU>
__global__ void __launch_bounds__(256) k(float *out, int iters, float seed)
{
    float a[8], b[8][NF > 0 ? NF : 1];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        a[i] = seed + 0.001f * (threadIdx.x + i);
#pragma unroll
        for (int j = 0; j < (NF > 0 ? NF : 1); j++) b[i][j] = seed * (j + 1) + i;
    }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (USE_MUFU) a[i] = ex2(a[i]);
#pragma unroll
            for (int j = 0; j < NF; j++) b[i][j] = fmaf(b[i][j], 0.999f, 0.001f);
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        s += a[i];
#pragma unroll
        for (int j = 0; j < (NF > 0 ? NF : 1); j++) s += b[i][j];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NF, bool M>
void run(int ctas)
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float *out; cudaMalloc(&out, sizeof(float) * sms * 8 * 256);
    const int iters = 20000, grid = sms * ctas;
    k<NF, M><<<grid, 256>>>(out, 100, 0.5f);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    k<NF, M><<<grid, 256>>>(out, iters, 0.5f);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double inner_per_smsp = (double)ctas * 8 * iters * 8 / 4;     // warp-level inner steps per SMSP
    printf("mufu=%d  FFMA per step=%2d  warps/SM=%2d: %.2f cycles per step per SMSP\n", (int)M, NF, ctas * 8,
           ms * 1e-3 * khz * 1e3 / inner_per_smsp);
    cudaFree(out);
}

int main()
{
    for (int c : {2, 8}) {
        run<0, true>(c); run<2, true>(c); run<4, true>(c); run<6, true>(c); run<8, true>(c);
        run<12, true>(c); run<16, true>(c); run<24, true>(c);
        run<8, false>(c); run<16, false>(c);
    }
    return 0;
}
