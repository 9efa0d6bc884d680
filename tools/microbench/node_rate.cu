// Throughput of the decoder's node updates in isolation (registers only): cycles per check node
// / per variable node per warp on one SM sub-partition, against the XU-pipe floor (15 MUFU x 8
// cycles = 120 cycles for a degree-5 check).
#include <cstdio>
#include <cuda_runtime.h>
#include "../../gr-ldpc_ece535a_b200/csrc/spa_math.cuh"
using namespace ldpc535;

template <int D, int ILP>
__global__ void __launch_bounds__(256) chk(float *out, int iters, float seed)
{
    float m[ILP][D];
#pragma unroll
    for (int i = 0; i < ILP; i++)
#pragma unroll
        for (int s = 0; s < D; s++) m[i][s] = seed * (s + 1) - 0.37f * i + 0.001f * threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            check_node_spa<D>(m[i]);
#pragma unroll
            for (int s = 0; s < D; s++) m[i][s] = m[i][s] * 0.5f + seed;   // keep values in range (1 FFMA per edge)
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++)
#pragma unroll
        for (int k = 0; k < D; k++) s += m[i][k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int D, int ILP>
void run(int ctas, int threads = 256)
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float *out; cudaMalloc(&out, sizeof(float) * sms * 8 * 256);
    const int iters = 4000, grid = sms * ctas;
    chk<D, ILP><<<grid, threads>>>(out, 10, 0.7f);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    chk<D, ILP><<<grid, threads>>>(out, iters, 0.7f);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double nodes_per_smsp = (double)ctas * (threads / 32) * iters * ILP / 4;
    const double cyc = ms * 1e-3 * khz * 1e3 / nodes_per_smsp;
    printf("check degree %d  ILP %d  warps/SM %2d: %.1f cycles per check per SMSP (XU floor %d) -> %.2f edge-it/clk/SM\n",
           D, ILP, ctas * threads / 32, cyc, 3 * D * 8, 4.0 * 32 * D / cyc);
    cudaFree(out);
}

int main()
{
    run<5, 1>(1, 128); run<5, 2>(1, 128); run<5, 3>(1, 128); run<5, 4>(1, 128); run<5, 6>(1, 128); run<5, 8>(1, 128);
    run<5, 1>(1); run<5, 2>(1); run<5, 4>(1);
    run<5, 2>(1, 384); run<5, 4>(1, 384);
    return 0;
}
