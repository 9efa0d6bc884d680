// Round 2: cost of the node updates in registers under alternative formulations -- cycles per node
// per warp on one SM sub-partition, two warps per sub-partition (the c4 thread kernel's occupancy).
//
//   A  check node, current form (spa_math.cuh): 2 lg2 per edge
//   B  the same arithmetic on TWO checks at once with packed fp32 (FFMA2 / FMUL2 / FADD2)
//   C  check node in the likelihood-ratio domain: lambda_E = (ev/od)^(+-1), 1 rcp per edge
//   D  C on two checks at once, packed
//   E  variable node, current form: 1 ex2 per edge
//   F  variable node in the likelihood-ratio domain: products, ONE rcp per bit
//   G  F on two bits at once, packed
// plus the raw rate of MUFU.RCP and of FFMA2.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o node_rate2 node_rate2.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../../gr-ldpc_ece535a_b200/csrc/spa_math.cuh"
using namespace ldpc535;

typedef unsigned long long f2;
__device__ __forceinline__ f2 pack2(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(f2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { f2 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// ---- B: two checks, current arithmetic, packed ----
template <int DC>
__device__ __forceinline__ void check_node_spa_x2(float (&ma)[DC], float (&mb)[DC])
{
    f2 e[DC];
    uint32_t sxa = 0, sxb = 0;
#pragma unroll
    for (int s = 0; s < DC; s++) {
        sxa ^= __float_as_uint(ma[s]); sxb ^= __float_as_uint(mb[s]);
        e[s] = pack2(fabsf(ma[s]), fabsf(mb[s]));
    }
    const f2 one = pack2(1.f, 1.f), zero = pack2(0.f, 0.f);
    f2 pe[DC], po[DC], se[DC], so[DC];
    pe[0] = one; po[0] = zero;
    if constexpr (DC > 1) { pe[1] = one; po[1] = e[0]; }
#pragma unroll
    for (int s = 2; s < DC; s++) { pe[s] = fma2(e[s - 1], po[s - 1], pe[s - 1]); po[s] = fma2(e[s - 1], pe[s - 1], po[s - 1]); }
    se[DC - 1] = one; so[DC - 1] = zero;
    if constexpr (DC > 1) { se[DC - 2] = one; so[DC - 2] = e[DC - 1]; }
#pragma unroll
    for (int s = DC - 3; s >= 0; s--) { se[s] = fma2(e[s + 1], so[s + 1], se[s + 1]); so[s] = fma2(e[s + 1], se[s + 1], so[s + 1]); }
    const uint32_t ba = sxa & 0x80000000u, bb = sxb & 0x80000000u;
#pragma unroll
    for (int s = 0; s < DC; s++) {
        f2 ev, od;
        if (s == 0) { ev = se[0]; od = so[0]; }
        else if (s == DC - 1) { ev = pe[s]; od = po[s]; }
        else { ev = fma2(po[s], so[s], mul2(pe[s], se[s])); od = fma2(po[s], se[s], mul2(pe[s], so[s])); }
        float eva, evb, oda, odb;
        unpack2(ev, eva, evb); unpack2(od, oda, odb);
        const f2 mag = sub2(pack2(lg2_approx(eva), lg2_approx(evb)), pack2(lg2_approx(oda), lg2_approx(odb)));
        float mga, mgb;
        unpack2(mag, mga, mgb);
        ma[s] = __uint_as_float(__float_as_uint(mga) ^ ((__float_as_uint(ma[s]) & 0x80000000u) ^ ba));
        mb[s] = __uint_as_float(__float_as_uint(mgb) ^ ((__float_as_uint(mb[s]) & 0x80000000u) ^ bb));
    }
}

// ---- C: likelihood-ratio check node: in t = copysign(e, sign), out lambda_E > 0 ----
template <int DC>
__device__ __forceinline__ void check_node_lr(float (&m)[DC])
{
    float e[DC];
    uint32_t sx = 0;
#pragma unroll
    for (int s = 0; s < DC; s++) { sx ^= __float_as_uint(m[s]); e[s] = fabsf(m[s]); }
    float pe[DC], po[DC], se[DC], so[DC];
    pe[0] = 1.f; po[0] = 0.f;
    if constexpr (DC > 1) { pe[1] = 1.f; po[1] = e[0]; }
#pragma unroll
    for (int s = 2; s < DC; s++) { pe[s] = __fmaf_rn(e[s - 1], po[s - 1], pe[s - 1]); po[s] = __fmaf_rn(e[s - 1], pe[s - 1], po[s - 1]); }
    se[DC - 1] = 1.f; so[DC - 1] = 0.f;
    if constexpr (DC > 1) { se[DC - 2] = 1.f; so[DC - 2] = e[DC - 1]; }
#pragma unroll
    for (int s = DC - 3; s >= 0; s--) { se[s] = __fmaf_rn(e[s + 1], so[s + 1], se[s + 1]); so[s] = __fmaf_rn(e[s + 1], se[s + 1], so[s + 1]); }
#pragma unroll
    for (int s = 0; s < DC; s++) {
        float ev, od;
        if (s == 0) { ev = se[0]; od = so[0]; }
        else if (s == DC - 1) { ev = pe[s]; od = po[s]; }
        else { ev = __fmaf_rn(po[s], so[s], __fmul_rn(pe[s], se[s])); od = __fmaf_rn(po[s], se[s], __fmul_rn(pe[s], so[s])); }
        const bool neg = (int)(__float_as_uint(m[s]) ^ sx) < 0;
        const float num = neg ? od : ev, den = neg ? ev : od;
        m[s] = __fmul_rn(num, rcp_approx(den));
    }
}

// ---- D: two checks in the likelihood-ratio domain, packed ----
template <int DC>
__device__ __forceinline__ void check_node_lr_x2(float (&ma)[DC], float (&mb)[DC])
{
    f2 e[DC];
    uint32_t sxa = 0, sxb = 0;
#pragma unroll
    for (int s = 0; s < DC; s++) {
        sxa ^= __float_as_uint(ma[s]); sxb ^= __float_as_uint(mb[s]);
        e[s] = pack2(fabsf(ma[s]), fabsf(mb[s]));
    }
    const f2 one = pack2(1.f, 1.f), zero = pack2(0.f, 0.f);
    f2 pe[DC], po[DC], se[DC], so[DC];
    pe[0] = one; po[0] = zero;
    if constexpr (DC > 1) { pe[1] = one; po[1] = e[0]; }
#pragma unroll
    for (int s = 2; s < DC; s++) { pe[s] = fma2(e[s - 1], po[s - 1], pe[s - 1]); po[s] = fma2(e[s - 1], pe[s - 1], po[s - 1]); }
    se[DC - 1] = one; so[DC - 1] = zero;
    if constexpr (DC > 1) { se[DC - 2] = one; so[DC - 2] = e[DC - 1]; }
#pragma unroll
    for (int s = DC - 3; s >= 0; s--) { se[s] = fma2(e[s + 1], so[s + 1], se[s + 1]); so[s] = fma2(e[s + 1], se[s + 1], so[s + 1]); }
#pragma unroll
    for (int s = 0; s < DC; s++) {
        f2 ev, od;
        if (s == 0) { ev = se[0]; od = so[0]; }
        else if (s == DC - 1) { ev = pe[s]; od = po[s]; }
        else { ev = fma2(po[s], so[s], mul2(pe[s], se[s])); od = fma2(po[s], se[s], mul2(pe[s], so[s])); }
        float eva, evb, oda, odb;
        unpack2(ev, eva, evb); unpack2(od, oda, odb);
        const bool na = (int)(__float_as_uint(ma[s]) ^ sxa) < 0, nb = (int)(__float_as_uint(mb[s]) ^ sxb) < 0;
        const f2 num = pack2(na ? oda : eva, nb ? odb : evb);
        const f2 rd = pack2(rcp_approx(na ? eva : oda), rcp_approx(nb ? evb : odb));
        unpack2(mul2(num, rd), ma[s], mb[s]);
    }
}

// ---- F: likelihood-ratio variable node: in lambda_E[k], lambda_r; out t[k]; returns Lambda ----
template <int DV>
__device__ __forceinline__ float var_node_lr(float (&x)[DV], float lr)
{
    float q[DV], pre[DV], suf[DV];
#pragma unroll
    for (int k = 0; k < DV; k++) q[k] = __fmul_rn(x[k], lr);
    pre[0] = q[0];
#pragma unroll
    for (int k = 1; k < DV; k++) pre[k] = __fmul_rn(pre[k - 1], q[k]);
    suf[DV - 1] = q[DV - 1];
#pragma unroll
    for (int k = DV - 2; k >= 0; k--) suf[k] = __fmul_rn(q[k], suf[k + 1]);
    const float Lam = pre[DV - 1];
    const float R = rcp_approx(Lam);
#pragma unroll
    for (int k = 0; k < DV; k++) {
        const float out = (k == 0) ? suf[1] : (k == DV - 1) ? pre[DV - 2] : __fmul_rn(pre[k - 1], suf[k + 1]);
        const float inv = __fmul_rn(R, q[k]);
        const float e = fminf(out, inv);
        x[k] = __uint_as_float(__float_as_uint(e) | (__float_as_uint(__fsub_rn(out, inv)) & 0x80000000u));
    }
    return Lam;
}

// ---- G: two bits in the likelihood-ratio domain, packed ----
template <int DV>
__device__ __forceinline__ void var_node_lr_x2(float (&xa)[DV], float (&xb)[DV], float lra, float lrb, float &La, float &Lb)
{
    f2 q[DV], pre[DV], suf[DV];
    const f2 lr = pack2(lra, lrb);
#pragma unroll
    for (int k = 0; k < DV; k++) q[k] = mul2(pack2(xa[k], xb[k]), lr);
    pre[0] = q[0];
#pragma unroll
    for (int k = 1; k < DV; k++) pre[k] = mul2(pre[k - 1], q[k]);
    suf[DV - 1] = q[DV - 1];
#pragma unroll
    for (int k = DV - 2; k >= 0; k--) suf[k] = mul2(q[k], suf[k + 1]);
    unpack2(pre[DV - 1], La, Lb);
    const f2 R = pack2(rcp_approx(La), rcp_approx(Lb));
#pragma unroll
    for (int k = 0; k < DV; k++) {
        const f2 out = (k == 0) ? suf[1] : (k == DV - 1) ? pre[DV - 2] : mul2(pre[k - 1], suf[k + 1]);
        const f2 inv = mul2(R, q[k]);
        const f2 dif = sub2(out, inv);
        float oa, ob, ia, ib, da, db;
        unpack2(out, oa, ob); unpack2(inv, ia, ib); unpack2(dif, da, db);
        xa[k] = __uint_as_float(__float_as_uint(fminf(oa, ia)) | (__float_as_uint(da) & 0x80000000u));
        xb[k] = __uint_as_float(__float_as_uint(fminf(ob, ib)) | (__float_as_uint(db) & 0x80000000u));
    }
}

// feedback that keeps the values in range: one FFMA + one LOP3 per edge in every variant
__device__ __forceinline__ float fb_t(float v, float seed)     // anything -> t in (0.1, 0.9) with a sign
{
    const float a = fminf(fabsf(v), 1.f) * 0.5f + seed;
    return __uint_as_float(__float_as_uint(a) | (__float_as_uint(v) << 31));
}
__device__ __forceinline__ float fb_l(float v, float seed)     // t -> lambda_E in (0.5, 2.5)
{
    return fabsf(v) * 2.f + seed;
}

template <int MODE, int D, int ILP>
__global__ void __launch_bounds__(256) node(float *out, int iters, float seed)
{
    float m[ILP][D];
#pragma unroll
    for (int i = 0; i < ILP; i++)
#pragma unroll
        for (int s = 0; s < D; s++) m[i][s] = (seed * 0.1f * (s + 1) + 0.01f * i + 0.0001f * threadIdx.x) * ((s + i) & 1 ? -1.f : 1.f);
    float acc = 0.f;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i += 2) {
            if (MODE == 0) { check_node_spa<D>(m[i]); check_node_spa<D>(m[i + 1]); }
            if (MODE == 1) { check_node_spa_x2<D>(m[i], m[i + 1]); }
            if (MODE == 2) { check_node_lr<D>(m[i]); check_node_lr<D>(m[i + 1]); }
            if (MODE == 3) { check_node_lr_x2<D>(m[i], m[i + 1]); }
            if (MODE == 4) {
                acc += var_node_spa<D>(m[i], D, seed); acc += var_node_spa<D>(m[i + 1], D, seed);
#pragma unroll
                for (int s = 0; s < D; s++) { m[i][s] = to_check_msg(m[i][s]); m[i + 1][s] = to_check_msg(m[i + 1][s]); }
            }
            if (MODE == 5) { acc += var_node_lr<D>(m[i], seed + 0.5f); acc += var_node_lr<D>(m[i + 1], seed + 0.5f); }
            if (MODE == 6) { float la, lb; var_node_lr_x2<D>(m[i], m[i + 1], seed + 0.5f, seed + 0.5f, la, lb); acc += la + lb; }
#pragma unroll
            for (int s = 0; s < D; s++) {
                if (MODE <= 3) { m[i][s] = fb_t(m[i][s], seed); m[i + 1][s] = fb_t(m[i + 1][s], seed); }
                else { m[i][s] = fb_l(m[i][s], seed); m[i + 1][s] = fb_l(m[i + 1][s], seed); }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < ILP; i++)
#pragma unroll
        for (int k = 0; k < D; k++) acc += m[i][k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
__global__ void __launch_bounds__(256) raw(float *out, int iters, float seed)
{
    float a[8];
    f2 p[4];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = seed + 0.001f * (threadIdx.x + i);
#pragma unroll
    for (int i = 0; i < 4; i++) p[i] = pack2(a[2 * i], a[2 * i + 1]);
    const f2 c = pack2(seed, 0.5f);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) a[i] = rcp_approx(a[i]);
            if (MODE == 1) a[i] = __fmaf_rn(a[i], seed, 0.5f);
        }
        if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 4; i++) p[i] = fma2(p[i], c, c);
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i];
#pragma unroll
    for (int i = 0; i < 4; i++) { float x, y; unpack2(p[i], x, y); s += x + y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static double time_ms(void (*launch)(float *, int, int), float *out, int grid, int iters)
{
    launch(out, grid, 10);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    launch(out, grid, iters);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms;
}

template <int MODE, int D, int ILP>
void run(const char *name, int mufu_per_node, int threads = 256)
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float *out; cudaMalloc(&out, sizeof(float) * sms * 1024);
    const int iters = 4000;
    node<MODE, D, ILP><<<sms, threads>>>(out, 10, 0.7f);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    node<MODE, D, ILP><<<sms, threads>>>(out, iters, 0.7f);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double nodes_per_smsp = (double)(threads / 32) * iters * ILP / 4;
    const double cyc = ms * 1e-3 * khz * 1e3 / nodes_per_smsp;
    printf("%-44s degree %d ILP %d warps/SM %2d: %6.1f cycles per node per SMSP = %5.2f per edge (XU floor %4.1f per edge)\n",
           name, D, ILP, threads / 32, cyc, cyc / D, 8.0 * mufu_per_node / D);
    cudaFree(out);
}

template <int MODE>
void run_raw(const char *name, int per_iter)
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float *out; cudaMalloc(&out, sizeof(float) * sms * 8 * 256);
    const int iters = 20000;
    for (int ctas : {1, 2, 8}) {
        raw<MODE><<<sms * ctas, 256>>>(out, 100, 0.5f);
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        cudaDeviceSynchronize();
        cudaEventRecord(a);
        raw<MODE><<<sms * ctas, 256>>>(out, iters, 0.5f);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        const double warp_instr_per_smsp = (double)ctas * 8 * iters * per_iter / 4;
        printf("%-24s warps/SM %2d: %.2f cycles per warp instruction per SMSP\n", name, ctas * 8,
               ms * 1e-3 * khz * 1e3 / warp_instr_per_smsp);
    }
    cudaFree(out);
}

int main()
{
    run_raw<0>("MUFU.RCP", 8);
    run_raw<1>("FFMA", 8);
    run_raw<2>("FFMA2", 4);
    run<0, 5, 2>("A check, current (2 lg2/edge)", 10);
    run<0, 5, 4>("A check, current (2 lg2/edge)", 10);
    run<1, 5, 2>("B check x2 packed, current arithmetic", 10);
    run<1, 5, 4>("B check x2 packed, current arithmetic", 10);
    run<2, 5, 2>("C check, likelihood ratio (1 rcp/edge)", 5);
    run<2, 5, 4>("C check, likelihood ratio (1 rcp/edge)", 5);
    run<3, 5, 2>("D check x2 packed, likelihood ratio", 5);
    run<3, 5, 4>("D check x2 packed, likelihood ratio", 5);
    run<3, 6, 4>("D check x2 packed, likelihood ratio", 6);
    run<4, 3, 4>("E variable, current (1 ex2/edge)", 3);
    run<5, 3, 4>("F variable, likelihood ratio (1 rcp/bit)", 1);
    run<6, 3, 4>("G variable x2 packed, likelihood ratio", 1);
    run<6, 3, 8>("G variable x2 packed, likelihood ratio", 1);
    return 0;
}
