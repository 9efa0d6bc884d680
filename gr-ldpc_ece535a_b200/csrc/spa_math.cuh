// Sum-product node updates shared by every decoder kernel (sm_100a).
//
// Reference arithmetic (lib/ldpc_decoder_cb_impl.cc:503-553, fp64):
//     E_ji = log((1 + T) / (1 - T)),  T = prod_{k != i} tanh(M_jk / 2)
//     L_i  = sum_j (E_ji + r_i)           (the intrinsic term is added once PER CHECK)
//     M_ji = sum_{k != j} (E_ki + r_i)
//
// fp32 formulation used here.  With e_k = exp(-|M_jk|):  tanh(|M_jk|/2) = (1-e_k)/(1+e_k).
// Expand prod_{k != i} (1 + e_k x) = even_i + odd_i x  (x^2 = 1; even/odd elementary
// symmetric sums of the e_k).  Then  prod(1+e_k) = even+odd,  prod(1-e_k) = even-odd and
//     (1 + |T|) / (1 - |T|) = even_i / odd_i         =>  |E_ji| = ln(even_i) - ln(odd_i).
// Every term of even/odd is non-negative, so there is no cancellation anywhere: the
// naive fp32 `1 - T` loses all accuracy once |E| > ~11, this form keeps ~1e-6 relative
// accuracy until exp(-|M|) underflows (|M| > 87).  The "all but one" products come from
// prefix/suffix recurrences (2 FMA per step), not from dividing a total (unstable when
// the excluded edge is the weakest).  Cost per edge and iteration: 3 MUFU (1 ex2 +
// 2 lg2) instead of the 4 (tanh/exp, rcp, div, log) of the direct form.
// sign(T) = XOR of the sign bits of the other inputs.  M = 0 gives e = 1, even = odd,
// E = 0, like tanh(0) = 0 does in the reference.  A padded slot holds e = 0 (M = +inf),
// the identity of the recurrence.
//
// Where the exponential is taken: a bit->check message is STORED as t = copysign(exp(-|M|), M)
// (to_check_msg), i.e. the ex2 runs in the variable phase, right where M is produced, and the
// check phase starts from t.  Same operations, same values -- but the MUFU work is spread over
// both phases (2 lg2 per edge in the check phase, 1 ex2 in the variable phase) instead of
// leaving the XU pipe idle for the whole variable phase (measured +20 % on B200, profiles/).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ldpc535 {

__device__ __forceinline__ float ex2_approx(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_approx(float x)
{
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// The sum-product kernels hold every message, intrinsic value and sum in units of ln 2
// ("bits"): x' = x * log2(e).  exp(-|M|) is then ex2(-|M'|) with the negate/abs folded into the
// MUFU operand, and |E'| = lg2(even) - lg2(odd) needs no multiply by ln 2 -- two FMULs per
// edge and iteration less than in natural units.  The variable update is linear, decisions
// only look at signs, so nothing else changes; message dumps are scaled back by kSpaUnscale.
constexpr float kSpaScale = 1.4426950408889634f;     // log2(e)
constexpr float kSpaUnscale = 0.6931471805599453f;   // ln 2

// bit->check message M (units of ln 2) -> stored form t = copysign(2^-|M|, M)
__device__ __forceinline__ float to_check_msg(float M)
{
    const float e = ex2_approx(-fabsf(M));
    return __uint_as_float(__float_as_uint(e) | (__float_as_uint(M) & 0x80000000u));
}

// In: m[s] = stored bit->check messages t of one check (padded slots +0).  Out: m[s] = E.
// Returns the XOR of the input bit patterns.  TAGGED: bit 30 of every input carries the hard
// decision of its bit (|t| <= 1 leaves that bit free); it is masked off here, and bit 30 of the
// returned word is the parity of the check -- the syndrome test costs one AND per edge.
template <int DC, bool TAGGED = false>
__device__ __forceinline__ uint32_t check_node_spa(float (&m)[DC])
{
    float e[DC];
    uint32_t sx = 0;
#pragma unroll
    for (int s = 0; s < DC; s++) {
        sx ^= __float_as_uint(m[s]);
        e[s] = TAGGED ? __uint_as_float(__float_as_uint(m[s]) & 0x3fffffffu) : fabsf(m[s]);
    }
    // prefix (pe, po)[s] = expansion over slots < s ; suffix (se, so)[s] over slots > s.
    // The first step of each recurrence and the two end combinations below are written out
    // (fma(e, 0, 1) = 1, fma(e, 1, 0) = e: exact), so a check evaluated with its exact degree
    // and the same check padded with e = 0 slots give the same bits.
    //
    // Every operation is spelled with a round-to-nearest intrinsic (__fmaf_rn / __fmul_rn / __fsub_rn),
    // which the compiler may neither fuse nor split.  With plain operators it did: in an instantiation
    // whose degree is a compile-time constant it knows se[s] == 1 for the last-but-one slot, rewrites
    // fma(po, 1, pe * so) as po + pe * so and contracts that into ONE fma(pe, so, po) -- a single
    // rounding where the padded instantiation rounds twice.  One ulp in od there moved a message by
    // 5e-6 relative and, on 1 frame in 10^7 at 2 dB, a hard decision (found by diffing the kernel
    // families on 10 M frames, tools/diff_kernels.py).
    float pe[DC], po[DC], se[DC], so[DC];
    pe[0] = 1.f; po[0] = 0.f;
    if constexpr (DC > 1) { pe[1] = 1.f; po[1] = e[0]; }
#pragma unroll
    for (int s = 2; s < DC; s++) {
        pe[s] = __fmaf_rn(e[s - 1], po[s - 1], pe[s - 1]);
        po[s] = __fmaf_rn(e[s - 1], pe[s - 1], po[s - 1]);
    }
    se[DC - 1] = 1.f; so[DC - 1] = 0.f;
    if constexpr (DC > 1) { se[DC - 2] = 1.f; so[DC - 2] = e[DC - 1]; }
#pragma unroll
    for (int s = DC - 3; s >= 0; s--) {
        se[s] = __fmaf_rn(e[s + 1], so[s + 1], se[s + 1]);
        so[s] = __fmaf_rn(e[s + 1], se[s + 1], so[s + 1]);
    }
    const uint32_t base = sx & 0x80000000u;                 // sign of the product over all inputs
#pragma unroll
    for (int s = 0; s < DC; s++) {
        float ev, od;
        if (s == 0) { ev = se[0]; od = so[0]; }
        else if (s == DC - 1) { ev = pe[s]; od = po[s]; }
        else {
            // pe[1] == 1 and se[DC - 2] == 1: x * 1 is x, exactly -- but the rn intrinsic keeps the multiply
            // (3 FMULs by one per check in the SASS), so those products are written out as their value
            const float pese = (s == 1) ? se[s] : (s == DC - 2) ? pe[s] : __fmul_rn(pe[s], se[s]);
            const float peso = (s == 1) ? so[s] : __fmul_rn(pe[s], so[s]);
            ev = __fmaf_rn(po[s], so[s], pese);
            od = __fmaf_rn(po[s], se[s], peso);
        }
        const float mag = __fsub_rn(lg2_approx(ev), lg2_approx(od));
        // mag >= 0 (even >= odd): attach sign(all inputs) ^ sign(own input)
        m[s] = __uint_as_float(__float_as_uint(mag) ^ ((__float_as_uint(m[s]) & 0x80000000u) ^ base));
    }
    return sx;
}

// Min-sum check update (decodeLogDomainSimple, lib/ldpc_decoder_cb_impl.cc:349-376):
//     Lr_ji = (prod_k sign(Lq_jk)) * sign(Lq_ji) * min_{k != i} |Lq_jk|,  sign(0) = 0.
// Padded slots hold +inf (sign +1, never the minimum).
//
// Min-sum runs in fp64 (T = double) like the reference: its only operations are +, -, min and
// sign, all exactly rounded in IEEE arithmetic and kept in the reference's order, so the
// messages -- not just the decisions -- equal the reference's bit for bit at any iteration
// count.  (An fp32 min-sum drifts on non-converging frames after a few dozen iterations.)
//
// How it is evaluated (round 2; the first version tracked the two smallest magnitudes and their
// position and multiplied by an integer sign: 150 of the warp kernel's 263 instructions per
// iteration, and that kernel is bound by instruction issue).  "All but one" minima come from
// prefix / suffix minima, like the products of the sum-product update: 3 (DC - 2) two-way minima
// and no position bookkeeping.  Signs are handled on the high words: the sign of the product of
// the OTHER inputs is the XOR of all sign bits and the own one.  sign(0) = 0 makes the reference's
// product over ALL inputs vanish as soon as one input is zero, and then every output of the check
// is zero: one test of the overall minimum, applied as a mask.  Values are exactly the reference's
// (a minimum of the same doubles, a sign flip); a zero result is +0 where the reference computes
// 0 * min = +0 as well.
template <int DC, typename T>
__device__ __forceinline__ void check_node_minsum(T (&m)[DC])
{
    static_assert(DC >= 2, "needs two slots");
    if constexpr (sizeof(T) == 8) {
        double a[DC];
        uint32_t sx = 0;
#pragma unroll
        for (int s = 0; s < DC; s++) {
            const uint32_t hi = (uint32_t)__double2hiint((double)m[s]);
            sx ^= hi;
            a[s] = __hiloint2double((int)(hi & 0x7fffffffu), __double2loint((double)m[s]));
        }
        auto mn2 = [](double x, double y) { return x < y ? x : y; };
        double pre[DC], suf[DC];            // pre[s] = min over slots < s, suf[s] = min over slots > s
        pre[1] = a[0];
#pragma unroll
        for (int s = 2; s < DC; s++) pre[s] = mn2(pre[s - 1], a[s - 1]);
        suf[DC - 2] = a[DC - 1];
#pragma unroll
        for (int s = DC - 3; s >= 0; s--) suf[s] = mn2(suf[s + 1], a[s + 1]);
        const bool anyzero = mn2(pre[DC - 1], a[DC - 1]) == 0.0;
        const uint32_t zmask = anyzero ? 0u : 0xffffffffu, smask = anyzero ? 0u : 0x80000000u;
#pragma unroll
        for (int s = 0; s < DC; s++) {
            const double mn = s == 0 ? suf[0] : s == DC - 1 ? pre[DC - 1] : mn2(pre[s], suf[s]);
            const uint32_t hi = ((uint32_t)__double2hiint(mn) & zmask) | ((sx ^ (uint32_t)__double2hiint((double)m[s])) & smask);
            m[s] = (T)__hiloint2double((int)hi, (int)((uint32_t)__double2loint(mn) & zmask));
        }
    } else {
        T min1 = (T)__int_as_float(0x7f800000), min2 = min1;   // two smallest magnitudes
        int arg1 = -1;
        int prod = 1;
#pragma unroll
        for (int s = 0; s < DC; s++) {
            const T a = m[s] < (T)0 ? -m[s] : m[s];
            prod *= (m[s] > (T)0) - (m[s] < (T)0);
            if (a < min1) { min2 = min1; min1 = a; arg1 = s; }
            else if (a < min2) { min2 = a; }
        }
#pragma unroll
        for (int s = 0; s < DC; s++) {
            const int sg = prod * ((m[s] > (T)0) - (m[s] < (T)0));
            const T mn = (s == arg1) ? min2 : min1;
            m[s] = (T)sg * mn;
        }
    }
}

// Variable update.  q_k = E_k + r for the dv real edges (checks ascending), 0 beyond.
//   L   = q_0 + q_1 + ...            (lib/ldpc_decoder_cb_impl.cc:519-525)
//   M_k = sum_{p != k} q_p            (:540-553)
// as prefix + suffix sums; for dv <= 3 this is the reference's own order of additions.
// Same update with a per-edge intrinsic value rk[k] (r on a real edge, 0 on an unused slot whose
// x[k] reads as 0): no degree predicates -- an unused slot contributes q = 0 like above and its
// outgoing message goes to a dummy location.
template <int DV>
__device__ __forceinline__ float var_node_spa_rk(float (&x)[DV], const float (&rk)[DV])
{
    float q[DV], pre[DV], suf[DV];
#pragma unroll
    for (int k = 0; k < DV; k++) q[k] = __fadd_rn(x[k], rk[k]);      // never fused with the product that made r
    pre[0] = q[0];
#pragma unroll
    for (int k = 1; k < DV; k++) pre[k] = __fadd_rn(pre[k - 1], q[k]);
    suf[DV - 1] = q[DV - 1];
#pragma unroll
    for (int k = DV - 2; k >= 0; k--) suf[k] = __fadd_rn(q[k], suf[k + 1]);
    if constexpr (DV == 1) {
        x[0] = 0.f;
    } else {
        x[0] = suf[1];
        x[DV - 1] = pre[DV - 2];
#pragma unroll
        for (int k = 1; k < DV - 1; k++) x[k] = __fadd_rn(pre[k - 1], suf[k + 1]);
    }
    return pre[DV - 1];
}

template <int DV>
__device__ __forceinline__ float var_node_spa(float (&x)[DV], int dv, float r)
{
    float q[DV], pre[DV], suf[DV];     // pre[k] = q_0+..+q_k ; suf[k] = q_k+..+q_{DV-1}
#pragma unroll
    for (int k = 0; k < DV; k++) q[k] = (k < dv) ? __fadd_rn(x[k], r) : 0.f;
    pre[0] = q[0];
#pragma unroll
    for (int k = 1; k < DV; k++) pre[k] = __fadd_rn(pre[k - 1], q[k]);
    suf[DV - 1] = q[DV - 1];
#pragma unroll
    for (int k = DV - 2; k >= 0; k--) suf[k] = __fadd_rn(q[k], suf[k + 1]);
    if constexpr (DV == 1) {
        x[0] = 0.f;
    } else {
        x[0] = suf[1];
        x[DV - 1] = pre[DV - 2];
#pragma unroll
        for (int k = 1; k < DV - 1; k++) x[k] = __fadd_rn(pre[k - 1], suf[k + 1]);
    }
    return pre[DV - 1];
}

// Min-sum variable update (lib/ldpc_decoder_cb_impl.cc:378-403):
//   sum = Lr_0 + Lr_1 + ... ;  Lq_k = (Lc + sum) - Lr_k ;  LQ = Lc + sum.
template <int DV, typename T>
__device__ __forceinline__ T var_node_minsum(T (&x)[DV], int dv, T lc)
{
    T sum = (T)0;
#pragma unroll
    for (int k = 0; k < DV; k++) sum += (k < dv) ? x[k] : (T)0;
    const T LQ = lc + sum;
#pragma unroll
    for (int k = 0; k < DV; k++) x[k] = LQ - x[k];
    return LQ;
}

}  // namespace ldpc535
