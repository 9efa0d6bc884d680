// CTA-per-codeword sum-product kernel for REGULAR codes (every check of degree DC, every bit of
// degree DV; BASELINE config 4's (3,6)-regular n = 8192 code): the generic decode_block_kernel
// with everything a regular code does not need removed -- no degree table, no padded-slot
// predicates -- plus
//   * the bit's DV message addresses packed into one 8-byte table row (one LDS.64 per bit
//     instead of DV 16-bit loads); that table is staged into shared memory once per CTA by a
//     1-D TMA bulk copy and amortised over all the codewords the CTA decodes;
//   * early stop from hard-decision tags carried in bit 30 of the stored messages (see
//     decode_block_kernel): the parity of a check falls out of the sign-XOR the check update
//     computes anyway, so there is no syndrome pass and only two barriers per iteration.
// Arithmetic and operation order are spa_math.cuh's: results equal the generic kernels' bit for
// bit.  Shared memory: msg[DC*M] | r[N] | hard[N/32] | red | var table [N][4] u16.
#pragma once
#include "decode_kernels.cuh"

namespace ldpc535 {

__host__ __device__ inline size_t regular_smem_bytes(int dc, int M, int N)
{
    size_t b = 16 + 4 * (size_t)dc * M + 4 * (size_t)N + 4 * (size_t)((N + 31) >> 5) + 16;
    b = (b + 15) & ~(size_t)15;
    return b + 8 * (size_t)N;
}

// MF, NF > 0: code size fixed at compile time (1024 threads): every stride is an immediate and
// the per-thread loops (MF/1024 checks, NF/1024 bits) unroll without bounds checks.  0: run-time.
template <int DC, int DV, int MF = 0, int NF = 0>
__global__ void __launch_bounds__(1024, 1)
decode_regular_kernel(const DecodeParams p, const uint16_t *__restrict__ var_row4)
{
    static_assert(DV <= 4, "the packed bit table holds four addresses");
    static_assert(MF % 1024 == 0 && NF % 1024 == 0, "fixed sizes are whole multiples of the CTA");
    constexpr bool kFixed = MF > 0 && NF > 0;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int M = kFixed ? MF : p.M, N = kFixed ? NF : p.N;
    const int tid = threadIdx.x, nt = kFixed ? 1024 : (int)blockDim.x, lane = tid & 31;
    const int nwords = (N + 31) >> 5;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
    float *msg = reinterpret_cast<float *>(smem_raw + 16);
    float *r = msg + (size_t)DC * M;
    uint32_t *hard = reinterpret_cast<uint32_t *>(r + N);
    int *red = reinterpret_cast<int *>(hard + nwords);
    const size_t tab_off = ((size_t)16 + 4 * (size_t)DC * M + 4 * (size_t)N + 4 * (size_t)nwords + 16 + 15) & ~(size_t)15;
    const uint2 *vt = reinterpret_cast<const uint2 *>(smem_raw + tab_off);

    // stage the packed bit table (8 N bytes) once per CTA: TMA bulk copy, mbarrier completion
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        const uint32_t bytes = 8u * (uint32_t)N;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                     :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
        tma_bulk_g2s(smem_raw + tab_off, var_row4, bytes, bar);
    }
    {
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done) : "r"(smem_u32(bar)) : "memory");
        }
    }

    const bool tag = p.early_stop != 0;
    for (long long w = blockIdx.x; w < p.n_win; w += gridDim.x) {
        const long long off = p.win_offset ? p.win_offset[w] : w * (long long)N;
        const float pol = p.polarity ? (float)p.polarity[w] : 1.f;
        const bool ok = off >= 0 && off + N <= p.n_sym;
        __syncthreads();                                   // previous window fully drained
        for (int i = tid; i < N; i += nt) r[i] = ok ? __fmul_rn(-pol * kSpaScale, load_re(p, off + i)) : 0.f;
        if (tid == 0) red[0] = 0;
        __syncthreads();
        // M_ji = r_i on every edge, stored as t = copysign(2^-|M|, M), no tags yet
        for (int i = tid; i < N; i += nt) {
            const uint2 a = vt[i];
            const float t = to_check_msg(r[i]);
            msg[a.x & 0xffffu] = t;
            if (DV > 1) msg[a.x >> 16] = t;
            if (DV > 2) msg[a.y & 0xffffu] = t;
            if (DV > 3) msg[a.y >> 16] = t;
        }
        __syncthreads();

        int iters = p.max_iters;
        bool clean = false;
        for (int h = 0; h < p.max_iters; h++) {
            // ---- check nodes (+ parity of the previous iteration's decisions from the tags) ----
            uint32_t tagbad = 0;
            auto do_check = [&](int j) {
                float m[DC];
#pragma unroll
                for (int s = 0; s < DC; s++) m[s] = msg[s * M + j];
                tagbad |= check_node_spa<DC, true>(m);
#pragma unroll
                for (int s = 0; s < DC; s++) msg[s * M + j] = m[s];
            };
            if constexpr (kFixed) {
#pragma unroll
                for (int q = 0; q < MF / 1024; q++) do_check(q * 1024 + tid);
            } else {
                for (int j = tid; j < M; j += nt) do_check(j);
            }
            if (tag) {
                const int bad = __syncthreads_or((int)(tagbad & 0x40000000u));
                if (h > 0 && !bad) { iters = h; clean = true; break; }
            } else {
                __syncthreads();
            }
            // ---- variable nodes: L, hard decision, next messages (tagged with the decision) ----
            const uint32_t tagmask = tag ? 0x40000000u : 0u;
            auto do_var = [&](int i) -> bool {
                const uint2 a = vt[i];
                int idx[4] = {(int)(a.x & 0xffffu), (int)(a.x >> 16), (int)(a.y & 0xffffu), (int)(a.y >> 16)};
                float x[DV];
#pragma unroll
                for (int k = 0; k < DV; k++) x[k] = msg[idx[k]];
                const float L = var_node_spa<DV>(x, DV, r[i]);
                const bool b = (L <= 0.f);
                const uint32_t tagbit = b ? tagmask : 0u;
#pragma unroll
                for (int k = 0; k < DV; k++)
                    msg[idx[k]] = __uint_as_float(__float_as_uint(to_check_msg(x[k])) | tagbit);
                return b;
            };
            if constexpr (kFixed) {
#pragma unroll
                for (int q = 0; q < NF / 1024; q++) {
                    const int i = q * 1024 + tid;
                    const uint32_t wd = __ballot_sync(0xffffffffu, do_var(i));
                    if (lane == 0) hard[i >> 5] = wd;
                }
            } else {
                for (int base = 0; base < N; base += nt) {
                    const int i = base + tid;
                    const bool b = (i < N) ? do_var(i) : false;
                    const uint32_t wd = __ballot_sync(0xffffffffu, b);
                    if (lane == 0 && i < N) hard[i >> 5] = wd;
                }
            }
            __syncthreads();
        }

        // ---- saturating syndrome weight of the final decision (checkFrame, :236-253) ----
        int cnt = 0;
        if (!clean) {
            for (int j = tid; j < M; j += nt) {
                uint32_t pj = 0;
#pragma unroll
                for (int s = 0; s < DC; s++) {
                    const int v = __ldg(p.chk_var + s * M + j);
                    pj ^= hard[v >> 5] >> (v & 31);
                }
                cnt += (int)(pj & 1u);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
            if (lane == 0 && cnt) atomicAdd(&red[0], cnt);
            __syncthreads();
            cnt = red[0];
        }
        // ---- outputs: data bits M .. N-1, MSB first (:207-219) ----
        for (int b = tid; b < p.nbytes; b += nt) {
            uint32_t bits = 0;
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const int v = M + 8 * b + q;
                if (v < N) bits |= ((hard[v >> 5] >> (v & 31)) & 1u) << q;
            }
            p.out_bytes[w * p.nbytes + b] = ok ? pack_msb_first(bits) : (uint8_t)0;
        }
        if (tid == 0) {
            if (p.out_synd) p.out_synd[w] = ok ? (uint8_t)min(cnt, p.thr + 1) : (uint8_t)255;
            if (p.out_iters) p.out_iters[w] = ok ? (uint8_t)min(iters, 255) : (uint8_t)255;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Register-table variant (compile-time sizes only): the kernel BASELINE config 4's code runs on.
// Same arithmetic, same message layout, but the per-bit state that never changes during a launch
// lives in REGISTERS instead of shared memory: 512 threads (<= 128 registers each), a thread owns
// NF/512 bits and MF/512 checks for the whole launch and keeps the packed address row of each of
// its bits (loaded once per CTA) and the bit's intrinsic value r (loaded once per codeword) in
// registers.  Per iteration and bit this drops the LDS.64 of the table row, its unpacking and the
// LDS of r: 27.2 -> 23.2 instructions and 7.5 -> 6.5 shared-memory wavefronts per edge-iteration,
// and fewer entries in the MIO queue that LDS/STS share with MUFU (the top stall of the
// 1024-thread kernel).  Measured on B200: 8.17e11 -> 8.96e11 edge-iterations/s at fixed
// iterations.  What was tried on top and rejected is in profiles/r1_microbench.txt (two codewords
// per CTA half an iteration apart, explicit load/compute/store batching, 256 and 1024 threads,
// r or the table back in shared memory).
// Shared memory: msg[DC*M] | r[N] | hard[N/32] | red.
template <int DC, int MF, int NF>
__host__ __device__ constexpr size_t regular_rt_smem_bytes()
{
    return 4 * (size_t)DC * MF + 4 * (size_t)NF + 4 * (size_t)(NF / 32) + 16;
}

template <int DC, int DV, int MF, int NF>
__global__ void __launch_bounds__(512, 1)
decode_regular_rt_kernel(const DecodeParams p, const uint16_t *__restrict__ var_row4, const uint32_t *__restrict__ chk_color)
{
    static_assert(DV == 3, "three message addresses per bit");
    static_assert(MF % 512 == 0 && NF % 512 == 0, "fixed sizes are whole multiples of the CTA");
    constexpr int NT = 512, CQ = MF / NT, BQ = NF / NT, M = MF, N = NF;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *msg = reinterpret_cast<float *>(smem_raw);
    float *rsm = msg + (size_t)DC * M;                     // intrinsic values, word q * NT + tid = bit of this thread
    uint32_t *hard = reinterpret_cast<uint32_t *>(rsm + N);
    int *red = reinterpret_cast<int *>(hard + N / 32);
    const int tid = threadIdx.x, lane = tid & 31;

    // the address rows of this thread's bits: once per CTA, coalesced 8-byte loads
    // three 16-bit message addresses per bit, packed: vx[q] = a0 | a1 << 16, vy[q / 2] = a2 of bits q, q + 1
    uint32_t vx[BQ], vy[BQ / 2];
#pragma unroll
    for (int q = 0; q < BQ; q += 2) {
        const uint2 ta = __ldg(reinterpret_cast<const uint2 *>(var_row4) + q * NT + tid);
        const uint2 tb = __ldg(reinterpret_cast<const uint2 *>(var_row4) + (q + 1) * NT + tid);
        vx[q] = ta.x; vx[q + 1] = tb.x;
        vy[q / 2] = (ta.y & 0xffffu) | (tb.y << 16);
    }
    auto a0 = [&](int q) { return vx[q] & 0xffffu; };
    auto a1 = [&](int q) { return vx[q] >> 16; };
    auto a2 = [&](int q) { return (q & 1) ? (vy[q / 2] >> 16) : (vy[q / 2] & 0xffffu); };
    // Coloured rows (code_tables.cpp: color_regular_rows): the word of (column j, slot s) sits in j's
    // row of 32 columns at the bank its edge was coloured with, so the check phase reads a permutation
    // of a row and the variable phase 32 different banks -- no bank conflict in either phase (the
    // variable phase ran at 2.0 wavefronts per access with whole columns placed by bank).  The byte
    // offsets of this thread's 8 x 6 check words (row start + 4 x colour; the slot's s * M * 4 is an
    // immediate) are kept two per register; the intrinsic values moved to shared memory to make room.
    uint32_t ck[CQ][DC / 2];
#pragma unroll
    for (int q = 0; q < CQ; q++) {
        const uint32_t cw = __ldg(chk_color + q * NT + tid);         // slot s at bits 5s+2 .. 5s+6
        const uint32_t rowb = (uint32_t)(q * NT + (tid - lane)) * 4u;
#pragma unroll
        for (int s = 0; s < DC; s += 2)
            ck[q][s / 2] = (rowb + ((cw >> (5 * s)) & 0x7cu)) | ((rowb + ((cw >> (5 * s + 5)) & 0x7cu)) << 16);
    }
    auto chk_word = [&](int q, int s) -> float * {
        const uint32_t off = (s & 1) ? (ck[q][s / 2] >> 16) : (ck[q][s / 2] & 0xffffu);
        return reinterpret_cast<float *>(reinterpret_cast<char *>(msg + s * M) + off);
    };

    const bool tag = p.early_stop != 0;
    const uint32_t tagmask = tag ? 0x40000000u : 0u;
    // Windows are handed out by an atomic cursor, one ahead: iteration counts differ from codeword to
    // codeword under early stop, so fixed shares would leave SMs idle at the end of the batch (the
    // mean over 54 codewords per CTA still spreads by a few percent; small batches by much more).
    // red[1] carries the next window of the CTA; window k < gridDim.x goes to CTA k without an atomic.
    long long w = blockIdx.x;
    while (w < p.n_win) {
        if (tid == 0) red[1] = (int)(gridDim.x + atomicAdd(p.cursor, 1u));
        const long long off = p.win_offset ? p.win_offset[w] : w * (long long)N;
        const float pol = p.polarity ? (float)p.polarity[w] : 1.f;
        const bool ok = off >= 0 && off + N <= p.n_sym;
        float r[BQ];
#pragma unroll
        for (int q = 0; q < BQ; q++) r[q] = ok ? __fmul_rn(-pol * kSpaScale, load_re(p, off + q * NT + tid)) : 0.f;
#pragma unroll
        for (int q = 0; q < BQ; q++) rsm[q * NT + tid] = r[q];        // read back by this thread only
        __syncthreads();                                   // previous window fully drained; red[1] visible
        const long long w_next = (long long)(unsigned int)red[1];
        // pull the next window of this CTA towards L2 while this one iterates (one 128-byte line per thread)
        if (w_next < p.n_win) {
            const long long offn = p.win_offset ? p.win_offset[w_next] : w_next * (long long)N;
            if (offn >= 0 && offn + N <= p.n_sym) {
                const char *base = p.sym_re ? reinterpret_cast<const char *>(p.sym_re + offn)
                                            : reinterpret_cast<const char *>(p.sym + offn);
                const int bytes = p.sym_re ? 4 * N : 8 * N;
                if (tid * 128 < bytes) asm volatile("prefetch.global.L2 [%0];" :: "l"(base + tid * 128));
            }
        }
        if (tid == 0) red[0] = 0;
        // M_ji = r_i on every edge, stored as t = copysign(2^-|M|, M), no tags yet
#pragma unroll
        for (int q = 0; q < BQ; q++) {
            const float t = to_check_msg(r[q]);
            msg[a0(q)] = t;
            msg[a1(q)] = t;
            msg[a2(q)] = t;
        }
        __syncthreads();

        int iters = p.max_iters;
        bool clean = false;
        for (int h = 0; h < p.max_iters; h++) {
            // ---- check nodes (+ parity of the previous iteration's decisions from the tags) ----
            uint32_t tagbad = 0;
            // (the same hand pipelining as in the variable phase below: +4 % once the bank conflicts were gone)
            float mn[DC];
#pragma unroll
            for (int s = 0; s < DC; s++) mn[s] = *chk_word(0, s);
#pragma unroll
            for (int q = 0; q < CQ; q++) {
                float m[DC];
#pragma unroll
                for (int s = 0; s < DC; s++) m[s] = mn[s];
                if (q + 1 < CQ) {
#pragma unroll
                    for (int s = 0; s < DC; s++) mn[s] = *chk_word(q + 1, s);
                }
                tagbad |= check_node_spa<DC, true>(m);
#pragma unroll
                for (int s = 0; s < DC; s++) *chk_word(q, s) = m[s];
            }
            if (tag) {
                const int bad = __syncthreads_or((int)(tagbad & 0x40000000u));
                if (h > 0 && !bad) { iters = h; clean = true; break; }
            } else {
                __syncthreads();
            }
            // ---- variable nodes: L, hard decision, next messages (tagged with the decision) ----
            // Software-pipelined by hand: the messages of bit q + 1 are loaded BEFORE those of bit q are
            // stored.  The compiler cannot do it -- to it the stores may alias the next loads -- and
            // without it every bit paid the full shared-memory latency between its predecessor's stores
            // and its own loads (ncu source view, profiles/r2b: the first FADD after each load group held
            // 53 % of the variable phase's samples).  Distinct bits never share a message word.
            // HARD = false: early stop off and not the last iteration -- nobody looks at the decisions: no total L,
            // no comparison, no ballot, no tag (one instantiation of the loop for each case, picked per iteration).
            auto var_phase = [&](auto hardc) {
                constexpr bool HARD = decltype(hardc)::value;
                float xn[DV] = {msg[a0(0)], msg[a1(0)], msg[a2(0)]};
                float rn = rsm[tid];
#pragma unroll
                for (int q = 0; q < BQ; q++) {
                    const int i0 = (int)a0(q), i1 = (int)a1(q), i2 = (int)a2(q);
                    float x[DV] = {xn[0], xn[1], xn[2]};
                    const float rq = rn;
                    if (q + 1 < BQ) {            // (two bits ahead measured no better)
                        xn[0] = msg[a0(q + 1)]; xn[1] = msg[a1(q + 1)]; xn[2] = msg[a2(q + 1)];
                        rn = rsm[(q + 1) * NT + tid];
                    }
                    const float L = var_node_spa<DV>(x, DV, rq);
                    if constexpr (HARD) {
                        const bool b = (L <= 0.f);
                        const uint32_t tagbit = b ? tagmask : 0u;
                        msg[i0] = __uint_as_float(__float_as_uint(to_check_msg(x[0])) | tagbit);
                        msg[i1] = __uint_as_float(__float_as_uint(to_check_msg(x[1])) | tagbit);
                        msg[i2] = __uint_as_float(__float_as_uint(to_check_msg(x[2])) | tagbit);
                        const uint32_t wd = __ballot_sync(0xffffffffu, b);
                        if (lane == 0) hard[(q * NT + tid) >> 5] = wd;
                    } else {
                        msg[i0] = to_check_msg(x[0]);
                        msg[i1] = to_check_msg(x[1]);
                        msg[i2] = to_check_msg(x[2]);
                    }
                }
            };
            if (tag || h + 1 == p.max_iters) var_phase(std::true_type{}); else var_phase(std::false_type{});
            __syncthreads();
        }

        // ---- saturating syndrome weight of the final decision (checkFrame, :236-253) ----
        int cnt = 0;
        if (!clean) {
            for (int j = tid; j < M; j += NT) {
                uint32_t pj = 0;
#pragma unroll
                for (int s = 0; s < DC; s++) {
                    const int v = __ldg(p.chk_var + s * M + j);
                    pj ^= hard[v >> 5] >> (v & 31);
                }
                cnt += (int)(pj & 1u);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
            if (lane == 0 && cnt) atomicAdd(&red[0], cnt);
            __syncthreads();
            cnt = red[0];
        }
        // ---- outputs: data bits M .. N-1, MSB first (:207-219) ----
        for (int b = tid; b < p.nbytes; b += NT) {
            uint32_t bits = 0;
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const int v = M + 8 * b + q;
                if (v < N) bits |= ((hard[v >> 5] >> (v & 31)) & 1u) << q;
            }
            p.out_bytes[w * p.nbytes + b] = ok ? pack_msb_first(bits) : (uint8_t)0;
        }
        if (tid == 0) {
            if (p.out_synd) p.out_synd[w] = ok ? (uint8_t)min(cnt, p.thr + 1) : (uint8_t)255;
            if (p.out_iters) p.out_iters[w] = ok ? (uint8_t)min(iters, 255) : (uint8_t)255;
        }
        w = w_next;
    }
}

}  // namespace ldpc535
