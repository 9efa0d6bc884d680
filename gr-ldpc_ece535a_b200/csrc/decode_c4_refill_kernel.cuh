// Thread-per-codeword sum-product kernel for the shipped 32x64 code WITH EARLY STOP: the kernel the
// reference block's own operating point runs on (5 iterations, exit on a zero syndrome,
// lib/ldpc_decoder_cb_impl.cc:39-40, :535-537).
//
// decode_c4_thread_kernel (decode_c4_kernel.cuh) starts 256 codewords together and can only stop
// them together, so early stop buys it nothing.  Here every thread is a persistent decoder slot
// with its own iteration counter:
//   * a codeword that passes its syndrome test (or reaches max_iters) is finished at once: its
//     outputs are parked in registers and written when the slot is refilled;
//   * a slot is refilled from an atomic window cursor -- warp-aggregated, one atomicAdd per warp
//     and refill -- when at least `refill_min` lanes of its warp are waiting (or none is active):
//     the refill path is divergent code that costs the warp the same time for 1 lane as for 32,
//     so it is taken when enough lanes share it (2 dB: most lanes need all 5 iterations and the
//     warp refills once per 5; 6 dB: most lanes finish after 2 and refills run every iteration);
//   * the NEXT window of every slot is already in shared memory when the slot needs it: right after a
//     slot is refilled the warp fetches the 64 real parts of the window after that with cp.async,
//     cooperatively (32 consecutive symbols per instruction), one whole decode ahead, so the refill
//     path never waits for HBM.
// The iteration itself -- compile-time Tanner graph, 100 messages in registers and 68 in a private
// shared-memory column, CTA barriers every few nodes to keep the warps inside one instruction-cache
// window -- is decode_c4_thread_kernel's, built from the same spa_math.cuh functions in the same
// order, so decisions, iteration counts and syndrome weights equal every other kernel family's.
// Shared memory per CTA: 68 message columns + per warp two (current / next) 64 x 33-word blocks of
// intrinsic values = 200 KB.
#pragma once
#include "decode_c4_kernel.cuh"

namespace ldpc535 {
namespace c4 {

constexpr int kRfThreads = 256;
// intrinsic values of one warp's 32 slots: [buffer][i][lane] with a row stride of 33 words, so that
// both access patterns are bank-conflict free -- the iteration reads row i across lanes (banks
// i + lane), the cooperative fetch writes column j down the rows (banks 33 i + j = i + j)
constexpr int kRfRow = 33;
constexpr int kRfWarpWords = 2 * kN * kRfRow;
constexpr size_t kRfSmemBytes = ((size_t)kRfThreads * kSmemEdges + (size_t)(kRfThreads / 32) * kRfWarpWords) * sizeof(float);

__global__ void __launch_bounds__(kRfThreads, 1)
decode_c4_refill_kernel(const DecodeParams p, unsigned int *__restrict__ cursor, const int refill_min)
{
    extern __shared__ float c4_smem[];
    constexpr int NT = kRfThreads;
    const int lane = threadIdx.x & 31;
    float *ms = c4_smem + threadIdx.x;                    // message e at ms[e * NT], e < kSmemEdges
    float *rw = c4_smem + kSmemEdges * NT + (threadIdx.x >> 5) * kRfWarpWords;   // this warp's r block
    float mreg[kRegEdges];
    auto msg_ld = [&](auto ec) -> float {
        constexpr int e = decltype(ec)::value;
        if constexpr (e < kSmemEdges) return ms[e * NT]; else return mreg[e - kSmemEdges];
    };
    auto msg_st = [&](auto ec, float v) {
        constexpr int e = decltype(ec)::value;
        if constexpr (e < kSmemEdges) ms[e * NT] = v; else mreg[e - kSmemEdges] = v;
    };

    // ---- slot state ----
    long long w = -1;                 // window being decoded
    long long wn = -1;                // next window: its symbols are being fetched into the idle r buffer
    bool ok = false, okn = false;     // their offsets are inside the symbol buffer
    int cur = 0;                      // which r buffer holds the current window
    int h = 0;                        // iterations done on the current window
    bool waiting = true;              // slot has no running decode (finished or never started)
    unsigned out_h1 = 0;              // parked outputs of a finished decode
    int out_cnt = 0, out_iters = 0;
    bool have_out = false;
    long long w_out = -1;
    bool ok_out = false;
    // the window after the next one is claimed one refill AHEAD: the atomicAdd is issued in one refill
    // and its result is first read in the next, so no warp ever waits for the L2 round trip
    unsigned pend_base = 0, pend_mask = 0;   // claim in flight: leader lane holds the cursor value
    long long parked = -1;                   // a claimed window this slot could not use yet
    bool has_parked = false;

    auto claim_issue = [&](bool want) {
        pend_mask = __ballot_sync(0xffffffffu, want);
        if (pend_mask && lane == (__ffs(pend_mask) - 1)) pend_base = atomicAdd(cursor, (unsigned)__popc(pend_mask));
    };
    auto claim_take = [&]() -> long long {          // the window this lane claimed in claim_issue, or -1
        if (pend_mask == 0) return -1;
        const unsigned base = __shfl_sync(0xffffffffu, pend_base, __ffs(pend_mask) - 1);
        const bool mine = (pend_mask >> lane) & 1u;
        const long long idx = (long long)base + __popc(pend_mask & ((1u << lane) - 1u));
        return (mine && idx < p.n_win) ? idx : -1;
    };
    // Start fetching the real parts of window wn of every slot that asks, into that slot's idle r
    // buffer.  Cooperative: for slot j the whole warp issues two 4-byte cp.async copies per lane (rows
    // lane and lane + 32 of column j), so the global side of an instruction is 32 consecutive symbols
    // (256 contiguous bytes of gr_complex, 128 of packed reals) instead of 32 different codewords.
    // (symbol offsets fit 32 bits here: the launcher sends longer buffers to the other kernels)
    const int elt = p.sym_re ? 1 : 2;               // floats per symbol
    const float *gsrc = (p.sym_re ? p.sym_re : reinterpret_cast<const float *>(p.sym)) + lane * elt;
    float *const rdst = rw + lane * kRfRow;
    auto prefetch_next = [&](bool ask) {
        bool good = false;
        long long off = 0;
        if (ask && wn >= 0) {
            off = p.win_offset ? p.win_offset[wn] : wn * (long long)kN;
            good = off >= 0 && off + kN <= p.n_sym;
        }
        if (ask) okn = good;
        unsigned m = __ballot_sync(0xffffffffu, good);
        const unsigned bufs = __ballot_sync(0xffffffffu, (cur ^ 1) != 0);     // idle buffer of every lane
        const unsigned lo = (unsigned)off;
        while (m) {
            const int j = __ffs(m) - 1;
            m &= m - 1;
            const unsigned oj = __shfl_sync(0xffffffffu, lo, j);
            float *dst = rdst + ((bufs >> j) & 1u) * (kN * kRfRow) + j;
            const float *s = gsrc + (size_t)oj * elt;
            cp_async_4(dst, s);
            cp_async_4(dst + 32 * kRfRow, s + 32 * elt);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    // ---- Step 1: check messages of all checks, in place (:503-516) ----
    auto check_all = [&]() {
        static_for<kM>([&](auto jc) {
            constexpr int j = decltype(jc)::value;
            constexpr int a = kT.row_ptr[j];
            constexpr int d = kT.row_ptr[j + 1] - a;
            float m[d];
            static_for<d>([&](auto sc) {
                constexpr int s = decltype(sc)::value;
                m[s] = msg_ld(std::integral_constant<int, a + s>{});
            });
            check_node_spa<d>(m);
            static_for<d>([&](auto sc) {
                constexpr int s = decltype(sc)::value;
                msg_st(std::integral_constant<int, a + s>{}, m[s]);
            });
            if constexpr (j % kChkGroup == kChkGroup - 1) __syncthreads();
        });
    };
    // ---- Test (:519-532) fused with Step 2 (:540-553); returns the syndrome weight (:535) ----
    auto var_all = [&](const float *rcur, unsigned &h1_out) -> int {
        unsigned h0 = 0, h1 = 0;
        static_for<kN>([&](auto ic) {
            constexpr int i = decltype(ic)::value;
            constexpr int a = kT.col_ptr[i];
            constexpr int dv = kT.col_ptr[i + 1] - a;
            float L = 0.f;
            if constexpr (dv > 0) {
                float x[dv];
                static_for<dv>([&](auto kc) {
                    constexpr int k = decltype(kc)::value;
                    constexpr int e = kT.edge_of_col[a + k];
                    x[k] = msg_ld(std::integral_constant<int, e>{});
                });
                L = var_node_spa<dv>(x, dv, rcur[i * kRfRow]);
                static_for<dv>([&](auto kc) {
                    constexpr int k = decltype(kc)::value;
                    constexpr int e = kT.edge_of_col[a + k];
                    msg_st(std::integral_constant<int, e>{}, to_check_msg(x[k]));
                });
            }
            const unsigned bit = (L <= 0.f) ? 1u : 0u;          // :527
            if constexpr (i < 32) h0 |= bit << i; else h1 |= bit << (i - 32);
            if constexpr (i % kVarGroup == kVarGroup - 1) __syncthreads();
        });
        int cnt = 0;
        static_for<kM>([&](auto jc) {
            constexpr int j = decltype(jc)::value;
            constexpr unsigned lo = kT.row_lo[j], hi = kT.row_hi[j];
            cnt += __popc((h0 & lo) ^ (h1 & hi)) & 1;
        });
        h1_out = h1;
        return cnt;
    };

    // every slot starts empty with its first window on the way (into buffer cur ^ 1) and the second claimed
    claim_issue(true);
    wn = claim_take();
    prefetch_next(true);
    claim_issue(wn >= 0);

    for (;;) {
        // ---- refill: finished / empty slots take their next window when enough lanes share the cost ----
        {
            // a waiting slot has something to do here if it holds parked outputs or a window to start;
            // a slot whose claims have run dry has neither and never asks again
            const unsigned want = __ballot_sync(0xffffffffu, waiting && (have_out || wn >= 0));
            const unsigned busy = __ballot_sync(0xffffffffu, !waiting);
            if (want && (__popc(want) >= refill_min || busy == 0)) {
                if (waiting && have_out) {
                    // outputs of the decode that finished in this slot: bits 32..63 MSB first
                    // (:207-219), checkFrame weight (:236-253), iteration count
                    if (ok_out) {
                        *reinterpret_cast<unsigned *>(p.out_bytes + w_out * 4) = __byte_perm(__brev(out_h1), 0, 0x0123);
                        if (p.out_synd) p.out_synd[w_out] = (uint8_t)min(out_cnt, p.thr + 1);
                        if (p.out_iters) p.out_iters[w_out] = (uint8_t)min(out_iters, 255);
                    } else {
                        *reinterpret_cast<unsigned *>(p.out_bytes + w_out * 4) = 0u;
                        if (p.out_synd) p.out_synd[w_out] = 255;
                        if (p.out_iters) p.out_iters[w_out] = 255;
                    }
                    have_out = false;
                }
                // Slots that start a window: the fetched window becomes current, the window claimed at the
                // previous refill becomes the next one (its fetch starts below) and a new claim is issued
                // for the refill after this.  Only slots that start a window hold a claim, and a slot whose
                // claim came back empty stops asking: the cursor never runs more than two claims per slot
                // past n_win.  The warp holds ONE claim word in flight, for the lanes that started at the
                // previous refill; it is taken here by all of them, and a lane that does not start now parks
                // its window until it does.
                const bool start = waiting && wn >= 0;
                const long long claimed = claim_take();     // collective
                if (((pend_mask >> lane) & 1u) && !start) { parked = claimed; has_parked = true; }
                long long nxt = -1;
                if (start) {
                    w = wn; ok = okn; cur ^= 1;
                    if (has_parked) { nxt = parked; has_parked = false; }
                    else if ((pend_mask >> lane) & 1u) nxt = claimed;
                    wn = nxt;
                }
                // the symbols of the windows that start now were requested a whole decode ago; every lane
                // waits for its share of those copies, the warp barrier makes all shares visible
                asm volatile("cp.async.wait_group %0;" :: "n"(0) : "memory");
                __syncwarp();
                prefetch_next(start);
                claim_issue(start && wn >= 0);
                if (start) {
                    // r_i = -pol * Re(sym_i) in units of ln 2 (:149-153, :486); M_ji = r_i on every
                    // edge (:489-496), stored as t = copysign(2^-|M|, M)
                    float *rc = rw + cur * (kN * kRfRow) + lane;
                    const float npol = (p.polarity ? -(float)p.polarity[w] : -1.f) * kSpaScale;
                    float t[kN];
#pragma unroll
                    for (int i = 0; i < kN; i++) {
                        const float ri = ok ? __fmul_rn(npol, rc[i * kRfRow]) : 0.f;
                        rc[i * kRfRow] = ri;
                        t[i] = to_check_msg(ri);
                    }
                    static_for<kE>([&](auto ec) {
                        constexpr int e = decltype(ec)::value;
                        constexpr int c = kT.col_idx[e];
                        msg_st(ec, t[c]);
                    });
                    h = 0;
                    waiting = false;
                }
            }
        }
        // every slot idle and nothing left to start: done (uniform across the CTA)
        if (__syncthreads_and(waiting && wn < 0 && !have_out)) break;

        check_all();
        asm volatile("" ::: "memory");
        unsigned h1;
        const int cnt = var_all(rw + cur * (kN * kRfRow) + lane, h1);
        asm volatile("" ::: "memory");
        if (!waiting) {
            h++;
            // the syndrome is tested every iteration, the last included (:535)
            if ((p.early_stop && cnt == 0) || h == p.max_iters) {
                out_h1 = h1; out_cnt = cnt; out_iters = h;
                w_out = w; ok_out = ok; have_out = true;
                waiting = true;
                w = -1;
            }
        }
    }
}

}  // namespace c4

inline cudaError_t launch_c4_refill(const DecodeParams &p, unsigned int *cursor, int refill_min, int sm_count,
                                    cudaStream_t st)
{
    auto kern = c4::decode_c4_refill_kernel;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c4::kRfSmemBytes);
    if (e != cudaSuccess) return e;
    const long long ctas = (p.n_win + c4::kRfThreads - 1) / c4::kRfThreads;
    const long long grid = ctas < sm_count ? ctas : sm_count;
    kern<<<(int)grid, c4::kRfThreads, c4::kRfSmemBytes, st>>>(p, cursor, refill_min);
    return cudaGetLastError();
}

}  // namespace ldpc535
