/* -*- c++ -*- */
/*
 * Host side of ldpc_decoder_cb: the reference's frame/phase synchronisation state machine
 * (lib/ldpc_decoder_cb_impl.cc:147-225 of the reference), replayed over per-window decode
 * results that arrive in batches.
 *
 * In the reference every loop step decodes ONE 64-symbol window, looks at the syndrome weight
 * and then decides where the next window starts (64 symbols on when in sync, 1 symbol on while
 * searching, the same window with the opposite sign on a failed frame while out of sync).
 * "Decode window (offset, polarity)" is a pure function of the input, so the machine here runs
 * the very same decisions in the very same order, but asks a WindowSource for results; on a
 * miss the source decodes a whole batch of the windows the machine is most likely to ask for
 * next (the next frames at the frame stride when in sync; the next offsets in both polarities
 * when searching) in one GPU call.  Speculation only ever costs unused results -- the output
 * bytes, the consumed count and the state transitions are those of the sequential loop.
 */
#ifndef INCLUDED_LDPC_ECE535A_SYNC_REPLAY_H
#define INCLUDED_LDPC_ECE535A_SYNC_REPLAY_H

#include <cstdint>
#include <cstring>
#include <vector>

namespace gr {
namespace ldpc_ece535a {

enum sync_state { OUT_OF_SYNC = 0, IN_SYNC = 1, IN_SYNC_INVERTED = 2 };
enum sync_event { EV_IN_SYNC = 1, EV_IN_SYNC_INVERTED = 2, EV_MAX_ERRORS = 3 };

struct window_result {
    const uint8_t *bytes;   // K/8 packed data bytes
    int synd;               // min(#unsatisfied checks, threshold + 1)
};

// Decodes windows of one general_work() input on request.  Implementations: the GPU batcher
// in ldpc_decoder_cb_impl.cc; a table of precomputed results in the CPU tests.
class window_source
{
public:
    virtual ~window_source() {}
    // The window starting at symbol `offset` multiplied by `polarity` (+1 / -1).  `tracking`
    // tells the source which windows to prefetch on a miss: true = the following frames at
    // the frame stride with this polarity, false = the following offsets in both polarities.
    virtual window_result get(long offset, int polarity, bool tracking) = 0;
};

struct sync_machine {
    int state = OUT_OF_SYNC;
    unsigned errors = 0;

    // One general_work(): returns items produced, sets *consumed.  on_event(ev) is called for
    // each state change the reference prints ("IN SYNC", "IN SYNC; PHASE INVERTED",
    // "MAX ERRORS; OUT OF SYNC").
    template <class OnEvent>
    int run(window_source &src, long ninput, int noutput, int N, int nbytes, int threshold,
            uint8_t *out, long *consumed, OnEvent on_event)
    {
        long ic = 0;
        int op = 0;
        while (ninput - ic >= N && noutput - op >= nbytes) {
            const bool in_sync_before = (state != OUT_OF_SYNC);
            const int pol = (state == IN_SYNC_INVERTED) ? -1 : 1;
            window_result r = src.get(ic, pol, in_sync_before);
            if (r.synd > threshold) {
                if (state == IN_SYNC || state == IN_SYNC_INVERTED) {
                    errors++;
                    if (errors > 10) {
                        errors = 0;
                        state = OUT_OF_SYNC;
                        on_event(EV_MAX_ERRORS);
                    }
                }
                if (state == OUT_OF_SYNC) {
                    // the same window with the opposite sign of what was just tried
                    window_result r2 = src.get(ic, -pol, false);
                    if (r2.synd <= threshold) {
                        on_event(EV_IN_SYNC_INVERTED);
                        state = IN_SYNC_INVERTED;
                        errors = 0;
                        r = r2;
                    } else {
                        ic += 1;   // slide one symbol and try again
                    }
                }
            } else if (state == OUT_OF_SYNC) {
                on_event(EV_IN_SYNC);
                state = IN_SYNC;
                errors = 0;
            }
            if (state == IN_SYNC || state == IN_SYNC_INVERTED) {
                std::memcpy(out + op, r.bytes, (size_t)nbytes);
                ic += N;
                op += nbytes;
            }
        }
        *consumed = ic;
        return op;
    }
};

}  // namespace ldpc_ece535a
}  // namespace gr

#endif /* INCLUDED_LDPC_ECE535A_SYNC_REPLAY_H */
