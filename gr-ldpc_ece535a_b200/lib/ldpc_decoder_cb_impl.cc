/* -*- c++ -*- */
/*
 * ldpc_decoder_cb: GNU Radio block, gr_complex in, unsigned char out.
 *
 * Behaviour follows the reference block (lib/ldpc_decoder_cb_impl.cc:35-234): 64 symbols per
 * frame, only the real part is used, the hard-wired 32x64 code, 5 iterations, frames with at
 * most M/8 = 4 unsatisfied checks are accepted, the OUT_OF_SYNC / IN_SYNC / IN_SYNC_INVERTED
 * machine slides one symbol at a time while searching, and the 32 data bits leave as 4 bytes
 * MSB first.  What differs is where the arithmetic runs: windows are decoded in batches by the
 * sm_100a kernels behind ldpc535_decode_batch (include/ldpc535.h); sync_replay.h explains why
 * batching cannot change the output.  No CPU decoder exists here: if the GPU library fails the
 * constructor throws and general_work() returns WORK_DONE.
 */
#ifdef HAVE_CONFIG_H
#include "config.h"
#endif

#include "ldpc_decoder_cb_impl.h"

#include <gnuradio/io_signature.h>

#include <algorithm>
#include <cstdlib>
#include <cstdio>
#include <stdexcept>
#include <string>

namespace gr {
namespace ldpc_ece535a {

namespace {
// console lines the reference prints with std::cout; C stdio keeps the module independent of
// which libstdc++ the host process (e.g. Python) happens to have loaded
void say(const char *line)
{
    std::fputs(line, stdout);
    std::fputc('\n', stdout);
    std::fflush(stdout);
}
}  // namespace

namespace {
int device_from_env()
{
    const char *s = std::getenv("LDPC535_DEVICE");
    return s ? std::atoi(s) : 0;
}
// How far the batcher speculates on a miss.  A GPU call costs ~0.15-0.4 ms of latency whatever its
// size while a window costs ~1 ns of kernel time, so the batcher buys as many windows per call as it
// may need:
//   tracking miss, link quiet      the next frames at the frame stride, current polarity
//   search miss / link turbulent   a DENSE span: every offset from here on in both polarities.  A dense
//                                  span answers every later request that falls inside it -- the search
//                                  slides, the frames that follow a re-lock (offset - base is a multiple
//                                  of 1), the -tx retries -- so a stream that keeps losing sync (Eb/N0
//                                  0-2 dB: a loss every ~13 frames) costs one call per span instead of two
//                                  per loss.  The span starts at 4096 offsets and doubles with every dense
//                                  miss in a turbulent stretch (at most kSearchMax), so a clean
//                                  acquisition stays cheap and a noisy buffer is covered in ~log2 calls.
const long kTrackBatch = 1 << 16;   // frames at the frame stride
const size_t kMaxEventLog = 65536;  // most recent sync events kept in d_events
const long kSearchMin = 4096;       // offsets of the first dense span, each in both polarities
const long kSearchMax = 1 << 18;    // 2 x 262 144 windows: ~0.6 ms of kernel, 2.6 MB of results
}  // namespace

ldpc_decoder_cb::sptr ldpc_decoder_cb::make(const int method)
{
    return gnuradio::get_initial_sptr(new ldpc_decoder_cb_impl(method));
}

ldpc_decoder_cb_impl::ldpc_decoder_cb_impl(const int method)
    : gr::block("ldpc_decoder_cb", gr::io_signature::make(1, 1, sizeof(gr_complex)),
                gr::io_signature::make(1, 1, sizeof(unsigned char))),
      d_method(method), d_M(0), d_N(0), d_nbytes(0), d_iterations(LDPC535_REF_ITERATIONS),
      d_early_stop(true), d_threshold(0), d_code(NULL), d_in(NULL), d_ninput(0), d_max_frames(0),
      d_base(0), d_stride(1), d_count(0), d_pol_mask(0), d_span(kSearchMin), d_turbulent(0),
      d_batches(0), d_windows(0)
{
    const int st = ldpc535_code_create_default(device_from_env(), &d_code);
    if (st != LDPC535_OK)
        throw std::runtime_error(std::string("ldpc_decoder_cb: cannot set up the GPU decoder: ") +
                                 ldpc535_strerror(st) + " (" + ldpc535_last_error() + ")");
    int K = 0;
    ldpc535_code_info(d_code, &d_M, &d_N, &K, NULL, NULL);
    d_nbytes = K / 8;
    d_threshold = d_M / 8;

    // the console line of the reference constructor, verbatim (:108-116)
    if (d_method == 3) say("Method: Hard");
    else if (d_method == 2) say("Method: BitFlip");
    else if (d_method == 1) say("Method: SumProduct");
    else say("Method: LogDomain");

    // One output byte needs 16 input symbols; whole frames only.  Hints for the scheduler's
    // buffer sizing, they do not change what is produced.
    set_output_multiple(d_nbytes);
    set_relative_rate((double)d_nbytes / (double)d_N);
}

ldpc_decoder_cb_impl::~ldpc_decoder_cb_impl() { ldpc535_code_destroy(d_code); }

void ldpc_decoder_cb_impl::forecast(int noutput_items, gr_vector_int &ninput_items_required)
{
    // the reference asks for noutput_items * N (:126-130); a frame of N symbols yields
    // N/16 bytes, so that request is kept (it over-asks, which only helps batching)
    ninput_items_required[0] = noutput_items * d_N;
}

void ldpc_decoder_cb_impl::fetch(long offset, int polarity, bool tracking)
{
    d_off.clear();
    d_pol.clear();
    d_base = offset;
    if (tracking && d_turbulent == 0) {
        long n = (d_ninput - offset) / d_N;
        n = std::min(n, std::min(d_max_frames, kTrackBatch));
        n = std::max(n, 1L);
        d_stride = d_N;
        d_count = n;
        d_pol_mask = polarity > 0 ? 1 : 2;
        for (long k = 0; k < n; k++) {
            d_off.push_back(offset + k * d_N);
            d_pol.push_back((int8_t)polarity);
        }
    } else {
        long n = d_ninput - d_N - offset + 1;       // windows that still fit the input
        n = std::max(1L, std::min(n, d_span));
        d_span = std::min(kSearchMax, d_span * 2);  // the next dense miss of this stretch reaches twice as far
        d_stride = 1;
        d_count = n;
        d_pol_mask = 3;
        for (int p = 0; p < 2; p++)
            for (long k = 0; k < n; k++) {
                d_off.push_back(offset + k);
                d_pol.push_back((int8_t)(p == 0 ? 1 : -1));
            }
    }
    const size_t nw = d_off.size();
    d_bytes.resize(nw * (size_t)d_nbytes);
    d_synd.resize(nw);
    const int st = ldpc535_decode_batch(d_code, d_in, (size_t)d_ninput, d_off.data(), d_pol.data(), nw,
                                        d_method, d_iterations, d_early_stop ? 1 : 0, d_threshold,
                                        d_bytes.data(), d_synd.data(), NULL);
    if (st != LDPC535_OK)
        throw std::runtime_error(std::string("ldpc_decoder_cb: GPU decode failed: ") +
                                 ldpc535_strerror(st) + " (" + ldpc535_last_error() + ")");
    d_batches++;
    d_windows += nw;
}

window_result ldpc_decoder_cb_impl::get(long offset, int polarity, bool tracking)
{
    const int bit = polarity > 0 ? 1 : 2;
    long k = -1;
    if (d_count > 0 && (d_pol_mask & bit) && offset >= d_base && (offset - d_base) % d_stride == 0) {
        k = (offset - d_base) / d_stride;
        if (k >= d_count) k = -1;
    }
    if (k < 0) {
        fetch(offset, polarity, tracking);
        k = 0;
    }
    const size_t idx = (d_pol_mask == 3 && polarity < 0) ? (size_t)(d_count + k) : (size_t)k;
    window_result r;
    r.bytes = d_bytes.data() + idx * (size_t)d_nbytes;
    r.synd = d_synd[idx];
    return r;
}

int ldpc_decoder_cb_impl::general_work(int noutput_items, gr_vector_int &ninput_items,
                                       gr_vector_const_void_star &input_items,
                                       gr_vector_void_star &output_items)
{
    const gr_complex *in = (const gr_complex *)input_items[0];
    unsigned char *out = (unsigned char *)output_items[0];

    d_in = reinterpret_cast<const float *>(in);
    d_ninput = ninput_items[0];
    d_max_frames = noutput_items / d_nbytes;
    d_count = 0;                                    // results never outlive the input buffer

    // a call that saw no sync loss ends a turbulent stretch: back to frame-stride batches and short spans
    if (d_turbulent > 0) d_turbulent--;
    if (d_turbulent == 0) d_span = kSearchMin;

    long consumed = 0;
    int produced = 0;
    try {
        produced = d_sync.run(*this, d_ninput, noutput_items, d_N, d_nbytes, d_threshold, out,
                              &consumed, [this](int ev) {
                                  if (ev == EV_MAX_ERRORS) d_turbulent = 2;   // this call and the next one
                                  // bounded log for tests / monitoring: a flowgraph that never
                                  // reads it must not grow without limit
                                  if (d_events.size() >= kMaxEventLog)
                                      d_events.erase(d_events.begin(), d_events.begin() + kMaxEventLog / 2);
                                  d_events.push_back(ev);
                                  if (ev == EV_IN_SYNC) say("IN SYNC");
                                  else if (ev == EV_IN_SYNC_INVERTED) say("IN SYNC; PHASE INVERTED");
                                  else say("MAX ERRORS; OUT OF SYNC");
                              });
    } catch (const std::exception &e) {
        std::fprintf(stderr, "%s\n", e.what());     // no CPU fallback: stop the flowgraph
        return WORK_DONE;
    }
    d_in = NULL;

    // Tell runtime system how many input items we consumed and how many we produced.
    consume_each((int)consumed);
    return produced;
}

}  // namespace ldpc_ece535a
}  // namespace gr
