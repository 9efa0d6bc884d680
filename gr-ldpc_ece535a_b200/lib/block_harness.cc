/* -*- c++ -*- */
/*
 * C entry points that stand in for the GNU Radio scheduler: build a block through its public
 * make(), call forecast()/general_work() on caller-supplied buffers and read back what the
 * block consumed -- exactly what gr::block_executor does around a work call.  Used by the
 * Python package (ldpc_ece535a.ldpc_decoder_cb / ldpc_encoder_bc outside GNU Radio) and by the
 * tests; a GNU Radio build does not need this file.
 */
#include <ldpc_ece535a/ldpc_decoder_cb.h>
#include <ldpc_ece535a/image_sink.h>
#include <ldpc_ece535a/ldpc_encoder_bc.h>

#include <cstring>
#include <exception>
#include <string>

#include "image_sink_impl.h"
#include "ldpc_decoder_cb_impl.h"
#include "sync_replay.h"

using namespace gr::ldpc_ece535a;

namespace {
thread_local std::string g_err;

struct blk {
    boost::shared_ptr<gr::block> b;
    ldpc_decoder_cb_impl *dec = nullptr;
};

// window_source over a table of precomputed results (CPU tests of the sync replay)
struct table_source : window_source {
    const uint8_t *bytes_pos, *bytes_neg, *synd_pos, *synd_neg;
    int nbytes;
    long n_off;
    long lookups = 0;
    window_result get(long offset, int polarity, bool) override
    {
        lookups++;
        window_result r;
        r.bytes = (polarity > 0 ? bytes_pos : bytes_neg) + (size_t)offset * nbytes;
        r.synd = (polarity > 0 ? synd_pos : synd_neg)[offset];
        return r;
    }
};
}  // namespace

#define HARNESS_API extern "C" __attribute__((visibility("default")))

HARNESS_API const char *ldpc535_blk_last_error(void) { return g_err.c_str(); }

HARNESS_API void *ldpc535_blk_decoder_new(int method)
{
    try {
        blk *h = new blk();
        h->b = ldpc_decoder_cb::make(method);
        h->dec = dynamic_cast<ldpc_decoder_cb_impl *>(h->b.get());
        return h;
    } catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}

HARNESS_API void *ldpc535_blk_encoder_new(void)
{
    try {
        blk *h = new blk();
        h->b = ldpc_encoder_bc::make();
        return h;
    } catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}

HARNESS_API void *ldpc535_blk_image_sink_new(void)
{
    try {
        blk *h = new blk();
        h->b = image_sink::make();
        return h;
    } catch (const std::exception &e) {
        g_err = e.what();
        return nullptr;
    }
}

HARNESS_API unsigned ldpc535_blk_image_sink_files(void *p)
{
    image_sink_impl *s = dynamic_cast<image_sink_impl *>(static_cast<blk *>(p)->b.get());
    return s ? s->files_written() : 0;
}

HARNESS_API void ldpc535_blk_free(void *p) { delete static_cast<blk *>(p); }

HARNESS_API const char *ldpc535_blk_name(void *p)
{
    static thread_local std::string s;
    s = static_cast<blk *>(p)->b->name();
    return s.c_str();
}

HARNESS_API int ldpc535_blk_item_size(void *p, int output)
{
    blk *h = static_cast<blk *>(p);
    return output ? h->b->output_signature()->sizeof_stream_item(0)
                  : h->b->input_signature()->sizeof_stream_item(0);
}

HARNESS_API int ldpc535_blk_forecast(void *p, int noutput_items)
{
    gr_vector_int req(1, 0);
    static_cast<blk *>(p)->b->forecast(noutput_items, req);
    return req[0];
}

// returns items produced (or a negative WORK_* code); *consumed = items the block consumed
HARNESS_API int ldpc535_blk_general_work(void *p, int noutput_items, int ninput_items, const void *in,
                                         void *out, int *consumed)
{
    blk *h = static_cast<blk *>(p);
    gr_vector_int nin(1, ninput_items);
    gr_vector_const_void_star ins(1, in);
    gr_vector_void_star outs(1, out);
    int r;
    try {
        r = h->b->general_work(noutput_items, nin, ins, outs);
    } catch (const std::exception &e) {
        g_err = e.what();
        r = -3;
    }
    *consumed = h->b->compat_take_consumed();
    return r;
}

HARNESS_API int ldpc535_blk_decoder_set(void *p, int max_iterations, int early_stop)
{
    blk *h = static_cast<blk *>(p);
    if (!h->dec) return 1;
    h->dec->set_max_iterations(max_iterations);
    h->dec->set_early_stop(early_stop != 0);
    return 0;
}

// state[0] = sync state, state[1] = error count, state[2] = GPU batches, state[3] = windows decoded
HARNESS_API int ldpc535_blk_decoder_state(void *p, unsigned long *state)
{
    blk *h = static_cast<blk *>(p);
    if (!h->dec) return 1;
    state[0] = (unsigned long)h->dec->sync_state_now();
    state[1] = h->dec->sync_errors_now();
    state[2] = h->dec->gpu_batches();
    state[3] = h->dec->gpu_windows();
    return 0;
}

// copies up to cap events (sync_event values) and clears the log; returns how many there were
HARNESS_API int ldpc535_blk_decoder_events(void *p, int *events, int cap)
{
    blk *h = static_cast<blk *>(p);
    if (!h->dec) return 0;
    const int n = (int)h->dec->d_events.size();
    for (int i = 0; i < n && i < cap; i++) events[i] = h->dec->d_events[i];
    h->dec->d_events.clear();
    return n;
}

// The sync replay alone, over a table of per-(offset, polarity) results: host logic, no GPU.
// state_io[0] = state, state_io[1] = errors (in/out).
HARNESS_API int ldpc535_sync_replay_table(const uint8_t *bytes_pos, const uint8_t *bytes_neg,
                                          const uint8_t *synd_pos, const uint8_t *synd_neg, long n_off,
                                          long ninput, int noutput, int N, int nbytes, int threshold,
                                          unsigned *state_io, uint8_t *out, long *consumed, int *events,
                                          int cap, int *n_events)
{
    table_source src;
    src.bytes_pos = bytes_pos; src.bytes_neg = bytes_neg;
    src.synd_pos = synd_pos; src.synd_neg = synd_neg;
    src.nbytes = nbytes; src.n_off = n_off;
    sync_machine m;
    m.state = (int)state_io[0];
    m.errors = state_io[1];
    int ne = 0;
    const int produced = m.run(src, ninput, noutput, N, nbytes, threshold, out, consumed, [&](int ev) {
        if (ne < cap) events[ne] = ev;
        ne++;
    });
    state_io[0] = (unsigned)m.state;
    state_io[1] = m.errors;
    *n_events = ne;
    return produced;
}
