/* -*- c++ -*- */
#ifndef INCLUDED_LDPC_ECE535A_LDPC_DECODER_CB_IMPL_H
#define INCLUDED_LDPC_ECE535A_LDPC_DECODER_CB_IMPL_H

#include <ldpc_ece535a/ldpc_decoder_cb.h>

#include <cstdint>
#include <vector>

#include "ldpc535.h"
#include "sync_replay.h"

namespace gr {
namespace ldpc_ece535a {

// Replaces the reference's ldpc_decoder_cb_impl (lib/ldpc_decoder_cb_impl.h:22-66): the
// private decode* members are gone (they are kernels behind ldpc535_decode_batch), the sync
// state (d_state, d_errors) lives in sync_machine.
class ldpc_decoder_cb_impl : public ldpc_decoder_cb, private window_source
{
private:
    int d_method;
    int d_M, d_N, d_nbytes;
    int d_iterations;        // reference: 5
    bool d_early_stop;       // reference: always
    int d_threshold;         // reference: M / 8 unsatisfied checks tolerated
    ldpc535_code *d_code;
    sync_machine d_sync;

    // ---- window batcher (one general_work() at a time) ----
    const float *d_in;       // interleaved re, im
    long d_ninput;
    long d_max_frames;       // frames the output buffer can still take, bounds speculation
    long d_base;             // first offset of the cached batch
    long d_stride;           // N (tracking batch) or 1 (search batch)
    long d_count;            // windows per polarity in the cached batch
    int d_pol_mask;          // bit 0: +1 cached, bit 1: -1 cached
    long d_span;             // offsets of the next dense span (doubles per dense miss in a turbulent stretch)
    int d_turbulent;         // > 0: a sync loss happened in this general_work() or the one before
    std::vector<int64_t> d_off;
    std::vector<int8_t> d_pol;
    std::vector<uint8_t> d_bytes, d_synd;
    unsigned long d_batches, d_windows;   // statistics

    window_result get(long offset, int polarity, bool tracking) override;
    void fetch(long offset, int polarity, bool tracking);

public:
    ldpc_decoder_cb_impl(const int method);
    ~ldpc_decoder_cb_impl();

    void set_max_iterations(int iterations) override { d_iterations = iterations < 1 ? 1 : iterations; }
    int max_iterations() const override { return d_iterations; }
    void set_early_stop(bool on) override { d_early_stop = on; }
    bool early_stop() const override { return d_early_stop; }

    void forecast(int noutput_items, gr_vector_int &ninput_items_required);

    int general_work(int noutput_items, gr_vector_int &ninput_items,
                     gr_vector_const_void_star &input_items, gr_vector_void_star &output_items);

    // introspection for tests / logging
    int sync_state_now() const { return d_sync.state; }
    unsigned sync_errors_now() const { return d_sync.errors; }
    unsigned long gpu_batches() const { return d_batches; }
    unsigned long gpu_windows() const { return d_windows; }
    std::vector<int> d_events;   // sync_event values in order of occurrence
};

}  // namespace ldpc_ece535a
}  // namespace gr

#endif /* INCLUDED_LDPC_ECE535A_LDPC_DECODER_CB_IMPL_H */
