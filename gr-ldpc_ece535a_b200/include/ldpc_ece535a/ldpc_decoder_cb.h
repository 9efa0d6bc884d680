/* -*- c++ -*- */
/*
 * gr::ldpc_ece535a::ldpc_decoder_cb -- complex symbols in, decoded bytes out.
 *
 * Public surface of the reference block (include/ldpc_ece535a/ldpc_decoder_cb.h:22-36):
 * same namespace, class name, base class, sptr typedef and make(method) factory, so
 * swig/ldpc_ece535a_swig.i, grc/ldpc_ece535a_ldpc_decoder_cb.xml and the example flowgraphs
 * bind to it unchanged.  The work is done by sm_100a kernels behind include/ldpc535.h; there
 * is no CPU decoder in this module.
 *
 * method: 0 LogDomain (min-sum), 1 SumProduct, 2 BitFlip, 3 Hard; anything else LogDomain
 * (lib/ldpc_decoder_cb_impl.cc:155-164 of the reference).
 *
 * Additions (defaults reproduce the reference exactly): the iteration limit and the early
 * stop the reference hard-codes (d_iterations = 5, :39-40; always stop on a zero syndrome,
 * :535-537) can be changed before the flowgraph starts.
 */
#ifndef INCLUDED_LDPC_ECE535A_LDPC_DECODER_CB_H
#define INCLUDED_LDPC_ECE535A_LDPC_DECODER_CB_H

#include <ldpc_ece535a/api.h>
#include <gnuradio/block.h>

namespace gr {
namespace ldpc_ece535a {

class LDPC_ECE535A_API ldpc_decoder_cb : virtual public gr::block
{
public:
    typedef boost::shared_ptr<ldpc_decoder_cb> sptr;

    static sptr make(const int method);

    virtual void set_max_iterations(int iterations) = 0;   // reference: 5
    virtual int max_iterations() const = 0;
    virtual void set_early_stop(bool on) = 0;               // reference: true
    virtual bool early_stop() const = 0;
};

}  // namespace ldpc_ece535a
}  // namespace gr

#endif /* INCLUDED_LDPC_ECE535A_LDPC_DECODER_CB_H */
