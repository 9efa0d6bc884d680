/* -*- c++ -*- */
/*
 * gr::ldpc_ece535a::image_sink -- byte sink that reassembles BMP files from the decoded stream.
 * Public surface of the reference block (include/ldpc_ece535a/image_sink.h:21-35): same
 * namespace, class name, base class (gr::sync_block), sptr typedef and make(), so
 * examples/receiver.grc keeps loading.  Host file I/O only; nothing here touches the GPU.
 */
#ifndef INCLUDED_LDPC_ECE535A_IMAGE_SINK_H
#define INCLUDED_LDPC_ECE535A_IMAGE_SINK_H

#include <ldpc_ece535a/api.h>
#include <gnuradio/sync_block.h>

namespace gr {
namespace ldpc_ece535a {

class LDPC_ECE535A_API image_sink : virtual public gr::sync_block
{
public:
    typedef boost::shared_ptr<image_sink> sptr;

    static sptr make();
};

}  // namespace ldpc_ece535a
}  // namespace gr

#endif /* INCLUDED_LDPC_ECE535A_IMAGE_SINK_H */
