// Stand-in for <gnuradio/io_signature.h> (oracle/_ref build only).
#pragma once
#include <boost/shared_ptr.hpp>
namespace gr {
class io_signature
{
    int d_min, d_max, d_size;
    io_signature(int mn, int mx, int sz) : d_min(mn), d_max(mx), d_size(sz) {}

public:
    typedef boost::shared_ptr<io_signature> sptr;
    static sptr make(int min_streams, int max_streams, int sizeof_stream_item)
    {
        return sptr(new io_signature(min_streams, max_streams, sizeof_stream_item));
    }
    int min_streams() const { return d_min; }
    int max_streams() const { return d_max; }
    int sizeof_stream_item(int) const { return d_size; }
};
}
