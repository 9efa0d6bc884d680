// Stand-in for <gnuradio/block.h> (oracle/_ref build only; GNU Radio is not in this image).
// Exactly what the reference's blocks touch: the (name, in, out) constructor, the virtual
// forecast / general_work pair and consume_each.  ref_take_consumed() is how the driver
// (oracle/ref_driver.cc) reads back what the scheduler would.
#pragma once
#include <complex>
#include <string>
#include <vector>

#include <boost/shared_ptr.hpp>
#include <gnuradio/io_signature.h>

typedef std::complex<float> gr_complex;
typedef std::vector<int> gr_vector_int;
typedef std::vector<const void *> gr_vector_const_void_star;
typedef std::vector<void *> gr_vector_void_star;

namespace gr {

class block
{
    std::string d_name;
    io_signature::sptr d_in, d_out;
    long d_consumed;

protected:
    block() : d_consumed(0) {}                     // pure-interface subclasses (virtual base)
    block(const std::string &name, io_signature::sptr in, io_signature::sptr out)
        : d_name(name), d_in(in), d_out(out), d_consumed(0)
    {
    }

public:
    virtual ~block() {}
    std::string name() const { return d_name; }
    io_signature::sptr input_signature() const { return d_in; }
    io_signature::sptr output_signature() const { return d_out; }

    virtual void forecast(int noutput_items, gr_vector_int &ninput_items_required)
    {
        for (std::size_t i = 0; i < ninput_items_required.size(); i++)
            ninput_items_required[i] = noutput_items;
    }
    virtual int general_work(int noutput_items, gr_vector_int &ninput_items,
                             gr_vector_const_void_star &input_items,
                             gr_vector_void_star &output_items) = 0;

    void consume_each(int how_many_items) { d_consumed += how_many_items; }

    long ref_take_consumed()
    {
        const long n = d_consumed;
        d_consumed = 0;
        return n;
    }
};

}  // namespace gr

namespace gnuradio {
template <class T> boost::shared_ptr<T> get_initial_sptr(T *p) { return boost::shared_ptr<T>(p); }
}
