// Minimal stand-in for <gnuradio/block.h>, used ONLY when this module is built without GNU
// Radio (this image has none).  It carries exactly the part of gr::block the two LDPC blocks
// touch -- name, io signatures, forecast/general_work, consume_each and the buffer-sizing
// hints -- so the block sources compile unchanged against a real GNU Radio 3.7 and against
// this shim, and so tests can drive general_work() the way the scheduler does.
#pragma once
#include <string>

#include <boost/shared_ptr.hpp>
#include <gnuradio/io_signature.h>
#include <gnuradio/types.h>

namespace gr {

class block
{
    std::string d_name;
    io_signature::sptr d_in, d_out;
    int d_consumed = 0;
    int d_output_multiple = 1;
    double d_relative_rate = 1.0;

protected:
    block(const std::string &name, io_signature::sptr in, io_signature::sptr out)
        : d_name(name), d_in(in), d_out(out)
    {
    }

public:
    enum { WORK_CALLED_PRODUCE = -2, WORK_DONE = -1 };
    virtual ~block() {}
    std::string name() const { return d_name; }
    io_signature::sptr input_signature() const { return d_in; }
    io_signature::sptr output_signature() const { return d_out; }

    virtual void forecast(int noutput_items, gr_vector_int &ninput_items_required)
    {
        for (auto &n : ninput_items_required) n = noutput_items;
    }
    virtual int general_work(int noutput_items, gr_vector_int &ninput_items,
                             gr_vector_const_void_star &input_items,
                             gr_vector_void_star &output_items) = 0;

    void consume(int, int how_many_items) { d_consumed += how_many_items; }
    void consume_each(int how_many_items) { d_consumed += how_many_items; }
    void set_output_multiple(int m) { d_output_multiple = m; }
    int output_multiple() const { return d_output_multiple; }
    void set_relative_rate(double r) { d_relative_rate = r; }
    double relative_rate() const { return d_relative_rate; }

    // shim only: what the scheduler would read back after a work call
    int compat_take_consumed()
    {
        const int n = d_consumed;
        d_consumed = 0;
        return n;
    }
};

class sync_block : public block
{
protected:
    sync_block(const std::string &name, io_signature::sptr in, io_signature::sptr out)
        : block(name, in, out)
    {
    }

public:
    virtual int work(int noutput_items, gr_vector_const_void_star &input_items,
                     gr_vector_void_star &output_items) = 0;
    int general_work(int noutput_items, gr_vector_int &, gr_vector_const_void_star &input_items,
                     gr_vector_void_star &output_items) override
    {
        const int r = work(noutput_items, input_items, output_items);
        if (r > 0) consume_each(r);
        return r;
    }
};

}  // namespace gr

namespace gnuradio {
template <class T>
boost::shared_ptr<T> get_initial_sptr(T *p)
{
    return boost::shared_ptr<T>(p);
}
}  // namespace gnuradio
