/* -*- c++ -*- */
#ifndef INCLUDED_LDPC_ECE535A_IMAGE_SINK_IMPL_H
#define INCLUDED_LDPC_ECE535A_IMAGE_SINK_IMPL_H

#include <ldpc_ece535a/image_sink.h>

#include <string>
#include <vector>

namespace gr {
namespace ldpc_ece535a {

class image_sink_impl : public image_sink
{
private:
    std::vector<unsigned char> d_pending;   // bytes since the last header
    unsigned int d_file_size;               // size field of the last header, 0 before the first
    std::string d_path;                     // LDPC535_IMAGE_PATH, default "result.bmp" like the reference
    bool d_display;                         // spawn /usr/bin/display after writing (reference: always)
    unsigned d_files_written;

    static bool header_at(const unsigned char *p);
    void flush_file();

public:
    image_sink_impl();
    ~image_sink_impl();

    int work(int noutput_items, gr_vector_const_void_star &input_items,
             gr_vector_void_star &output_items);

    unsigned files_written() const { return d_files_written; }
};

}  // namespace ldpc_ece535a
}  // namespace gr

#endif /* INCLUDED_LDPC_ECE535A_IMAGE_SINK_IMPL_H */
