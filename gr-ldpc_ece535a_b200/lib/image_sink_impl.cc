/* -*- c++ -*- */
/*
 * image_sink: behaviour of the reference block (lib/image_sink_impl.cc:46-84).
 *
 * Every input byte is appended to a pending buffer.  When a BMP file header starts at the
 * current byte -- 'B' 'M', four zero reserved bytes at +6..+9 and a known DIB header size at
 * +14, and at least 19 bytes of this work() call remain so the test can look ahead -- the
 * pending buffer, if it already holds a whole file of the previously announced size, is written
 * out (only the announced number of bytes), and the buffer restarts with the new header, whose
 * little-endian size field at +2..+5 is remembered.  A file is therefore written when the NEXT
 * header arrives, exactly like upstream.  Differences: the output path can be redirected
 * (LDPC535_IMAGE_PATH) and the viewer is only spawned when /usr/bin/display exists and
 * LDPC535_IMAGE_DISPLAY is not "0" (headless boxes).
 */
#ifdef HAVE_CONFIG_H
#include "config.h"
#endif

#include "image_sink_impl.h"

#include <gnuradio/io_signature.h>
#include <sys/types.h>
#include <sys/wait.h>
#include <unistd.h>

#include <cstdlib>
#include <cstring>
#include <cstdio>

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

image_sink::sptr image_sink::make() { return gnuradio::get_initial_sptr(new image_sink_impl()); }

image_sink_impl::image_sink_impl()
    : gr::sync_block("image_sink", gr::io_signature::make(1, 1, sizeof(unsigned char)),
                     gr::io_signature::make(0, 0, 0)),
      d_file_size(0), d_path("result.bmp"), d_display(true), d_files_written(0)
{
    if (const char *p = std::getenv("LDPC535_IMAGE_PATH")) d_path = p;
    const char *d = std::getenv("LDPC535_IMAGE_DISPLAY");
    d_display = !(d && std::strcmp(d, "0") == 0) && access("/usr/bin/display", X_OK) == 0;
}

image_sink_impl::~image_sink_impl() {}

bool image_sink_impl::header_at(const unsigned char *p)
{
    static const unsigned char dib_sizes[] = {12, 40, 52, 56, 64, 108, 124};
    if (p[0] != 'B' || p[1] != 'M') return false;
    if (p[6] | p[7] | p[8] | p[9]) return false;
    for (unsigned char s : dib_sizes)
        if (p[14] == s) return true;
    return false;
}

void image_sink_impl::flush_file()
{
    if (d_file_size == 0 || d_pending.size() < d_file_size) return;
    std::FILE *file = std::fopen(d_path.c_str(), "wb");
    if (!file) return;
    std::fwrite(d_pending.data(), 1, d_file_size, file);
    std::fclose(file);
    d_files_written++;
    say("File written");
    if (d_display) {
        // the reference runs `/usr/bin/display result.bmp &` through system() (:68); the path can
        // come from the environment here, so no shell is involved: double fork + execl, the
        // viewer is re-parented to init and never becomes a zombie of the flowgraph
        const pid_t pid = fork();
        if (pid == 0) {
            if (fork() == 0) {
                execl("/usr/bin/display", "display", d_path.c_str(), (char *)NULL);
                _exit(127);
            }
            _exit(0);
        } else if (pid > 0) {
            int status = 0;
            waitpid(pid, &status, 0);
        } else {
            std::fputs("image_sink: could not start the viewer\n", stderr);
        }
    }
}

int image_sink_impl::work(int noutput_items, gr_vector_const_void_star &input_items, gr_vector_void_star &)
{
    const unsigned char *in = (const unsigned char *)input_items[0];
    for (int i = 0; i < noutput_items; i++) {
        if (i < noutput_items - 18 && header_at(in + i)) {
            flush_file();
            d_pending.clear();
            d_file_size = ((unsigned)in[i + 5] << 24) | ((unsigned)in[i + 4] << 16) |
                          ((unsigned)in[i + 3] << 8) | (unsigned)in[i + 2];
            char line[64];
            std::snprintf(line, sizeof(line), "BMP Header Found: fileSize=%u", d_file_size);
            say(line);
        }
        d_pending.push_back(in[i]);
    }
    return noutput_items;   // a sink consumes everything it is shown
}

}  // namespace ldpc_ece535a
}  // namespace gr
