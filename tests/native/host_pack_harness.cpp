// CPU test harness for csrc/host_pack.cpp (the persistent packing team of ldpc535_decode_batch):
// dst[i] = src[2 i] for ragged sizes, any team size, repeated jobs on one team (the spin-then-sleep
// hand-over), no write past the end.  Built and run by tests/test_host_pack.py.
#include "host_pack.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

using namespace ldpc535;

int main(int argc, char **argv)
{
    const int threads = argc > 1 ? std::atoi(argv[1]) : 4;
    const size_t nmax = ((size_t)1 << 21) + 77;                 // 16 MiB of complex symbols
    std::vector<float> src(2 * nmax), dst(nmax + 16);
    for (size_t i = 0; i < 2 * nmax; i++) src[i] = (float)((i * 2654435761u) % 1000003u) - 500000.f;
    PackPool pool(threads);
    if (pool.threads() != (threads < 1 ? 1 : threads)) { std::printf("FAIL: team size\n"); return 1; }
    const size_t sizes[] = {0, 1, 7, 8, 9, 1000, 65535, 65536, 65537, 3 * 65536 + 17, (size_t)1 << 20, nmax};
    for (int round = 0; round < 3; round++) {
        for (size_t n : sizes) {
            for (size_t off : {(size_t)0, (size_t)3}) {          // destination not 32-byte aligned as well
                for (auto &v : dst) v = -7.f;
                if (n + off > nmax) continue;
                pool.pack(src.data() + 2 * off, dst.data() + off, n);
                for (size_t i = 0; i < n; i++)
                    if (dst[off + i] != src[2 * (off + i)]) { std::printf("FAIL: n=%zu off=%zu i=%zu\n", n, off, i); return 1; }
                if (dst[off + n] != -7.f || (off && dst[off - 1] != -7.f)) { std::printf("FAIL: overrun n=%zu off=%zu\n", n, off); return 1; }
            }
        }
        if (round == 1) std::this_thread::sleep_for(std::chrono::milliseconds(5));   // let the workers fall asleep once
    }
    pack_real_parts_serial(src.data(), dst.data(), 1000);
    for (size_t i = 0; i < 1000; i++) if (dst[i] != src[2 * i]) { std::printf("FAIL: serial\n"); return 1; }
    std::printf("ok %d\n", threads);
    return 0;
}
