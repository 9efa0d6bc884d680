// CPU test harness for the look-up encoder's tile choice (csrc/encode_m4r_tiles.h): for many batch sizes
// the tiling covers every frame with no empty tile, respects the cap, is never worse (in the cost model the
// host uses) than full tiles, and fills whole waves where a full-tile split would leave SMs idle.
#include "encode_m4r_tiles.h"

#include <cstdio>

using namespace ldpc535;

int main()
{
    const long long sms = 148, rbs = 4;
    int checked = 0;
    for (long long slots : {124ll, 128ll})
        for (long long tpf : {7ll, 8ll}) {
            const long long cap = slots * tpf;
            for (long long nf = 1; nf <= 2000000; nf = nf < 5000 ? nf + 37 : nf + nf / 9) {
                const M4rTiling t = m4r_choose_tiling(nf, rbs, sms, slots, cap);
                if (t.tile_frames < 1 || t.tile_frames > cap) { std::printf("FAIL: tile %lld for %lld frames\n", t.tile_frames, nf); return 1; }
                if (t.tiles * t.tile_frames < nf || (t.tiles - 1) * t.tile_frames >= nf) { std::printf("FAIL: cover %lld\n", nf); return 1; }
                if (t.units != t.tiles * rbs) { std::printf("FAIL: units\n"); return 1; }
                const long long full = (nf + cap - 1) / cap;
                const double c_full = m4r_tiling_cost(full, (nf + full - 1) / full, rbs, sms, slots);
                const double c_pick = m4r_tiling_cost(t.tiles, t.tile_frames, rbs, sms, slots);
                if (c_pick > c_full + 1e-9) { std::printf("FAIL: %lld frames: picked cost %.3f > full tiles %.3f\n", nf, c_pick, c_full); return 1; }
                checked++;
            }
            // the cases DESIGN.md quotes
            const M4rTiling a = m4r_choose_tiling(200000, rbs, sms, slots, cap);
            if (a.units % sms != 0) { std::printf("FAIL: 200 000 frames should fill whole waves (%lld units)\n", a.units); return 1; }
            const M4rTiling f = m4r_choose_tiling(200000, rbs, sms, slots, cap, 500);
            if (f.tile_frames != 500 || f.tiles != 400) { std::printf("FAIL: forced tile\n"); return 1; }
            const M4rTiling s = m4r_choose_tiling(2048, rbs, sms, slots, cap);
            if (s.units < sms - rbs || s.units > sms) { std::printf("FAIL: 2 048 frames should give every SM one unit (%lld)\n", s.units); return 1; }
        }
    std::printf("ok %d\n", checked);
    return 0;
}
