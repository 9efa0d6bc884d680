// Host-side choice of the look-up encoder's tile size (plain C++, no CUDA; tested on the CPU by
// tests/test_encoder_tiles.py).  A unit of work is (tile of frames, row block); a unit costs
// a + b * frames / slots (a: the table stages of the row block, b: look-ups and stores of `slots`
// frames; measured a : b = 2.85) and the slowest SM works on ceil(units / SMs) of them, so take
// the tile count that minimises ceil(tiles * RB / SMs) * (2.85 + frames / slots).
#pragma once
#include <algorithm>

namespace ldpc535 {

struct M4rTiling {
    long long tile_frames;   // frames per tile, 1 .. cap
    long long tiles;         // ceil(n_frames / tile_frames): no tile is empty
    long long units;         // tiles * row_blocks
};

inline double m4r_tiling_cost(long long tiles, long long frames_per_tile, long long row_blocks, long long sms, long long slots)
{
    return (double)((tiles * row_blocks + sms - 1) / sms) * (2.85 + (double)frames_per_tile / (double)slots);
}

// n_frames >= 1; slots = frame slots of a CTA (128 or 124), cap = slots * frames per slot; forced_tile > 0 overrides
inline M4rTiling m4r_choose_tiling(long long n_frames, long long row_blocks, long long sms, long long slots, long long cap,
                                   long long forced_tile = 0)
{
    long long tiles = (n_frames + cap - 1) / cap;
    if (forced_tile > 0) {
        const long long tf = std::min(cap, forced_tile);
        tiles = (n_frames + tf - 1) / tf;
    } else {
        double best = 1e300;
        for (long long nt = tiles, last = tiles + 2 * sms; nt <= last && nt <= n_frames; nt++) {
            const double cost = m4r_tiling_cost(nt, (n_frames + nt - 1) / nt, row_blocks, sms, slots);
            if (cost < best - 1e-9) { best = cost; tiles = nt; }
        }
    }
    M4rTiling t;
    t.tile_frames = std::min(cap, (n_frames + tiles - 1) / tiles);
    t.tiles = (n_frames + t.tile_frames - 1) / t.tile_frames;
    t.units = t.tiles * row_blocks;
    return t;
}

}  // namespace ldpc535
