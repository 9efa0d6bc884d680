"""CPU: the decoder block's host logic -- the sync state machine replayed over per-window
results (lib/sync_replay.h) -- against the block-level oracle
(lib/ldpc_decoder_cb_impl.cc:132-234 restated in oracle/ldpc_oracle.c:orc_decoder_work).
Window results come from the oracle here (no GPU); the GPU tests repeat this end to end."""
import numpy as np
import pytest

from ldpc_ece535a import blocks as B
from oracle import oracle as O
import util

N, NB, THR = 64, 4, 4


def window_table(stream, Hp, method):
    """Decode every window offset in both polarities with the oracle."""
    n_off = stream.size - N + 1
    wins = np.lib.stride_tricks.sliding_window_view(stream, N)[:n_off]
    bp, _, sp, _ = O.decode_frames(np.ascontiguousarray(wins), Hp, method=method, iterations=5,
                                   early_stop=True, threshold=THR, threads=8)
    bn, _, sn, _ = O.decode_frames(np.ascontiguousarray(-wins), Hp, method=method, iterations=5,
                                   early_stop=True, threshold=THR, threads=8)
    return bp, bn, sp, sn


def make_stream(shipped, n_frames, ebn0, seed, lead=0, invert=False, burst=None):
    rng = np.random.default_rng(seed)
    data = rng.integers(0, 256, 4 * n_frames).astype(np.uint8)
    frames, _ = O.encoder_work(shipped["Hp"], shipped["L"], shipped["U"], data, n_frames * N)
    s = util.awgn(frames, ebn0, rng)
    if burst:                                     # a stretch of pure noise: loses sync
        a, b = burst
        s[a * N:b * N] = (rng.standard_normal((b - a) * N) * 1.5).astype(np.float32)
    if invert:
        s = -s
    if lead:
        s = np.concatenate([(rng.standard_normal(lead) * 0.7).astype(np.complex64), s])
    return s.astype(np.complex64), data


@pytest.mark.parametrize("case", [
    dict(n_frames=30, ebn0=None, seed=1),
    dict(n_frames=30, ebn0=None, seed=2, lead=7, invert=True),
    dict(n_frames=40, ebn0=6.0, seed=3, lead=13),
    dict(n_frames=60, ebn0=4.0, seed=4, lead=5, burst=(10, 30)),
    dict(n_frames=50, ebn0=2.0, seed=5, invert=True, lead=70),
    dict(n_frames=40, ebn0=0.0, seed=6),
])
@pytest.mark.parametrize("method", [1, 0])
def test_replay_equals_block_oracle(shipped, case, method):
    stream, _ = make_stream(shipped, **case)
    Hp = shipped["Hp"]
    bp, bn, sp, sn = window_table(stream, Hp, method)
    noutput = (stream.size // N) * NB
    got, consumed, events, state = B.sync_replay_table(bp, bn, sp, sn, stream.size, noutput, N, NB, THR)
    blk = O.DecoderBlock(Hp, method)
    want, wconsumed = blk.work(stream, noutput)
    assert consumed == wconsumed
    assert np.array_equal(got, want)
    assert events == blk.events
    assert state == (blk.st.state, blk.st.errors)


def test_replay_chunked_calls_carry_state(shipped):
    """Scheduler-style chunking: unconsumed symbols are presented again; state carries over."""
    stream, _ = make_stream(shipped, n_frames=60, ebn0=4.0, seed=7, lead=9, burst=(20, 36))
    Hp = shipped["Hp"]
    blk = O.DecoderBlock(Hp, 1)
    want, _ = blk.work(stream, 60 * NB)
    got, pos, state, events = [], 0, (0, 0), []
    rng = np.random.default_rng(0)
    while True:
        take = int(rng.integers(1, 700))
        chunk = stream[pos:pos + take]
        if chunk.size >= N:
            bp, bn, sp, sn = window_table(chunk, Hp, 1)
            nout = int(rng.integers(0, 12)) * NB
            out, consumed, ev, state = B.sync_replay_table(bp, bn, sp, sn, chunk.size, nout, N, NB, THR,
                                                           state)
            got += list(out)
            events += ev
            pos += consumed
        if pos + N > stream.size and take >= stream.size - pos:
            break
    assert got == list(want)
    assert events == blk.events


def test_output_space_limits_production(shipped):
    stream, _ = make_stream(shipped, n_frames=20, ebn0=None, seed=8)
    bp, bn, sp, sn = window_table(stream, shipped["Hp"], 1)
    out, consumed, ev, st = B.sync_replay_table(bp, bn, sp, sn, stream.size, 3 * NB + 3, N, NB, THR)
    assert out.size == 3 * NB and consumed == 3 * N
    out, consumed, ev, st = B.sync_replay_table(bp, bn, sp, sn, stream.size, 3, N, NB, THR)
    assert out.size == 0 and consumed == 0
