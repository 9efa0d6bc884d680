"""GPU: the two GNU Radio blocks, driven through general_work() the way the scheduler does
(lib/block_harness.cc), against the block-level oracle
(lib/ldpc_encoder_bc_impl.cc:118-178, lib/ldpc_decoder_cb_impl.cc:132-234 restated in
oracle/ldpc_oracle.c).  Byte streams, consumed counts, sync events and final sync state must
be identical; the decoder's windows run on the sm_100a kernels in speculative batches."""
import numpy as np
import pytest

import ldpc_ece535a as L
from ldpc_ece535a import blocks as B
from oracle import oracle as O
import util
from test_sync_replay import make_stream, N, NB

pytestmark = pytest.mark.gpu


def test_block_identity_and_forecast():
    enc, dec = L.ldpc_encoder_bc(), L.ldpc_decoder_cb(1)
    assert enc.name() == "ldpc_encoder_bc" and dec.name() == "ldpc_decoder_cb"
    assert enc.item_sizes() == (1, 8)            # unsigned char in, gr_complex out
    assert dec.item_sizes() == (8, 1)            # gr_complex in, unsigned char out
    assert enc.forecast(64) == 4 and enc.forecast(65) == 5      # ceil(n / 16)
    assert dec.forecast(4) == 256                               # noutput_items * 64


@pytest.mark.parametrize("nbytes_in,nout", [(4000, 64000), (4003, 64000), (4000, 6400 + 63), (3, 64),
                                            (4, 63), (0, 640), (40, 0)])
def test_encoder_block_vs_oracle(shipped, nbytes_in, nout):
    rng = np.random.default_rng(nbytes_in + nout)
    data = rng.integers(0, 256, nbytes_in).astype(np.uint8)
    enc = L.ldpc_encoder_bc()
    out, consumed = enc.general_work(data, nout)
    want, wconsumed = O.encoder_work(shipped["Hp"], shipped["L"], shipped["U"], data, nout)
    assert consumed == wconsumed
    assert np.array_equal(out, want)


CASES = [
    dict(n_frames=200, ebn0=None, seed=11),
    dict(n_frames=120, ebn0=None, seed=12, lead=7, invert=True),
    dict(n_frames=300, ebn0=6.0, seed=13, lead=13),
    dict(n_frames=300, ebn0=4.0, seed=14, lead=5, burst=(50, 90)),
    dict(n_frames=200, ebn0=2.0, seed=15, invert=True, lead=70),
    dict(n_frames=150, ebn0=0.0, seed=16),
]


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("method", [1, 0, 2, 3])
def test_decoder_block_vs_oracle(shipped, case, method):
    stream, _ = make_stream(shipped, **case)
    noutput = (stream.size // N) * NB
    dec = L.ldpc_decoder_cb(method)
    got, consumed = dec.general_work(stream, noutput)
    blk = O.DecoderBlock(shipped["Hp"], method)
    want, wconsumed = blk.work(stream, noutput)
    assert consumed == wconsumed
    assert np.array_equal(got, want)
    assert dec.take_events() == blk.events
    st = dec.state()
    assert (st["state"], st["errors"]) == (blk.st.state, blk.st.errors)
    assert st["gpu_windows"] >= consumed // N       # everything was decoded on the GPU


def test_decoder_block_chunked_like_the_scheduler(shipped):
    """Small, irregular work calls (GNU Radio's default buffers hold ~64 frames): unconsumed
    symbols come back in the next call, output space is sometimes short."""
    stream, data = make_stream(shipped, n_frames=400, ebn0=5.0, seed=21, lead=33, burst=(100, 140))
    blk = O.DecoderBlock(shipped["Hp"], 1)
    want, _ = blk.work(stream, 400 * NB)
    dec = L.ldpc_decoder_cb(1)
    rng = np.random.default_rng(1)
    got, pos = [], 0
    while stream.size - pos >= N:
        take = int(rng.integers(N, 4096))
        nout = int(rng.integers(0, 64)) * NB + int(rng.integers(0, 4))
        out, consumed = dec.general_work(stream[pos:pos + take], nout)
        got += list(out)
        pos += consumed
        if consumed == 0 and take >= stream.size - pos and nout >= NB:
            break
    assert got == list(want)
    assert dec.take_events() == blk.events


def test_transmit_receive_pipeline(shipped):
    """BASELINE config 2's shape: an image-sized byte file (19 270 bytes, as examples/mandril.bmp)
    through encoder block -> AWGN -> decoder block at Eb/N0 = 0..4 dB; at every SNR the decoded
    stream equals the block-level oracle's, and the noiseless run returns the file."""
    rng = np.random.default_rng(31)
    payload = rng.integers(0, 256, 19270).astype(np.uint8)
    payload[:2] = (0x42, 0x4D)                               # 'BM'
    enc = L.ldpc_encoder_bc()
    sym, consumed = enc.general_work(payload, (payload.size // 4) * 64)
    assert consumed == 19268 and sym.size == 4817 * 64       # two tail bytes never encoded
    dec = L.ldpc_decoder_cb(1)
    out, c = dec.general_work(sym, 4817 * 4)
    assert np.array_equal(out, payload[:19268]) and c == sym.size
    for ebn0 in (0.0, 1.0, 2.0, 3.0, 4.0):
        noisy = util.awgn(sym, ebn0, np.random.default_rng(int(ebn0) + 40))
        dec = L.ldpc_decoder_cb(1)
        got, consumed = dec.general_work(noisy, 4817 * 4)
        blk = O.DecoderBlock(shipped["Hp"], 1)
        want, wconsumed = blk.work(noisy, 4817 * 4)
        assert consumed == wconsumed and np.array_equal(got, want), ebn0
        assert dec.take_events() == blk.events


def test_decoder_block_iteration_knob(shipped):
    """max_iterations / early_stop are additions; the defaults are the reference's 5 / on."""
    stream, _ = make_stream(shipped, n_frames=100, ebn0=3.0, seed=51)
    dec = L.ldpc_decoder_cb(1)
    dec.set_max_iterations(50, early_stop=True)
    got, consumed = dec.general_work(stream, 400)
    blk = O.DecoderBlock(shipped["Hp"], 1, iterations=50)
    want, wconsumed = blk.work(stream, 400)
    assert consumed == wconsumed and np.array_equal(got, want)


def test_headless_txrx_writes_the_image(shipped, tmp_path):
    """Config 2 end to end: file -> encoder block -> AWGN -> decoder block -> image_sink, the
    headless stand-in for transmitter.grc + receiver.grc (examples/headless_txrx.py).  Sent twice
    back to back; noiseless the sink writes back exactly the file, and at every Eb/N0 the decoded
    stream equals the block-level oracle's."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("headless_txrx", os.path.join(root, "examples", "headless_txrx.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    from test_image_sink import bmp
    image = bmp(19268, 77)                 # mandril.bmp's size, rounded down to whole frames
    payload = np.tile(image, 2)
    out = tmp_path / "result.bmp"
    r = mod.run_chain(payload, None, 1, out_path=str(out), buf_frames=300)
    assert np.array_equal(r["decoded"], payload)
    assert r["files"] == 1 and out.read_bytes() == image.tobytes()
    assert r["events"] == [1]
    for ebn0 in (0.0, 2.0, 4.0):
        r = mod.run_chain(payload, ebn0, 1, out_path=str(out), buf_frames=300, seed=int(ebn0) + 3)
        # the same chain through the oracle (encoder exact, so reuse the noisy symbols' recipe)
        sym, _ = O.encoder_work(shipped["Hp"], shipped["L"], shipped["U"], payload, payload.size * 16)
        rng = np.random.default_rng(int(ebn0) + 3)
        sigma = np.float32(np.sqrt(10.0 ** (-ebn0 / 10.0)))
        sym = sym.copy()
        sym.real += rng.standard_normal(sym.size, dtype=np.float32) * sigma
        sym.imag += rng.standard_normal(sym.size, dtype=np.float32) * sigma
        blk = O.DecoderBlock(shipped["Hp"], 1)
        want, _ = blk.work(sym, payload.size)
        assert np.array_equal(r["decoded"], want), ebn0
        assert r["events"] == blk.events


def test_headless_txrx_with_the_reference_image(shipped, tmp_path):
    """BASELINE config 2 with the file it names: the reference's examples/mandril.bmp (committed as
    tests/golden/mandril.bmp, 19 270 bytes) sent twice through encoder block -> AWGN -> decoder block
    -> image_sink at Eb/N0 = noiseless, 0, 1, 2, 3, 4 dB.  Noiseless the sink writes the image back
    byte for byte; at every Eb/N0 the decoded stream and the sync events equal the block-level
    oracle's and -- where the prebuilt oracle/_ref travelled along -- the reference's own decoder
    block's, driven over the same symbols."""
    import importlib.util
    import os
    from oracle import ref as R
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("headless_txrx", os.path.join(root, "examples", "headless_txrx.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    image = np.fromfile(os.path.join(root, "tests", "golden", "mandril.bmp"), np.uint8)
    assert image.size == 19270
    payload = np.tile(image, 2)                      # 38 540 bytes = 9 635 whole frames
    out = tmp_path / "result.bmp"
    r = mod.run_chain(payload, None, 1, out_path=str(out), buf_frames=4096)
    assert np.array_equal(r["decoded"], payload)
    assert r["files"] == 1 and out.read_bytes() == image.tobytes()
    assert r["events"] == [1]
    sym0, _ = O.encoder_work(shipped["Hp"], shipped["L"], shipped["U"], payload, payload.size * 16)
    for ebn0 in (0.0, 1.0, 2.0, 3.0, 4.0):
        r = mod.run_chain(payload, ebn0, 1, out_path=str(out), buf_frames=4096, seed=int(ebn0) + 30)
        rng = np.random.default_rng(int(ebn0) + 30)
        sigma = np.float32(np.sqrt(10.0 ** (-ebn0 / 10.0)))
        sym = sym0.copy()
        sym.real += rng.standard_normal(sym.size, dtype=np.float32) * sigma
        sym.imag += rng.standard_normal(sym.size, dtype=np.float32) * sigma
        blk = O.DecoderBlock(shipped["Hp"], 1)
        want, _ = blk.work(sym, payload.size)
        assert np.array_equal(r["decoded"], want), ebn0
        assert r["events"] == blk.events
        if R.available() and ebn0 in (2.0, 4.0):      # the reference block decodes one window at a time on one core
            ref = R.RefDecoder(1)
            rout, _ = ref.work(sym, payload.size)
            assert np.array_equal(r["decoded"], rout), ebn0
            assert r["events"] == ref.events
            ref.close()
