"""GPU: the sm_100a path, through the C ABI and through the block layer, against the REFERENCE'S
OWN CODE -- not only against the restatement in oracle/ldpc_oracle.c.

* goldens: tests/golden/ref_build_golden.npz holds outputs of oracle/_ref (the reference's
  decoder/encoder block sources compiled from /root/reference against oracle/refshim/, see
  tools/gen_ref_golden.py) on stored inputs;
* live: oracle/_ref/libldpc_ref.so itself -- it is built in the build container and travels to the
  GPU box with the snapshot (git-ignored, not gpurun-ignored); skipped if it is not there.
  Nothing here reads /root/reference at run time.

Bars: encoder symbols, hard decisions, syndrome weights, block byte streams, consumed/produced
counts, sync events and final sync state identical.
"""
import os

import numpy as np
import pytest

import ldpc_ece535a as L
from oracle import ref as R
import util
from test_reference_build import run_chunks, CODES

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
live = pytest.mark.skipif(not R.available(), reason="oracle/_ref/libldpc_ref.so did not travel")


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLD, "ref_build_golden.npz"))


@pytest.fixture(scope="module")
def c4():
    code = L.Code(None, device=0)
    yield code
    code.close()


def data_bits(bytes_, K):
    return util.unpack_bits_msb(bytes_, K)


# ---------------------------------------------------------------------------------------------
# committed outputs of the reference build
# ---------------------------------------------------------------------------------------------

def test_encoder_matches_reference_build(c4, g):
    sym = c4.encode(g["enc_bytes"].reshape(-1, 4))
    assert np.array_equal(sym.real.reshape(-1), g["enc_symbols_re"].astype(np.float32))
    assert not sym.imag.any()


@pytest.mark.parametrize("name", CODES)
def test_encoder_parity_matches_reference_build_all_codes(ref_codes, g, name):
    H = ref_codes[name]["H"]
    M, N = H.shape
    if (N - M) % 8:
        pytest.skip("K = %d is not a whole number of bytes" % (N - M))
    code = L.Code(H, device=0)
    d = g["par_%s_d" % name]
    sym = code.encode(util.pack_bits_msb(d))
    bits = (sym.real.reshape(len(d), N) > 0).astype(np.int8)
    assert np.array_equal(bits[:, :M], g["par_%s_c" % name])
    assert np.array_equal(bits[:, M:], d)
    code.close()


@pytest.mark.parametrize("kernel", ["warp", "block", "c4-thread"])
@pytest.mark.parametrize("tag", ["clean", "0dB", "2dB", "4dB", "6dB"])
@pytest.mark.parametrize("iters", [5, 50])
def test_decisions_match_reference_build(c4, g, kernel, tag, iters):
    """Same frames through the CUDA decoder: data bytes and the syndrome weight of the decision
    equal what the reference's decode members + checkFrame(v, 4) returned."""
    try:
        c4.set_kernel(kernel)
    except L.Ldpc535Error as e:
        pytest.skip(str(e))
    rx = g["cw_%s_rx" % tag]
    sym = rx.astype(np.complex64)
    for method in (1, 0, 2, 3):
        key = "cw_%s_m%d_it%d_vhat" % (tag, method, iters)
        if key not in g or (kernel == "c4-thread" and method != 1):
            continue
        want = util.pack_bits_msb(g[key][:, 32:])
        b, sy, _ = c4.decode(sym, method=method, max_iters=iters, early_stop=True)
        assert np.array_equal(b, want), (method, tag, iters)
        assert np.array_equal(sy, g["cw_%s_m%d_it%d_synd" % (tag, method, iters)].astype(np.uint8))
    c4.set_kernel(None)


@pytest.mark.parametrize("name", ["hData3", "hData5", "hData2"])
def test_decisions_other_codes_match_reference_build(ref_codes, g, name):
    H = ref_codes[name]["H"]
    M, N = H.shape
    if (N - M) % 8:
        pytest.skip("K = %d is not a whole number of bytes" % (N - M))
    code = L.Code(H, device=0)
    sym = g["cwx_%s_rx" % name].astype(np.complex64)
    for method in (0, 1, 2, 3):
        for iters in (5, 20):
            want = util.pack_bits_msb(g["cwx_%s_m%d_it%d_vhat" % (name, method, iters)][:, M:])
            b, _, _ = code.decode(sym, method=method, max_iters=iters, early_stop=True)
            assert np.array_equal(b, want), (name, method, iters)
    code.close()


@pytest.mark.parametrize("tag,method", [("blk_m0_4dB", 0), ("blk_m1_4dB", 1), ("blk_m1_1dB", 1),
                                        ("blk_m2_7dB", 2), ("blk_m3_7dB", 3)])
def test_decoder_block_matches_reference_build(g, tag, method):
    """ldpc_decoder_cb.general_work on the GPU == the reference block's general_work, call by call."""
    stream = g[tag + "_stream_re"].astype(np.complex64)
    dec = L.ldpc_decoder_cb(method)

    class Blk:
        def work(self, sym, nout):
            return dec.general_work(sym, nout)

    out, cons, prod = run_chunks(Blk(), stream, g["chunks"], g["nout"])
    assert np.array_equal(cons, g[tag + "_consumed"])
    assert np.array_equal(prod, g[tag + "_produced"])
    assert np.array_equal(out, g[tag + "_bytes"])
    assert dec.take_events() == list(g[tag + "_events"])
    st = dec.state()
    assert [st["state"], st["errors"]] == list(g[tag + "_final"])


def test_block_forecasts_match_reference_build(g):
    enc, dec = L.ldpc_encoder_bc(), L.ldpc_decoder_cb(1)
    assert [dec.forecast(n) for n in (1, 4, 64, 4096)] == list(g["dec_forecast"])
    assert [enc.forecast(n) for n in (1, 16, 17, 64, 640, 4096)] == list(g["enc_forecast"])
    s2, c2 = enc.general_work(g["enc_bytes"][:11], 64 * 2 + 32)
    assert [s2.size, c2] == list(g["enc_ragged"])


# ---------------------------------------------------------------------------------------------
# BASELINE config 4's n = 8192 code against the reference's own decode members (committed outputs)
# ---------------------------------------------------------------------------------------------

@pytest.mark.parametrize("variant", ["1", "0"])
def test_c8k_decisions_match_reference_build(monkeypatch, variant):
    """decodeSumProductSoft / decodeLogDomainSimple of lib/ldpc_decoder_cb_impl.cc on the 4096 x 8192
    matrix (tools/gen_ref_golden_c8k.py) vs the CTA-per-codeword kernels: data bytes of the decision and
    checkFrame(v, M/8), for both (3,6)-regular kernel variants and the generic block kernel (min-sum)."""
    g8 = np.load(os.path.join(GOLD, "ref_build_c8k.npz"))
    monkeypatch.setenv("LDPC535_REGULAR_VARIANT", variant)
    rp, ci, M, N = L.codes.regular_code(8192, 3, 6, int(g8["seed"]))
    code = L.Code((rp, ci, M, N), device=0)
    sym = g8["rx"].astype(np.complex64)
    sent = np.unpackbits(g8["sent"], axis=1)[:, :N]
    assert np.array_equal(code.encode(np.packbits(sent[:, M:], axis=1))[:, :].real.reshape(-1, N) > 0, sent == 1)
    for name, method, iters in (("spa3", 1, 3), ("spa12", 1, 12), ("minsum8", 0, 8)):
        want = np.unpackbits(g8[name + "_vhat"], axis=1)[:, :N]
        b, sy, _ = code.decode(sym, method=method, max_iters=iters, early_stop=True)
        assert np.array_equal(b, np.packbits(want[:, M:], axis=1)), name
        # one byte per window: the C ABI caps the threshold at 253, so the weight saturates at 254
        assert np.array_equal(sy.astype(np.int32), np.minimum(g8[name + "_synd"], 254)), name
    code.close()


@pytest.mark.parametrize("kernel", ["regular-rt", "regular-1024", "block"])
def test_c8k_decisions_match_reference_build_50_iterations(monkeypatch, kernel):
    """The reference's own decodeSumProductSoft run to its exit or to 50 iterations on 36 frames at
    1.0 / 1.5 / 2.0 dB -- 12 of them never converge (tools/gen_ref_golden_c8k_50it.py ->
    ref_build_c8k_50it.npz): data bytes of the decision and checkFrame(v, M/8) identical for every
    CTA-per-codeword kernel; iteration counts equal the restatement's (the reference does not
    return its own)."""
    g = np.load(os.path.join(GOLD, "ref_build_c8k_50it.npz"))
    monkeypatch.setenv("LDPC535_REGULAR_VARIANT", "0" if kernel == "regular-1024" else "1")
    rp, ci, M, N = L.codes.regular_code(8192, 3, 6, int(g["seed"]))
    code = L.Code((rp, ci, M, N), device=0)
    code.set_kernel("block" if kernel == "block" else "regular")
    want = np.unpackbits(g["spa50_vhat"], axis=1)[:, :N]
    b, sy, it = code.decode(g["rx"].astype(np.complex64), method=1, max_iters=50, early_stop=True)
    code.close()
    assert np.array_equal(b, np.packbits(want[:, M:], axis=1))
    # one byte per window: the C ABI caps the threshold at 253, so the weight saturates at 254
    assert np.array_equal(sy.astype(np.int32), np.minimum(g["spa50_synd"], 254))
    assert np.array_equal(it.astype(np.int32), g["iters_oracle"])


# ---------------------------------------------------------------------------------------------
# live: the prebuilt reference library on the GPU box
# ---------------------------------------------------------------------------------------------

@live
@pytest.mark.parametrize("method,ebn0", [(1, 2.0), (1, 4.0), (0, 4.0), (2, 8.0), (3, 8.0)])
def test_live_decisions_20k_frames(c4, method, ebn0):
    """20 000 seeded frames: reference block encodes, reference decode members decide, the CUDA
    decoder must return the same bytes and syndrome weights (5 iterations, reference defaults)."""
    n = 20_000 if method != 2 else 4_000          # the reference's bit flipping is O(M N^2)
    rng = np.random.default_rng(535 + method)
    data = rng.integers(0, 256, 4 * n).astype(np.uint8)
    enc, dec = R.RefEncoder(), R.RefDecoder(method)
    sym, _ = enc.work(data, 64 * n)
    assert np.array_equal(c4.encode(data.reshape(-1, 4)).reshape(-1), sym)
    noisy = util.awgn(sym, ebn0, rng)
    b, sy, _ = c4.decode(noisy, method=method, max_iters=5, early_stop=True)
    rx = noisy.real.astype(np.float64).reshape(n, 64)
    want = np.array([dec.decode(r, method, 5) for r in rx], np.int32)
    wsy = np.array([dec.check_frame(v, 4) for v in want], np.uint8)
    assert np.array_equal(b, util.pack_bits_msb(want[:, 32:]))
    assert np.array_equal(sy, wsy)


@live
def test_live_ber_fer_sweep_matches_reference_simulator_convention(c4):
    """The reference's stand-alone simulator (apps/ldpc_lapack.cpp:533-714) sweeps Eb/N0, sends random
    frames through makeParityCheck -> 2u-1 -> + sqrt(N0) N(0,1) with N0 = 10^(-EbN0/10) (:626-642)
    and counts bit and frame errors of its four decoders (:647-714).  Same experiment here, -2 ... 8 dB
    in 1 dB steps, with the decoders of the reference build on one side and the CUDA decoders on the
    other: info-bit error counts and frame error counts must be EQUAL for every method at every point
    (they are sums over identical decisions), and the curve must behave (monotone FER for sum-product)."""
    enc = R.RefEncoder()
    n_of = {1: 1500, 0: 1500, 3: 1500, 2: 300}        # the reference's bit flipping is O(M N^2)
    rng = np.random.default_rng(5350)
    fer_spa = []
    for ebn0 in range(-2, 9):
        data = rng.integers(0, 256, 4 * 1500).astype(np.uint8)
        sym, _ = enc.work(data, 64 * 1500)
        sigma = np.float32(np.sqrt(10.0 ** (-ebn0 / 10.0)))
        noisy = sym.copy()
        noisy.real += rng.standard_normal(sym.size).astype(np.float32) * sigma
        bits = np.unpackbits(data.reshape(-1, 4), axis=1)
        for method in (1, 0, 2, 3):
            n = n_of[method]
            want, _, _ = R.decode_frames(noisy[:64 * n], method=method, iterations=5, threads=os.cpu_count() or 1)
            got, _, _ = c4.decode(noisy[:64 * n], method=method, max_iters=5, early_stop=True)
            eb_ref = int((np.unpackbits(want, axis=1) != bits[:n]).sum())
            eb_gpu = int((np.unpackbits(got, axis=1) != bits[:n]).sum())
            ef_ref = int((want != data.reshape(-1, 4)[:n]).any(axis=1).sum())
            ef_gpu = int((got != data.reshape(-1, 4)[:n]).any(axis=1).sum())
            assert (eb_gpu, ef_gpu) == (eb_ref, ef_ref), (ebn0, method)
            assert np.array_equal(got, want), (ebn0, method)
            if method == 1:
                fer_spa.append(ef_gpu / n)
    assert fer_spa[0] > 0.95 and fer_spa[-1] < 0.01
    assert all(a >= b - 0.03 for a, b in zip(fer_spa, fer_spa[1:]))


@live
@pytest.mark.parametrize("method", [1, 0, 3])
@pytest.mark.parametrize("ebn0", [None, 3.0, 5.0])
def test_live_block_streams(method, ebn0):
    """Fresh seeded streams through both decoder blocks (GPU block vs reference block)."""
    rng = np.random.default_rng(2000 + method * 10 + int(ebn0 or 9))
    nfr = 600
    data = rng.integers(0, 256, 4 * nfr).astype(np.uint8)
    sym, _ = R.RefEncoder().work(data, 64 * nfr)
    noisy = util.awgn(sym, ebn0, rng)
    lead = (rng.standard_normal(45) * 0.9).astype(np.float32).astype(np.complex64)
    stream = np.concatenate([lead, noisy[:64 * 200], -noisy[64 * 200:64 * 430], noisy[64 * 430:]])
    ref, dec = R.RefDecoder(method), L.ldpc_decoder_cb(method)

    class Blk:
        def work(self, s, nout):
            return dec.general_work(s, nout)

    chunks, nout = (4096, 64, 10000, 65, 130, 777), (256, 4, 700, 7, 9, 50)
    oa, ca, pa = run_chunks(ref, stream, chunks, nout)
    ob, cb, pb = run_chunks(Blk(), stream, chunks, nout)
    assert np.array_equal(ca, cb) and np.array_equal(pa, pb) and np.array_equal(oa, ob)
    assert dec.take_events() == ref.events
    st = dec.state()
    assert (st["state"], st["errors"]) == ref.state
