#!/usr/bin/env python
"""Generate tests/golden/ref_build_golden.npz from oracle/_ref -- the reference's OWN block
sources (lib/ldpc_decoder_cb_impl.cc, lib/ldpc_encoder_bc_impl.cc) compiled where they lie under
/root/reference against the stand-ins in oracle/refshim/ (see oracle/ref_driver.cc).

Run in the build container (the only place /root/reference exists):

    make -C oracle && python tools/gen_ref_golden.py

Everything stored is an OUTPUT OF THE REFERENCE'S CODE on seeded inputs (the inputs are stored
too, so nothing depends on a random generator staying stable):

  tab_<code>_{Hp,L,U}      reorderHMatrix on the five matrices of apps/test_data.h
  enc_*                    encoder general_work on the shipped 32x64 code (bytes -> symbols)
  par_<code>_{d,c}         makeParityCheck for every reference matrix
  cw_<tag>_*               the four private decode members, one codeword at a time, 5 and 50
                           iterations, Eb/N0 0..6 dB + noiseless (decisions; the syndrome weight
                           checkFrame(v, M/8) returns for them)
  blk_m<method>_*          decoder general_work through scheduler-style chunks on a stream with a
                           lead-in, an inverted stretch and noise: bytes, consumed/produced per
                           call, the std::cout sync lines as event codes, final d_state/d_errors
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import oracle as O          # noqa: E402  (only for the code fixtures)
from oracle import ref as R             # noqa: E402

CODES = ["hData1", "hData2", "hData3", "hData4", "hData5"]
CHUNKS = (4099, 640, 63, 64, 65, 1000, 129, 8192)   # symbols offered per scheduler call
NOUT = (200, 4, 3, 64, 1, 500, 16, 1000)            # output room per call


def awgn_real(sym, ebn0_db, rng):
    """apps/ldpc_lapack.cpp:626-642 convention: sigma^2 = 10^(-EbN0/10), real axis."""
    if ebn0_db is None:
        return sym.astype(np.complex64)
    sigma = np.float32(np.sqrt(10.0 ** (-ebn0_db / 10.0)))
    return (sym + rng.standard_normal(sym.shape).astype(np.float32) * sigma).astype(np.complex64)


def block_stream(dec, sym):
    """Drive general_work like the runtime: re-present what was not consumed."""
    pos, k = 0, 0
    out, consumed, produced = [], [], []
    idle = 0
    while pos < sym.size and idle < len(CHUNKS):
        n, room = CHUNKS[k % len(CHUNKS)], NOUT[k % len(NOUT)]
        k += 1
        o, c = dec.work(sym[pos:pos + n], room)
        out.append(o.copy())
        consumed.append(c)
        produced.append(len(o))
        pos += c
        idle = idle + 1 if (c == 0 and len(o) == 0) else 0
    return np.concatenate(out), np.array(consumed, np.int64), np.array(produced, np.int64)


def main():
    assert R.available(), "build oracle/_ref first (make -C oracle, needs /root/reference)"
    codes = O.load_ref_codes()
    g = {}
    rng = np.random.default_rng(535)

    enc = R.RefEncoder()
    dec = R.RefDecoder(1)
    Hs, Ls, Us = enc.tables()
    assert np.array_equal(dec.H, Hs)
    g["shipped_Hp"], g["shipped_L"], g["shipped_U"] = (a.astype(np.int8) for a in (Hs, Ls, Us))

    # encoder block on the shipped code (as constructed, nothing swapped in)
    data = rng.integers(0, 256, 4 * 512).astype(np.uint8)
    sym, consumed = enc.work(data, 64 * 512)
    assert consumed == data.size and sym.size == 64 * 512
    assert np.all(sym.imag == 0) and np.all(np.abs(sym.real) == 1)
    g["enc_bytes"] = data
    g["enc_symbols_re"] = sym.real.astype(np.int8)
    # ragged call: room for 2.5 frames, 11 bytes offered
    s2, c2 = enc.work(data[:11], 64 * 2 + 32)
    g["enc_ragged"] = np.array([s2.size, c2], np.int64)
    g["enc_forecast"] = np.array([enc.forecast(n) for n in (1, 16, 17, 64, 640, 4096)], np.int64)
    g["dec_forecast"] = np.array([dec.forecast(n) for n in (1, 4, 64, 4096)], np.int64)

    # tables + parity for every reference matrix
    for name in CODES:
        H = codes[name]["H"]
        e = R.RefEncoder()
        Hp, L, U = e.set_code(H)
        g["tab_%s_Hp" % name], g["tab_%s_L" % name], g["tab_%s_U" % name] = (
            a.astype(np.int8) for a in (Hp, L, U))
        d = rng.integers(0, 2, (64, H.shape[1] - H.shape[0])).astype(np.int32)
        g["par_%s_d" % name] = d.astype(np.int8)
        g["par_%s_c" % name] = np.array([e.make_parity(x) for x in d], np.int8)
        e.close()

    # codeword level, shipped code: all four private decode members
    cw_bits = np.concatenate([(sym.real.reshape(-1, 64) > 0).astype(np.int8)], axis=0)
    for tag, eb in (("clean", None), ("0dB", 0.0), ("2dB", 2.0), ("4dB", 4.0), ("6dB", 6.0)):
        n = 48
        rx = awgn_real(sym[:64 * n], eb, rng).real.astype(np.float32).reshape(n, 64)
        g["cw_%s_rx" % tag] = rx
        for iters in (5, 50):
            for method in (0, 1, 2, 3):
                if method == 3 and iters == 50:
                    continue
                v = np.array([dec.decode(r.astype(np.float64), method, iters) for r in rx], np.int8)
                sw = np.array([dec.check_frame(x, 4) for x in v], np.int8)
                g["cw_%s_m%d_it%d_vhat" % (tag, method, iters)] = v
                g["cw_%s_m%d_it%d_synd" % (tag, method, iters)] = sw
    g["cw_sent_bits"] = cw_bits[:48]

    # codeword level, the 8x16 QA code (hData3) and the 16x32 one: other degrees
    for name in ("hData3", "hData5", "hData2"):
        H = codes[name]["H"]
        M, N = H.shape
        d2 = R.RefDecoder(1)
        e2 = R.RefEncoder()
        d2.set_code(H)
        e2.set_code(H)
        n = 32
        db = rng.integers(0, 2, (n, N - M)).astype(np.int32)
        cb = np.array([np.concatenate([e2.make_parity(x), x]) for x in db], np.int32)
        rx = awgn_real((2.0 * cb - 1.0).astype(np.complex64), 3.0, rng).real.astype(np.float32)
        g["cwx_%s_rx" % name] = rx
        for method in (0, 1, 2, 3):
            for iters in (5, 20):
                v = np.array([d2.decode(r.astype(np.float64), method, iters) for r in rx], np.int8)
                g["cwx_%s_m%d_it%d_vhat" % (name, method, iters)] = v
        d2.close()
        e2.close()

    # block level: lead-in + inverted stretch + upright stretch, AWGN, chunked calls
    nfr = 160
    base = sym[:64 * nfr]
    for method, eb in ((0, 4.0), (1, 4.0), (1, 1.0), (2, 7.0), (3, 7.0)):
        lead = (rng.standard_normal(37) * 0.8).astype(np.float32).astype(np.complex64)
        noisy = awgn_real(base, eb, rng)
        stream = np.concatenate([lead, -noisy[:64 * 60], noisy[64 * 60:]]).astype(np.complex64)
        b = R.RefDecoder(method)
        out, cons, prod = block_stream(b, stream)
        tag = "blk_m%d_%ddB" % (method, int(eb))
        g[tag + "_stream_re"] = stream.real.astype(np.float32)
        g[tag + "_bytes"] = out
        g[tag + "_consumed"] = cons
        g[tag + "_produced"] = prod
        g[tag + "_events"] = np.array(b.events, np.int8)
        g[tag + "_final"] = np.array(b.state, np.int64)
        b.close()

    g["chunks"] = np.array(CHUNKS, np.int64)
    g["nout"] = np.array(NOUT, np.int64)
    path = os.path.join(ROOT, "tests", "golden", "ref_build_golden.npz")
    np.savez_compressed(path, **g)
    print("wrote %s: %d arrays, %d bytes" % (path, len(g), os.path.getsize(path)))


if __name__ == "__main__":
    main()
