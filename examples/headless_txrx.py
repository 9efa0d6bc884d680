#!/usr/bin/env python3
"""Headless equivalent of the reference's transmitter.grc + receiver.grc / example6.grc chain
(BASELINE.json config 2):

    file source -> ldpc_encoder_bc -> [AWGN channel] -> ldpc_decoder_cb(method) -> image_sink

The three blocks are the module's own C++ block classes, driven work()-call by work()-call like
the GNU Radio scheduler does (bounded buffers, unconsumed items presented again); the RRC /
USRP / clock-recovery chain of the .grc files is replaced by a symbol-rate AWGN channel in the
reference's own convention (sigma^2 = 10^(-EbN0/10), apps/ldpc_lapack.cpp:635-642).

  python examples/headless_txrx.py image.bmp --ebn0 0 1 2 3 4 --method 1 [--repeat 2]

The file is sent `repeat` times back to back (image_sink writes a file when the NEXT header
arrives).  Prints, per Eb/N0: decoded bytes, byte/bit error rate against the file, sync events,
files written, throughput.  Needs a B200: there is no CPU path.
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gr-ldpc_ece535a_b200", "python"))
import ldpc_ece535a as L   # noqa: E402


def run_chain(payload, ebn0_db, method, seed=535, buf_frames=4096, out_path=None):
    """-> dict(decoded, events, files, seconds).  buf_frames bounds every work() call."""
    if out_path:
        os.environ["LDPC535_IMAGE_PATH"] = out_path
    os.environ.setdefault("LDPC535_IMAGE_DISPLAY", "0")
    enc, dec, sink = L.ldpc_encoder_bc(), L.ldpc_decoder_cb(method), L.image_sink()
    rng = np.random.default_rng(seed)
    sigma = None if ebn0_db is None else np.float32(np.sqrt(10.0 ** (-ebn0_db / 10.0)))
    t0 = time.perf_counter()
    # transmitter: bytes -> symbols, one bounded work() call at a time
    tx, pos = [], 0
    while payload.size - pos >= 4:
        out, consumed = enc.general_work(payload[pos:pos + 4 * buf_frames], 64 * buf_frames)
        if consumed == 0:
            break
        tx.append(out)
        pos += consumed
    sym = np.concatenate(tx) if tx else np.zeros(0, np.complex64)
    # channel
    if sigma is not None:
        sym = sym.copy()
        sym.real += rng.standard_normal(sym.size, dtype=np.float32) * sigma
        sym.imag += rng.standard_normal(sym.size, dtype=np.float32) * sigma
    # receiver: symbols -> bytes -> image sink
    decoded, events, pos = [], [], 0
    while sym.size - pos >= 64:
        out, consumed = dec.general_work(sym[pos:pos + 64 * buf_frames], 4 * buf_frames)
        events += dec.take_events()
        if out.size:
            sink.work(out)
            decoded.append(out)
        if consumed == 0:
            break
        pos += consumed
    dt = time.perf_counter() - t0
    decoded = np.concatenate(decoded) if decoded else np.zeros(0, np.uint8)
    return {"decoded": decoded, "events": events, "files": sink.files_written(), "seconds": dt,
            "state": dec.state(), "symbols": sym.size}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("file")
    ap.add_argument("--ebn0", type=float, nargs="*", default=[0, 1, 2, 3, 4])
    ap.add_argument("--method", type=int, default=1)
    ap.add_argument("--repeat", type=int, default=2)
    ap.add_argument("--out", default="result.bmp")
    args = ap.parse_args()
    data = np.fromfile(args.file, np.uint8)
    payload = np.tile(data, args.repeat)
    print("file %s: %d bytes x %d, method %d" % (args.file, data.size, args.repeat, args.method))
    for e in [None] + list(args.ebn0):
        r = run_chain(payload, e, args.method, out_path=args.out)
        d = r["decoded"]
        n = min(d.size, payload.size)
        # with frame sync from symbol 0 the decoded stream is aligned with the payload
        byte_err = int((d[:n] != payload[:n]).sum()) if n else 0
        bit_err = int(np.unpackbits(d[:n] ^ payload[:n]).sum()) if n else 0
        print("Eb/N0 %-5s decoded %7d bytes  byte errors %6d  BER %.3e  events %-22s files %d  %.1f ms"
              % ("inf" if e is None else "%.1f" % e, d.size, byte_err, bit_err / max(8 * n, 1),
                 r["events"][:6], r["files"], r["seconds"] * 1e3))


if __name__ == "__main__":
    main()
