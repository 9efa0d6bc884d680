#!/usr/bin/env python
"""Generate tests/golden/ref_build_c8k_50it.npz: the reference's OWN decodeSumProductSoft
(oracle/_ref, lib/ldpc_decoder_cb_impl.cc:478-557 compiled unmodified) run to its exit or to
50 iterations on BASELINE config 4's (3,6)-regular n = 8192 code, on frames that converge AND
on frames that do not.

36 frames: 12 each at Eb/N0 = 1.0 dB (below the code's threshold: none converges, every one
runs all 50 iterations), 1.5 dB (late convergence, 18-45 iterations, some failures) and 2.0 dB
(config 4's operating point).  The reference scans dense 4096 x 8192 matrices: ~4 s per
iteration and codeword, so this takes ~10 minutes on 8 cores (one reference decoder per
process):

    make -C oracle && python tools/gen_ref_golden_c8k_50it.py

Stored: the noisy real parts (fp32), the transmitted words, and for every frame the reference's
50-iteration decision (bit-packed) and checkFrame(v, M/8).  The reference does not return its
iteration count; `iters_oracle` is the sparse restatement's (oracle/ldpc_oracle.c), stored so
the CPU suite can check restatement == reference on the decisions and the GPU suite can check
decisions against the reference and iteration counts against the restatement on the same frames.
"""
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gr-ldpc_ece535a_b200", "python"))

SEED = 535
EBN0 = (1.0, 1.5, 2.0)
PER_SNR = 12
ITERS = 50

_dec = None
_M = None


def _init():
    global _dec, _M
    from oracle import ref as R
    from ldpc_ece535a import codes
    rp, ci, M, N = codes.regular_code(8192, 3, 6, SEED)
    _dec = R.RefDecoder(1)
    _dec.set_code(codes.to_dense(rp, ci, M, N))
    _M = M


def _work(args):
    f, rx = args
    t0 = time.time()
    v = _dec.decode(rx.astype(np.float64), 1, ITERS)
    w = _dec.check_frame(v, _M // 8)
    print("frame %d: %.0f s, checkFrame %d" % (f, time.time() - t0, w), flush=True)
    return f, np.packbits(v.astype(np.uint8)), w


def main():
    from oracle import oracle as O
    from oracle import ref as R
    import ldpc_ece535a as L
    from ldpc_ece535a import codes
    assert R.available()
    code = L.Code(codes.regular_code(8192, 3, 6, SEED), device=L.DEVICE_NONE if hasattr(L, "DEVICE_NONE") else -1)
    M, N, K = code.M, code.N, code.K
    row_ptr, col_idx = code.h_csr()
    P = code.generator()
    Pbits = np.unpackbits(P.view(np.uint8).reshape(M, -1), axis=1, bitorder="little")[:, :K].astype(np.int64)
    rows = np.repeat(np.arange(M), np.diff(row_ptr))
    order = np.lexsort((rows, col_idx)).astype(np.int32)
    col_ptr = np.zeros(N + 1, np.int32)
    np.add.at(col_ptr, col_idx + 1, 1)
    tables = (row_ptr, col_idx, np.cumsum(col_ptr).astype(np.int32), order)

    rng = np.random.default_rng(81920)
    n = len(EBN0) * PER_SNR
    d = rng.integers(0, 2, (n, K)).astype(np.int64)
    cw = np.concatenate([(d @ Pbits.T) & 1, d], axis=1)
    synd = np.zeros((n, M), np.int64)
    np.add.at(synd.T, rows, cw[:, col_idx].T)
    assert not (synd & 1).any()
    rx = np.empty((n, N), np.float32)
    ebn0 = np.repeat(np.array(EBN0), PER_SNR)
    for f in range(n):
        sigma = np.sqrt(10.0 ** (-ebn0[f] / 10.0))
        rx[f] = ((2.0 * cw[f] - 1.0) + rng.standard_normal(N) * sigma).astype(np.float32)

    iters_o = np.zeros(n, np.int32)
    vhat_o = np.zeros((n, N // 8), np.uint8)
    for f in range(n):
        v, run = O.decode_spa_sparse(rx[f], tables, M, N, ITERS, True)
        iters_o[f] = run
        vhat_o[f] = np.packbits(v.astype(np.uint8))
    print("restatement iteration counts:", iters_o.tolist(), flush=True)

    # longest frames first so the pool drains evenly
    jobs = sorted(((f, rx[f]) for f in range(n)), key=lambda a: -iters_o[a[0]])
    vhat = np.zeros((n, N // 8), np.uint8)
    wts = np.zeros(n, np.int32)
    t0 = time.time()
    with mp.Pool(min(8, os.cpu_count()), initializer=_init) as pool:
        for f, v, w in pool.imap_unordered(_work, jobs):
            vhat[f] = v
            wts[f] = w
    print("reference build: %.0f s" % (time.time() - t0))
    print("restatement == reference on all decisions:", bool(np.array_equal(vhat, vhat_o)))
    path = os.path.join(ROOT, "tests", "golden", "ref_build_c8k_50it.npz")
    np.savez_compressed(path, seed=np.array(SEED), ebn0=ebn0.astype(np.float32), rx=rx,
                        sent=np.packbits(cw.astype(np.uint8), axis=1), spa50_vhat=vhat,
                        spa50_synd=wts, iters_oracle=iters_o)
    print("wrote %s, %d bytes" % (path, os.path.getsize(path)))


if __name__ == "__main__":
    main()
