#!/usr/bin/env python
"""Generate tests/golden/ref_build_c8k.npz: the reference's OWN code (oracle/_ref, see
tools/gen_ref_golden.py) run on BASELINE config 4's (3,6)-regular n = 8192 code.

The reference blocks are hard-wired to 32x64; oracle/ref_driver.cc swaps the matrix in and calls
the reference's private members, so this is reorderHMatrix, decodeSumProductSoft,
decodeLogDomainSimple and checkFrame of lib/ldpc_decoder_cb_impl.cc on 4096 x 8192 dense
matrices: ~10 s for the re-ordering, ~4 s per sum-product iteration and codeword.  Run in the
build container (a few minutes):

    make -C oracle && python tools/gen_ref_golden_c8k.py

Stored (small): SHA-256 of the re-ordered H and of the L, U factors (the matrices themselves are
4 MB each bit-packed), their non-zero counts; three noisy codewords (fp32 real parts) and the
reference's decisions for them (sum-product at 3 and 12 iterations, min-sum at 8), bit-packed,
with checkFrame(v, M/8).  The reference's encoder is NOT run at this size: its real-valued dgesv
detour overflows for n = 8192; codewords come from the oracle's GF(2) solve of the
reference's own L, U factors and are checked against H here.
"""
import hashlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gr-ldpc_ece535a_b200", "python"))

from oracle import oracle as O          # noqa: E402
from oracle import ref as R             # noqa: E402
from ldpc_ece535a import codes          # noqa: E402

SEED = 535


def sha(a):
    return hashlib.sha256(np.packbits(np.asarray(a, np.uint8), axis=1).tobytes()).hexdigest()


def main():
    assert R.available()
    rp, ci, M, N = codes.regular_code(8192, 3, 6, SEED)
    H = codes.to_dense(rp, ci, M, N)
    t0 = time.time()
    dec = R.RefDecoder(1)
    Hp, L, U = dec.set_code(H)
    print("reference reorderHMatrix: %.1f s" % (time.time() - t0), flush=True)
    g = {"seed": np.array(SEED), "Hp_sha256": np.array(sha(Hp)), "L_sha256": np.array(sha(L)),
         "U_sha256": np.array(sha(U)), "L_nnz": np.array(int(L.sum())), "U_nnz": np.array(int(U.sum())),
         "Hp_col_weight_head": Hp.sum(0)[:64].astype(np.int8)}

    rng = np.random.default_rng(8192)
    d = rng.integers(0, 2, (3, N - M)).astype(np.int32)
    cw = np.array([np.concatenate([O.make_parity_check(x, Hp, L, U, gf2=True)[0], x]) for x in d])
    assert not ((cw @ Hp.T) % 2).any()
    rx = np.empty(cw.shape, np.float32)
    for f, ebn0 in enumerate((2.0, 2.0, 4.0)):
        sigma = np.sqrt(10.0 ** (-ebn0 / 10.0))
        rx[f] = ((2.0 * cw[f] - 1.0) + rng.standard_normal(N) * sigma).astype(np.float32)
    g["rx"] = rx
    g["sent"] = np.packbits(cw.astype(np.uint8), axis=1)
    for name, method, iters in (("spa3", 1, 3), ("spa12", 1, 12), ("minsum8", 0, 8)):
        vs, ws = [], []
        for f in range(3):
            t0 = time.time()
            v = dec.decode(rx[f].astype(np.float64), method, iters)
            w = dec.check_frame(v, M // 8)
            print("%s frame %d: %.1f s, %d bit errors, checkFrame %d" % (name, f, time.time() - t0,
                                                                        int((v != cw[f]).sum()), w), flush=True)
            vs.append(np.packbits(v.astype(np.uint8)))
            ws.append(w)
        g[name + "_vhat"] = np.array(vs)
        g[name + "_synd"] = np.array(ws, np.int32)
    path = os.path.join(ROOT, "tests", "golden", "ref_build_c8k.npz")
    np.savez_compressed(path, **g)
    print("wrote %s, %d bytes" % (path, os.path.getsize(path)))


if __name__ == "__main__":
    main()
