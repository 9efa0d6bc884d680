#!/usr/bin/env python3
"""Freeze oracle outputs for the shipped 32x64 code as a small golden fixture.

The reference's own tests pin nothing for this code, so these goldens come
from the oracle AFTER it reproduced the reference's two QA known-answer tests on the
8x16 code (tests/test_oracle_kat.py).  CPU tests check the oracle still reproduces the
fixture (drift guard); GPU tests check the CUDA path against it.

Writes tests/golden/c4_golden.npz.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as O   # noqa: E402
import util                      # noqa: E402

POINTS = (("2dB_5it", 2.0, 5, 1), ("4dB_5it", 4.0, 5, 1), ("2dB_50it_noearly", 2.0, 50, 0),
          ("clean_5it", None, 5, 1))


def main():
    codes = O.load_ref_codes()
    H = codes["shipped"]["H"]
    Hp, L, U, piv = O.reorder_h(H)
    out = {"pivots": piv}
    # encoder: the reference's own deterministic source words (apps/test_data.h:180-213)
    src = codes["shipped"]["source_words"]
    out["enc_data_bits"] = src.astype(np.uint8)
    out["enc_codewords"] = util.oracle_encode_bits(src, Hp, L, U).astype(np.uint8)
    # decoder: 64 seeded frames per operating point
    for k, (tag, ebn0, iters, early) in enumerate(POINTS):
        data, cw, sym = util.synth_frames(Hp, L, U, 64, ebn0, seed=5350 + k)
        for method, mname in ((1, "spa"), (0, "minsum")):
            b, it, sy, _ = O.decode_frames(sym, Hp, method=method, iterations=iters, early_stop=early)
            out["dec_%s_%s_bytes" % (tag, mname)] = b
            out["dec_%s_%s_iters" % (tag, mname)] = it
            out["dec_%s_%s_synd" % (tag, mname)] = sy
        out["dec_%s_sym" % tag] = sym
        out["dec_%s_data" % tag] = data.astype(np.uint8)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "c4_golden.npz"), **out)
    print("wrote c4_golden.npz with", len(out), "arrays")


if __name__ == "__main__":
    main()
