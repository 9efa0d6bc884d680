#!/usr/bin/env python3
"""Where does a launch of the thread-per-codeword kernel spend its time?  ms versus the iteration
count at fixed iterations (early stop off / on at 0 dB, where nothing converges): the slope is the
cost of one iteration, the intercept the per-codeword start-up (load, first messages) and output."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gr-ldpc_ece535a_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, torch
import ldpc_ece535a as L
from profile_kernels import timed, synth

stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); sp = C.c_void_p(stream.cuda_stream)
gen = torch.Generator(device="cuda"); gen.manual_seed(1)
c4 = L.Code(None, device=0)
n = 148 * 256 * 40
ob = torch.empty((n, 4), dtype=torch.uint8, device="cuda"); os_ = torch.empty(n, dtype=torch.uint8, device="cuda"); oi = torch.empty(n, dtype=torch.uint8, device="cuda")
data, sym = synth(c4, n, 2.0, sp, gen)
for kern in sys.argv[1:] or ["c4-thread", "warp"]:
    c4.set_kernel(kern)
    for early in (False, True):
        row = []
        for iters in (1, 2, 3, 5, 10, 20, 50):
            ms = timed(stream, lambda: c4.decode_dev(sym.data_ptr(), n * 64, n, ob.data_ptr(), os_.data_ptr(), oi.data_ptr(), max_iters=iters, early_stop=early, stream=sp), reps=3)
            row.append((iters, ms, oi.float().mean().item()))
        per_batch = [(it, ms * 1e3 / 40, mi) for it, ms, mi in row]
        print("%-10s early=%d  us per 256-codeword batch per SM (mean iters): %s" % (kern, early, "  ".join("%d: %.1f (%.2f)" % x for x in per_batch)))
