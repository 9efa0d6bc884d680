#!/usr/bin/env python3
"""Early-stop throughput of the C4 kernels versus Eb/N0 (which family should early-stop mode use?)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gr-ldpc_ece535a_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, torch
import ldpc_ece535a as L
from profile_kernels import timed, synth

stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); sp = C.c_void_p(stream.cuda_stream)
gen = torch.Generator(device="cuda"); gen.manual_seed(1)
c4 = L.Code(None, device=0)
n = 2_000_000
ob = torch.empty((n, 4), dtype=torch.uint8, device="cuda"); os_ = torch.empty(n, dtype=torch.uint8, device="cuda"); oi = torch.empty(n, dtype=torch.uint8, device="cuda")
for ebn0 in (0.0, 2.0, 4.0, 6.0, 8.0):
    data, sym = synth(c4, n, ebn0, sp, gen)
    for iters in (5, 50):
        row = []
        for kern in ("c4-thread", "warp"):
            c4.set_kernel(kern)
            ms = timed(stream, lambda: c4.decode_dev(sym.data_ptr(), n * 64, n, ob.data_ptr(), os_.data_ptr(), oi.data_ptr(), max_iters=iters, early_stop=True, stream=sp), reps=3)
            row.append("%s %.3f ms %.1f Gbit/s" % (kern, ms, n * 32 / ms / 1e6))
        print("Eb/N0 %.0f dB max_iters %2d mean iters %.2f: %s" % (ebn0, iters, oi.float().mean().item(), " | ".join(row)))
    del data, sym
