#!/usr/bin/env python3
"""One timed launch of a decoder family on the shipped code for ncu captures:
   python tools/warp_one.py <kernel> <method> <Eb/N0 dB> <max_iters> <early_stop> [codewords]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gr-ldpc_ece535a_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
import ldpc_ece535a as L
from profile_kernels import timed, synth

kern, method, ebn0, iters, early = sys.argv[1], int(sys.argv[2]), float(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
n = int(sys.argv[6]) if len(sys.argv) > 6 else 4_000_000
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); sp = C.c_void_p(stream.cuda_stream)
gen = torch.Generator(device="cuda"); gen.manual_seed(1)
c = L.Code(None, device=0)
if kern != "auto":
    c.set_kernel(kern)
data, sym = synth(c, n, ebn0, sp, gen)
ob = torch.empty((n, 4), dtype=torch.uint8, device="cuda"); os_ = torch.empty(n, dtype=torch.uint8, device="cuda"); oi = torch.empty(n, dtype=torch.uint8, device="cuda")
ms = timed(stream, lambda: c.decode_dev(sym.data_ptr(), n * 64, n, ob.data_ptr(), os_.data_ptr(), oi.data_ptr(), method=method,
                                        max_iters=iters, early_stop=bool(early), stream=sp), reps=3)
print("%s method %d %.0f dB iters %d early %d: %.3f ms %.2f Gbit/s mean iters %.2f (%s)"
      % (kern, method, ebn0, iters, early, ms, n * 32 / ms / 1e6, oi.float().mean().item(), c.kernel_name(method)))
