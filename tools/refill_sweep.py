#!/usr/bin/env python3
"""decode_c4_refill_kernel: throughput vs `refill_min` (lanes of a warp that must be waiting before
its slots are refilled), beside the warp-per-codeword kernel and the lock-step thread kernel, for the
reference block's settings (5 iterations max, early stop) and for 50 iterations max."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gr-ldpc_ece535a_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, torch
import ldpc_ece535a as L
from profile_kernels import timed, synth

stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); sp = C.c_void_p(stream.cuda_stream)
gen = torch.Generator(device="cuda"); gen.manual_seed(1)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
codes = {}
base = L.Code(None, device=0)
for name, env in [("warp", None), ("lockstep", None), ("adaptive", None)] + \
                 [("refill%d" % r, {"LDPC535_C4_REFILL_MIN": str(r)}) for r in (1, 8, 16)]:
    for k, v in (env or {}).items():
        os.environ[k] = v
    c = L.Code(None, device=0)
    for k in (env or {}):
        del os.environ[k]
    c.set_kernel("warp" if name == "warp" else "c4-thread" if name == "lockstep" else "c4-adaptive" if name == "adaptive" else "c4-refill")
    codes[name] = c
ob = torch.empty((n, 4), dtype=torch.uint8, device="cuda"); os_ = torch.empty(n, dtype=torch.uint8, device="cuda"); oi = torch.empty(n, dtype=torch.uint8, device="cuda")
ref = None
for ebn0, iters in ((0.0, 5), (2.0, 5), (3.0, 5), (4.0, 5), (6.0, 5), (8.0, 5), (2.0, 10), (2.0, 50), (6.0, 50)):
    data, sym = synth(base, n, ebn0, sp, gen)
    row, ref = [], None
    for name, c in codes.items():
        ms = timed(stream, lambda: c.decode_dev(sym.data_ptr(), n * 64, n, ob.data_ptr(), os_.data_ptr(), oi.data_ptr(), max_iters=iters, early_stop=True, stream=sp), reps=3)
        sig = (int(ob.view(torch.int32).to(torch.int64).sum().item()), int(oi.to(torch.int64).sum().item()), int(os_.to(torch.int64).sum().item()))
        if ref is None:
            ref = sig
        row.append("%s %.1f%s" % (name, n * 32 / ms / 1e6, "" if sig == ref else " MISMATCH"))
    mean_it = oi.float().mean().item()
    print("Eb/N0 %.0f dB max_iters %2d mean iters %5.2f Gbit/s: %s" % (ebn0, iters, mean_it, " | ".join(row)), flush=True)
    del data, sym
