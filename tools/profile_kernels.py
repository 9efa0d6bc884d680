#!/usr/bin/env python3
"""One launch of every hot kernel on device-resident synthetic data, for ncu captures
(profiles/).  Prints CUDA-event times so the same command can be checked without ncu.

  python tools/profile_kernels.py [--which decode_c4,decode_warp,decode_block,encode_small,encode_generic]
"""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gr-ldpc_ece535a_b200", "python"))

import numpy as np       # noqa: E402
import torch             # noqa: E402
import ldpc_ece535a as L  # noqa: E402


def timed(stream, fn, reps=1):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn()
    torch.cuda.synchronize()
    a.record(stream)
    for _ in range(reps):
        fn()
    b.record(stream)
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def synth(code, n, ebn0, sp, gen):
    data = torch.randint(0, 256, (n, code.nbytes), dtype=torch.uint8, device="cuda", generator=gen)
    sym = torch.empty((n, code.N, 2), dtype=torch.float32, device="cuda")
    code.encode_dev(data.data_ptr(), n, sym.data_ptr(), stream=sp)
    sigma = float(np.sqrt(10.0 ** (-ebn0 / 10.0)))
    sym[:, :, 0] += sigma * torch.randn((n, code.N), device="cuda", generator=gen)
    torch.cuda.synchronize()
    return data, sym


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--which", default="decode_c4,decode_warp,methods,decode_block,encode_small,encode_generic")
    ap.add_argument("--c4-codewords", type=int, default=2_000_000)
    ap.add_argument("--c8k-codewords", type=int, default=20_000)
    args = ap.parse_args()
    which = args.which.split(",")
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    sp = C.c_void_p(stream.cuda_stream)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(535)

    c4 = L.Code(None, device=0)
    n = args.c4_codewords
    data, sym = synth(c4, n, 2.0, sp, gen)
    ob = torch.empty((n, 4), dtype=torch.uint8, device="cuda")
    os_ = torch.empty(n, dtype=torch.uint8, device="cuda")
    oi = torch.empty(n, dtype=torch.uint8, device="cuda")
    for kern, tag in (("c4-thread", "decode_c4"), ("warp", "decode_warp")):
        if tag not in which:
            continue
        c4.set_kernel(kern)
        for iters, early in ((50, False), (5, True)):
            ms = timed(stream, lambda: c4.decode_dev(sym.data_ptr(), n * 64, n, ob.data_ptr(), os_.data_ptr(),
                                                     oi.data_ptr(), max_iters=iters, early_stop=early, stream=sp))
            it = oi.float().mean().item()
            print("%-14s C4 n=%d iters=%d early=%d: %.3f ms  %.3f Gbit/s info  mean iters %.2f  %.3e edge-it/s"
                  % (tag, n, iters, early, ms, n * 32 / ms / 1e6, it, n * 168 * it / ms * 1e3))
    c4.set_kernel(None)
    if "methods" in which:
        for method, name in ((0, "min-sum fp64"), (2, "bit-flip"), (3, "hard")):
            for iters, early in ((5, True), (50, False)):
                ms = timed(stream, lambda: c4.decode_dev(sym.data_ptr(), n * 64, n, ob.data_ptr(), os_.data_ptr(),
                                                         oi.data_ptr(), method=method, max_iters=iters,
                                                         early_stop=early, stream=sp))
                print("method %d %-13s C4 n=%d iters=%d early=%d (%s kernel): %.3f ms  %.3f Gbit/s info  %.1f GB/s"
                      % (method, name, n, iters, early, c4.kernel_name(method), ms, n * 32 / ms / 1e6, n * 518 / ms / 1e6))
    if "encode_small" in which:
        ne = 10_000_000
        d = torch.randint(0, 256, (ne, 4), dtype=torch.uint8, device="cuda", generator=gen)
        o = torch.empty((ne, 64, 2), dtype=torch.float32, device="cuda")
        ms = timed(stream, lambda: c4.encode_dev(d.data_ptr(), ne, o.data_ptr(), stream=sp))
        print("encode_small   C4 n=%d: %.3f ms  %.1f GB/s (516 B/frame)  %.2f Gbit/s info"
              % (ne, ms, ne * 516 / ms / 1e6, ne * 32 / ms / 1e6))
        del d, o
    del sym, data

    if "decode_block" in which or "encode_generic" in which:
        c8, seed = L.codes.first_invertible(n=8192, seed=535, device=0)
        n8 = args.c8k_codewords
        data8, sym8 = synth(c8, n8, 2.0, sp, gen)
        if "encode_generic" in which:
            o = torch.empty((n8, 8192, 2), dtype=torch.float32, device="cuda")
            ms = timed(stream, lambda: c8.encode_dev(data8.data_ptr(), n8, o.data_ptr(), stream=sp))
            print("encode_generic C8k n=%d: %.3f ms  %.1f GB/s (66048 B/frame)  %.2f Gbit/s info"
                  % (n8, ms, n8 * 66048 / ms / 1e6, n8 * 4096 / ms / 1e6))
            del o
        if "decode_block" in which:
            ob = torch.empty((n8, 512), dtype=torch.uint8, device="cuda")
            os_ = torch.empty(n8, dtype=torch.uint8, device="cuda")
            oi = torch.empty(n8, dtype=torch.uint8, device="cuda")
            for iters, early in ((50, True), (50, False)):
                nn = n8 if early else n8 // 4
                ms = timed(stream, lambda: c8.decode_dev(sym8.data_ptr(), nn * 8192, nn, ob.data_ptr(),
                                                         os_.data_ptr(), oi.data_ptr(), max_iters=iters,
                                                         early_stop=early, stream=sp))
                it = oi[:nn].float().mean().item()
                okf = (ob[:nn] == data8[:nn]).all(dim=1).float().mean().item()
                print("decode_block   C8k n=%d iters=%d early=%d: %.3f ms  %.3f Gbit/s info  mean iters %.2f  "
                      "%.3e edge-it/s  frames ok %.4f" % (nn, iters, early, ms, nn * 4096 / ms / 1e6, it,
                                                         nn * 24576 * it / ms * 1e3, okf))


if __name__ == "__main__":
    main()
