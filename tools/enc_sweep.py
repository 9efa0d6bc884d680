#!/usr/bin/env python3
"""Look-up encoder (encode_m4r_kernel) on the n = 8192 code: ring shapes, tile sizes and -- with the
timing-experiment build (tools/ab/libldpc535_exp.so, LDPC535_M4R_FLAGS) -- the kernel with parts
switched off, device-resident frames, CUDA-event times.  Every configuration that computes the
real thing is first checked against the AND/XOR scan encoder on a ragged batch.

  python tools/enc_sweep.py [--frames 200000,400000] [--configs "ring=42;ring=62;ring=63;ring=42,tile=1024"]
"""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gr-ldpc_ece535a_b200", "python"))

import torch              # noqa: E402
import ldpc_ece535a as L  # noqa: E402

ENV = {"ring": "LDPC535_M4R_RING", "tpf": "LDPC535_M4R_TPF", "tile": "LDPC535_M4R_TILE", "flags": "LDPC535_M4R_FLAGS", "enc": "LDPC535_ENCODER"}


def make(base, cfg):
    for k in ENV.values():
        os.environ.pop(k, None)
    for kv in cfg.split(","):
        if kv:
            k, v = kv.split("=")
            os.environ[ENV[k]] = v
    code = L.Code(base.h_csr() + (base.M, base.N), device=0)
    for k in ENV.values():
        os.environ.pop(k, None)
    return code


def timed(stream, fn, reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn()
    torch.cuda.synchronize()
    a.record(stream)
    for _ in range(reps):
        fn()
    b.record(stream)
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", default="200000,400000")
    ap.add_argument("--configs", default="ring=42,tile=1024;ring=42;ring=62;ring=63")
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    frames = [int(x) for x in args.frames.split(",")]
    stream = torch.cuda.Stream()          # the library launches on the stream it is handed; events go on the same one
    torch.cuda.set_stream(stream)
    sp = C.c_void_p(stream.cuda_stream)
    base, _ = L.codes.first_invertible(n=8192, seed=535, device=0)
    scan = make(base, "enc=generic")
    gen = torch.Generator(device="cuda")
    gen.manual_seed(19)
    nmax = max(frames + [6000])
    data = torch.randint(0, 256, (nmax, base.nbytes), dtype=torch.uint8, device="cuda", generator=gen)
    out = torch.empty((nmax, base.N, 2), dtype=torch.float32, device="cuda")
    ncheck = 148 * 32 + 1077
    want = torch.empty((ncheck, base.N, 2), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    scan.encode_dev(data.data_ptr(), ncheck, want.data_ptr(), stream=sp)
    torch.cuda.synchronize()
    for cfg in args.configs.split(";"):
        code = make(base, cfg)
        line = "%-28s" % cfg
        if "flags" not in cfg:
            out[:ncheck].zero_()
            torch.cuda.synchronize()
            code.encode_dev(data.data_ptr(), ncheck, out.data_ptr(), stream=sp)
            torch.cuda.synchronize()
            line += " check %s |" % ("ok" if torch.equal(out[:ncheck], want) else "MISMATCH")
        for n in frames:
            ms = timed(stream, lambda: code.encode_dev(data.data_ptr(), n, out.data_ptr(), stream=sp), args.reps)
            line += " n=%d: %.3f ms %.0f GB/s |" % (n, ms, n * 66048 / ms / 1e6)
        print(line, flush=True)
        code.close()


if __name__ == "__main__":
    main()
