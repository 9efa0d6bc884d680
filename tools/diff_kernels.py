#!/usr/bin/env python3
"""Find frames on which the C4 kernel families disagree (device-resident, 10 M frames) and ask the oracle."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gr-ldpc_ece535a_b200", "python")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, torch
import ldpc_ece535a as L
from oracle import oracle as O
from profile_kernels import synth

stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); sp = C.c_void_p(stream.cuda_stream)
gen = torch.Generator(device="cuda"); gen.manual_seed(1)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
ebn0 = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
codes = {}
for name, env in [("warp", None), ("lockstep", None), ("refill8", {"LDPC535_C4_REFILL_MIN": "8"}), ("refill1", {"LDPC535_C4_REFILL_MIN": "1"})]:
    for k, v in (env or {}).items():
        os.environ[k] = v
    c = L.Code(None, device=0)
    for k in (env or {}):
        del os.environ[k]
    c.set_kernel("warp" if name == "warp" else "c4-thread" if name == "lockstep" else "c4-refill")
    codes[name] = c
data, sym = synth(codes["warp"], n, ebn0, sp, gen)
res = {}
for name, c in codes.items():
    ob = torch.full((n, 4), 0xEE, dtype=torch.uint8, device="cuda"); os_ = torch.full((n,), 0xEE, dtype=torch.uint8, device="cuda"); oi = torch.full((n,), 0xEE, dtype=torch.uint8, device="cuda")
    c.decode_dev(sym.data_ptr(), n * 64, n, ob.data_ptr(), os_.data_ptr(), oi.data_ptr(), max_iters=iters, early_stop=True, stream=sp)
    torch.cuda.synchronize()
    res[name] = (ob, os_, oi)
Hp, Lm, Um, _ = O.reorder_h(O.load_ref_codes()["shipped"]["H"])
base = res["warp"]
for name in ("lockstep", "refill8", "refill1"):
    r = res[name]
    bad = ((r[0] != base[0]).any(dim=1) | (r[1] != base[1]) | (r[2] != base[2])).nonzero().flatten()
    print("%s vs warp: %d frames differ of %d" % (name, bad.numel(), n))
    for i in bad[:6].tolist():
        fr = sym[i, :, 0].cpu().numpy().astype(np.float64)
        v, run = O.decode_spa(fr, Hp, iters, True)
        wb = np.packbits(v[32:].astype(np.uint8)); ws = O.check_frame(v, Hp, 4)
        print("  frame %d: warp bytes %s synd %d it %d | %s bytes %s synd %d it %d | oracle bytes %s synd %d it %d" % (
            i, base[0][i].tolist(), int(base[1][i]), int(base[2][i]), name, r[0][i].tolist(), int(r[1][i]), int(r[2][i]), wb.tolist(), ws, run))

# message-level hunt on the first differing frame: where do the two kernels' messages first part?
r = res["lockstep"]
bad = ((r[0] != base[0]).any(dim=1) | (r[1] != base[1]) | (r[2] != base[2])).nonzero().flatten()
if bad.numel():
    i = int(bad[0])
    fr = sym[i].cpu().numpy().view(np.complex64).reshape(1, 64)
    np.save(os.path.join(ROOT, "gpurun_out", "diff_frame.npy"), fr)
    c = codes["warp"]
    c.set_kernel(None)
    row_ptr, col_idx = c.h_csr()
    for it in range(1, iters + 1):
        a = c.decode_debug(fr, max_iters=it, early_stop=False, kernel="warp")
        b = c.decode_debug(fr, max_iters=it, early_stop=False, kernel="c4-thread")
        d = c.decode_debug(fr, max_iters=it, early_stop=False, kernel="block")
        v, run, Lo, Eo, Mo = O.decode_spa(fr[0].real.astype(np.float64), Hp, it, False, debug=True)
        for name in ("E", "M", "L"):
            x, y, z = a[name][0], b[name][0], d[name][0]
            ne = np.nonzero(x.view(np.uint32) != y.view(np.uint32))[0]
            nz = np.nonzero(x.view(np.uint32) != z.view(np.uint32))[0]
            print("iteration %d %s: warp != c4-thread at %d positions %s ; warp != block at %d" % (it, name, ne.size, ne[:8].tolist(), nz.size))
            for k in ne[:4]:
                print("    pos %d: warp %.9g c4 %.9g block %.9g" % (k, x[k], y[k], z[k]))
        La, Lb = a["L"][0], b["L"][0]
        print("   L[39]: warp %.9g c4 %.9g oracle %.9g" % (La[39], Lb[39], Lo[39]))
