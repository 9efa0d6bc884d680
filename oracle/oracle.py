"""ctypes front-end of oracle/ldpc_oracle.c (the CPU restatement of the reference).

TEST INFRASTRUCTURE ONLY -- see the header of ldpc_oracle.c.  Imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs, never
by the product package.
"""
import ctypes as C
import os
import subprocess
import threading
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libldpc_oracle.so")


def build(force=False):
    src = os.path.join(_HERE, "ldpc_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None
_pi = C.POINTER(C.c_int)
_pd = C.POINTER(C.c_double)
_pf = C.POINTER(C.c_float)
_pb = C.POINTER(C.c_uint8)


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_decode_frames.argtypes = [_pf, C.c_long, _pi, C.c_int, C.c_int, C.c_int, C.c_int,
                                           C.c_int, C.c_int, _pb, _pb, _pb]
        _lib.orc_decode_frames.restype = None
    return _lib


def _i(a):
    return a.ctypes.data_as(_pi)


def _d(a):
    return None if a is None else a.ctypes.data_as(_pd)


def _ints(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def reorder_h(H):
    """-> (H_perm, L, U, chosen).  apps/ldpc_lapack.cpp:163-211."""
    H = _ints(H).copy()
    M, N = H.shape
    L = np.zeros((M, N - M), np.int32)
    U = np.zeros((M, N - M), np.int32)
    chosen = np.zeros(M, np.int32)
    lib().orc_reorder_h(_i(H), M, N, _i(L), _i(U), _i(chosen))
    return H, L, U, chosen


def make_parity_check(d, Hp, L, U, gf2=False):
    Hp, L, U, d = _ints(Hp), _ints(L), _ints(U), _ints(d)
    M, N = Hp.shape
    c = np.zeros(M, np.int32)
    bad = lib().orc_make_parity_check(_i(d), _i(Hp), _i(L), _i(U), M, N, int(gf2), _i(c))
    return c, bad


def check_frame(u, H, threshold):
    H, u = _ints(H), _ints(u)
    return lib().orc_check_frame(_i(u), _i(H), H.shape[0], H.shape[1], int(threshold))


def decode_spa(rx, H, iterations=5, early_stop=True, debug=False):
    """apps/ldpc_lapack.cpp:260-335.  -> vhat, iters_run[, L, E(dense), M(dense)]."""
    H = _ints(H)
    M, N = H.shape
    rx = np.ascontiguousarray(rx, np.float64)
    vhat = np.zeros(N, np.int32)
    run = C.c_int(0)
    Lv = np.zeros(N) if debug else None
    E = np.zeros((M, N)) if debug else None
    Mm = np.zeros((M, N)) if debug else None
    lib().orc_decode_spa(_d(rx), _i(H), M, N, int(iterations), int(early_stop), _i(vhat),
                         C.byref(run), _d(Lv), _d(E), _d(Mm))
    return (vhat, run.value, Lv, E, Mm) if debug else (vhat, run.value)


def csr_csc(H):
    """CSR (columns ascending) + per-column edge-id lists (rows ascending) of dense H."""
    H = np.asarray(H)
    M, N = H.shape
    rows, cols = np.nonzero(H)                 # row-major order == CSR order
    row_ptr = np.zeros(M + 1, np.int32)
    np.add.at(row_ptr, rows + 1, 1)
    row_ptr = np.cumsum(row_ptr).astype(np.int32)
    col_idx = cols.astype(np.int32)
    order = np.lexsort((rows, cols))           # by column, then row
    edge_of_col = order.astype(np.int32)
    col_ptr = np.zeros(N + 1, np.int32)
    np.add.at(col_ptr, cols + 1, 1)
    col_ptr = np.cumsum(col_ptr).astype(np.int32)
    return row_ptr, col_idx, col_ptr, edge_of_col


def decode_spa_sparse(rx, tables, M, N, iterations=5, early_stop=True, debug=False):
    row_ptr, col_idx, col_ptr, edge_of_col = tables
    rx = np.ascontiguousarray(rx, np.float64)
    ne = int(row_ptr[M])
    vhat = np.zeros(N, np.int32)
    run = C.c_int(0)
    Lv = np.zeros(N) if debug else None
    E = np.zeros(ne) if debug else None
    Mm = np.zeros(ne) if debug else None
    lib().orc_decode_spa_sparse(_d(rx), M, N, _i(row_ptr), _i(col_idx), _i(col_ptr),
                                _i(edge_of_col), int(iterations), int(early_stop), _i(vhat),
                                C.byref(run), _d(Lv), _d(E), _d(Mm))
    return (vhat, run.value, Lv, E, Mm) if debug else (vhat, run.value)


def decode_minsum(rx, H, iterations=5, early_stop=True):
    H = _ints(H)
    rx = np.ascontiguousarray(rx, np.float64)
    vhat = np.zeros(H.shape[1], np.int32)
    run = C.c_int(0)
    lib().orc_decode_minsum(_d(rx), _i(H), H.shape[0], H.shape[1], int(iterations),
                            int(early_stop), _i(vhat), C.byref(run))
    return vhat, run.value


def decode_hard(rx):
    rx = np.ascontiguousarray(rx, np.float64)
    vhat = np.zeros(rx.shape[0], np.int32)
    lib().orc_decode_hard(_d(rx), rx.shape[0], _i(vhat))
    return vhat


def decode_bitflip(rx, H, iterations=5):
    H = _ints(H)
    rx = np.ascontiguousarray(rx, np.float64)
    vhat = np.zeros(H.shape[1], np.int32)
    lib().orc_decode_bitflip(_d(rx), _i(H), H.shape[0], H.shape[1], int(iterations), _i(vhat))
    return vhat


def encoder_work(Hp, L, U, in_bytes, noutput_items, gf2=False):
    """lib/ldpc_encoder_bc_impl.cc:118-178 -> (complex64 out[:produced], consumed)."""
    Hp, L, U = _ints(Hp), _ints(L), _ints(U)
    M, N = Hp.shape
    in_bytes = np.ascontiguousarray(in_bytes, np.uint8)
    out = np.zeros(max(noutput_items, 1), np.complex64)
    consumed = C.c_int(0)
    prod = lib().orc_encoder_work(_i(Hp), _i(L), _i(U), M, N, int(gf2), int(noutput_items),
                                  int(in_bytes.size), in_bytes.ctypes.data_as(_pb),
                                  out.ctypes.data_as(_pf), C.byref(consumed))
    return out[:prod], consumed.value


class DecoderState(C.Structure):
    _fields_ = [("method", C.c_int), ("state", C.c_int), ("errors", C.c_uint),
                ("iterations", C.c_int)]


class DecoderBlock:
    """Block-level oracle: lib/ldpc_decoder_cb_impl.cc:132-234 incl. the sync machine."""

    def __init__(self, Hp, method, iterations=5):
        self.H = _ints(Hp)
        self.st = DecoderState(int(method), 0, 0, int(iterations))
        self.events = []

    def work(self, sym, noutput_items):
        """sym: complex64 -> (bytes out[:produced], consumed symbols)."""
        sym = np.ascontiguousarray(sym, np.complex64)
        M, N = self.H.shape
        out = np.zeros(max(noutput_items, 1), np.uint8)
        consumed = C.c_int(0)
        nev = C.c_int(0)
        cap = 4096
        ev = np.zeros(cap, np.int32)
        prod = lib().orc_decoder_work(C.byref(self.st), _i(self.H), M, N, int(noutput_items),
                                      int(sym.size), sym.ctypes.data_as(_pf),
                                      out.ctypes.data_as(_pb), C.byref(consumed), _i(ev), cap,
                                      C.byref(nev))
        self.events += [int(e) for e in ev[:min(nev.value, cap)]]
        return out[:prod], consumed.value


def decode_frames(sym, H, method=1, iterations=5, early_stop=True, threshold=None, threads=1,
                  pin=False):
    """Decode aligned frames with the dense restatement, `threads` worker threads on
    disjoint contiguous shards (ctypes drops the GIL).  -> bytes, iters, synd, seconds."""
    H = _ints(H)
    M, N = H.shape
    sym = np.ascontiguousarray(sym, np.complex64).reshape(-1)
    n_cw = sym.size // N
    nb = (N - M) // 8
    thr = M // 8 if threshold is None else threshold
    out = np.zeros(n_cw * nb, np.uint8)
    iters = np.zeros(n_cw, np.uint8)
    synd = np.zeros(n_cw, np.uint8)
    f = lib().orc_decode_frames
    bounds = [n_cw * t // threads for t in range(threads + 1)]
    symf = sym.view(np.float32)

    def worker(t):
        if pin:
            try:
                os.sched_setaffinity(0, {sorted(os.sched_getaffinity(0))[t % len(os.sched_getaffinity(0))]})
            except OSError:
                pass
        a, b = bounds[t], bounds[t + 1]
        if b > a:
            f(symf[a * N * 2:].ctypes.data_as(_pf), b - a, _i(H), M, N, int(method),
              int(iterations), int(early_stop), int(thr),
              out[a * nb:].ctypes.data_as(_pb), iters[a:].ctypes.data_as(_pb),
              synd[a:].ctypes.data_as(_pb))

    t0 = time.perf_counter()
    if threads == 1:
        worker(0)
    else:
        ths = [threading.Thread(target=worker, args=(t,)) for t in range(threads)]
        for th in ths:
            th.start()
        for th in ths:
            th.join()
    dt = time.perf_counter() - t0
    return out.reshape(n_cw, nb), iters, synd, dt


# ---------------------------------------------------------------------------
# fixtures
# ---------------------------------------------------------------------------

def load_ref_codes():
    import json
    p = os.path.join(os.path.dirname(_HERE), "tests", "golden", "ref_codes.json")
    with open(p) as fh:
        raw = json.load(fh)
    codes = {}
    for name, e in raw.items():
        if not isinstance(e, dict):
            continue
        H = np.zeros((e["M"], e["N"]), np.int32)
        for r, cols in enumerate(e["rows"]):
            H[r, cols] = 1
        codes[name] = {"H": H, "source_words": np.array(e.get("source_words", []), np.int32)}
    codes["shipped"] = codes[raw["shipped"]]
    return codes


class ImageSinkOracle:
    """lib/image_sink_impl.cc:46-84 restated: header scan with 19 bytes of look-ahead inside
    one work() call, file written (announced size only) when the NEXT header arrives."""
    DIB = (12, 40, 52, 56, 64, 108, 124)

    def __init__(self):
        self.buffer = bytearray()
        self.file_size = 0
        self.files = []            # every file the block would have written to result.bmp

    def work(self, data):
        d = bytes(bytearray(data))
        n = len(d)
        for i in range(n):
            if (i < n - 18 and d[i] == 0x42 and d[i + 1] == 0x4D and d[i + 6] == 0 and d[i + 7] == 0
                    and d[i + 8] == 0 and d[i + 9] == 0 and d[i + 14] in self.DIB):
                if self.file_size > 0 and len(self.buffer) >= self.file_size:
                    self.files.append(bytes(self.buffer[:self.file_size]))
                self.buffer = bytearray()
                self.file_size = (d[i + 5] << 24) | (d[i + 4] << 16) | (d[i + 3] << 8) | d[i + 2]
            self.buffer.append(d[i])
        return n
