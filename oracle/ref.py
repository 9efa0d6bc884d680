"""ctypes front-end of oracle/_ref/libldpc_ref.so: the REFERENCE's own decoder/encoder block
sources (lib/ldpc_decoder_cb_impl.cc, lib/ldpc_encoder_bc_impl.cc), compiled where they lie under
/root/reference against the dependency stand-ins of oracle/refshim/ (see oracle/ref_driver.cc).

TEST INFRASTRUCTURE ONLY.  Used by tests/ (to pin oracle/ldpc_oracle.c and, on the GPU box, to
check the CUDA path against the reference's own code), by tools/gen_ref_golden.py and by
bench.py's `--impl reference` leg; never by the product package.

The library is built by `make -C oracle` where /root/reference exists (this container); the GPU
box gets the prebuilt file with the snapshot.  available() says whether it can be loaded.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libldpc_ref.so")
REF_ROOT = "/root/reference"

_pi = C.POINTER(C.c_int)
_pd = C.POINTER(C.c_double)
_pf = C.POINTER(C.c_float)
_pb = C.POINTER(C.c_uint8)
_lib = None


def build():
    """(Re)build where the reference sources exist; a no-op elsewhere."""
    if os.path.exists(os.path.join(REF_ROOT, "lib", "ldpc_decoder_cb_impl.cc")):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def available():
    try:
        lib()
        return True
    except OSError:
        return False


def lib():
    global _lib
    if _lib is None:
        build()
        if not os.path.exists(_SO):
            raise OSError("oracle/_ref/libldpc_ref.so is not built (needs /root/reference)")
        L = C.CDLL(_SO)
        L.ref_version.restype = C.c_char_p
        L.ref_cout_end.restype = C.c_long
        L.ref_cout_end.argtypes = [C.c_char_p, C.c_long]
        for n in ("ref_decoder_new", "ref_encoder_new"):
            getattr(L, n).restype = C.c_void_p
        L.ref_decoder_new.argtypes = [C.c_int]
        L.ref_decoder_free.argtypes = [C.c_void_p]
        L.ref_encoder_free.argtypes = [C.c_void_p]
        L.ref_decoder_forecast.argtypes = [C.c_void_p, C.c_int]
        L.ref_encoder_forecast.argtypes = [C.c_void_p, C.c_int]
        L.ref_decoder_work.argtypes = [C.c_void_p, _pf, C.c_int, _pb, C.c_int, C.POINTER(C.c_long)]
        L.ref_encoder_work.argtypes = [C.c_void_p, _pb, C.c_int, _pf, C.c_int, C.POINTER(C.c_long)]
        pu = C.POINTER(C.c_uint)
        L.ref_decoder_get_state.argtypes = [C.c_void_p, _pi, pu, pu, pu, pu]
        L.ref_decoder_set_state.argtypes = [C.c_void_p, C.c_int, C.c_uint]
        L.ref_decoder_get_h.argtypes = [C.c_void_p, _pi]
        L.ref_decoder_set_code.argtypes = [C.c_void_p, _pi, C.c_int, C.c_int, C.c_int, _pi, _pi, _pi]
        L.ref_decoder_decode.argtypes = [C.c_void_p, C.c_int, _pd, C.c_int, _pi]
        L.ref_decoder_decode_frames.argtypes = [C.c_void_p, C.c_int, _pf, C.c_long, C.c_int, _pb, _pb]
        L.ref_decoder_decode_frames.restype = None
        L.ref_decoder_check_frame.argtypes = [C.c_void_p, _pi, C.c_int]
        L.ref_encoder_get_tables.argtypes = [C.c_void_p, _pi, _pi, _pi]
        L.ref_encoder_set_code.argtypes = [C.c_void_p, _pi, C.c_int, C.c_int, _pi, _pi, _pi]
        L.ref_encoder_make_parity.argtypes = [C.c_void_p, _pi, _pi]
        _lib = L
    return _lib


def _ints(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _ip(a):
    return a.ctypes.data_as(_pi)


class _Capture:
    """Collect what the blocks print to std::cout (their only diagnostics channel)."""

    def __enter__(self):
        lib().ref_cout_begin()
        self.text = ""
        return self

    def __exit__(self, *exc):
        buf = C.create_string_buffer(1 << 20)
        lib().ref_cout_end(buf, len(buf))
        self.text = buf.value.decode()
        return False


# what the decoder prints on a state change (lib/ldpc_decoder_cb_impl.cc:172,190,201), mapped to
# the event codes oracle.DecoderBlock records: 1 in sync, 2 in sync inverted, 3 out of sync
_EVENT_OF_LINE = {"IN SYNC": 1, "IN SYNC; PHASE INVERTED": 2, "MAX ERRORS; OUT OF SYNC": 3}


class RefDecoder:
    """The reference's ldpc_decoder_cb, driven the way the GNU Radio scheduler drives it."""

    def __init__(self, method, quiet=True):
        with _Capture() as cap:
            self.h = lib().ref_decoder_new(int(method))
        self.banner = cap.text
        self.events = []
        self.quiet = quiet
        self._dims()

    def _dims(self):
        M, N, it = C.c_uint(), C.c_uint(), C.c_uint()
        lib().ref_decoder_get_state(self.h, None, None, C.byref(M), C.byref(N), C.byref(it))
        self.M, self.N, self.iterations = M.value, N.value, it.value

    def close(self):
        if self.h:
            lib().ref_decoder_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def H(self):
        H = np.zeros((self.M, self.N), np.int32)
        lib().ref_decoder_get_h(self.h, _ip(H))
        return H

    @property
    def state(self):
        s, e = C.c_int(), C.c_uint()
        lib().ref_decoder_get_state(self.h, C.byref(s), C.byref(e), None, None, None)
        return s.value, e.value

    def set_code(self, H, iterations=0):
        """Swap the pasted-in literal for H (raw); runs the reference's reorderHMatrix on it."""
        H = _ints(H)
        M, N = H.shape
        Hp = np.zeros((M, N), np.int32)
        L = np.zeros((M, N - M), np.int32)
        U = np.zeros((M, N - M), np.int32)
        lib().ref_decoder_set_code(self.h, _ip(H), M, N, int(iterations), _ip(Hp), _ip(L), _ip(U))
        self._dims()
        return Hp, L, U

    def forecast(self, noutput_items):
        return lib().ref_decoder_forecast(self.h, int(noutput_items))

    def work(self, sym, noutput_items):
        """sym: complex64 -> (bytes out[:produced], consumed symbols)."""
        sym = np.ascontiguousarray(sym, np.complex64)
        out = np.zeros(max(int(noutput_items), 1), np.uint8)
        consumed = C.c_long(0)
        with _Capture() as cap:
            prod = lib().ref_decoder_work(self.h, sym.view(np.float32).ctypes.data_as(_pf),
                                          int(sym.size), out.ctypes.data_as(_pb),
                                          int(noutput_items), C.byref(consumed))
        for line in cap.text.splitlines():
            self.events.append(_EVENT_OF_LINE[line.strip()])
        return out[:prod], consumed.value

    def decode(self, rx, method, iterations):
        """One codeword through the private decode member (rx as general_work forms tx)."""
        rx = np.ascontiguousarray(rx, np.float64)
        v = np.zeros(self.N, np.int32)
        lib().ref_decoder_decode(self.h, int(method), rx.ctypes.data_as(_pd), int(iterations), _ip(v))
        return v

    def check_frame(self, u, threshold):
        return lib().ref_decoder_check_frame(self.h, _ip(_ints(u)), int(threshold))


def decode_frames(sym, method=1, iterations=5, threads=1, pin=False, H=None):
    """Aligned frames through the reference's private decode members (their built-in early exit
    included), `threads` reference decoder instances on disjoint contiguous shards.
    -> bytes (n, K/8), syndrome weights checkFrame(v, M/8), seconds."""
    import threading
    import time
    decs = [RefDecoder(method) for _ in range(threads)]
    if H is not None:
        for d in decs:
            d.set_code(H)
    M, N = decs[0].M, decs[0].N
    sym = np.ascontiguousarray(sym, np.complex64).reshape(-1)
    symf = sym.view(np.float32)
    n = sym.size // N
    nb = (N - M) // 8
    out = np.zeros((n, nb), np.uint8)
    synd = np.zeros(n, np.uint8)
    bounds = [n * t // threads for t in range(threads + 1)]
    f = lib().ref_decoder_decode_frames

    def worker(t):
        if pin:
            try:
                cpus = sorted(os.sched_getaffinity(0))
                os.sched_setaffinity(0, {cpus[t % len(cpus)]})
            except OSError:
                pass
        a, b = bounds[t], bounds[t + 1]
        if b > a:
            f(decs[t].h, int(method), symf[a * N * 2:].ctypes.data_as(_pf), b - a, int(iterations),
              out[a:].ctypes.data_as(_pb), synd[a:].ctypes.data_as(_pb))

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
    for d in decs:
        d.close()
    return out, synd, dt


class RefEncoder:
    """The reference's ldpc_encoder_bc."""

    def __init__(self):
        self.h = lib().ref_encoder_new()
        self.M, self.N = 32, 64

    def close(self):
        if self.h:
            lib().ref_encoder_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def tables(self):
        H = np.zeros((self.M, self.N), np.int32)
        L = np.zeros((self.M, self.N - self.M), np.int32)
        U = np.zeros_like(L)
        lib().ref_encoder_get_tables(self.h, _ip(H), _ip(L), _ip(U))
        return H, L, U

    def set_code(self, H):
        H = _ints(H)
        M, N = H.shape
        Hp = np.zeros((M, N), np.int32)
        L = np.zeros((M, N - M), np.int32)
        U = np.zeros_like(L)
        lib().ref_encoder_set_code(self.h, _ip(H), M, N, _ip(Hp), _ip(L), _ip(U))
        self.M, self.N = M, N
        return Hp, L, U

    def forecast(self, noutput_items):
        return lib().ref_encoder_forecast(self.h, int(noutput_items))

    def work(self, in_bytes, noutput_items):
        """bytes -> (complex64 out[:produced], consumed bytes)."""
        in_bytes = np.ascontiguousarray(in_bytes, np.uint8)
        out = np.zeros(max(int(noutput_items), 1), np.complex64)
        consumed = C.c_long(0)
        prod = lib().ref_encoder_work(self.h, in_bytes.ctypes.data_as(_pb), int(in_bytes.size),
                                      out.view(np.float32).ctypes.data_as(_pf),
                                      int(noutput_items), C.byref(consumed))
        return out[:prod], consumed.value

    def make_parity(self, d):
        d = _ints(d)
        c = np.zeros(self.M, np.int32)
        lib().ref_encoder_make_parity(self.h, _ip(d), _ip(c))
        return c
