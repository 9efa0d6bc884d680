"""The two GNU Radio blocks, drivable without GNU Radio.

`ldpc_decoder_cb(method)` and `ldpc_encoder_bc()` are the constructors the reference's Python
package exports from its SWIG module (python/__init__.py:45, swig/ldpc_ece535a_swig.i:17-22).
Inside a GNU Radio install the SWIG module provides them; here they wrap the very same C++
block classes (lib/ldpc_*_impl.cc, built against the compat headers) through the scheduler
stand-in of lib/block_harness.cc: `general_work(items, noutput_items)` is one scheduler call.
"""
import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
BLOCKS_LIB_PATH = os.environ.get("LDPC535_BLOCKS_LIB") or os.path.normpath(
    os.path.join(_PKG, "..", "..", "libgnuradio-ldpc_ece535a.so"))

EV_IN_SYNC, EV_IN_SYNC_INVERTED, EV_MAX_ERRORS = 1, 2, 3
OUT_OF_SYNC, IN_SYNC, IN_SYNC_INVERTED = 0, 1, 2

_lib = None


def blocks_lib():
    global _lib
    if _lib is None:
        if not os.path.exists(BLOCKS_LIB_PATH):
            raise OSError("libgnuradio-ldpc_ece535a.so not built: run `make -C gr-ldpc_ece535a_b200`")
        h = C.CDLL(BLOCKS_LIB_PATH)
        vp, i = C.c_void_p, C.c_int
        h.ldpc535_blk_last_error.restype = C.c_char_p
        h.ldpc535_blk_decoder_new.restype = vp
        h.ldpc535_blk_decoder_new.argtypes = [i]
        h.ldpc535_blk_encoder_new.restype = vp
        h.ldpc535_blk_image_sink_new.restype = vp
        h.ldpc535_blk_image_sink_files.argtypes = [vp]
        h.ldpc535_blk_image_sink_files.restype = C.c_uint
        h.ldpc535_blk_free.argtypes = [vp]
        h.ldpc535_blk_name.restype = C.c_char_p
        h.ldpc535_blk_name.argtypes = [vp]
        h.ldpc535_blk_item_size.argtypes = [vp, i]
        h.ldpc535_blk_forecast.argtypes = [vp, i]
        h.ldpc535_blk_general_work.argtypes = [vp, i, i, vp, vp, C.POINTER(i)]
        h.ldpc535_blk_decoder_set.argtypes = [vp, i, i]
        h.ldpc535_blk_decoder_state.argtypes = [vp, C.POINTER(C.c_ulong)]
        h.ldpc535_blk_decoder_events.argtypes = [vp, C.POINTER(i), i]
        h.ldpc535_sync_replay_table.argtypes = [vp, vp, vp, vp, C.c_long, C.c_long, i, i, i, i,
                                                C.POINTER(C.c_uint), vp, C.POINTER(C.c_long),
                                                C.POINTER(i), i, C.POINTER(i)]
        _lib = h
    return _lib


class _Block:
    _in_dtype = _out_dtype = None

    def __init__(self, handle):
        if not handle:
            raise RuntimeError(blocks_lib().ldpc535_blk_last_error().decode())
        self._h = C.c_void_p(handle)

    def close(self):
        if getattr(self, "_h", None):
            blocks_lib().ldpc535_blk_free(self._h)
            self._h = None

    __del__ = close

    def name(self):
        return blocks_lib().ldpc535_blk_name(self._h).decode()

    def item_sizes(self):
        L = blocks_lib()
        return L.ldpc535_blk_item_size(self._h, 0), L.ldpc535_blk_item_size(self._h, 1)

    def forecast(self, noutput_items):
        return blocks_lib().ldpc535_blk_forecast(self._h, int(noutput_items))

    def general_work(self, items, noutput_items):
        """One scheduler call: -> (output items produced, input items consumed)."""
        items = np.ascontiguousarray(items, self._in_dtype)
        out = np.zeros(max(int(noutput_items), 1), self._out_dtype)
        consumed = C.c_int(0)
        n = blocks_lib().ldpc535_blk_general_work(self._h, int(noutput_items), items.size,
                                                  items.ctypes.data_as(C.c_void_p),
                                                  out.ctypes.data_as(C.c_void_p), C.byref(consumed))
        if n < 0:
            raise RuntimeError("general_work returned %d: %s" % (
                n, blocks_lib().ldpc535_blk_last_error().decode()))
        return out[:n], consumed.value


class ldpc_decoder_cb(_Block):
    """gr_complex in, bytes out.  method: 0 LogDomain (min-sum), 1 SumProduct, 2 BitFlip, 3 Hard."""
    _in_dtype, _out_dtype = np.complex64, np.uint8

    def __init__(self, method):
        super().__init__(blocks_lib().ldpc535_blk_decoder_new(int(method)))

    def set_max_iterations(self, n, early_stop=True):
        blocks_lib().ldpc535_blk_decoder_set(self._h, int(n), int(bool(early_stop)))

    def state(self):
        st = (C.c_ulong * 4)()
        blocks_lib().ldpc535_blk_decoder_state(self._h, st)
        return {"state": st[0], "errors": st[1], "gpu_batches": st[2], "gpu_windows": st[3]}

    def take_events(self):
        buf = (C.c_int * 65536)()
        n = blocks_lib().ldpc535_blk_decoder_events(self._h, buf, 65536)
        return list(buf[:min(n, 65536)])


class ldpc_encoder_bc(_Block):
    """bytes in, gr_complex out."""
    _in_dtype, _out_dtype = np.uint8, np.complex64

    def __init__(self):
        super().__init__(blocks_lib().ldpc535_blk_encoder_new())


class image_sink(_Block):
    """bytes in, nothing out: reassembles BMP files from the decoded stream (host I/O only)."""
    _in_dtype, _out_dtype = np.uint8, np.uint8

    def __init__(self):
        super().__init__(blocks_lib().ldpc535_blk_image_sink_new())

    def work(self, items):
        """One scheduler call of a sync sink: consumes everything it is shown."""
        items = np.ascontiguousarray(items, np.uint8)
        _, consumed = self.general_work(items, items.size)
        return consumed

    def files_written(self):
        return int(blocks_lib().ldpc535_blk_image_sink_files(self._h))


def sync_replay_table(bytes_pos, bytes_neg, synd_pos, synd_neg, ninput, noutput, N, nbytes,
                      threshold, state=(0, 0)):
    """The decoder block's sync machine alone over a table of per-(offset, polarity) results
    (host logic; no GPU).  -> (out bytes, consumed, events, (state, errors))."""
    L = blocks_lib()
    bp = np.ascontiguousarray(bytes_pos, np.uint8)
    bn = np.ascontiguousarray(bytes_neg, np.uint8)
    sp = np.ascontiguousarray(synd_pos, np.uint8)
    sn = np.ascontiguousarray(synd_neg, np.uint8)
    out = np.zeros(max(noutput, 1), np.uint8)
    st = (C.c_uint * 2)(*state)
    consumed = C.c_long(0)
    ev = (C.c_int * 4096)()
    nev = C.c_int(0)
    vp = C.c_void_p
    n = L.ldpc535_sync_replay_table(bp.ctypes.data_as(vp), bn.ctypes.data_as(vp), sp.ctypes.data_as(vp),
                                    sn.ctypes.data_as(vp), sp.size, int(ninput), int(noutput), int(N),
                                    int(nbytes), int(threshold), st, out.ctypes.data_as(vp),
                                    C.byref(consumed), ev, 4096, C.byref(nev))
    return out[:n], consumed.value, list(ev[:min(nev.value, 4096)]), (st[0], st[1])
