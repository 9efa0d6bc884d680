"""`Code`: one parity-check code on one B200, through the C ABI (include/ldpc535.h)."""
import ctypes as C

import numpy as np

from . import _abi
from ._abi import check, lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Code:
    """Handle over ldpc535_code.  `H`: dense 0/1 (M, N) array, or (row_ptr, col_idx) CSR,
    or None for the 32x64 code the reference blocks are hard-wired to."""

    def __init__(self, H=None, device=0):
        self._h = C.c_void_p()
        L = lib()
        if H is None:
            check(L.ldpc535_code_create_default(int(device), C.byref(self._h)), "code_create_default")
        elif isinstance(H, tuple):
            row_ptr = np.ascontiguousarray(H[0], np.int32)
            col_idx = np.ascontiguousarray(H[1], np.int32)
            M, N = int(H[2]), int(H[3])
            check(L.ldpc535_code_create_sparse(_ptr(row_ptr), _ptr(col_idx), M, N, int(device),
                                               C.byref(self._h)), "code_create_sparse")
        else:
            Hd = np.ascontiguousarray(H, np.int32)
            check(L.ldpc535_code_create(_ptr(Hd), Hd.shape[0], Hd.shape[1], int(device),
                                        C.byref(self._h)), "code_create")
        m, n, k, e, d = (C.c_int() for _ in range(5))
        check(L.ldpc535_code_info(self._h, m, n, k, e, d), "code_info")
        self.M, self.N, self.K, self.E, self.device = m.value, n.value, k.value, e.value, d.value
        self.nbytes = (self.K + 7) // 8

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            lib().ldpc535_code_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- introspection -----------------------------------------------------------
    @property
    def handle(self):
        return self._h

    def pivots(self):
        out = np.zeros(self.M, np.int32)
        check(lib().ldpc535_code_get_pivots(self._h, _ptr(out)), "get_pivots")
        return out

    def h_csr(self):
        row_ptr = np.zeros(self.M + 1, np.int32)
        col_idx = np.zeros(self.E, np.int32)
        check(lib().ldpc535_code_get_h(self._h, _ptr(row_ptr), _ptr(col_idx)), "get_h")
        return row_ptr, col_idx

    def h_dense(self):
        row_ptr, col_idx = self.h_csr()
        H = np.zeros((self.M, self.N), np.int32)
        for j in range(self.M):
            H[j, col_idx[row_ptr[j]:row_ptr[j + 1]]] = 1
        return H

    def generator(self):
        kw = (self.K + 31) // 32
        P = np.zeros((self.M, kw), np.uint32)
        check(lib().ldpc535_code_get_generator(self._h, _ptr(P)), "get_generator")
        return P

    def kernel_name(self, method=_abi.METHOD_SUMPRODUCT):
        return lib().ldpc535_code_kernel_name(self._h, int(method)).decode()

    def kernel_for(self, method, early_stop, n_win):
        return lib().ldpc535_code_kernel_for(self._h, int(method), int(bool(early_stop)), int(n_win)).decode()

    def set_kernel(self, name):
        check(lib().ldpc535_code_set_kernel(self._h, None if name is None else name.encode()),
              "set_kernel")

    def host_path(self):
        """How decode() moves host symbols: {'pack_pinned': 0 raw | 1 packed | 2 chunk by chunk, 'pack_threads': int}."""
        a, b = C.c_int(), C.c_int()
        check(lib().ldpc535_code_host_path(self._h, a, b), "host_path")
        return {"pack_pinned": a.value, "pack_threads": b.value}

    def host_stats(self):
        """Cumulative {'chunks_packed', 'chunks_raw', 'h2d_bytes'} of decode() on this handle."""
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        check(lib().ldpc535_code_host_stats(self._h, a, b, c), "host_stats")
        return {"chunks_packed": a.value, "chunks_raw": b.value, "h2d_bytes": c.value}

    def set_host_path(self, pack_pinned=-1, pack_threads=0):
        """pack_threads >= 1 (0 keeps it); pack_pinned 0 raw / 1 packed / 2 chunk by chunk, -1 = default (2)."""
        check(lib().ldpc535_code_set_host_path(self._h, int(pack_pinned), int(pack_threads)), "set_host_path")

    def launch_count(self):
        return int(lib().ldpc535_launch_count(self._h))

    # ---- host-buffer API ------------------------------------------------------------
    def encode(self, data_bytes, out=None):
        """bytes (n_frames * K/8, MSB first) -> complex64 (n_frames, N), parity then data."""
        data = np.ascontiguousarray(data_bytes, np.uint8).reshape(-1)
        if data.size % self.nbytes:
            raise ValueError("input is not a whole number of %d-byte frames" % self.nbytes)
        n = data.size // self.nbytes
        if out is None:
            out = np.empty((n, self.N), np.complex64)
        check(lib().ldpc535_encode_batch(self._h, _ptr(data), n, _ptr(out)), "encode_batch")
        return out

    def decode(self, sym, method=_abi.METHOD_SUMPRODUCT, max_iters=5, early_stop=True,
               synd_threshold=None, win_offset=None, polarity=None, n_win=None, out=None):
        """complex64 symbols -> (bytes (n_win, K/8), synd (n_win,), iters (n_win,)).

        Defaults are the reference block's constants (5 iterations, early stop,
        threshold M/8; lib/ldpc_decoder_cb_impl.cc:39-40, :141-142)."""
        sym = np.ascontiguousarray(sym, np.complex64).reshape(-1)
        off = pol = None
        if win_offset is not None:
            off = np.ascontiguousarray(win_offset, np.int64)
            n_win = off.size
        elif n_win is None:
            n_win = sym.size // self.N
        if polarity is not None:
            pol = np.ascontiguousarray(polarity, np.int8)
            if pol.size != n_win:
                raise ValueError("polarity must have one entry per window")
        thr = self.M // 8 if synd_threshold is None else int(synd_threshold)
        if out is None:
            out = (np.empty((n_win, self.nbytes), np.uint8), np.empty(n_win, np.uint8),
                   np.empty(n_win, np.uint8))
        ob, os_, oi = out
        check(lib().ldpc535_decode_batch(self._h, _ptr(sym), sym.size, _ptr(off), _ptr(pol), n_win,
                                         int(method), int(max_iters), int(bool(early_stop)), thr,
                                         _ptr(ob), _ptr(os_), _ptr(oi)), "decode_batch")
        return ob, os_, oi

    def decode_debug(self, sym, max_iters=5, early_stop=True, kernel=None):
        """Sum-product with message dumps -> dict(L, E, M, bytes, iters); E/M in CSR edge
        order of the re-ordered H."""
        sym = np.ascontiguousarray(sym, np.complex64).reshape(-1)
        n_win = sym.size // self.N
        L = np.zeros((n_win, self.N), np.float32)
        E = np.zeros((n_win, self.E), np.float32)
        M = np.zeros((n_win, self.E), np.float32)
        ob = np.zeros((n_win, self.nbytes), np.uint8)
        oi = np.zeros(n_win, np.uint8)
        check(lib().ldpc535_decode_debug(self._h, _ptr(sym), n_win, int(max_iters),
                                         int(bool(early_stop)),
                                         None if kernel is None else kernel.encode(),
                                         _ptr(L), _ptr(E), _ptr(M), _ptr(ob), _ptr(oi)),
              "decode_debug")
        return {"L": L, "E": E, "M": M, "bytes": ob, "iters": oi}

    # ---- device-resident API (raw device pointers, e.g. torch tensors' data_ptr()) -----
    def encode_dev(self, d_in, n_frames, d_out, stream=None):
        check(lib().ldpc535_encode_batch_dev(self._h, d_in, int(n_frames), d_out, stream),
              "encode_batch_dev")

    def decode_dev(self, d_sym, n_sym, n_win, d_out_bytes, d_out_synd=None, d_out_iters=None,
                   method=_abi.METHOD_SUMPRODUCT, max_iters=5, early_stop=True,
                   synd_threshold=None, d_win_offset=None, d_polarity=None, stream=None):
        thr = self.M // 8 if synd_threshold is None else int(synd_threshold)
        check(lib().ldpc535_decode_batch_dev(self._h, d_sym, int(n_sym), d_win_offset, d_polarity,
                                             int(n_win), int(method), int(max_iters),
                                             int(bool(early_stop)), thr, d_out_bytes, d_out_synd,
                                             d_out_iters, stream), "decode_batch_dev")

    def synth_bytes_dev(self, seed, first_frame, n_frames, d_bytes, stream=None):
        check(lib().ldpc535_synth_bytes_dev(self._h, int(seed), int(first_frame), int(n_frames), d_bytes, stream),
              "synth_bytes_dev")

    def synth_awgn_dev(self, seed, first_frame, n_frames, sigma, d_sym, stream=None):
        check(lib().ldpc535_synth_awgn_dev(self._h, int(seed), int(first_frame), int(n_frames), float(sigma),
                                           d_sym, stream), "synth_awgn_dev")

    def probe_pipe_peak(self, which):
        """Measured thread-level ops/s of an SM pipe (_abi.PIPE_MUFU / PIPE_FP64)."""
        v = C.c_double()
        check(lib().ldpc535_probe_pipe_peak(self._h, int(which), C.byref(v)), "probe_pipe_peak")
        return v.value

    def sync(self, stream=None):
        check(lib().ldpc535_stream_sync(self._h, stream), "stream_sync")


class Pool:
    """Several GPUs from one process (ldpc535_pool): contiguous shards, one host thread per
    device, host-side gather.  devices may repeat (two handles on one GPU)."""

    def __init__(self, devices, H=None):
        self._p = C.c_void_p()
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        if H is None:
            check(lib().ldpc535_pool_create(None, 0, 0, devs, len(devices), C.byref(self._p)), "pool_create")
        else:
            Hd = np.ascontiguousarray(H, np.int32)
            check(lib().ldpc535_pool_create(_ptr(Hd), Hd.shape[0], Hd.shape[1], devs, len(devices),
                                            C.byref(self._p)), "pool_create")
        m, n, k = C.c_int(), C.c_int(), C.c_int()
        check(lib().ldpc535_code_info(lib().ldpc535_pool_code(self._p, 0), m, n, k, None, None), "code_info")
        self.M, self.N, self.K = m.value, n.value, k.value
        self.nbytes = (self.K + 7) // 8
        self.size = lib().ldpc535_pool_size(self._p)

    def close(self):
        if getattr(self, "_p", None):
            lib().ldpc535_pool_destroy(self._p)
            self._p = None

    __del__ = close

    def encode(self, data_bytes):
        data = np.ascontiguousarray(data_bytes, np.uint8).reshape(-1)
        n = data.size // self.nbytes
        out = np.empty((n, self.N), np.complex64)
        check(lib().ldpc535_pool_encode_batch(self._p, _ptr(data), n, _ptr(out)), "pool_encode_batch")
        return out

    def decode(self, sym, method=_abi.METHOD_SUMPRODUCT, max_iters=5, early_stop=True, win_offset=None,
               polarity=None):
        sym = np.ascontiguousarray(sym, np.complex64).reshape(-1)
        off = None if win_offset is None else np.ascontiguousarray(win_offset, np.int64)
        pol = None if polarity is None else np.ascontiguousarray(polarity, np.int8)
        n_win = sym.size // self.N if off is None else off.size
        ob = np.empty((n_win, self.nbytes), np.uint8)
        os_ = np.empty(n_win, np.uint8)
        oi = np.empty(n_win, np.uint8)
        check(lib().ldpc535_pool_decode_batch(self._p, _ptr(sym), sym.size, _ptr(off), _ptr(pol), n_win,
                                              int(method), int(max_iters), int(bool(early_stop)), self.M // 8,
                                              _ptr(ob), _ptr(os_), _ptr(oi)), "pool_decode_batch")
        return ob, os_, oi


def device_count():
    n = C.c_int()
    check(lib().ldpc535_device_count(n), "device_count")
    return n.value


def device_info(device=0):
    name = C.create_string_buffer(128)
    mj, mn, sms, khz = (C.c_int() for _ in range(4))
    check(lib().ldpc535_device_info(int(device), name, mj, mn, sms, khz), "device_info")
    return {"name": name.value.decode(), "sm": (mj.value, mn.value), "sm_count": sms.value,
            "sm_clock_khz": khz.value}
