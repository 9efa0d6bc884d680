"""ctypes declarations for libldpc535.so (include/ldpc535.h).

The library is the product: if it is missing, or no sm_100 GPU is usable, everything
here raises -- there is no CPU path in this package.
"""
import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LDPC535_LIB") or os.path.normpath(os.path.join(_PKG, "..", "..", "libldpc535.so"))

OK, ERR_INVALID, ERR_NO_DEVICE, ERR_CUDA, ERR_SINGULAR, ERR_UNSUPPORTED, ERR_NOMEM = range(7)

METHOD_LOGDOMAIN, METHOD_SUMPRODUCT, METHOD_BITFLIP, METHOD_HARD = 0, 1, 2, 3
PIPE_MUFU, PIPE_FP64 = 0, 1

# every symbol include/ldpc535.h declares: name -> (restype, argtypes)
_vp, _i, _sz, _u64 = C.c_void_p, C.c_int, C.c_size_t, C.c_uint64
_pi = C.POINTER(C.c_int)
SYMBOLS = {
    "ldpc535_version": (C.c_char_p, []),
    "ldpc535_strerror": (C.c_char_p, [_i]),
    "ldpc535_last_error": (C.c_char_p, []),
    "ldpc535_device_count": (_i, [_pi]),
    "ldpc535_device_info": (_i, [_i, C.c_char_p, _pi, _pi, _pi, _pi]),
    "ldpc535_code_create": (_i, [_vp, _i, _i, _i, C.POINTER(_vp)]),
    "ldpc535_code_create_sparse": (_i, [_vp, _vp, _i, _i, _i, C.POINTER(_vp)]),
    "ldpc535_code_create_default": (_i, [_i, C.POINTER(_vp)]),
    "ldpc535_code_destroy": (None, [_vp]),
    "ldpc535_code_info": (_i, [_vp, _pi, _pi, _pi, _pi, _pi]),
    "ldpc535_code_get_pivots": (_i, [_vp, _vp]),
    "ldpc535_code_get_h": (_i, [_vp, _vp, _vp]),
    "ldpc535_code_get_generator": (_i, [_vp, _vp]),
    "ldpc535_code_kernel_name": (C.c_char_p, [_vp, _i]),
    "ldpc535_code_set_kernel": (_i, [_vp, C.c_char_p]),
    "ldpc535_code_kernel_for": (C.c_char_p, [_vp, _i, _i, _sz]),
    "ldpc535_launch_count": (_u64, [_vp]),
    "ldpc535_pool_create": (_i, [_vp, _i, _i, _pi, _i, C.POINTER(_vp)]),
    "ldpc535_pool_destroy": (None, [_vp]),
    "ldpc535_pool_size": (_i, [_vp]),
    "ldpc535_pool_code": (_vp, [_vp, _i]),
    "ldpc535_pool_decode_batch": (_i, [_vp, _vp, _sz, _vp, _vp, _sz, _i, _i, _i, _i, _vp, _vp, _vp]),
    "ldpc535_pool_encode_batch": (_i, [_vp, _vp, _sz, _vp]),
    "ldpc535_code_host_path": (_i, [_vp, _pi, _pi]),
    "ldpc535_code_set_host_path": (_i, [_vp, _i, _i]),
    "ldpc535_code_host_stats": (_i, [_vp, C.POINTER(_u64), C.POINTER(_u64), C.POINTER(_u64)]),
    "ldpc535_synth_bytes_dev": (_i, [_vp, _u64, _u64, _sz, _vp, _vp]),
    "ldpc535_synth_awgn_dev": (_i, [_vp, _u64, _u64, _sz, C.c_float, _vp, _vp]),
    "ldpc535_probe_pipe_peak": (_i, [_vp, _i, C.POINTER(C.c_double)]),
    "ldpc535_host_alloc": (_i, [_sz, C.POINTER(_vp)]),
    "ldpc535_host_free": (_i, [_vp]),
    "ldpc535_encode_batch": (_i, [_vp, _vp, _sz, _vp]),
    "ldpc535_decode_batch": (_i, [_vp, _vp, _sz, _vp, _vp, _sz, _i, _i, _i, _i, _vp, _vp, _vp]),
    "ldpc535_encode_batch_dev": (_i, [_vp, _vp, _sz, _vp, _vp]),
    "ldpc535_decode_batch_dev": (_i, [_vp, _vp, _sz, _vp, _vp, _sz, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "ldpc535_stream_sync": (_i, [_vp, _vp]),
    "ldpc535_dev_alloc": (_i, [_vp, _sz, C.POINTER(_vp)]),
    "ldpc535_dev_free": (_i, [_vp, _vp]),
    "ldpc535_memcpy_h2d": (_i, [_vp, _vp, _vp, _sz]),
    "ldpc535_memcpy_d2h": (_i, [_vp, _vp, _vp, _sz]),
    "ldpc535_decode_debug": (_i, [_vp, _vp, _sz, _i, _i, C.c_char_p, _vp, _vp, _vp, _vp, _vp]),
}

_lib = None


class Ldpc535Error(RuntimeError):
    def __init__(self, status, where, detail):
        self.status = status
        super().__init__("%s failed: status %d (%s)" % (where, status, detail))


def lib():
    """Load libldpc535.so; raises OSError if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OSError("libldpc535.so not built: run `make -C gr-ldpc_ece535a_b200` "
                          "(or __graft_entry__.build()); there is no CPU fallback")
        h = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(h, name)
            fn.restype = res
            fn.argtypes = args
        _lib = h
    return _lib


def check(status, where):
    if status != OK:
        detail = lib().ldpc535_last_error().decode() or lib().ldpc535_strerror(status).decode()
        raise Ldpc535Error(status, where, detail)
