"""ldpc_ece535a -- B200-native LDPC encode/decode behind the reference module's names.

The reference's Python package re-exports its SWIG module (python/__init__.py:45):
`ldpc_ece535a.ldpc_decoder_cb(method)`, `ldpc_ece535a.ldpc_encoder_bc()`.  This package
keeps those two constructors (block-level objects driven through the same C++ block
sources GNU Radio would load) and adds the codeword-level `Code` API over the C ABI.
There is no CPU path: importing works anywhere, computing needs libldpc535.so and a B200.
"""
from ._abi import (METHOD_BITFLIP, METHOD_HARD, METHOD_LOGDOMAIN, METHOD_SUMPRODUCT, LIB_PATH,
                   Ldpc535Error)
from .code import Code, Pool, device_count, device_info
from .blocks import image_sink, ldpc_decoder_cb, ldpc_encoder_bc
from . import codes

__all__ = ["ldpc_decoder_cb", "ldpc_encoder_bc", "image_sink", "Code", "Pool", "codes", "device_count", "device_info", "Ldpc535Error", "LIB_PATH",
           "METHOD_LOGDOMAIN", "METHOD_SUMPRODUCT", "METHOD_BITFLIP", "METHOD_HARD"]
