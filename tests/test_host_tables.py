"""CPU: the library's host-side code setup (GF(2) re-ordering on bit-packed rows, the
generator P, adjacency tables) against the oracle's dense restatement, through the C ABI
with LDPC535_DEVICE_NONE (tables only -- no compute happens here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import ldpc_ece535a as L
from ldpc_ece535a import _abi
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def parity_from_generator(P, d):
    K = d.size
    dw = np.zeros(P.shape[1] * 32, np.uint64)
    dw[:K] = d
    words = (dw.reshape(-1, 32) << np.arange(32, dtype=np.uint64)).sum(1).astype(np.uint32)
    return np.array([sum(bin(int(x)).count("1") for x in (row & words)) & 1 for row in P])


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "ldpc535.h")).read()
    declared = set(re.findall(r"LDPC535_API\s+[\w\s\*]+?\b(ldpc535_\w+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = _abi.lib()
    for name in sorted(declared):
        assert hasattr(lib, name), "library does not export " + name
    assert declared == set(_abi.SYMBOLS), "binding and header disagree: %s" % (
        declared ^ set(_abi.SYMBOLS))
    assert b"sm_100a" in lib.ldpc535_version()


@pytest.mark.parametrize("name", ["hData1", "hData2", "hData3", "hData4", "hData5"])
def test_tables_match_oracle(ref_codes, name):
    H = ref_codes[name]["H"]
    code = L.Code(H, device=-1)
    Hp, Lm, Um, piv = O.reorder_h(H)
    assert np.array_equal(code.pivots(), piv)
    assert np.array_equal(code.h_dense(), Hp)
    P = code.generator()
    rng = np.random.default_rng(7)
    for _ in range(40):
        d = rng.integers(0, 2, code.K).astype(np.int32)
        c, _ = O.make_parity_check(d, Hp, Lm, Um)
        assert np.array_equal(parity_from_generator(P, d), c)


def test_default_code_is_the_shipped_literal(shipped):
    code = L.Code(None, device=-1)
    assert (code.M, code.N, code.K, code.E) == (32, 64, 32, 168)
    assert np.array_equal(code.h_dense(), shipped["Hp"])
    assert np.array_equal(code.pivots(), shipped["pivots"])


def test_medium_regular_code_matches_oracle():
    row_ptr, col_idx, M, N = L.codes.regular_code(n=512, seed=535)
    code = None
    for seed in range(535, 545):
        r = L.codes.regular_code(n=512, seed=seed)
        try:
            code = L.Code(r, device=-1)
            break
        except L.Ldpc535Error as e:
            assert e.status == _abi.ERR_SINGULAR
    assert code is not None
    H = L.codes.to_dense(*r)
    assert (H.sum(0) == 3).all() and (H.sum(1) == 6).all()
    Hp, Lm, Um, piv = O.reorder_h(H)
    assert np.array_equal(code.pivots(), piv)
    assert np.array_equal(code.h_dense(), Hp)
    P = code.generator()
    rng = np.random.default_rng(1)
    for _ in range(5):
        d = rng.integers(0, 2, code.K).astype(np.int32)
        c, _ = O.make_parity_check(d, Hp, Lm, Um, gf2=True)
        assert np.array_equal(parity_from_generator(P, d), c)


def test_8192_code_structure():
    """Config 4's code: too big for the dense oracle re-ordering, so check structure:
    H_perm is a column permutation of H and P's parity satisfies every check."""
    r = L.codes.regular_code(n=8192, seed=535)
    code = L.Code(r, device=-1)
    assert (code.M, code.N, code.K, code.E) == (4096, 8192, 4096, 24576)
    row_ptr, col_idx = code.h_csr()
    assert (np.diff(row_ptr) == 6).all()
    assert (np.bincount(col_idx, minlength=8192) == 3).all()
    assert code.kernel_name() in ("block", "unsupported")   # no device: family only
    P = code.generator()
    rng = np.random.default_rng(2)
    rows = np.repeat(np.arange(4096), 6)
    for _ in range(3):
        d = rng.integers(0, 2, 4096).astype(np.int64)
        dw = (d.reshape(-1, 32).astype(np.uint64) << np.arange(32, dtype=np.uint64)).sum(1).astype(np.uint32)
        anded = P & dw
        c = np.zeros(4096, np.int64)
        for sh in range(32):
            c ^= ((anded >> np.uint32(sh)) & np.uint32(1)).sum(1).astype(np.int64) & 1
        x = np.concatenate([c & 1, d])
        synd = np.bincount(rows, weights=x[col_idx], minlength=4096).astype(np.int64) & 1
        assert not synd.any()


def test_singular_and_invalid_inputs():
    H = np.zeros((4, 8), np.int32)
    H[0, 0] = H[1, 1] = H[2, 2] = 1            # row 3 empty -> no pivot
    with pytest.raises(L.Ldpc535Error) as e:
        L.Code(H, device=-1)
    assert e.value.status == _abi.ERR_SINGULAR
    with pytest.raises(L.Ldpc535Error) as e:
        L.Code(np.ones((4, 4), np.int32), device=-1)     # N <= M
    assert e.value.status == _abi.ERR_INVALID


def test_compute_refuses_without_device():
    code = L.Code(None, device=-1)
    with pytest.raises(L.Ldpc535Error) as e:
        code.encode(np.zeros(4, np.uint8))
    assert e.value.status == _abi.ERR_NO_DEVICE
    with pytest.raises(L.Ldpc535Error) as e:
        code.decode(np.zeros(64, np.complex64))
    assert e.value.status == _abi.ERR_NO_DEVICE
