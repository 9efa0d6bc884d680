"""Parity-check matrix constructors.

The reference holds no code constructor (its matrices were pasted in from MATLAB's
makeLdpc, lib/ldpc_decoder_cb_impl.cc:60-62), so the (3,6)-regular n = 8192 code of
BASELINE.json's config 4 comes from this seeded generator.  Host-side numpy only.
"""
import numpy as np


def regular_code(n=8192, dv=3, dc=6, seed=535):
    """Seeded (dv, dc)-regular parity-check matrix as CSR: (row_ptr, col_idx, M, N).

    Socket-permutation (configuration-model) construction; double edges are repaired
    by swapping sockets with random partners.  Deterministic for a given (n, dv, dc,
    seed) -- numpy PCG64 streams are stable across versions.  The caller (Code(...))
    runs the reference's column re-ordering; if it reports a singular matrix, use the
    next seed (`first_invertible`)."""
    if (n * dv) % dc:
        raise ValueError("n * dv must be divisible by dc")
    M = n * dv // dc
    rng = np.random.Generator(np.random.PCG64(seed))
    var_of_socket = np.repeat(np.arange(n, dtype=np.int64), dv)
    perm = rng.permutation(n * dv)
    for _ in range(1000):
        chk = perm // dc                          # check of every variable socket
        key = var_of_socket * M + chk[np.arange(n * dv)]
        order = np.argsort(key, kind="stable")
        dup = np.zeros(n * dv, bool)
        dup[order[1:]] = key[order[1:]] == key[order[:-1]]
        bad = np.nonzero(dup)[0]
        if bad.size == 0:
            break
        partners = rng.integers(0, n * dv, size=bad.size)
        for a, b in zip(bad, partners):
            perm[a], perm[b] = perm[b], perm[a]
    else:
        raise RuntimeError("could not remove double edges")
    chk = perm // dc
    order = np.lexsort((var_of_socket, chk))      # by check, then variable
    col_idx = var_of_socket[order].astype(np.int32)
    row_ptr = (np.arange(M + 1, dtype=np.int64) * dc).astype(np.int32)
    return row_ptr, col_idx, M, n


def to_dense(row_ptr, col_idx, M, N):
    H = np.zeros((M, N), np.int32)
    for j in range(M):
        H[j, col_idx[row_ptr[j]:row_ptr[j + 1]]] = 1
    return H


def first_invertible(n=8192, dv=3, dc=6, seed=535, tries=64, device=0):
    """First seed >= `seed` whose matrix the reference's re-ordering can factor; returns
    (Code, seed)."""
    from ._abi import ERR_SINGULAR, Ldpc535Error
    from .code import Code
    for s in range(seed, seed + tries):
        try:
            return Code(regular_code(n, dv, dc, s), device=device), s
        except Ldpc535Error as e:
            if e.status != ERR_SINGULAR:
                raise
    raise RuntimeError("no invertible code in %d seeds" % tries)
