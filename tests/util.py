"""Shared helpers for the tests: seeded synthetic frames in the reference's conventions."""
import numpy as np

from oracle import oracle as O


def pack_bits_msb(bits):
    """(n, K) 0/1 -> (n, ceil(K/8)) bytes, MSB first (lib/ldpc_encoder_bc_impl.cc:138-147)."""
    bits = np.asarray(bits, np.uint8)
    n, K = bits.shape
    pad = (-K) % 8
    if pad:
        bits = np.concatenate([bits, np.zeros((n, pad), np.uint8)], axis=1)
    return np.packbits(bits, axis=1)


def unpack_bits_msb(bytes_, K):
    return np.unpackbits(np.asarray(bytes_, np.uint8), axis=1)[:, :K]


def oracle_encode_bits(data_bits, Hp, L, U, gf2=False):
    """(n, K) data bits -> (n, N) codeword bits [parity | data] via the oracle."""
    out = []
    for d in np.asarray(data_bits, np.int32):
        c, bad = O.make_parity_check(d, Hp, L, U, gf2=gf2)
        assert not bad
        out.append(np.concatenate([c, d]))
    return np.array(out, np.int32)


def bpsk(bits):
    """bit 1 -> +1+0j, bit 0 -> -1+0j (lib/ldpc_encoder_bc_impl.cc:154-165)."""
    return (2.0 * np.asarray(bits, np.float32) - 1.0).astype(np.complex64)


def awgn(sym, ebn0_db, rng, imag_noise=True):
    """Reference AWGN convention: sigma^2 = N0 = 10^(-EbN0/10) on the real axis
    (apps/ldpc_lapack.cpp:626-642).  Imaginary noise is added too; the decoder ignores it."""
    if ebn0_db is None:
        return sym.astype(np.complex64)
    sigma = np.sqrt(10.0 ** (-ebn0_db / 10.0))
    re = rng.standard_normal(sym.shape).astype(np.float32) * np.float32(sigma)
    im = rng.standard_normal(sym.shape).astype(np.float32) * np.float32(sigma) if imag_noise else 0
    return (sym + re + 1j * im).astype(np.complex64)


def synth_frames(Hp, L, U, n, ebn0_db, seed, gf2=False):
    """Seeded random data -> oracle encode -> BPSK -> AWGN.  -> (data_bits, codeword_bits, sym)."""
    rng = np.random.default_rng(seed)
    M, N = Hp.shape
    data = rng.integers(0, 2, size=(n, N - M)).astype(np.int32)
    cw = oracle_encode_bits(data, Hp, L, U, gf2=gf2)
    return data, cw, awgn(bpsk(cw), ebn0_db, rng)


def oracle_spa_batch(sym, Hp, iterations, early_stop, threshold=None):
    """Dense oracle over aligned frames -> (bytes, iters, synd)."""
    b, it, sy, _ = O.decode_frames(sym, Hp, method=1, iterations=iterations, early_stop=early_stop,
                                   threshold=threshold, threads=4)
    return b, it, sy


def csr_to_edges(row_ptr, col_idx):
    rows = np.repeat(np.arange(len(row_ptr) - 1), np.diff(row_ptr))
    return rows, np.asarray(col_idx)


def sparse_tables(row_ptr, col_idx, M, N):
    """CSR + per-column edge lists in the form oracle.decode_spa_sparse takes."""
    rows = np.repeat(np.arange(M), np.diff(row_ptr))
    order = np.lexsort((rows, col_idx))
    col_ptr = np.zeros(N + 1, np.int32)
    np.add.at(col_ptr, np.asarray(col_idx) + 1, 1)
    return (np.asarray(row_ptr, np.int32), np.asarray(col_idx, np.int32),
            np.cumsum(col_ptr).astype(np.int32), order.astype(np.int32))


def oracle_spa_sparse_batch(rx, tables, M, N, iterations, early_stop, threads=None):
    """Sparse restatement over a batch of real-valued frames, one frame per task on a thread pool
    (ctypes drops the GIL).  -> (vhat (n, N) uint8, iters (n,) int32)."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    rx = np.asarray(rx)
    n = rx.shape[0]
    vhat = np.zeros((n, N), np.uint8)
    iters = np.zeros(n, np.int32)

    def one(f):
        v, run = O.decode_spa_sparse(rx[f], tables, M, N, iterations, early_stop)
        vhat[f] = v
        iters[f] = run

    with ThreadPoolExecutor(threads or min(32, os.cpu_count() or 1)) as ex:
        list(ex.map(one, range(n)))
    return vhat, iters


def syndrome_weights(vhat, row_ptr, col_idx, M):
    """Number of unsatisfied checks of every row of vhat (n, N)."""
    rows = np.repeat(np.arange(M), np.diff(row_ptr))
    s = np.zeros((vhat.shape[0], M), np.int64)
    np.add.at(s.T, rows, vhat[:, col_idx].T.astype(np.int64))
    return (s & 1).sum(1)
