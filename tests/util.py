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
