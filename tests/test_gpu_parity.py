"""GPU parity tests proper: the sm_100a kernels, called through the C ABI
(include/ldpc535.h via the ctypes host package), against the CPU oracle on the same seeded
inputs, against the committed golden fixtures, and -- at sizes the oracle cannot reach --
through size-independent properties (H c = 0, encode -> decode round trips, linearity).

Bars: encoder / syndromes / hard decisions / iteration counts bit-exact; sum-product
messages |gpu - ref| <= 1e-4 |ref| + 2e-6 against the fp64 oracle (the kernels compute in
fp32; the absolute floor covers messages that are ~0)."""
import json
import os

import numpy as np
import pytest

import ldpc_ece535a as L
from oracle import oracle as O
import util

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
RTOL, ATOL = 1e-4, 2e-6

C4_KERNELS = ["warp", "block", "c4-thread"]


@pytest.fixture(scope="module")
def c4():
    code = L.Code(None, device=0)
    yield code
    code.close()


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "c4_golden.npz"))


def with_kernel(code, name):
    try:
        code.set_kernel(name)
    except L.Ldpc535Error as e:
        pytest.skip("kernel family %s not available: %s" % (name, e))


def bits_to_cw_sym(cw_bits):
    return util.bpsk(cw_bits)


def sym_to_bits(sym):
    s = np.asarray(sym).reshape(-1)
    assert np.all(s.imag == 0) and np.all(np.abs(s.real) == 1)
    return (s.real > 0).astype(np.uint8)


# ---------------------------------------------------------------------------------------
# encoder
# ---------------------------------------------------------------------------------------
def test_encoder_reference_qa_kat(ref_codes):
    """The reference's own encoder known-answer test (python/qa_ldpc_encoder_bc.py:23-46),
    on the 8x16 code it was written for, through the CUDA path."""
    with open(os.path.join(GOLD, "ref_qa_kat.json")) as f:
        k = json.load(f)
    frames = []
    for c, d in zip(k["mod_check"], k["mod_data"]):
        frames += c + d
    code = L.Code(ref_codes[k["code"]]["H"], device=0)
    out = code.encode(np.array(k["data_bytes"], np.uint8))
    assert np.array_equal(out.reshape(-1), np.array(frames, np.complex64))


def test_encoder_golden_source_words(c4, gold):
    """apps/test_data.h:180-213 source words -> frozen oracle codewords."""
    out = c4.encode(util.pack_bits_msb(gold["enc_data_bits"]))
    assert out.shape == (30, 64)
    assert np.array_equal(sym_to_bits(out).reshape(30, 64), gold["enc_codewords"])


@pytest.mark.parametrize("name", ["hData1", "hData2", "hData3", "hData4", "hData5"])
def test_encoder_vs_oracle_all_reference_codes(ref_codes, name):
    H = ref_codes[name]["H"]
    Hp, Lm, Um, _ = O.reorder_h(H)
    M, N = H.shape
    K = N - M
    code = L.Code(H, device=0)
    rng = np.random.default_rng(11)
    n = 257                                               # ragged: not a multiple of any tile
    data = rng.integers(0, 2, (n, K)).astype(np.int32)
    if K % 8:
        pytest.skip("K not a whole number of bytes: the reference block cannot frame it")
    out = code.encode(util.pack_bits_msb(data))
    want = util.oracle_encode_bits(data, Hp, Lm, Um)
    assert np.array_equal(sym_to_bits(out).reshape(n, N), want)
    # every codeword satisfies every check of the re-ordered H
    assert not ((want @ Hp.T) % 2).any()
    code.close()


def test_encoder_block_level_matches_oracle_work(c4, shipped):
    """Byte stream -> symbol stream exactly as ldpc_encoder_bc_impl::general_work emits it."""
    rng = np.random.default_rng(5)
    data = rng.integers(0, 256, 4 * 1000).astype(np.uint8)
    want, consumed = O.encoder_work(shipped["Hp"], shipped["L"], shipped["U"], data, 1000 * 64)
    assert consumed == data.size
    got = c4.encode(data).reshape(-1)
    assert np.array_equal(got, want)


def test_encoder_empty_and_single(c4):
    assert c4.encode(np.zeros(0, np.uint8)).shape == (0, 64)
    one = c4.encode(np.zeros(4, np.uint8))
    assert np.array_equal(one.reshape(-1), np.full(64, -1 + 0j, np.complex64))
    with pytest.raises(ValueError):
        c4.encode(np.zeros(3, np.uint8))


def test_encoder_linearity_and_syndrome_1m(c4, shipped):
    """1 Mi random words: H c = 0 for every frame and enc(a ^ b) = enc(a) ^ enc(b)."""
    rng = np.random.default_rng(6)
    n = 1 << 20
    a = rng.integers(0, 256, (n, 4)).astype(np.uint8)
    b = rng.integers(0, 256, (n, 4)).astype(np.uint8)
    ca = (c4.encode(a).real > 0)
    cb = (c4.encode(b).real > 0)
    cab = (c4.encode(a ^ b).real > 0)
    assert np.array_equal(ca ^ cb, cab)
    Hp = shipped["Hp"].astype(np.uint8)
    synd = (ca.astype(np.uint8) @ Hp.T) & 1
    assert not synd.any()
    # data bits are carried verbatim in the second half
    assert np.array_equal(np.packbits(ca[:, 32:], axis=1), a)


# ---------------------------------------------------------------------------------------
# decoder: goldens and oracle parity on the shipped code
# ---------------------------------------------------------------------------------------
POINTS = (("2dB_5it", 5, 1), ("4dB_5it", 5, 1), ("2dB_50it_noearly", 50, 0), ("clean_5it", 5, 1))


@pytest.mark.parametrize("kernel", C4_KERNELS)
@pytest.mark.parametrize("tag,iters,early", POINTS)
def test_decoder_spa_goldens(c4, gold, kernel, tag, iters, early):
    with_kernel(c4, kernel)
    b, sy, it = c4.decode(gold["dec_%s_sym" % tag], method=1, max_iters=iters, early_stop=early)
    assert np.array_equal(b, gold["dec_%s_spa_bytes" % tag])
    assert np.array_equal(it, gold["dec_%s_spa_iters" % tag])
    assert np.array_equal(sy, gold["dec_%s_spa_synd" % tag])
    c4.set_kernel(None)


@pytest.mark.parametrize("kernel", ["warp", "block"])
@pytest.mark.parametrize("tag,iters,early", POINTS)
def test_decoder_minsum_goldens(c4, gold, kernel, tag, iters, early):
    with_kernel(c4, kernel)
    b, sy, it = c4.decode(gold["dec_%s_sym" % tag], method=0, max_iters=iters, early_stop=early)
    assert np.array_equal(b, gold["dec_%s_minsum_bytes" % tag])
    assert np.array_equal(it, gold["dec_%s_minsum_iters" % tag])
    assert np.array_equal(sy, gold["dec_%s_minsum_synd" % tag])
    c4.set_kernel(None)


@pytest.mark.parametrize("kernel", C4_KERNELS)
@pytest.mark.parametrize("ebn0,iters,early", [(0.0, 5, 1), (2.0, 5, 1), (4.0, 5, 1), (6.0, 5, 1),
                                              (2.0, 50, 0), (2.0, 50, 1), (4.0, 1, 1)])
def test_decoder_spa_vs_oracle_fixed_seed(c4, shipped, kernel, ebn0, iters, early):
    """Config 1: 1 000 seeded codewords, encode -> BPSK + AWGN -> sum-product; hard
    decisions, iteration counts and saturating syndrome weights identical to the oracle."""
    with_kernel(c4, kernel)
    n = 1000 if iters <= 5 else 300
    _, _, sym = util.synth_frames(shipped["Hp"], shipped["L"], shipped["U"], n, ebn0,
                                  seed=int(ebn0 * 10) + iters)
    wb, wit, wsy = util.oracle_spa_batch(sym, shipped["Hp"], iters, early)
    b, sy, it = c4.decode(sym, method=1, max_iters=iters, early_stop=early)
    assert np.array_equal(b, wb)
    assert np.array_equal(it, wit)
    assert np.array_equal(sy, wsy)
    c4.set_kernel(None)


@pytest.mark.parametrize("method", [0, 2, 3, 7])
@pytest.mark.parametrize("kernel", ["warp", "block"])
def test_decoder_other_methods_vs_oracle(c4, shipped, method, kernel):
    """min-sum (0 and any unknown value), bit-flip (2), hard (3)."""
    with_kernel(c4, kernel)
    _, _, sym = util.synth_frames(shipped["Hp"], shipped["L"], shipped["U"], 500, 3.0, seed=77)
    wb, wit, wsy, _ = O.decode_frames(sym, shipped["Hp"], method=method, iterations=5,
                                      early_stop=True, threads=4)
    b, sy, it = c4.decode(sym, method=method, max_iters=5)
    assert np.array_equal(b, wb)
    assert np.array_equal(sy, wsy)
    if method in (0, 7):
        assert np.array_equal(it, wit)
    c4.set_kernel(None)


@pytest.mark.parametrize("kernel", C4_KERNELS)
def test_decoder_messages_within_tolerance(c4, shipped, kernel):
    """L, E, M after 1..5 iterations against the fp64 oracle: rtol 1e-4 (+ 2e-6 floor)."""
    with_kernel(c4, kernel)
    Hp = shipped["Hp"]
    rows, cols = np.nonzero(Hp)
    _, _, sym = util.synth_frames(Hp, shipped["L"], shipped["U"], 48, 2.0, seed=123)
    for iters, early in ((1, 0), (2, 0), (5, 0), (5, 1)):
        got = c4.decode_debug(sym, max_iters=iters, early_stop=early, kernel=kernel)
        for f in range(sym.shape[0]):
            vhat, run, Lr, Er, Mr = O.decode_spa(sym[f].real, Hp, iters, early, debug=True)
            assert got["iters"][f] == run
            for name, ref in (("L", Lr), ("E", Er[rows, cols]), ("M", Mr[rows, cols])):
                g = got[name][f].astype(np.float64)
                assert np.all(np.abs(g - ref) <= RTOL * np.abs(ref) + ATOL), (
                    kernel, iters, early, f, name, np.max(np.abs(g - ref)))
    c4.set_kernel(None)


@pytest.mark.parametrize("kernel", C4_KERNELS)
def test_decoder_reference_qa_kat_shape_noiseless(c4, shipped, kernel):
    """Noiseless frames decode to their data in one iteration with zero syndrome."""
    with_kernel(c4, kernel)
    data, cw, sym = util.synth_frames(shipped["Hp"], shipped["L"], shipped["U"], 333, None, seed=1)
    b, sy, it = c4.decode(sym, method=1)
    assert np.array_equal(b, util.pack_bits_msb(data))
    assert (sy == 0).all() and (it == 1).all()
    c4.set_kernel(None)


def test_decoder_reference_qa_kat(ref_codes):
    """python/qa_ldpc_decoder_cb.py:20-43 on the 8x16 code, every method."""
    with open(os.path.join(GOLD, "ref_qa_kat.json")) as f:
        k = json.load(f)
    frames = []
    for c, d in zip(k["mod_check"], k["mod_data"]):
        frames += c + d
    code = L.Code(ref_codes[k["code"]]["H"], device=0)
    for method in (0, 1, 2, 3):
        b, sy, it = code.decode(np.array(frames, np.complex64), method=method)
        assert list(b.reshape(-1)) == k["data_bytes"]
        assert (sy == 0).all()
    code.close()


@pytest.mark.parametrize("kernel", C4_KERNELS)
def test_decoder_windows_and_polarity(c4, shipped, kernel):
    """Arbitrary symbol offsets and -1 polarity (the sync machine's slide / inverted retry,
    lib/ldpc_decoder_cb_impl.cc:151-152, :178-199) equal the oracle on the shifted, negated
    window."""
    with_kernel(c4, kernel)
    Hp = shipped["Hp"]
    _, _, sym = util.synth_frames(Hp, shipped["L"], shipped["U"], 40, 3.0, seed=8)
    stream = sym.reshape(-1)
    rng = np.random.default_rng(9)
    offs = rng.integers(0, stream.size - 64 + 1, 300).astype(np.int64)
    offs[:3] = (0, stream.size - 64, 1)
    pol = rng.choice(np.array([-1, 1], np.int8), 300)
    wins = np.stack([stream[o:o + 64] * p for o, p in zip(offs, pol)])
    wb, wit, wsy = util.oracle_spa_batch(wins, Hp, 5, True)
    b, sy, it = c4.decode(stream, method=1, win_offset=offs, polarity=pol)
    assert np.array_equal(b, wb) and np.array_equal(it, wit) and np.array_equal(sy, wsy)
    with pytest.raises(L.Ldpc535Error):
        c4.decode(stream, win_offset=np.array([stream.size - 63], np.int64))
    with pytest.raises(L.Ldpc535Error):
        c4.decode(stream, win_offset=np.array([-1], np.int64))
    c4.set_kernel(None)


@pytest.mark.parametrize("kernel", C4_KERNELS)
@pytest.mark.parametrize("n", [0, 1, 31, 33, 1025])
def test_decoder_ragged_batches(c4, shipped, kernel, n):
    with_kernel(c4, kernel)
    _, _, sym = util.synth_frames(shipped["Hp"], shipped["L"], shipped["U"], max(n, 1), 2.0, seed=n)
    sym = sym[:n]
    b, sy, it = c4.decode(sym.reshape(-1), method=1, n_win=n)
    if n:
        wb, wit, wsy = util.oracle_spa_batch(sym, shipped["Hp"], 5, True)
        assert np.array_equal(b, wb) and np.array_equal(it, wit) and np.array_equal(sy, wsy)
    else:
        assert b.shape == (0, 4)
    c4.set_kernel(None)


def test_decoder_kernel_families_agree_1m(c4, shipped):
    """1 Mi noisy frames: every kernel family returns the same bytes / iterations /
    syndromes (a checksum of checksums across implementations), and the noiseless
    round trip returns the data."""
    rng = np.random.default_rng(10)
    n = 1 << 20
    data = rng.integers(0, 256, (n, 4)).astype(np.uint8)
    clean = c4.encode(data)
    b, sy, it = c4.decode(clean, method=1)
    assert np.array_equal(b, data) and not sy.any() and (it == 1).all()
    sigma = np.float32(np.sqrt(10.0 ** (-2.0 / 10.0)))
    noisy = clean.copy()
    noisy.real += rng.standard_normal(clean.shape, dtype=np.float32) * sigma
    res = {}
    for kern in C4_KERNELS:
        try:
            c4.set_kernel(kern)
        except L.Ldpc535Error:
            continue
        res[kern] = c4.decode(noisy, method=1)
    c4.set_kernel(None)
    assert len(res) >= 2
    base = res.pop("warp")
    for kern, r in res.items():
        for a, b2 in zip(base, r):
            assert np.array_equal(a, b2), kern
    # spot-check 2 000 of them against the oracle
    idx = rng.choice(n, 2000, replace=False)
    wb, wit, wsy = util.oracle_spa_batch(noisy[idx], shipped["Hp"], 5, True)
    assert np.array_equal(base[0][idx], wb)
    assert np.array_equal(base[2][idx], wit)
    assert np.array_equal(base[1][idx], wsy)


# ---------------------------------------------------------------------------------------
# other reference codes and the n = 8192 code (CTA-per-codeword kernel)
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["hData1", "hData2", "hData5"])
def test_decoder_other_reference_codes(ref_codes, name):
    H = ref_codes[name]["H"]
    Hp, Lm, Um, _ = O.reorder_h(H)
    if (H.shape[1] - H.shape[0]) % 8:
        pytest.skip("K not a whole number of bytes")
    code = L.Code(H, device=0)
    _, _, sym = util.synth_frames(Hp, Lm, Um, 200, 4.0, seed=3)
    for method in (1, 0):
        wb, wit, wsy, _ = O.decode_frames(sym, Hp, method=method, iterations=5, early_stop=True,
                                          threads=4)
        b, sy, it = code.decode(sym, method=method)
        assert np.array_equal(b, wb) and np.array_equal(it, wit) and np.array_equal(sy, wsy)
    code.close()


@pytest.fixture(scope="module")
def c8k():
    code, seed = L.codes.first_invertible(n=8192, seed=535, device=0)
    yield code
    code.close()


def c8k_tables(code):
    row_ptr, col_idx = code.h_csr()
    rows = np.repeat(np.arange(code.M), np.diff(row_ptr))
    order = np.lexsort((rows, col_idx))
    col_ptr = np.zeros(code.N + 1, np.int32)
    np.add.at(col_ptr, col_idx + 1, 1)
    return (row_ptr, col_idx, np.cumsum(col_ptr).astype(np.int32), order.astype(np.int32)), rows


def test_c8k_encode_satisfies_every_check(c8k):
    rng = np.random.default_rng(12)
    n = 300
    data = rng.integers(0, 256, (n, c8k.nbytes)).astype(np.uint8)
    out = c8k.encode(data)
    bits = sym_to_bits(out).reshape(n, c8k.N)
    (row_ptr, col_idx, _, _), rows = c8k_tables(c8k)
    for f in range(n):
        synd = np.bincount(rows, weights=bits[f][col_idx], minlength=c8k.M).astype(np.int64) & 1
        assert not synd.any()
    assert np.array_equal(np.packbits(bits[:, c8k.M:], axis=1), data)
    # generator parity equals the host table product
    P = c8k.generator()
    d = np.unpackbits(data[0]).astype(np.uint64)
    dw = (d.reshape(-1, 32) << np.arange(32, dtype=np.uint64)).sum(1).astype(np.uint32)
    anded = P & dw
    c = np.zeros(c8k.M, np.int64)
    for sh in range(32):
        c ^= ((anded >> np.uint32(sh)) & np.uint32(1)).sum(1).astype(np.int64) & 1
    assert np.array_equal(c, bits[0][:c8k.M])


@pytest.mark.parametrize("kernel", ["regular", "block"])
def test_c8k_decode_vs_sparse_oracle(c8k, kernel):
    """Config 4's code: decisions and iteration counts vs the sparse restatement (checked
    bit-identical to the dense one on the shipped code by the CPU suite)."""
    with_kernel(c8k, kernel)
    assert c8k.kernel_name() == kernel
    rng = np.random.default_rng(13)
    n = 12
    data = rng.integers(0, 256, (n, c8k.nbytes)).astype(np.uint8)
    clean = c8k.encode(data)
    tables, _ = c8k_tables(c8k)
    for ebn0, iters in ((4.0, 20), (2.0, 50)):
        sigma = np.float32(np.sqrt(10.0 ** (-ebn0 / 10.0)))
        noisy = clean.copy()
        noisy.real += rng.standard_normal(clean.shape, dtype=np.float32) * sigma
        b, sy, it = c8k.decode(noisy, method=1, max_iters=iters, early_stop=True)
        for f in range(n):
            vhat, run = O.decode_spa_sparse(noisy[f].real, tables, c8k.M, c8k.N, iters, True)
            assert it[f] == run, (ebn0, f)
            assert np.array_equal(np.packbits(vhat[c8k.M:].astype(np.uint8)), b[f]), (ebn0, f)
    b, sy, it = c8k.decode(clean, method=1)
    assert np.array_equal(b, data) and not sy.any() and (it == 1).all()
    # fixed iterations (no early stop) and a short limit that most frames do not reach
    sigma = np.float32(np.sqrt(10.0 ** (-2.0 / 10.0)))
    noisy = clean.copy()
    noisy.real += rng.standard_normal(clean.shape, dtype=np.float32) * sigma
    for iters, early in ((6, False), (9, True)):
        b, sy, it = c8k.decode(noisy, method=1, max_iters=iters, early_stop=early)
        for f in range(4):
            vhat, run = O.decode_spa_sparse(noisy[f].real, tables, c8k.M, c8k.N, iters, early)
            assert it[f] == run
            assert np.array_equal(np.packbits(vhat[c8k.M:].astype(np.uint8)), b[f])
    c8k.set_kernel(None)


def test_c8k_kernels_agree(c8k):
    """regular vs generic CTA kernel on the same noisy frames: identical bytes, syndrome weights
    (saturating) and iteration counts."""
    rng = np.random.default_rng(16)
    n = 300
    data = rng.integers(0, 256, (n, c8k.nbytes)).astype(np.uint8)
    noisy = c8k.encode(data)
    noisy.real += rng.standard_normal(noisy.shape, dtype=np.float32) * np.float32(0.8)
    res = {}
    for kern in ("regular", "block"):
        c8k.set_kernel(kern)
        res[kern] = [c8k.decode(noisy, method=1, max_iters=mi, early_stop=es) for mi, es in ((50, True), (8, True), (5, False))]
    c8k.set_kernel(None)
    for a, b in zip(res["regular"], res["block"]):
        for x, y in zip(a, b):
            assert np.array_equal(x, y)


@pytest.mark.parametrize("variant", [None, "60/8", "40/7", "42", "62"])
def test_c8k_lookup_encoder_equals_scan_encoder(c8k, monkeypatch, variant):
    """The table look-up encoders (default for this size: the free-running encode_m4r_fr_kernel with a
    6-slot ring and 7 frames per slot; LDPC535_M4R_RING / LDPC535_M4R_TPF select the 4-slot ring, 8
    frames per slot and the barrier kernel encode_m4r_kernel) against the AND/XOR scan
    (encode_generic_kernel, LDPC535_ENCODER=generic): identical symbols for ragged batch sizes --
    partial tiles, tiles smaller than a CTA, fewer units than SMs, several units per CTA (the parked
    parity of one unit drained during the next) -- device-resident for the large ones."""
    import torch
    monkeypatch.setenv("LDPC535_ENCODER", "generic")
    scan = L.Code(c8k.h_csr() + (c8k.M, c8k.N), device=0)
    monkeypatch.delenv("LDPC535_ENCODER")
    if variant is None:
        look = c8k
    else:
        ring, _, tpf = variant.partition("/")
        monkeypatch.setenv("LDPC535_M4R_RING", ring)
        if tpf:
            monkeypatch.setenv("LDPC535_M4R_TPF", tpf)
        look = L.Code(c8k.h_csr() + (c8k.M, c8k.N), device=0)
        monkeypatch.delenv("LDPC535_M4R_RING")
        monkeypatch.delenv("LDPC535_M4R_TPF", raising=False)
    rng = np.random.default_rng(18)
    for n in (1, 257, 2047, 2048, 2049, 3073, 4100):       # the look-up kernel takes over at 2048 frames
        data = rng.integers(0, 256, (n, c8k.nbytes)).astype(np.uint8)
        assert np.array_equal(look.encode(data), scan.encode(data)), n
    sms = L.device_info(0)["sm_count"]
    gen = torch.Generator(device="cuda")
    gen.manual_seed(18)
    for n in (sms * 32 * 2 + 37, sms * 64 * 4 + 1, sms * 64 * 8 + 129, sms * 128 * 8 + 5, sms * 217 * 3 + 11):
        d = torch.randint(0, 256, (n, c8k.nbytes), dtype=torch.uint8, device="cuda", generator=gen)
        a = torch.empty((n, c8k.N, 2), dtype=torch.float32, device="cuda")
        b = torch.empty((n, c8k.N, 2), dtype=torch.float32, device="cuda")
        torch.cuda.synchronize()
        look.encode_dev(d.data_ptr(), n, a.data_ptr())
        scan.encode_dev(d.data_ptr(), n, b.data_ptr())
        look.sync()
        scan.sync()
        assert torch.equal(a, b), n
        del a, b, d
    scan.close()
    if look is not c8k:
        look.close()


def test_c8k_regular_kernel_variants_agree(c8k, monkeypatch):
    """The 512-thread register-table kernel (default for this size) and the 1024-thread kernel with
    the tables in shared memory (LDPC535_REGULAR_VARIANT=0): identical bytes, syndrome weights and
    iteration counts, for batch sizes around the grid size, with search windows and polarity."""
    monkeypatch.setenv("LDPC535_REGULAR_VARIANT", "0")
    other = L.Code(c8k.h_csr() + (c8k.M, c8k.N), device=0)
    monkeypatch.delenv("LDPC535_REGULAR_VARIANT")
    assert c8k.kernel_name() == "regular" and other.kernel_name() == "regular"
    rng = np.random.default_rng(17)
    n = 300
    data = rng.integers(0, 256, (n, c8k.nbytes)).astype(np.uint8)
    noisy = c8k.encode(data)
    assert np.array_equal(noisy, other.encode(data))
    noisy.real += rng.standard_normal(noisy.shape, dtype=np.float32) * np.float32(0.8)
    for cnt in (1, 2, 147, 148, 149, 300):
        for mi, es in ((50, True), (7, False)):
            a = c8k.decode(noisy[:cnt], method=1, max_iters=mi, early_stop=es)
            b = other.decode(noisy[:cnt], method=1, max_iters=mi, early_stop=es)
            for x, y in zip(a, b):
                assert np.array_equal(x, y), (cnt, mi, es)
    # windows at arbitrary offsets with both polarities
    flat = noisy[:4].reshape(-1)
    offs = np.array([0, 8192, 5, 3 * 8192, 3 * 8192 - 1, 2 * 8192 - 7], np.int64)
    pol = np.array([1, -1, 1, -1, 1, -1], np.int8)
    a = c8k.decode(flat, method=1, max_iters=10, early_stop=True, win_offset=offs, polarity=pol)
    b = other.decode(flat, method=1, max_iters=10, early_stop=True, win_offset=offs, polarity=pol)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    with pytest.raises(L.Ldpc535Error):          # a window past the end is refused by the host API
        c8k.decode(flat, method=1, win_offset=np.array([3 * 8192 + 1], np.int64))
    other.close()


@pytest.fixture(scope="module")
def c8k_low_snr(c8k):
    """504 frames of BASELINE config 4's code at Eb/N0 = 1.0 / 1.25 / 1.5 dB -- at and below the
    code's threshold: every frame at 1.0 dB runs all 50 iterations without converging, 1.25 dB is
    mixed, 1.5 dB converges late (18-45 iterations) -- with the sparse restatement's decisions and
    iteration counts (50 iterations max, early stop on: the reference's decodeSumProductSoft
    semantics, lib/ldpc_decoder_cb_impl.cc:478-557)."""
    rng = np.random.default_rng(8192)
    per = 168
    data = rng.integers(0, 256, (3 * per, c8k.nbytes)).astype(np.uint8)
    noisy = c8k.encode(data)
    for k, ebn0 in enumerate((1.0, 1.25, 1.5)):
        sigma = np.float32(np.sqrt(10.0 ** (-ebn0 / 10.0)))
        noisy[k * per:(k + 1) * per].real += rng.standard_normal((per, c8k.N), dtype=np.float32) * sigma
    row_ptr, col_idx = c8k.h_csr()
    tables = util.sparse_tables(row_ptr, col_idx, c8k.M, c8k.N)
    vhat, iters = util.oracle_spa_sparse_batch(noisy.real, tables, c8k.M, c8k.N, 50, True)
    synd = util.syndrome_weights(vhat, row_ptr, col_idx, c8k.M)
    return {"sym": noisy, "vhat": vhat, "iters": iters, "synd": synd, "per": per, "tables": tables}


@pytest.mark.parametrize("kernel", ["regular-rt", "regular-1024", "block"])
def test_c8k_low_snr_500_frames_vs_sparse_oracle(c8k, c8k_low_snr, kernel, monkeypatch):
    """Non-converging and late-converging frames (the domain VERDICT r1 flagged as unmeasured): hard
    decisions of ALL N bits' data half, iteration counts and saturating syndrome weights identical
    to the fp64 restatement on 504 frames, for both (3,6)-regular kernels and the generic CTA kernel."""
    g = c8k_low_snr
    assert (g["synd"][:g["per"]] > 0).sum() > 0.9 * g["per"]                              # 1.0 dB: hardly any converges
    assert ((g["iters"] == 50) == (g["synd"] > 0)).all() or (g["iters"][g["synd"] == 0] <= 50).all()
    assert (g["iters"][2 * g["per"]:] < 50).sum() > g["per"] // 2                         # 1.5 dB: most do
    code = c8k
    if kernel == "regular-1024":
        monkeypatch.setenv("LDPC535_REGULAR_VARIANT", "0")
        code = L.Code(c8k.h_csr() + (c8k.M, c8k.N), device=0)
        monkeypatch.delenv("LDPC535_REGULAR_VARIANT")
    code.set_kernel("block" if kernel == "block" else "regular")
    b, sy, it = code.decode(g["sym"], method=1, max_iters=50, early_stop=True, synd_threshold=253)
    code.set_kernel(None)
    if code is not c8k:
        code.close()
    want = np.packbits(g["vhat"][:, c8k.M:], axis=1)
    bad_frames = np.nonzero((b != want).any(axis=1))[0]
    assert bad_frames.size == 0, ("decisions differ on frames", bad_frames[:10], it[bad_frames[:10]])
    assert np.array_equal(it, g["iters"]), np.nonzero(it != g["iters"])[0][:10]
    assert np.array_equal(sy.astype(np.int64), np.minimum(g["synd"], 254))


def test_c8k_early_stop_off_saturated_domain(c8k):
    """Early stop OFF on the n = 8192 code: converged frames keep iterating, the reference's fp64
    tanh rounds to 1 and its messages become +-inf (all 24 576 of them a few iterations after
    convergence) while the fp32 kernels saturate later.  Signs agree at every bit, so
    no inf - inf arises and the decisions are the same; checked at 30 and 60 iterations."""
    rng = np.random.default_rng(8193)
    n = 48
    data = rng.integers(0, 256, (n, c8k.nbytes)).astype(np.uint8)
    noisy = c8k.encode(data)
    for k, ebn0 in enumerate((1.5, 2.0, 3.0)):
        sigma = np.float32(np.sqrt(10.0 ** (-ebn0 / 10.0)))
        noisy[k * 16:(k + 1) * 16].real += rng.standard_normal((16, c8k.N), dtype=np.float32) * sigma
    row_ptr, col_idx = c8k.h_csr()
    tables = util.sparse_tables(row_ptr, col_idx, c8k.M, c8k.N)
    for iters in (30, 60):
        vhat, run = util.oracle_spa_sparse_batch(noisy.real, tables, c8k.M, c8k.N, iters, False)
        assert (run == iters).all()
        for kern in ("regular", "block"):
            c8k.set_kernel(kern)
            b, sy, it = c8k.decode(noisy, method=1, max_iters=iters, early_stop=False, synd_threshold=253)
            c8k.set_kernel(None)
            assert np.array_equal(b, np.packbits(vhat[:, c8k.M:], axis=1)), (iters, kern)
            assert (it == iters).all()
            assert np.array_equal(sy.astype(np.int64),
                                  np.minimum(util.syndrome_weights(vhat, row_ptr, col_idx, c8k.M), 254))


def test_c8k_messages_within_tolerance(c8k):
    rng = np.random.default_rng(14)
    data = rng.integers(0, 256, (2, c8k.nbytes)).astype(np.uint8)
    noisy = c8k.encode(data)
    noisy.real += rng.standard_normal(noisy.shape, dtype=np.float32) * np.float32(0.7)
    tables, _ = c8k_tables(c8k)
    got = c8k.decode_debug(noisy, max_iters=3, early_stop=False)
    for f in range(2):
        vhat, run, Lr, Er, Mr = O.decode_spa_sparse(noisy[f].real, tables, c8k.M, c8k.N, 3, False,
                                                    debug=True)
        for name, ref in (("L", Lr), ("E", Er), ("M", Mr)):
            g = got[name][f].astype(np.float64)
            assert np.all(np.abs(g - ref) <= RTOL * np.abs(ref) + ATOL), (f, name)


# ---------------------------------------------------------------------------------------
# device-resident API and bookkeeping
# ---------------------------------------------------------------------------------------
def test_device_api_matches_host_api(c4, shipped):
    import torch
    rng = np.random.default_rng(15)
    n = 5000
    data = rng.integers(0, 256, (n, 4)).astype(np.uint8)
    d_in = torch.from_numpy(data).cuda()
    d_sym = torch.empty((n, 64, 2), dtype=torch.float32, device="cuda")
    before = c4.launch_count()
    c4.encode_dev(d_in.data_ptr(), n, d_sym.data_ptr())
    c4.sync()
    assert c4.launch_count() == before + 1
    host = c4.encode(data)
    assert np.array_equal(d_sym.cpu().numpy().view(np.complex64).reshape(n, 64), host)
    d_sym[:, :, 0] += 0.8 * torch.randn((n, 64), device="cuda")
    d_b = torch.empty((n, 4), dtype=torch.uint8, device="cuda")
    d_s = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_i = torch.empty(n, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    c4.decode_dev(d_sym.data_ptr(), n * 64, n, d_b.data_ptr(), d_s.data_ptr(), d_i.data_ptr())
    c4.sync()
    hb, hs, hi = c4.decode(d_sym.cpu().numpy().view(np.complex64).reshape(n, 64))
    assert np.array_equal(d_b.cpu().numpy(), hb)
    assert np.array_equal(d_s.cpu().numpy(), hs)
    assert np.array_equal(d_i.cpu().numpy(), hi)


def test_device_info_is_b200():
    assert L.device_count() >= 1
    info = L.device_info(0)
    assert info["sm"][0] == 10, info


# ---------------------------------------------------------------------------------------
# larger fixed-seed sets and full-size properties
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,ebn0,iters,early", [(100_000, 2.0, 5, 1), (100_000, 4.0, 5, 1), (20_000, 2.0, 50, 0)])
def test_decoder_large_fixed_seed_set_vs_oracle(c4, shipped, n, ebn0, iters, early):
    """fp32 kernels against the fp64 oracle on a large seeded set: every hard decision, iteration
    count and syndrome weight identical (the bar BASELINE.json sets for the fixed-seed test set)."""
    rng = np.random.default_rng(2024 + iters)
    data = rng.integers(0, 256, (n, 4)).astype(np.uint8)
    sym = c4.encode(data)
    sym.real += rng.standard_normal(sym.shape, dtype=np.float32) * np.float32(np.sqrt(10.0 ** (-ebn0 / 10.0)))
    b, sy, it = c4.decode(sym, method=1, max_iters=iters, early_stop=early)
    wb, wit, wsy, _ = O.decode_frames(sym, shipped["Hp"], method=1, iterations=iters, early_stop=bool(early),
                                      threads=os.cpu_count() or 4)
    bad = np.nonzero((b != wb).any(axis=1) | (it != wit) | (sy != wsy))[0]
    assert bad.size == 0, "%d of %d frames differ, first %s" % (bad.size, n, bad[:5])


def test_config3_full_size_round_trip(c4):
    """BASELINE config 3's size (10 M codewords), device-resident: encode -> decode returns the
    data for every frame in one iteration (noiseless), and with the bench's noise every frame
    runs exactly the fixed 50 iterations."""
    import torch
    n = 10_000_000
    g = torch.Generator(device="cuda")
    g.manual_seed(7)
    d = torch.randint(0, 256, (n, 4), dtype=torch.uint8, device="cuda", generator=g)
    sym = torch.empty((n, 64, 2), dtype=torch.float32, device="cuda")
    ob = torch.empty((n, 4), dtype=torch.uint8, device="cuda")
    os_ = torch.empty(n, dtype=torch.uint8, device="cuda")
    oi = torch.empty(n, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    c4.encode_dev(d.data_ptr(), n, sym.data_ptr())
    c4.decode_dev(sym.data_ptr(), n * 64, n, ob.data_ptr(), os_.data_ptr(), oi.data_ptr())
    c4.sync()
    assert bool((ob == d).all()) and int(os_.max()) == 0 and int(oi.min()) == 1 and int(oi.max()) == 1
    # linearity of the code: the XOR of two codewords is the codeword of the XORed data
    d2 = torch.randint(0, 256, (n, 4), dtype=torch.uint8, device="cuda", generator=g)
    sym2 = torch.empty_like(sym)
    c4.encode_dev(d2.data_ptr(), n, sym2.data_ptr())
    c4.sync()
    prod = sym[:, :, 0] * sym2[:, :, 0]                       # +1 where bits equal: BPSK of ~(a ^ b)
    sym2[:, :, 0] = -prod                                     # BPSK of a ^ b
    del prod
    torch.cuda.synchronize()                                  # torch's stream -> the handle's stream
    c4.decode_dev(sym2.data_ptr(), n * 64, n, ob.data_ptr(), os_.data_ptr(), oi.data_ptr())
    c4.sync()
    assert bool((ob == (d ^ d2)).all()) and int(os_.max()) == 0


def test_host_path_modes_agree_on_pinned_input():
    """ldpc535_decode_batch on a PINNED caller buffer: the copy engine reading it as it is (0), the host team
    packing the real parts first (1) and the measured choice between the two (2, the default) give identical
    bytes, syndrome weights and iteration counts -- several 128 MiB chunks, so the measured mode has estimates
    to act on -- and ldpc535_code_host_stats accounts for every chunk and every PCIe byte."""
    import ctypes as C
    from ldpc_ece535a import _abi
    code = L.Code(None, device=0)
    n = 600_000                                    # 307 MB of complex symbols: 3 chunks
    rng = np.random.default_rng(77)
    data = rng.integers(0, 256, (n, 4)).astype(np.uint8)
    sym = code.encode(data)
    sym.real += rng.standard_normal(sym.shape, dtype=np.float32) * np.float32(0.8)
    p = C.c_void_p()
    _abi.check(_abi.lib().ldpc535_host_alloc(sym.nbytes, C.byref(p)), "host_alloc")
    try:
        pinned = np.frombuffer((C.c_uint8 * sym.nbytes).from_address(p.value), dtype=np.complex64).reshape(sym.shape)
        pinned[...] = sym
        assert code.host_path()["pack_pinned"] == 2
        res = {}
        for mode in (0, 1, 2, 2):
            code.set_host_path(mode, 0)
            s0 = code.host_stats()
            res[mode] = code.decode(pinned, method=1, max_iters=5, early_stop=True)
            s1 = code.host_stats()
            packed, raw = s1["chunks_packed"] - s0["chunks_packed"], s1["chunks_raw"] - s0["chunks_raw"]
            assert packed + raw >= 3
            assert s1["h2d_bytes"] - s0["h2d_bytes"] in range(n * 64 * 4, n * 64 * 8 + 1)
            if mode == 0:
                assert packed == 0 and s1["h2d_bytes"] - s0["h2d_bytes"] == n * 64 * 8
            if mode == 1:
                assert raw == 0 and s1["h2d_bytes"] - s0["h2d_bytes"] == n * 64 * 4
        for mode in (1, 2):
            for x, y in zip(res[0], res[mode]):
                assert np.array_equal(x, y), mode
        want = code.decode(sym, method=1, max_iters=5, early_stop=True)       # pageable input: always packed
        for x, y in zip(res[0], want):
            assert np.array_equal(x, y)
    finally:
        _abi.lib().ldpc535_host_free(p)
        code.close()


@pytest.mark.parametrize("devices", [[0], [0, 0], [0, 0, 0]])
def test_pool_shards_and_gathers(c4, devices):
    """ldpc535_pool: contiguous shards over several handles (here on one GPU; on a multi-GPU box
    list every device), host-side gather; results equal the single-handle call."""
    if L.device_count() > 1 and len(devices) > 1:
        devices = [i % L.device_count() for i in range(len(devices))]
    pool = L.Pool(devices)
    assert pool.size == len(devices)
    rng = np.random.default_rng(21)
    for n in (0, 1, 2, 1001):
        data = rng.integers(0, 256, (n, 4)).astype(np.uint8)
        sym = pool.encode(data)
        assert np.array_equal(sym, c4.encode(data))
        noisy = sym.copy()
        noisy.real += rng.standard_normal(sym.shape, dtype=np.float32) * np.float32(0.7)
        for a, b in zip(pool.decode(noisy), c4.decode(noisy)):
            assert np.array_equal(a, b)
    stream = noisy.reshape(-1)
    offs = rng.integers(0, stream.size - 63, 777).astype(np.int64)
    pol = rng.choice(np.array([-1, 1], np.int8), 777)
    for a, b in zip(pool.decode(stream, win_offset=offs, polarity=pol), c4.decode(stream, win_offset=offs, polarity=pol)):
        assert np.array_equal(a, b)
    pool.close()


@pytest.mark.parametrize("n", [1, 3, 4, 5, 1000, 4099])
def test_hard_decision_kernel_vs_oracle(c4, shipped, n):
    """Method 3 on the default dispatch (the dedicated HBM-bound kernel): aligned frames, then
    arbitrary window offsets with both polarities."""
    assert c4.kernel_name(3) == "hard64"
    assert c4.kernel_name(2) == "hard64"          # no bit of the shipped code can flip (degree <= M/2)
    fb, fs, _ = c4.decode(util.synth_frames(shipped["Hp"], shipped["L"], shipped["U"], n, 1.0, seed=300 + n)[2], method=2)
    _, _, sym = util.synth_frames(shipped["Hp"], shipped["L"], shipped["U"], n, 1.0, seed=300 + n)
    wb, wit, wsy, _ = O.decode_frames(sym, shipped["Hp"], method=3, iterations=5, early_stop=True, threads=4)
    b, sy, it = c4.decode(sym, method=3)
    assert np.array_equal(b, wb) and np.array_equal(sy, wsy) and not it.any()
    w2, _, s2, _ = O.decode_frames(sym, shipped["Hp"], method=2, iterations=5, early_stop=True, threads=4)
    assert np.array_equal(fb, w2) and np.array_equal(fs, s2)
    stream = sym.reshape(-1)
    if stream.size > 64:
        rng = np.random.default_rng(n)
        offs = rng.integers(0, stream.size - 63, 500).astype(np.int64)
        pol = rng.choice(np.array([-1, 1], np.int8), 500)
        wins = np.stack([stream[o:o + 64] * q for o, q in zip(offs, pol)])
        wb, _, wsy, _ = O.decode_frames(wins, shipped["Hp"], method=3, iterations=5, early_stop=True, threads=4)
        b, sy, it = c4.decode(stream, method=3, win_offset=offs, polarity=pol)
        assert np.array_equal(b, wb) and np.array_equal(sy, wsy)


def _dense_code(M, N, col_deg, seed):
    """Random H = [A | B] with A invertible-by-column-pivoting and dense columns (exercises the
    <16, 8> kernel instantiations and bit-flip decisions that really flip)."""
    rng = np.random.default_rng(seed)
    for _ in range(200):
        H = np.zeros((M, N), np.int32)
        for c in range(N):
            rows = rng.choice(M, size=min(M, col_deg[c % len(col_deg)]), replace=False)
            H[rows, c] = 1
        if H.sum(1).max() > 16 or H.sum(1).min() < 2:
            continue
        try:
            L.Code(H, device=-1)
            return H
        except L.Ldpc535Error:
            continue
    raise RuntimeError("no invertible dense code found")


@pytest.mark.parametrize("M,N,col_deg", [(8, 16, (5, 3, 4, 6)), (16, 32, (3, 5, 2, 4, 7)), (24, 48, (3, 2, 4, 5)),
                                         (32, 64, (3, 2, 4, 8))])
def test_dense_small_codes_all_methods(M, N, col_deg):
    """Rate-1/2 codes (the reference's L/U layout needs N = 2M) denser than it ships: check degree up to 16, bit degree up to 8, a bit
    degree above M/2 so that decodeBitFlipping (:439-473) actually flips.  Every method vs the
    oracle, warp and CTA kernels."""
    H = _dense_code(M, N, col_deg, seed=M * 100 + N)
    Hp, Lm, Um, _ = O.reorder_h(H)
    code = L.Code(H, device=0)
    assert (code.M, code.N) == (M, N)
    _, cw, sym = util.synth_frames(Hp, Lm, Um, 400, 5.0, seed=M)
    enc = code.encode(util.pack_bits_msb(cw[:, M:]))
    assert np.array_equal(sym_to_bits(enc).reshape(400, N), cw)
    flips = 0
    for kern in ("warp", "block"):
        code.set_kernel(kern)
        for method in (1, 0, 2, 3):
            for iters in (5, 12):
                wb, wit, wsy, _ = O.decode_frames(sym, Hp, method=method, iterations=iters, early_stop=True,
                                                  threads=4)
                b, sy, it = code.decode(sym, method=method, max_iters=iters)
                assert np.array_equal(b, wb), (kern, method, iters)
                assert np.array_equal(sy, wsy), (kern, method, iters)
                if method in (0, 1):
                    assert np.array_equal(it, wit), (kern, method, iters)
        hb, _, _ = code.decode(sym, method=3)
        fb, _, _ = code.decode(sym, method=2)
        flips += int((hb != fb).any(axis=1).sum())
    if max(col_deg) > M // 2:
        assert flips > 0, "bit-flip never flipped: the test does not exercise :464-465"
    code.close()


@pytest.mark.parametrize("M,N,col_deg", [(24, 64, (3, 2, 4)), (16, 64, (2, 3)), (32, 64, (3, 2, 4))])
def test_hard_and_bitflip_with_more_than_32_data_bits(M, N, col_deg):
    """64-symbol codes with K = N - M > 32 data bits (5 or 6 bytes per frame): the hard-decision fast
    path keeps one 32-bit word per codeword and must not be dispatched here; methods 3
    and 2 against the oracle's decodeHard / decodeBitFlipping on the library's re-ordered H."""
    H = _dense_code(M, N, col_deg, seed=M * 1000 + N)
    code = L.Code(H, device=0)
    Hp = code.h_dense()
    assert code.nbytes == (N - M) // 8
    if N - M > 32:
        assert code.kernel_name(3) != "hard64" and code.kernel_name(2) != "hard64"
    rng = np.random.default_rng(M)
    n = 777
    sym = (rng.choice(np.array([-1.0, 1.0], np.float32), (n, N)) +
           rng.standard_normal((n, N)).astype(np.float32) * np.float32(0.6)).astype(np.complex64)
    for method in (3, 2):
        wb, _, wsy, _ = O.decode_frames(sym, Hp, method=method, iterations=5, early_stop=True, threads=4)
        b, sy, it = code.decode(sym, method=method)
        assert np.array_equal(b, wb), method
        assert np.array_equal(sy, wsy), method
    code.close()


def test_device_api_refuses_misaligned_pointers(c4):
    """ldpc535_*_batch_dev read symbols 16 bytes at a time and store packed bytes as 32-bit words: a
    pointer that is not aligned that way is refused with LDPC535_ERR_INVALID, not faulted on."""
    import torch
    n = 64
    sym = torch.zeros((n + 1) * 64 * 2, dtype=torch.float32, device="cuda")
    out = torch.zeros(n * 4 + 16, dtype=torch.uint8, device="cuda")
    for sym_off, out_off in ((8, 0), (4, 0), (0, 1), (0, 2)):
        with pytest.raises(L.Ldpc535Error) as e:
            c4.decode_dev(sym.data_ptr() + sym_off, n * 64, n, out.data_ptr() + out_off)
        assert e.value.status == L._abi.ERR_INVALID
    data = torch.zeros(n * 4 + 16, dtype=torch.uint8, device="cuda")
    with pytest.raises(L.Ldpc535Error):
        c4.encode_dev(data.data_ptr() + 1, n, sym.data_ptr())
    with pytest.raises(L.Ldpc535Error):
        c4.encode_dev(data.data_ptr(), n, sym.data_ptr() + 8)
    c4.decode_dev(sym.data_ptr(), n * 64, n, out.data_ptr())          # aligned: accepted
    c4.sync()


def _philox4x32(ctr, key):
    """numpy Philox4x32-10 (Salmon et al.): ctr (n, 4) uint32, key (2,) -> (n, 4) uint32."""
    c = ctr.astype(np.uint64)
    k0, k1 = np.uint64(key[0]), np.uint64(key[1])
    M0, M1, MASK = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c[:, 0], M1 * c[:, 2]
        c = np.stack([(p1 >> np.uint64(32)) ^ c[:, 1] ^ k0, p1 & MASK,
                      (p0 >> np.uint64(32)) ^ c[:, 3] ^ k1, p0 & MASK], axis=1)
        k0 = (k0 + np.uint64(0x9E3779B9)) & MASK
        k1 = (k1 + np.uint64(0xBB67AE85)) & MASK
    return c.astype(np.uint32)


def test_counter_based_synthetic_inputs(c4):
    """bench.py's inputs: Philox4x32-10, seed 535, counter = GLOBAL frame index.  The data bytes equal
    a numpy restatement of the generator, a shard generated at first_frame = a equals rows a.. of a
    bigger run (what makes the shards of an N-GPU job the same frames at any N), and the channel
    noise has the requested sigma."""
    import torch
    n, first = 5000, 123_456_789_012
    d = torch.empty((n, 4), dtype=torch.uint8, device="cuda")
    c4.synth_bytes_dev(535, first, n, d.data_ptr())
    c4.sync()
    f = np.arange(first, first + n, dtype=np.uint64)
    ctr = np.stack([f & np.uint64(0xFFFFFFFF), f >> np.uint64(32), np.zeros(n, np.uint64), np.zeros(n, np.uint64)], axis=1)
    want = _philox4x32(ctr, (535, 0))[:, 0].astype("<u4").view(np.uint8).reshape(n, 4)
    assert np.array_equal(d.cpu().numpy(), want)
    part = torch.empty((1000, 4), dtype=torch.uint8, device="cuda")
    c4.synth_bytes_dev(535, first + 3000, 1000, part.data_ptr())
    c4.sync()
    assert torch.equal(part, d[3000:4000])
    sym = torch.zeros((n, 64, 2), dtype=torch.float32, device="cuda")
    sub = torch.zeros((1000, 64, 2), dtype=torch.float32, device="cuda")
    c4.synth_awgn_dev(535, first, n, 0.5, sym.data_ptr())
    c4.synth_awgn_dev(535, first + 3000, 1000, 0.5, sub.data_ptr())
    c4.sync()
    assert torch.equal(sub, sym[3000:4000])
    re = sym[:, :, 0].double()
    assert float(sym[:, :, 1].abs().max()) == 0.0
    assert abs(float(re.mean())) < 0.005 and abs(float(re.std()) - 0.5) < 0.005
    assert abs(float((re ** 4).mean()) / 0.5 ** 4 - 3.0) < 0.1          # Gaussian kurtosis
    # first normal of frame `first`: Box-Muller on the first two words of Philox(ctr = (f, q = 0, 1))
    x = _philox4x32(np.array([[first & 0xFFFFFFFF, first >> 32, 0, 1]], np.uint64), (535, 0))[0].astype(np.float64)
    u1, u2 = (x[0] + 0.5) / 2 ** 32, (x[1] + 0.5) / 2 ** 32
    assert abs(float(re[0, 0]) - 0.5 * np.sqrt(-2 * np.log(u1)) * np.cos(2 * np.pi * u2)) < 1e-4
    # the C8k configuration uses 512-byte frames: 32 counter blocks per frame
    mufu = c4.probe_pipe_peak(L._abi.PIPE_MUFU)
    assert 2e12 < mufu < 6e12, mufu                                       # ~16 lanes/clk/SM x 148 SMs x ~1.9 GHz


def test_c4_refill_kernel_matches_warp_kernel_and_oracle(c4, shipped):
    """Early-stop sum-product on the shipped code with the opt-in persistent-slot kernel
    (set_kernel("c4-refill"): decode_c4_refill_kernel -- per-codeword stop, slots refilled from an
    atomic cursor, next window prefetched by cp.async), on batches large enough for it: bytes, syndrome weights and iteration counts equal the
    warp-per-codeword kernel's on every frame at 2 / 6 / 10 dB for 5 and 50 iterations max, for ragged
    batch sizes, with window offsets and polarities; a 30 000-frame slice equals the oracle."""
    sms = L.device_info(0)["sm_count"]
    big = sms * 256 * 2
    n = big + 70_001
    _, cw, _ = util.synth_frames(shipped["Hp"], shipped["L"], shipped["U"], 4096, None, seed=91)
    rng = np.random.default_rng(92)
    clean = util.bpsk(cw)[rng.integers(0, 4096, n)]
    for ebn0, iters in ((2.0, 5), (6.0, 5), (10.0, 5), (2.0, 50), (6.0, 3)):
        sigma = np.float32(np.sqrt(10.0 ** (-ebn0 / 10.0)))
        noisy = clean.copy()
        noisy.real += rng.standard_normal(clean.shape, dtype=np.float32) * sigma
        noisy.imag = rng.standard_normal(clean.shape, dtype=np.float32)          # ignored by the decoder
        for cnt in (n, big, big + 1):
            c4.set_kernel("warp")
            want = c4.decode(noisy[:cnt], method=1, max_iters=iters, early_stop=True)
            c4.set_kernel("c4-refill")
            got = c4.decode(noisy[:cnt], method=1, max_iters=iters, early_stop=True)
            for x, y, name in zip(got, want, ("bytes", "synd", "iters")):
                assert np.array_equal(x, y), (ebn0, iters, cnt, name)
        m = 30_000 if iters <= 5 else 2_000
        assert c4.kernel_for(1, True, n) == "c4-refill" and c4.kernel_for(1, True, 1000) == "warp"
        assert c4.kernel_for(1, False, n) == "c4-thread"
        wb, wit, wsy = util.oracle_spa_batch(noisy[:m], shipped["Hp"], iters, True)
        assert np.array_equal(got[0][:m], wb) and np.array_equal(got[2][:m], wit) and np.array_equal(got[1][:m], wsy)
    # windows at arbitrary symbol offsets, both polarities
    flat = noisy.reshape(-1)
    offs = rng.integers(0, flat.size - 64, big + 333).astype(np.int64)
    pol = rng.choice(np.array([-1, 1], np.int8), offs.size)
    c4.set_kernel("warp")
    want = c4.decode(flat, method=1, max_iters=5, early_stop=True, win_offset=offs, polarity=pol)
    c4.set_kernel("c4-refill")
    got = c4.decode(flat, method=1, max_iters=5, early_stop=True, win_offset=offs, polarity=pol)
    c4.set_kernel(None)
    assert c4.kernel_for(1, True, n) == "warp"          # the default dispatch for early stop
    for x, y in zip(got, want):
        assert np.array_equal(x, y)


def test_c4_adaptive_family_equals_warp_kernel(c4, shipped):
    """Early stop on a batch that fills the GPU several times over ("c4-adaptive", the default there):
    a 16 384-window sample goes through the warp kernel, pick_family_kernel chooses on the device,
    and the rest runs on the lock-step thread kernel (noisy channel) or the warp kernel (clean
    channel).  Bytes, syndrome weights and iteration counts equal the warp kernel's on every frame
    whichever family was picked, for ragged sizes, with window offsets and polarities."""
    sms = L.device_info(0)["sm_count"]
    n = sms * 256 * 4 + 16384 + 4099
    assert c4.kernel_for(1, True, n) == "c4-adaptive" and c4.kernel_for(1, True, n - 4100) == "warp"
    assert c4.kernel_for(1, False, n) == "c4-thread"
    _, cw, _ = util.synth_frames(shipped["Hp"], shipped["L"], shipped["U"], 4096, None, seed=93)
    rng = np.random.default_rng(94)
    clean = util.bpsk(cw)[rng.integers(0, 4096, n)]
    for ebn0, iters in ((1.0, 5), (9.0, 5), (4.0, 7), (2.0, 20)):
        sigma = np.float32(np.sqrt(10.0 ** (-ebn0 / 10.0)))
        noisy = clean.copy()
        noisy.real += rng.standard_normal(clean.shape, dtype=np.float32) * sigma
        c4.set_kernel("warp")
        want = c4.decode(noisy, method=1, max_iters=iters, early_stop=True)
        c4.set_kernel(None)
        before = c4.launch_count()
        got = c4.decode(noisy, method=1, max_iters=iters, early_stop=True)
        assert c4.launch_count() == before + 4          # sample, pick, and both families queued for the rest
        for x, y, name in zip(got, want, ("bytes", "synd", "iters")):
            assert np.array_equal(x, y), (ebn0, iters, name)
    flat = noisy.reshape(-1)
    offs = rng.integers(0, flat.size - 64, n).astype(np.int64)
    pol = rng.choice(np.array([-1, 1], np.int8), offs.size)
    c4.set_kernel("warp")
    want = c4.decode(flat, method=1, max_iters=5, early_stop=True, win_offset=offs[:-1], polarity=pol[:-1])
    c4.set_kernel(None)
    got = c4.decode(flat, method=1, max_iters=5, early_stop=True, win_offset=offs[:-1], polarity=pol[:-1])
    for x, y in zip(got, want):
        assert np.array_equal(x, y)
