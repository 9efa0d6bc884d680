"""CPU: pin the oracle (oracle/ldpc_oracle.c) to the REFERENCE'S OWN CODE.

oracle/_ref/libldpc_ref.so is lib/ldpc_decoder_cb_impl.cc + lib/ldpc_encoder_bc_impl.cc of the
reference, compiled unmodified from /root/reference against the dependency stand-ins in
oracle/refshim/ (oracle/ref_driver.cc, oracle/Makefile).  Two layers:

* tests on tests/golden/ref_build_golden.npz -- outputs of that library on stored seeded inputs
  (tools/gen_ref_golden.py) -- run everywhere, including where /root/reference does not exist;
* `live` tests call the library itself on wider sweeps and are skipped where it is not built.
"""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
from oracle import ref as R
import util

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CODES = ["hData1", "hData2", "hData3", "hData4", "hData5"]
live = pytest.mark.skipif(not R.available(), reason="oracle/_ref not built (no /root/reference)")


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(GOLD, "ref_build_golden.npz"))


def run_chunks(blk, stream, chunks, nout):
    pos, k, idle = 0, 0, 0
    out, cons, prod = [], [], []
    while pos < stream.size and idle < len(chunks):
        n, room = int(chunks[k % len(chunks)]), int(nout[k % len(nout)])
        k += 1
        o, c = blk.work(stream[pos:pos + n], room)
        out.append(np.array(o, np.uint8))
        cons.append(c)
        prod.append(len(o))
        pos += c
        idle = idle + 1 if (c == 0 and len(o) == 0) else 0
    return np.concatenate(out), np.array(cons), np.array(prod)


# ---------------------------------------------------------------------------------------------
# oracle vs outputs of the reference's code (committed goldens)
# ---------------------------------------------------------------------------------------------

@pytest.mark.parametrize("name", CODES)
def test_reorder_matches_reference_build(ref_codes, g, name):
    Hp, L, U, _ = O.reorder_h(ref_codes[name]["H"])
    assert np.array_equal(Hp, g["tab_%s_Hp" % name])
    assert np.array_equal(L, g["tab_%s_L" % name])
    assert np.array_equal(U, g["tab_%s_U" % name])


def test_shipped_tables_match_reference_constructors(shipped, g):
    assert np.array_equal(shipped["Hp"], g["shipped_Hp"])
    assert np.array_equal(shipped["L"], g["shipped_L"])
    assert np.array_equal(shipped["U"], g["shipped_U"])


@pytest.mark.parametrize("name", CODES)
def test_parity_matches_reference_build(ref_codes, g, name):
    """makeParityCheck through dgesv (the reference) == the oracle's real-valued and GF(2) solves."""
    Hp, L, U, _ = O.reorder_h(ref_codes[name]["H"])
    for d, c in zip(g["par_%s_d" % name], g["par_%s_c" % name]):
        for gf2 in (False, True):
            got, bad = O.make_parity_check(d.astype(np.int32), Hp, L, U, gf2=gf2)
            assert not bad and np.array_equal(got, c)


def test_encoder_block_matches_reference_build(shipped, g):
    out, consumed = O.encoder_work(shipped["Hp"], shipped["L"], shipped["U"], g["enc_bytes"],
                                   64 * 512)
    assert consumed == g["enc_bytes"].size
    assert np.array_equal(out.real, g["enc_symbols_re"].astype(np.float32)) and not out.imag.any()
    s2, c2 = O.encoder_work(shipped["Hp"], shipped["L"], shipped["U"], g["enc_bytes"][:11],
                            64 * 2 + 32)
    assert [s2.size, c2] == list(g["enc_ragged"])


@pytest.mark.parametrize("tag", ["clean", "0dB", "2dB", "4dB", "6dB"])
@pytest.mark.parametrize("iters", [5, 50])
def test_codeword_decoders_match_reference_build(shipped, g, tag, iters):
    Hp = shipped["Hp"]
    rx = g["cw_%s_rx" % tag].astype(np.float64)
    for method in (0, 1, 2, 3):
        key = "cw_%s_m%d_it%d_vhat" % (tag, method, iters)
        if key not in g:
            continue
        want, wsyn = g[key], g["cw_%s_m%d_it%d_synd" % (tag, method, iters)]
        for r, w, ws in zip(rx, want, wsyn):
            if method == 1:
                v = O.decode_spa(r, Hp, iters)[0]
            elif method == 0:
                v = O.decode_minsum(r, Hp, iters)[0]
            elif method == 2:
                v = O.decode_bitflip(r, Hp, iters)
            else:
                v = O.decode_hard(r)
            assert np.array_equal(np.asarray(v).reshape(-1), w), (method, tag, iters)
            assert O.check_frame(np.asarray(v, np.int32).reshape(-1), Hp, 4) == ws


@pytest.mark.parametrize("name", ["hData3", "hData5", "hData2"])
def test_codeword_decoders_other_codes_match_reference_build(ref_codes, g, name):
    Hp, _, _, _ = O.reorder_h(ref_codes[name]["H"])
    rx = g["cwx_%s_rx" % name].astype(np.float64)
    for method in (0, 1, 2, 3):
        for iters in (5, 20):
            want = g["cwx_%s_m%d_it%d_vhat" % (name, method, iters)]
            for r, w in zip(rx, want):
                if method == 1:
                    v = O.decode_spa(r, Hp, iters)[0]
                elif method == 0:
                    v = O.decode_minsum(r, Hp, iters)[0]
                elif method == 2:
                    v = O.decode_bitflip(r, Hp, iters)
                else:
                    v = O.decode_hard(r)
                assert np.array_equal(np.asarray(v).reshape(-1), w), (name, method, iters)


@pytest.mark.parametrize("tag,method", [("blk_m0_4dB", 0), ("blk_m1_4dB", 1), ("blk_m1_1dB", 1),
                                        ("blk_m2_7dB", 2), ("blk_m3_7dB", 3)])
def test_decoder_block_matches_reference_build(shipped, g, tag, method):
    """general_work incl. the sync machine: bytes, per-call consumed/produced, printed sync
    lines and final (d_state, d_errors)."""
    stream = g[tag + "_stream_re"].astype(np.complex64)
    blk = O.DecoderBlock(shipped["Hp"], method)
    out, cons, prod = run_chunks(blk, stream, g["chunks"], g["nout"])
    assert np.array_equal(cons, g[tag + "_consumed"])
    assert np.array_equal(prod, g[tag + "_produced"])
    assert np.array_equal(out, g[tag + "_bytes"])
    assert blk.events == list(g[tag + "_events"])
    assert [blk.st.state, blk.st.errors] == list(g[tag + "_final"])


def test_forecasts_match_reference_build(g):
    # decoder asks noutput*N (lib/ldpc_decoder_cb_impl.cc:126-130), encoder ceil(nout/16) (:111-116)
    assert list(g["dec_forecast"]) == [n * 64 for n in (1, 4, 64, 4096)]
    assert list(g["enc_forecast"]) == [int(np.ceil(n / 16.0)) for n in (1, 16, 17, 64, 640, 4096)]


# ---------------------------------------------------------------------------------------------
# BASELINE config 4's (3,6)-regular n = 8192 code: the reference's reorderHMatrix and decode members on
# 4096 x 8192 dense matrices (tools/gen_ref_golden_c8k.py; ~4 s per iteration and codeword there)
# ---------------------------------------------------------------------------------------------

def sha_bits(a):
    return hashlib.sha256(np.packbits(np.asarray(a, np.uint8), axis=1).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def g8k():
    return np.load(os.path.join(GOLD, "ref_build_c8k.npz"))


@pytest.fixture(scope="module")
def c8k_host(g8k):
    """The product's host tables for the same seeded code (LDPC535_DEVICE_NONE: no compute)."""
    import ldpc_ece535a as L
    rp, ci, M, N = L.codes.regular_code(8192, 3, 6, int(g8k["seed"]))
    code = L.Code((rp, ci, M, N), device=-1)
    yield code, L.codes.to_dense(rp, ci, M, N)
    code.close()


def test_c8k_reorder_matches_reference_build(g8k, c8k_host):
    """Column re-ordering and LU factors of the 4096 x 8192 matrix: the library's bit-packed GF(2)
    elimination and the oracle's dense one against the reference's reorderHMatrix."""
    code, H = c8k_host
    assert sha_bits(code.h_dense()) == str(g8k["Hp_sha256"])
    assert np.array_equal(code.h_dense().sum(0)[:64], g8k["Hp_col_weight_head"])
    Hp, Lm, Um, piv = O.reorder_h(H)
    assert sha_bits(Hp) == str(g8k["Hp_sha256"])
    assert sha_bits(Lm) == str(g8k["L_sha256"]) and int(Lm.sum()) == int(g8k["L_nnz"])
    assert sha_bits(Um) == str(g8k["U_sha256"]) and int(Um.sum()) == int(g8k["U_nnz"])
    assert np.array_equal(code.pivots(), piv)


@pytest.mark.parametrize("name,method,iters", [("spa3", 1, 3), ("spa12", 1, 12), ("minsum8", 0, 8)])
def test_c8k_sparse_oracle_matches_reference_build(g8k, c8k_host, name, method, iters):
    """Decisions of the reference's dense decodeSumProductSoft / decodeLogDomainSimple on the n = 8192
    code == the sparse restatement the GPU tests use at this size."""
    code, _ = c8k_host
    if method != 1:
        pytest.skip("the sparse restatement covers sum-product; min-sum at this size is checked on the GPU")
    row_ptr, col_idx = code.h_csr()
    rows = np.repeat(np.arange(code.M), np.diff(row_ptr))
    order = np.lexsort((rows, col_idx))
    col_ptr = np.zeros(code.N + 1, np.int32)
    np.add.at(col_ptr, col_idx + 1, 1)
    tables = (row_ptr, col_idx, np.cumsum(col_ptr).astype(np.int32), order.astype(np.int32))
    for f in range(g8k["rx"].shape[0]):
        vhat, run = O.decode_spa_sparse(g8k["rx"][f].astype(np.float64), tables, code.M, code.N, iters, True)
        assert np.array_equal(np.packbits(vhat.astype(np.uint8)), g8k[name + "_vhat"][f]), (name, f)


def test_c8k_sparse_oracle_matches_reference_build_50_iterations(c8k_host):
    """36 frames at 1.0 / 1.5 / 2.0 dB through the reference's decodeSumProductSoft to its exit or
    to 50 iterations (tools/gen_ref_golden_c8k_50it.py; 12 frames never converge, the others take
    14-45 iterations): the sparse restatement gives the same 8192 decisions on every one, the
    same saturating syndrome weight, and stops at the stored iteration counts."""
    g = np.load(os.path.join(GOLD, "ref_build_c8k_50it.npz"))
    code, _ = c8k_host
    row_ptr, col_idx = code.h_csr()
    tables = util.sparse_tables(row_ptr, col_idx, code.M, code.N)
    vhat, iters = util.oracle_spa_sparse_batch(g["rx"], tables, code.M, code.N, 50, True)
    assert np.array_equal(np.packbits(vhat, axis=1), g["spa50_vhat"])
    assert np.array_equal(iters, g["iters_oracle"])
    synd = util.syndrome_weights(vhat, row_ptr, col_idx, code.M)
    # checkFrame(v, M/8) stops counting at M/8 + 1 (lib/ldpc_decoder_cb_impl.cc:236-253)
    assert np.array_equal(np.minimum(synd, code.M // 8 + 1), g["spa50_synd"])
    assert (iters[:12] == 50).all() and (synd[:12] > 0).all()      # the 1.0 dB frames do not converge
    assert (synd[iters < 50] == 0).all()


# ---------------------------------------------------------------------------------------------
# live: the library itself
# ---------------------------------------------------------------------------------------------

@live
def test_goldens_are_reproducible(g, shipped):
    """The committed fixture is what the library gives now (guards a stale fixture)."""
    enc = R.RefEncoder()
    sym, _ = enc.work(g["enc_bytes"], 64 * 512)
    assert np.array_equal(sym.real, g["enc_symbols_re"].astype(np.float32))
    dec = R.RefDecoder(1)
    for r, w in zip(g["cw_2dB_rx"], g["cw_2dB_m1_it50_vhat"]):
        assert np.array_equal(dec.decode(r.astype(np.float64), 1, 50), w)


@live
def test_reference_qa_kats_on_reference_build(ref_codes):
    """python/qa_ldpc_encoder_bc.py:23-46 and python/qa_ldpc_decoder_cb.py:20-43 (8x16 code)
    pass on the reference build once the 8x16 matrix replaces the pasted-in literal."""
    with open(os.path.join(GOLD, "ref_qa_kat.json")) as f:
        k = json.load(f)
    frames = []
    for c, d in zip(k["mod_check"], k["mod_data"]):
        frames += c + d
    frames = np.array(frames, np.complex64)
    enc = R.RefEncoder()
    enc.set_code(ref_codes[k["code"]]["H"])
    out, consumed = enc.work(np.array(k["data_bytes"], np.uint8), 8 * 16)
    assert consumed == 8 and np.array_equal(out, frames)
    for method in (0, 1, 2, 3):
        dec = R.RefDecoder(method)
        dec.set_code(ref_codes[k["code"]]["H"])
        got, consumed = dec.work(frames, 8)
        assert list(got) == k["data_bytes"] and consumed == 8 * 16 and dec.events == [1]


@live
@pytest.mark.parametrize("method", [0, 1, 2, 3])
@pytest.mark.parametrize("ebn0", [None, 0.0, 2.0, 4.0, 6.0])
def test_live_block_streams(shipped, method, ebn0):
    """Fresh seeded streams (lead-in, inverted stretch, noise), oracle block vs reference block."""
    rng = np.random.default_rng(1000 + method * 10 + int(ebn0 or 9))
    # bit flipping that never locks slides one symbol at a time with two O(M N^2) decodes per
    # step (in the reference and in the oracle alike): keep those streams short
    nfr = 30 if (method == 2 and ebn0 is not None and ebn0 < 6) else 300
    data = rng.integers(0, 256, 4 * nfr).astype(np.uint8)
    sym, _ = O.encoder_work(shipped["Hp"], shipped["L"], shipped["U"], data, 64 * nfr)
    noisy = util.awgn(sym, ebn0, rng)
    lead = (rng.standard_normal(91) * 0.9).astype(np.float32).astype(np.complex64)
    a3, b3 = 64 * (nfr // 3), 64 * (2 * nfr // 3)
    stream = np.concatenate([lead, noisy[:a3], -noisy[a3:b3], noisy[b3:]])
    a, b = O.DecoderBlock(shipped["Hp"], method), R.RefDecoder(method)
    chunks, nout = (777, 64, 5000, 65, 130), (50, 4, 400, 7, 9)
    oa, ca, pa = run_chunks(a, stream, chunks, nout)
    ob, cb, pb = run_chunks(b, stream, chunks, nout)
    assert np.array_equal(ca, cb) and np.array_equal(pa, pb) and np.array_equal(oa, ob)
    assert a.events == b.events
    assert (a.st.state, a.st.errors) == b.state


@live
@pytest.mark.parametrize("name", CODES)
def test_live_codeword_level_all_codes(ref_codes, name):
    """Every reference matrix, every method, 1/5/30 iterations, 1..5 dB: identical decisions."""
    H = ref_codes[name]["H"]
    M, N = H.shape
    Hp, L, U, _ = O.reorder_h(H)
    dec, enc = R.RefDecoder(1), R.RefEncoder()
    rHp, rL, rU = enc.set_code(H)
    dec.set_code(H)
    assert np.array_equal(rHp, Hp) and np.array_equal(rL, L) and np.array_equal(rU, U)
    assert np.array_equal(dec.H, Hp)
    rng = np.random.default_rng(77)
    for ebn0 in (1.0, 3.0, 5.0):
        d = rng.integers(0, 2, (12, N - M)).astype(np.int32)
        cw = np.array([np.concatenate([enc.make_parity(x), x]) for x in d])
        assert np.array_equal(cw, util.oracle_encode_bits(d, Hp, L, U))
        rx = util.awgn(util.bpsk(cw), ebn0, rng).real.astype(np.float64)
        for r in rx:
            for iters in (1, 5, 30):
                assert np.array_equal(dec.decode(r, 1, iters), O.decode_spa(r, Hp, iters)[0])
                assert np.array_equal(dec.decode(r, 0, iters), O.decode_minsum(r, Hp, iters)[0])
                assert np.array_equal(dec.decode(r, 2, iters), O.decode_bitflip(r, Hp, iters))
            v = dec.decode(r, 3, 1)
            assert np.array_equal(v, O.decode_hard(r))
            for thr in (0, M // 8):
                assert dec.check_frame(v, thr) == O.check_frame(v, Hp, thr)
