"""CPU: image_sink (host file I/O, no GPU) against the restated reference behaviour
(lib/image_sink_impl.cc:46-84): what gets written, and when, for arbitrary chunking."""
import os

import numpy as np
import pytest

from oracle import oracle as O


def bmp(size, seed, dib=40):
    rng = np.random.default_rng(seed)
    b = rng.integers(0, 256, size).astype(np.uint8)
    b[0:2] = (0x42, 0x4D)
    b[2:6] = np.frombuffer(np.uint32(size).tobytes(), np.uint8)
    b[6:10] = 0
    b[14] = dib
    # no accidental second header inside the payload
    for i in range(2, size - 18):
        if b[i] == 0x42 and b[i + 1] == 0x4D:
            b[i + 1] = 0
    return b


@pytest.fixture()
def sink(tmp_path, monkeypatch):
    monkeypatch.setenv("LDPC535_IMAGE_PATH", str(tmp_path / "result.bmp"))
    monkeypatch.setenv("LDPC535_IMAGE_DISPLAY", "0")
    import ldpc_ece535a as L
    s = L.image_sink()
    assert s.name() == "image_sink" and s.item_sizes()[0] == 1
    return s, tmp_path / "result.bmp"


@pytest.mark.parametrize("chunks", [[100000], [4096] * 30, [17, 5000, 19, 1, 18, 30000, 70000], [19270, 19270, 19270, 50]])
def test_files_match_reference_behaviour(sink, chunks):
    s, path = sink
    stream = np.concatenate([np.arange(50, dtype=np.uint8), bmp(19270, 1), bmp(3000, 2, dib=124), bmp(19270, 3),
                             bmp(60, 4, dib=12)])
    orc = O.ImageSinkOracle()
    pos, written = 0, []
    for c in chunks:
        part = stream[pos:pos + c]
        if part.size == 0:
            break
        assert s.work(part) == part.size
        orc.work(part)
        pos += part.size
        if s.files_written() > len(written):
            written.append(path.read_bytes())
    assert s.files_written() == len(orc.files)
    assert written and written[-1] == orc.files[-1]
    # every file seen on disk is one the reference would have written, in order
    it = iter(orc.files)
    assert all(any(w == f for f in it) for w in written)


def test_header_split_across_calls_is_missed_like_upstream(sink):
    """The look-ahead stays inside one work() call: a header in the last 18 bytes is not seen."""
    s, path = sink
    a, b = bmp(500, 5), bmp(400, 6)
    stream = np.concatenate([a, b])
    cut = 500 + 10                       # second header straddles the call boundary
    orc = O.ImageSinkOracle()
    for part in (stream[:cut], stream[cut:], bmp(100, 7)):
        s.work(part)
        orc.work(part)
    assert s.files_written() == len(orc.files)
    if orc.files:
        assert path.read_bytes() == orc.files[-1]
