"""CPU: the host packing team of ldpc535_decode_batch (gr-ldpc_ece535a_b200/csrc/host_pack.cpp) -- real
parts of interleaved complex symbols into pinned staging, blocks handed out from a shared counter,
workers that spin and then sleep between jobs.  The harness (tests/native/host_pack_harness.cpp) is
compiled here with g++ against the library's own source file."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "gr-ldpc_ece535a_b200", "csrc")


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("host_pack") / "host_pack_harness")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-I", CSRC,
                           os.path.join(ROOT, "tests", "native", "host_pack_harness.cpp"),
                           os.path.join(CSRC, "host_pack.cpp"), "-o", exe])
    return exe


@pytest.mark.parametrize("threads", [1, 2, 3, 8, 17])
def test_pack_team_copies_the_real_parts(harness, threads):
    out = subprocess.run([harness, str(threads)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.strip() == "ok %d" % threads
