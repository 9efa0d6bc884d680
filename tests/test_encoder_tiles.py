"""CPU: the look-up encoder's host-side tile choice (gr-ldpc_ece535a_b200/csrc/encode_m4r_tiles.h): every
frame covered, no empty tile, the cap respected, never worse than full tiles in the host's cost model, whole
waves at the batch sizes DESIGN.md quotes.  Native harness: tests/native/encoder_tiles_harness.cpp."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "gr-ldpc_ece535a_b200", "csrc")


def test_tile_choice(tmp_path):
    exe = str(tmp_path / "encoder_tiles_harness")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", CSRC,
                           os.path.join(ROOT, "tests", "native", "encoder_tiles_harness.cpp"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.startswith("ok "), out.stdout
