"""CPU: the shared-memory layouts the decoder kernels rely on (gr-ldpc_ece535a_b200/csrc/code_tables.cpp).
The warp-per-codeword kernel's strip (color_warp_layout): every check-phase and every variable-phase access
of 32 lanes touches 32 different banks, unused slots included (the free bank they were given), and real edges
map one-to-one onto words.  The n = 8192 register-table kernel's coloured rows (color_regular_rows): the same
for rows of 32 checks and groups of 32 bits of a (3,6)-regular code.  Checked on the shipped 32x64 code, on
random sparse codes with uneven degrees and on (3,6)-regular 256x512 codes by a native harness
(tests/native/warp_layout_harness.cpp) compiled here against the library's own source file."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "gr-ldpc_ece535a_b200", "csrc")


def test_strip_layouts_are_bank_conflict_free(tmp_path):
    exe = str(tmp_path / "warp_layout_harness")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", CSRC,
                           os.path.join(ROOT, "tests", "native", "warp_layout_harness.cpp"),
                           os.path.join(CSRC, "code_tables.cpp"), "-o", exe])
    out = subprocess.run([exe, "200"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.startswith("ok "), out.stdout
    assert int(out.stdout.split()[1]) >= 50          # the shipped code, most of the random ones, two regular ones
