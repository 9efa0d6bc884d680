import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gr-ldpc_ece535a_b200", "python")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def ref_codes():
    from oracle import oracle as O
    return O.load_ref_codes()


@pytest.fixture(scope="session")
def shipped(ref_codes):
    """Shipped 32x64 code, re-ordered by the oracle: dict(H, Hp, L, U, pivots)."""
    from oracle import oracle as O
    H = ref_codes["shipped"]["H"]
    Hp, L, U, ch = O.reorder_h(H)
    return {"H": H, "Hp": Hp, "L": L, "U": U, "pivots": ch}
