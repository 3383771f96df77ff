"""pytest configuration: the `gpu` marker and shared loaders.

`-m "not gpu"` covers the oracle against the golden vectors, the host logic and
the C-ABI symbol table; `-m gpu` holds the parity tests proper (CUDA path vs
oracle / goldens through the C ABI).
"""
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def nb():
    """The product package (ppa-nbody-collisions_b200), built in tree."""
    import __graft_entry__ as G
    return G.load_package()


@pytest.fixture(scope="session")
def golden_dir():
    return ROOT / "tests" / "golden"
