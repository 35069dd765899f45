import gzip
import json
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _build_native_once():
    # in-tree builds; both are no-ops when up to date
    import importlib.util
    spec = importlib.util.spec_from_file_location("_qr_build", ROOT / "qrusty_b200" / "build.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build()
    from oracle import oracle
    oracle.build()


_build_native_once()


@pytest.fixture(scope="session")
def fixtures():
    """Reference Hamiltonian inputs (tests/golden/make_fixtures.py)."""
    with gzip.open(ROOT / "tests" / "golden" / "h_fixtures.json.gz") as f:
        raw = json.load(f)
    return {k: (v["labels"], [complex(a, b) for a, b in v["coeffs"]]) for k, v in raw.items()}


@pytest.fixture(scope="session")
def golden_sums():
    with open(ROOT / "tests" / "golden" / "csr_checksums.json") as f:
        return json.load(f)


def has_gpu():
    from qrusty_b200 import _ffi
    return _ffi.device_count() > 0
