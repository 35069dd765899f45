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


@pytest.fixture(scope="session")
def precond_vectors():
    """pyqrusty/tests/test_it.py:284-301: the reference's own precond fixture on H2 -- (dx, e,
    expected rv as printed there, 9 significant digits)."""
    import numpy as np
    dx = np.zeros(16, complex)
    dx[[6, 9, 10]] = [-9.57567359e-14, -9.57428581e-14, 1.74695127e-01]
    rv = np.zeros(16, complex)
    rv[[6, 9, 10]] = [-9.91433685e-14, -9.91289999e-14, 8.88073154e-02]
    return dx, -1.7037077186606393 + 0j, rv


def has_gpu():
    from qrusty_b200 import _ffi
    return _ffi.device_count() > 0
