"""CPU check of the rows kernel's index arithmetic (no GPU): the Python model in tools/rows_kernel_model.py walks the
batches thread by thread exactly as fill_rows_kernel does -- thread <-> group state, Gray-coded batches of 2^Q rows,
+-cnt slot steps, the extras table, two batch buffers -- and must reproduce the oracle's CSR bit for bit."""
import importlib.util
import sys
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle as O
from qrusty_b200 import hamiltonians as H

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def model():
    spec = importlib.util.spec_from_file_location("rows_kernel_model", ROOT / "tools" / "rows_kernel_model.py")
    mod = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["rows_kernel_model.py", "--no-run"]
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.argv = argv
    return mod


@pytest.mark.parametrize("NG,TH,Q,log2R,lo,hi", [(1, 64, 1, 5, 0, 1 << 10), (2, 32, 2, 6, 100, 1000), (3, 32, 1, 4, 0, 256), (1, 64, 3, 7, 3, 1021)])
def test_model_matches_oracle(model, NG, TH, Q, log2R, lo, hi):
    labels, coeffs = H.random_pauli_sum(10, 90, 60, 10, 7)
    n, params = O.make_params(labels, coeffs)
    ref = O.build_csr(params, n)
    hi = min(hi, 1 << n)
    s0, s1, ix, dt = model.emulate(params, n, lo, hi, NG, TH, Q, log2R)
    G = len(ix) // (hi - lo)
    assert G == 60 and s1 > s0
    a, b = (s0 - lo) * G, (s1 - lo) * G
    assert np.array_equal(ix[a:b], ref[1][s0 * G:s1 * G].astype(np.int64))
    assert np.array_equal(dt[a:b].view(np.uint64), ref[2][s0 * G:s1 * G].view(np.uint64))
