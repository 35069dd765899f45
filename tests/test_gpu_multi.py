"""Multi-process GPU test (needs >= 2 GPUs; skipped otherwise): one rank per GPU under torchrun,
row-sharded CSR build verified against the oracle, distributed matrix-free H.v through both the
NCCL all-gather path and the fused peer-memory path (qr_apply_p2p), which must agree bit for bit."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

from qrusty_b200 import _ffi

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cfg", ["xxz16", "C1", "H6", "H8", "xxz19"])
def test_two_ranks(cfg):
    if _ffi.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29577", str(ROOT / "tools" / "multi_gpu_config.py"), cfg]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert res.returncode == 0 and lines, res.stdout[-2000:] + res.stderr[-2000:]
    out = json.loads(lines[-1])
    assert out["n_gpus"] == 2 and out["rows_bad"] == 0
    assert out["hv_max_rel_err"] < 1e-12
    assert out["hv_p2p_equals_allgather"] is True


@pytest.mark.parametrize("cfg", ["xxz16", "H8"])
def test_two_ranks_lanczos(cfg):
    """The two-pass Lanczos loop (qr_apply_p2p_dot + device scalars, all-reduced in place) against the round-1 loop over the
    all-gather apply (host scalars) on two ranks: same alphas to rounding."""
    if _ffi.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29579", str(ROOT / "tools" / "lanczos_bench.py"), cfg, "20"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert res.returncode == 0 and lines, res.stdout[-2000:] + res.stderr[-2000:]
    out = json.loads(lines[-1])
    assert out["n_gpus"] == 2 and out["max_alpha_diff"] < 1e-9 * max(1.0, abs(out["beta0"]))
