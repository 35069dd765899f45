#!/usr/bin/env python
"""qr_build_host on C2 into pinned arrays for several row-window sizes (QR_HOST_WIN_MB) and host thread counts.  GPU box only."""
import ctypes as C, os, sys, time, json, subprocess
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
if len(sys.argv) == 1:
    for mb in ("64", "32", "16", "8", "128"):
        for th in ("", "8", "12"):
            env = dict(os.environ, QR_HOST_WIN_MB=mb)
            if th: env["QR_HOST_COPY_THREADS"] = th
            subprocess.run([sys.executable, __file__, "child"], env=env)
    sys.exit(0)
sys.path.insert(0, str(ROOT))
import numpy as np
import qrusty_b200 as Q
from qrusty_b200 import _ffi, hamiltonians as H
from qrusty_b200._ffi import call
from qrusty_b200._runtime import pinned_empty
labels, coeffs = H.xxz_chain(20, 1.0, 0.7)
op = Q.SparsePauliOp([Q.Pauli(l) for l in labels], coeffs)
plan = op.plan(); G, dim = plan.n_groups, plan.dim
ip = pinned_empty(dim + 1, np.uint64); ix = pinned_empty(dim * G, np.uint64); dt = pinned_empty(dim * G, np.complex128)
def T(f, reps=9):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0)
    return round(min(ts) * 1e3, 3), round(float(np.median(ts)) * 1e3, 3)
out = {"win_mb": os.environ.get("QR_HOST_WIN_MB"), "threads": os.environ.get("QR_HOST_COPY_THREADS") or "default"}
out["build_host_compact_ms"] = T(lambda: call("qr_build_host", plan.handle, 0, dim, ip.ctypes.data, ix.ctypes.data, dt.ctypes.data, 0))
out["full_e2e_step_ms"] = T(lambda: Q.SparsePauliOp.from_terms(20, op.terms()).to_matrix_mode("Cuda").export())
print(json.dumps(out), flush=True)
