#!/usr/bin/env python
"""qr_build_host on C2 into pinned arrays: the compact wire form against QR_HOST_WIDE, the raw D2H rate beside it, and a
host-only measurement of the column rebuild (numpy restatement is too slow to mean anything, so the C pool is timed by
running the compact form on a tiny PCIe payload: G = 21, data bytes unchanged).  Run per thread count:
  QR_HOST_COPY_THREADS=8 python tools/e2e_wire_probe.py"""
import ctypes as C, os, sys, time, json
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import qrusty_b200 as Q
from qrusty_b200 import _ffi, hamiltonians as H
from qrusty_b200._ffi import call
from qrusty_b200._runtime import DeviceBuffer, pinned_empty

def T(f, reps=7):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); f(); call("qr_stream_synchronize", None); ts.append(time.perf_counter() - t0)
    return round(min(ts) * 1e3, 3), round(float(np.median(ts)) * 1e3, 3)

labels, coeffs = H.xxz_chain(20, 1.0, 0.7)
op = Q.SparsePauliOp([Q.Pauli(l) for l in labels], coeffs)
plan = op.plan(); G, dim = plan.n_groups, plan.dim
ip = pinned_empty(dim + 1, np.uint64); ix = pinned_empty(dim * G, np.uint64); dt = pinned_empty(dim * G, np.complex128)
out = {"threads_env": os.environ.get("QR_HOST_COPY_THREADS"), "cpus": os.cpu_count(), "affinity": len(os.sched_getaffinity(0))}
d = DeviceBuffer(dim * G * 16)
out["d2h_352MB_pinned_ms"] = T(lambda: d.download(dt))
out["build_host_wide_ms"] = T(lambda: call("qr_build_host", plan.handle, 0, dim, ip.ctypes.data, ix.ctypes.data, dt.ctypes.data, _ffi.QR_HOST_WIDE))
out["build_host_compact_ms"] = T(lambda: call("qr_build_host", plan.handle, 0, dim, ip.ctypes.data, ix.ctypes.data, dt.ctypes.data, 0))
out["d2h_bytes_compact"] = _ffi.last_d2h_bytes()
for col in ("2", "4"):
    os.environ["QR_HOST_WIRE_COL"] = col
    out["build_host_compact_col%s_ms" % col] = T(lambda: call("qr_build_host", plan.handle, 0, dim, ip.ctypes.data, ix.ctypes.data, dt.ctypes.data, 0))
os.environ.pop("QR_HOST_WIRE_COL")
def full():
    return Q.SparsePauliOp.from_terms(20, op.terms()).to_matrix_mode("Cuda").export()
out["full_e2e_step_ms"] = T(full)
print(json.dumps(out), flush=True)
