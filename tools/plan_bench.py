#!/usr/bin/env python
"""Host-visible latency of the pieces around the fill for one operator: plan creation (H2D of the terms, K1, kernel choice),
the first build (lazy tables), a warm build, and the matrix-free diagonal.  GPU box only.   python tools/plan_bench.py H8 [--rows LOG2]"""
import argparse, ctypes as C, json, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tools"))
import qrusty_b200 as Q
from qrusty_b200._ffi import call
from qrusty_b200._runtime import DeviceBuffer
from fill_sweep import get_workload
ap = argparse.ArgumentParser(); ap.add_argument("workload"); ap.add_argument("--rows", type=int, default=16)
a = ap.parse_args()
labels, coeffs = get_workload(a.workload)
op = Q.SparsePauliOp([Q.Pauli(l) for l in labels], coeffs)
terms = op.terms(); n = len(labels[0])
def wall(f, reps=5):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); r = f(); call("qr_stream_synchronize", None); ts.append((time.perf_counter() - t0) * 1e3)
    return round(min(ts), 3), round(float(np.median(ts)), 3), r
out = {"workload": a.workload, "n": n, "T": len(labels)}
plan = Q.SparsePauliOp.from_terms(n, terms).plan()               # context, module load
tc, td = [], []
for _ in range(5):                                               # the ABI calls themselves (QR_PLAN_TRACE=1 prints the milestones)
    h = C.c_void_p(); t0 = time.perf_counter()
    call("qr_plan_create", n, terms.ctypes.data, len(terms), 0, 0, C.byref(h)); t1 = time.perf_counter()
    call("qr_plan_destroy", h); t2 = time.perf_counter()
    tc.append((t1 - t0) * 1e3); td.append((t2 - t1) * 1e3)
out["plan_create_ms"], out["plan_destroy_ms"] = round(min(tc), 3), round(min(td), 3)
G, dim = plan.n_groups, plan.dim
out["G"] = G; out["fill_kernel"] = plan.fill_kernel
rows = min(dim, 1 << a.rows)
ip, ix, dt = DeviceBuffer((rows + 1) * 8), DeviceBuffer(rows * G * 8), DeviceBuffer(rows * G * 16)
fresh = Q.SparsePauliOp.from_terms(n, terms).plan()
t0 = time.perf_counter(); call("qr_build_rows_device", fresh.handle, 0, rows, ip.ptr, ix.ptr, dt.ptr, 0, None); call("qr_stream_synchronize", None)
out["first_build_ms"] = round((time.perf_counter() - t0) * 1e3, 3)
out["warm_build_ms"] = wall(lambda: call("qr_build_rows_device", fresh.handle, 0, rows, ip.ptr, ix.ptr, dt.ptr, 0, None))[0]
dd = DeviceBuffer(rows * 16)
out["diagonal_ms"] = wall(lambda: call("qr_diagonal_device", fresh.handle, 0, rows, dd.ptr, None))[0]
out["rows"] = rows
print(json.dumps(out), flush=True)
