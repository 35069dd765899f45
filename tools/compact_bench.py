#!/usr/bin/env python
"""eliminate_zeros on a device-RESIDENT built shard (count_kept + scan + compact_rows, and the generic indptr-driven pair)
against the fused drop-zeros build of the same rows.  GPU box only.   python tools/compact_bench.py C2 [--rows LOG2] [--tol 1e-7]"""
import argparse, ctypes as C, json, os, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tools"))
import qrusty_b200 as Q
from qrusty_b200 import _ffi, hamiltonians as H
from qrusty_b200._ffi import call
from qrusty_b200._runtime import DeviceBuffer
from fill_sweep import get_workload
ap = argparse.ArgumentParser(); ap.add_argument("workload"); ap.add_argument("--rows", type=int, default=None); ap.add_argument("--tol", type=float, default=1e-7)
ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()
labels, coeffs = get_workload(a.workload)
plan = Q.SparsePauliOp([Q.Pauli(l) for l in labels], coeffs).plan(); dim, G = plan.dim, plan.n_groups
rows = dim if a.rows is None else min(dim, 1 << a.rows)
ip, ix, dt = DeviceBuffer((rows + 1) * 8), DeviceBuffer(rows * G * 8), DeviceBuffer(rows * G * 16)
call("qr_build_rows_device", plan.handle, 0, rows, ip.ptr, ix.ptr, dt.ptr, 0, None)
st = C.c_void_p(); call("qr_stream_create", C.byref(st))
e0, e1 = C.c_void_p(), C.c_void_p(); call("qr_event_create", C.byref(e0)); call("qr_event_create", C.byref(e1))
oip = DeviceBuffer((rows + 1) * 8); kept = C.c_uint64()
call("qr_count_kept_device", rows, G, dt.ptr, a.tol, oip.ptr, C.byref(kept), st)
oix, odt = DeviceBuffer(max(kept.value * 8, 16)), DeviceBuffer(max(kept.value * 16, 16))
def timed(fn):
    for _ in range(3): fn()
    call("qr_event_record", e0, st)
    for _ in range(a.reps): fn()
    call("qr_event_record", e1, st); call("qr_stream_synchronize", st)
    ms = C.c_float(); call("qr_event_elapsed_ms", e0, e1, C.byref(ms)); return ms.value / a.reps
def resident():
    k = C.c_uint64()
    call("qr_count_kept_device", rows, G, dt.ptr, a.tol, oip.ptr, C.byref(k), st)
    call("qr_compact_rows_device", rows, G, ix.ptr, dt.ptr, a.tol, oip.ptr, oix.ptr, odt.ptr, st)
def generic():
    k = C.c_uint64()
    call("qr_csr_count_kept_device", rows, ip.ptr, dt.ptr, a.tol, oip.ptr, C.byref(k), st)
    call("qr_csr_compact_device", rows, ip.ptr, ix.ptr, dt.ptr, a.tol, oip.ptr, oix.ptr, odt.ptr, st)
def fused():
    k = C.c_uint64()
    call("qr_build_compact_count", plan.handle, 0, rows, a.tol, oip.ptr, C.byref(k), st)
    call("qr_build_compact_fill", plan.handle, 0, rows, a.tol, oip.ptr, oix.ptr, odt.ptr, st)
nnz = rows * G
for name, fn in (("resident (count_kept + scan + compact_rows)", resident), ("generic (csr_count_kept + scan + csr_compact_rows)", generic), ("fused build", fused)):
    t = timed(fn)
    moved = nnz * 16 + nnz * 24 + kept.value * 24 if name != "fused build" else kept.value * 24
    print(json.dumps({"workload": a.workload, "path": name, "G": G, "rows": rows, "stored": nnz, "kept": kept.value, "ms": round(t, 4),
                      "GBps_moved": round(moved / t / 1e6, 1)}), flush=True)
