#!/usr/bin/env python
"""CSR SpMV on a device-resident built matrix (the reference's spmat_dot_densevec, accel.rs:338-370): U entries of a row in flight per thread
(QR_SPMV_UNROLL = 1: the round-1 loop).  GPU box only.   python tools/spmv_bench.py C2 [--rows LOG2]"""
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
ap = argparse.ArgumentParser(); ap.add_argument("workload"); ap.add_argument("--rows", type=int, default=None); ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()
labels, coeffs = get_workload(a.workload)
op = Q.SparsePauliOp([Q.Pauli(l) for l in labels], coeffs)
plan = op.plan(); dim, G = plan.dim, plan.n_groups
rows = dim if a.rows is None else min(dim, 1 << a.rows)
lo = 0
ip, ix, dt = DeviceBuffer((rows + 1) * 8), DeviceBuffer(rows * G * 8), DeviceBuffer(rows * G * 16)
call("qr_build_rows_device", plan.handle, lo, lo + rows, ip.ptr, ix.ptr, dt.ptr, 0, None)
dv, dy = DeviceBuffer(dim * 16), DeviceBuffer(rows * 16)
for c0 in range(0, dim, 1 << 22):
    v = H.lanczos_start_vector(c0, min(dim, c0 + (1 << 22)))
    call("qr_memcpy_h2d", dv.ptr + c0 * 16, v.ctypes.data, v.nbytes, None)
st = C.c_void_p(); call("qr_stream_create", C.byref(st))
e0, e1 = C.c_void_p(), C.c_void_p(); call("qr_event_create", C.byref(e0)); call("qr_event_create", C.byref(e1))
ref = None
for mode in ("1", "2", "4", "8"):
    os.environ["QR_SPMV_UNROLL"] = mode
    for _ in range(3): call("qr_spmv_device", rows, ip.ptr, ix.ptr, dt.ptr, dv.ptr, dy.ptr, st)
    call("qr_event_record", e0, st)
    for _ in range(a.reps): call("qr_spmv_device", rows, ip.ptr, ix.ptr, dt.ptr, dv.ptr, dy.ptr, st)
    call("qr_event_record", e1, st)
    ms = C.c_float(); call("qr_event_elapsed_ms", e0, e1, C.byref(ms)); t = ms.value / a.reps
    y = np.empty(min(rows, 1 << 18), dtype=np.complex128); call("qr_memcpy_d2h", y.ctypes.data, dy.ptr, y.nbytes, None)
    if ref is None: ref = y.copy()
    nnz = rows * G
    print(json.dumps({"workload": a.workload, "kernel": "spmv_csr_kernel" if mode == "1" else "spmv_csr_unrolled_kernel<%s>" % mode, "n": plan.n_qubits, "G": G, "rows": rows,
                      "ms": round(t, 4), "GBps_matrix": round(nnz * 24 / t / 1e6, 1), "GBps_40B": round(nnz * 40 / t / 1e6, 1), "Gnnz_s": round(nnz / t / 1e6, 2),
                      "same_bits": bool(np.array_equal(y.view(np.uint64), ref.view(np.uint64)))}), flush=True)
