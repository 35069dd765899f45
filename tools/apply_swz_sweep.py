#!/usr/bin/env python
"""CTA order of the gather H.v kernel (QR_APPLY_SWZ = f: the top f row-block bits vary fastest in time).  GPU box only.
  python tools/apply_swz_sweep.py C4 [--fs "0 3 4 5 6"]"""
import argparse, ctypes as C, json, os, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tools"))
import qrusty_b200 as Q
from qrusty_b200 import hamiltonians as H
from qrusty_b200._ffi import call
from qrusty_b200._runtime import DeviceBuffer
from fill_sweep import get_workload

ap = argparse.ArgumentParser(); ap.add_argument("workload"); ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--fs", default="0 2 3 4 5 6 7 8"); ap.add_argument("--var", default="QR_APPLY_SWZ")
a = ap.parse_args()
labels, coeffs = get_workload(a.workload)
op = Q.SparsePauliOp([Q.Pauli(l) for l in labels], coeffs)
plan = op.plan(); dim, G = plan.dim, plan.n_groups
dv, dy, dy0 = DeviceBuffer(dim * 16), DeviceBuffer(dim * 16), DeviceBuffer(dim * 16)
for c0 in range(0, dim, 1 << 22):
    v = H.lanczos_start_vector(c0, min(dim, c0 + (1 << 22)))
    call("qr_memcpy_h2d", dv.ptr + c0 * 16, v.ctypes.data, v.nbytes, None)
st = C.c_void_p(); call("qr_stream_create", C.byref(st))
e0, e1 = C.c_void_p(), C.c_void_p(); call("qr_event_create", C.byref(e0)); call("qr_event_create", C.byref(e1))
ref = None
for f in a.fs.split():
    os.environ[a.var] = f
    for _ in range(3):
        call("qr_apply_device", plan.handle, 0, dim, dv.ptr, dy.ptr, st)
    call("qr_event_record", e0, st)
    for _ in range(a.reps):
        call("qr_apply_device", plan.handle, 0, dim, dv.ptr, dy.ptr, st)
    call("qr_event_record", e1, st)
    ms = C.c_float(); call("qr_event_elapsed_ms", e0, e1, C.byref(ms)); t = ms.value / a.reps
    head = np.empty(1 << 16, dtype=np.complex128)
    call("qr_memcpy_d2h", head.ctypes.data, dy.ptr + (dim // 2) * 16, head.nbytes, None)
    if ref is None: ref = head.copy()
    print(json.dumps({"workload": a.workload, a.var: int(f), "kernel": plan.apply_kernel(), "n": plan.n_qubits, "G": G,
                      "ms": round(t, 4), "GBps_compulsory": round(32 * dim / t / 1e6, 1), "same_bits": bool(np.array_equal(head.view(np.uint64), ref.view(np.uint64)))}), flush=True)
