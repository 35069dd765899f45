#!/usr/bin/env python
"""The elementwise kernels either side of H.v (accel.rs:374-393, pyqrusty/src/lib.rs:436-468) on 2^25 elements, against
the bytes they must move.  GPU box only."""
import ctypes as C, json, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import qrusty_b200 as Q
from qrusty_b200._ffi import call
from qrusty_b200._runtime import DeviceBuffer
from qrusty_b200 import hamiltonians as H
n = 1 << 25
x, y, z = DeviceBuffer(n * 16), DeviceBuffer(n * 16), DeviceBuffer(n * 16)
for c0 in range(0, n, 1 << 22):
    v = H.lanczos_start_vector(c0, c0 + (1 << 22))
    call("qr_memcpy_h2d", x.ptr + c0 * 16, v.ctypes.data, v.nbytes, None); call("qr_memcpy_h2d", y.ptr + c0 * 16, v.ctypes.data, v.nbytes, None)
st = C.c_void_p(); call("qr_stream_create", C.byref(st))
e0, e1 = C.c_void_p(), C.c_void_p(); call("qr_event_create", C.byref(e0)); call("qr_event_create", C.byref(e1))
a = (C.c_double * 2)(0.3, -0.2); b = (C.c_double * 2)(1.1, 0.4); out = DeviceBuffer(64)
op = Q.SparsePauliOp([Q.Pauli(l) for l in H.CONFIGS["C4"][1]()[0]], H.CONFIGS["C4"][1]()[1]); plan = op.plan()
def timed(fn, reps=20):
    for _ in range(3): fn()
    call("qr_event_record", e0, st)
    for _ in range(reps): fn()
    call("qr_event_record", e1, st); call("qr_stream_synchronize", st)
    ms = C.c_float(); call("qr_event_elapsed_ms", e0, e1, C.byref(ms)); return ms.value / reps
cases = [("axpby", 48, lambda: call("qr_axpby_device", n, a, x.ptr, b, y.ptr, z.ptr, st)),
         ("axpy", 48, lambda: call("qr_axpy_device", n, a, x.ptr, y.ptr, z.ptr, st)),
         ("ax", 32, lambda: call("qr_ax_device", n, a, x.ptr, z.ptr, st)),
         ("dotc", 32, lambda: call("qr_dotc_device", n, x.ptr, y.ptr, out.ptr, st)),
         ("precond2", 48, lambda: call("qr_precond2_device", n, x.ptr, y.ptr, a, 1e-8, z.ptr, st)),
         ("lanczos_update", 64, lambda: call("qr_lanczos_update_device", n, a, b, z.ptr, x.ptr, y.ptr, z.ptr, out.ptr, st)),
         ("diagonal C4 (matrix-free, 40 Z-only terms)", 16, lambda: call("qr_diagonal_device", plan.handle, 0, n, z.ptr, st))]
for name, bpe, fn in cases:
    t = timed(fn)
    print(json.dumps({"kernel": name, "n": n, "ms": round(t, 4), "bytes_per_element": bpe, "GBps": round(n * bpe / t / 1e6, 1)}), flush=True)
