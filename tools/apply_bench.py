#!/usr/bin/env python
"""Times matrix-free H.v for one workload: tiled passes vs the v0 gather kernel.  GPU box only."""
import argparse, ctypes as C, json, os, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
import qrusty_b200 as Q
from qrusty_b200 import _ffi, hamiltonians as H
from qrusty_b200._ffi import call
from qrusty_b200._runtime import DeviceBuffer
from fill_sweep import get_workload

ap = argparse.ArgumentParser(); ap.add_argument("workload"); ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()
labels, coeffs = get_workload(a.workload)
op = Q.SparsePauliOp([Q.Pauli(l) for l in labels], coeffs)
plan = op.plan(); dim, G = plan.dim, plan.n_groups
dv, dy = DeviceBuffer(dim * 16), DeviceBuffer(dim * 16)
for c0 in range(0, dim, 1 << 22):
    v = H.lanczos_start_vector(c0, min(dim, c0 + (1 << 22)))
    call("qr_memcpy_h2d", dv.ptr + c0 * 16, v.ctypes.data, v.nbytes, None)
st = C.c_void_p(); call("qr_stream_create", C.byref(st))
e0, e1 = C.c_void_p(), C.c_void_p(); call("qr_event_create", C.byref(e0)); call("qr_event_create", C.byref(e1))
for mode in ("0", "1"):
    os.environ["QR_APPLY_V0"] = mode
    for _ in range(3):
        call("qr_apply_device", plan.handle, 0, dim, dv.ptr, dy.ptr, st)
    call("qr_event_record", e0, st)
    for _ in range(a.reps):
        call("qr_apply_device", plan.handle, 0, dim, dv.ptr, dy.ptr, st)
    call("qr_event_record", e1, st)
    ms = C.c_float(); call("qr_event_elapsed_ms", e0, e1, C.byref(ms)); t = ms.value / a.reps
    print(json.dumps({"workload": a.workload, "kernel": "v0_gather" if mode == "1" else "tiled_passes", "n": plan.n_qubits,
                      "T": len(labels), "G": G, "ms": round(t, 4), "GBps_compulsory": round(32 * dim / t / 1e6, 1),
                      "Grows_s": round(dim / t / 1e6, 2)}), flush=True)
