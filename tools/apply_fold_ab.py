#!/usr/bin/env python
"""A/B of the two matrix-free H.v kernels on one workload: QR_APPLY_FOLD=0 (gather, term by term) vs 1 (bucketed fold).
  python tools/apply_fold_ab.py H12 [--reps 5] [--rows LOG2]     workloads as tools/fill_sweep.py.  GPU box only."""
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

ap = argparse.ArgumentParser(); ap.add_argument("workload"); ap.add_argument("--reps", type=int, default=5); ap.add_argument("--modes", default="gather,fold,ptile11,ptile12,default")
a = ap.parse_args()
labels, coeffs = get_workload(a.workload)
dim = 1 << len(labels[0])
dv, dy = DeviceBuffer(dim * 16), DeviceBuffer(dim * 16)
for c0 in range(0, dim, 1 << 22):
    v = H.lanczos_start_vector(c0, min(dim, c0 + (1 << 22)))
    call("qr_memcpy_h2d", dv.ptr + c0 * 16, v.ctypes.data, v.nbytes, None)
st = C.c_void_p(); call("qr_stream_create", C.byref(st))
e0, e1 = C.c_void_p(), C.c_void_p(); call("qr_event_create", C.byref(e0)); call("qr_event_create", C.byref(e1))
ys = {}
MODES = {"gather": ("0", "0", "12"), "fold": ("1", "0", "12"), "ptile10": ("1", "1", "10"), "ptile11": ("1", "1", "11"), "ptile12": ("1", "1", "12"),
         "default": (None, None, None)}
for mode in (a.modes.split(",")):
    for key, val in zip(("QR_APPLY_FOLD", "QR_APPLY_PTILE", "QR_APPLY_PTILE_K"), MODES[mode]):
        os.environ.pop(key, None)
        if val is not None: os.environ[key] = val
    plan = Q.SparsePauliOp([Q.Pauli(l) for l in labels], coeffs).plan()
    for _ in range(2):
        call("qr_apply_device", plan.handle, 0, dim, dv.ptr, dy.ptr, st)
    call("qr_event_record", e0, st)
    for _ in range(a.reps):
        call("qr_apply_device", plan.handle, 0, dim, dv.ptr, dy.ptr, st)
    call("qr_event_record", e1, st)
    call("qr_stream_synchronize", st)
    ms = C.c_float(); call("qr_event_elapsed_ms", e0, e1, C.byref(ms)); t = ms.value / a.reps
    y = np.empty(min(dim, 1 << 20), np.complex128)
    call("qr_memcpy_d2h", y.ctypes.data, dy.ptr, y.nbytes, None); call("qr_stream_synchronize", None)
    ys[mode] = y
    print(json.dumps({"workload": a.workload, "mode": mode, "nbuf": os.environ.get("QR_APPLY_PTILE_NBUF"), "kernel": plan.apply_kernel(), "n": plan.n_qubits, "T": plan.n_terms, "G": plan.n_groups,
                      "ms": round(t, 4), "term_row_evals_per_s": plan.n_terms * dim / t * 1e3, "GBps_compulsory": round(32 * dim / t / 1e6, 1)}), flush=True)
k0 = list(ys)[0]
print(json.dumps({"max_abs_diff_first_2^20_rows_vs_" + k0: {k: float(np.abs(ys[k0] - ys[k]).max()) for k in ys}, "max_abs_y": float(np.abs(ys[k0]).max())}))
