#!/usr/bin/env python
"""Matrix-free H.v on one GPU: the single-pass gather kernel (QR_APPLY_D=0) against the two-pass form
(gather below bit d + shared-memory tiles from bit d up, apply_tile.cuh) for several cut bits d.
Each variant is checked against the single-pass result.  GPU box only.
  python tools/apply_sweep.py C4 [--cuts "0 19 20 21 22 23"] [--stages "3"] [--reps 20]
"""
import argparse, ctypes as C, json, os, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
import qrusty_b200 as Q
from qrusty_b200 import hamiltonians as H
from qrusty_b200._ffi import call
from qrusty_b200._runtime import DeviceBuffer
from fill_sweep import get_workload

ap = argparse.ArgumentParser()
ap.add_argument("workload"); ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--cuts", default="0 19 20 21 22 23", help="QR_APPLY_D values (needs QR_APPLY_TILE=1 in the environment)"); ap.add_argument("--stages", default="3")
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
sample = np.random.default_rng(1).integers(0, dim - 4096, 64)
ref = None
for stages in a.stages.split():
    os.environ["QR_APPLY_TILE_STAGES"] = stages
    for cut in a.cuts.split():
        os.environ["QR_APPLY_D"] = cut
        # a fresh plan per variant: tile plans are cached per (block, cut) and the stage count is read when they are made
        pl = Q.SparsePauliOp([Q.Pauli(l) for l in labels], coeffs).plan()
        call("qr_memset_device", dy.ptr, 0xFF, dim * 16, st)
        for _ in range(3):
            call("qr_apply_device", pl.handle, 0, dim, dv.ptr, dy.ptr, st)
        call("qr_event_record", e0, st)
        for _ in range(a.reps):
            call("qr_apply_device", pl.handle, 0, dim, dv.ptr, dy.ptr, st)
        call("qr_event_record", e1, st)
        ms = C.c_float(); call("qr_event_elapsed_ms", e0, e1, C.byref(ms)); t = ms.value / a.reps
        got = np.concatenate([dy.download(np.empty(4096, np.complex128), offset=int(o) * 16) for o in sample])
        if ref is None:
            ref = got
        err = float(np.abs(got - ref).max() / max(np.abs(ref).max(), 1e-300))
        print(json.dumps({"workload": a.workload, "n": plan.n_qubits, "G": G, "cut_bit": int(cut), "stages": int(stages),
                          "ms": round(t, 4), "GBps_compulsory": round(32 * dim / t / 1e6, 1),
                          "max_rel_diff_vs_first": err}), flush=True)
        del pl
