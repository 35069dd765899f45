#!/usr/bin/env python
"""Times the fill kernel (events on the launching stream) for one workload under several
(RW,GW) tile configurations, the direct kernel, and a row window size.  GPU box only.
  python tools/fill_sweep.py C2 | C4 | C3 | xxz<n> | H8 ...   [--rows LOG2] [--cfgs "2,4 1,8 direct lanes:8:8 blocked:32:1 rows:1:1:7"]
"""
import argparse, ctypes as C, gzip, json, os, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import qrusty_b200 as Q
from qrusty_b200 import _ffi, hamiltonians as H
from qrusty_b200._ffi import call
from qrusty_b200._runtime import DeviceBuffer


def get_workload(name):
    if name in H.CONFIGS:
        return H.CONFIGS[name][1]()
    if name.startswith("xxz"):
        return H.xxz_chain(int(name[3:]), 1.0, 0.7)
    if name.startswith("rand:"):                      # rand:<n>:<T>:<G>  (C3's generator at another shape)
        n, T, G = (int(v) for v in name.split(":")[1:4])
        return H.random_pauli_sum(n, T, G, min(100, (T - G) // 2), 24)
    fx = json.load(gzip.open(ROOT / "tests/golden/h_fixtures.json.gz"))
    return fx[name]["labels"], [complex(a, b) for a, b in fx[name]["coeffs"]]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload")
    ap.add_argument("--rows", type=int, default=None, help="log2 of the row window (default: all rows that fit 8 GB)")
    ap.add_argument("--cfgs", default="auto direct")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--max-gb", type=float, default=8.0, help="cap on the output size of one build")
    a = ap.parse_args()
    labels, coeffs = get_workload(a.workload)
    terms = Q.SparsePauliOp([Q.Pauli(l) for l in labels], coeffs).terms()
    n = len(labels[0])
    stream = C.c_void_p(); call("qr_stream_create", C.byref(stream))
    e0, e1 = C.c_void_p(), C.c_void_p(); call("qr_event_create", C.byref(e0)); call("qr_event_create", C.byref(e1))
    bufs = None
    for cfg in a.cfgs.split():
        flags = 0
        for k in ("QR_FILL_CFG", "QR_FILL_LANES", "QR_FILL_LANES_R", "QR_FILL_LANES_W", "QR_FILL_LANES_SYNC",
                  "QR_FILL_LANES_PERSIST", "QR_FILL_BLOCK", "QR_FILL_BLOCK_E", "QR_FILL_ROWS", "QR_FILL_ROWS_REGT", "QR_FILL_ROWS_HVS", "QR_FILL_ROWS_SL", "QR_FILL_ROWS_CL", "QR_FILL_ROWS_SPLIT",
                  "QR_FILL_ROWS_Q", "QR_FILL_ROWS_R", "QR_FILL_ROWS_HV"):      # the ones a cfg string sets
            os.environ.pop(k, None)
        os.environ.pop("QR_FILL_SWZ", None)
        if cfg.split(":")[0] in ("swz", "noswz"):    # staged kernel with / without the swizzled tile (rows kernel off) [:RW,GW]
            os.environ["QR_FILL_SWZ"] = "1" if cfg.startswith("swz") else "0"
            os.environ["QR_FILL_ROWS"] = "0"
            if ":" in cfg: os.environ["QR_FILL_CFG"] = cfg.split(":")[1]
        elif cfg == "direct":
            flags = _ffi.QR_FILL_DIRECT
        elif cfg.startswith("rows"):                  # rows[:regt 0|1[:log2(rows per batch)[:log2(rows per run)[:heavy threshold[:log2 heavy strip[:log2 sub-batches[:cluster size[:split S]]]]]]]], "" = default
            parts = cfg.split(":")
            os.environ["QR_FILL_ROWS"] = "1"
            for i, key in enumerate(("QR_FILL_ROWS_REGT", "QR_FILL_ROWS_Q", "QR_FILL_ROWS_R", "QR_FILL_ROWS_HV", "QR_FILL_ROWS_HVS", "QR_FILL_ROWS_SL", "QR_FILL_ROWS_CL", "QR_FILL_ROWS_SPLIT")):
                if len(parts) > i + 1 and parts[i + 1] != "": os.environ[key] = parts[i + 1]
        elif cfg.startswith("lanes"):                 # lanes[:log2R[:warps[:sync[:persist]]]]
            parts = cfg.split(":")
            os.environ["QR_FILL_LANES"] = "1"
            if len(parts) > 1: os.environ["QR_FILL_LANES_R"] = parts[1]
            if len(parts) > 2: os.environ["QR_FILL_LANES_W"] = parts[2]
            if len(parts) > 3: os.environ["QR_FILL_LANES_SYNC"] = parts[3]
            if len(parts) > 4: os.environ["QR_FILL_LANES_PERSIST"] = parts[4]
        elif cfg.startswith("blocked"):               # blocked[:S[:E]]
            parts = cfg.split(":")
            os.environ["QR_FILL_LANES"] = "0"
            os.environ["QR_FILL_BLOCK"] = parts[1] if len(parts) > 1 else "32"
            if len(parts) > 2: os.environ["QR_FILL_BLOCK_E"] = parts[2]
        elif cfg != "auto" and cfg.split(":")[0] not in ("swz", "noswz"):
            os.environ["QR_FILL_CFG"] = cfg
        op = Q.SparsePauliOp.from_terms(n, terms)
        plan = op.plan()
        G, dim = plan.n_groups, plan.dim
        rows = dim
        if a.rows is not None:
            rows = min(dim, 1 << a.rows)
        while rows * G * 24 > a.max_gb * 1e9:
            rows //= 2
        if bufs is None:
            bufs = (DeviceBuffer((rows + 1) * 8), DeviceBuffer(rows * G * 8), DeviceBuffer(rows * G * 16))
        ip, ix, dt = bufs
        lo = (dim // 2) // rows * rows if rows < dim else 0
        for _ in range(3):
            call("qr_build_rows_device", plan.handle, lo, lo + rows, ip.ptr, ix.ptr, dt.ptr, flags, stream)
        call("qr_event_record", e0, stream)
        for _ in range(a.reps):
            call("qr_build_rows_device", plan.handle, lo, lo + rows, ip.ptr, ix.ptr, dt.ptr, flags, stream)
        call("qr_event_record", e1, stream)
        ms = C.c_float(); call("qr_event_elapsed_ms", e0, e1, C.byref(ms))
        t = ms.value / a.reps
        nbytes = rows * G * 24 + (rows + 1) * 8
        print(json.dumps({"workload": a.workload, "cfg": cfg, "kernel": plan.fill_kernel, "n": n, "T": len(labels), "G": G, "rows": rows,
                          "ms": round(t, 5), "GBps": round(nbytes / t / 1e6, 1), "Gnnz_s": round(rows * G / t / 1e6, 2)}), flush=True)


if __name__ == "__main__":
    main()
