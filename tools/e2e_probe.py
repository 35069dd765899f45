#!/usr/bin/env python
"""Where does the end-to-end time go?  Times each piece of to_matrix_mode('Cuda').export() on C2."""
import ctypes as C, sys, time, os
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import qrusty_b200 as Q
from qrusty_b200 import _ffi, hamiltonians as H
from qrusty_b200._ffi import call
from qrusty_b200._runtime import DeviceBuffer, pinned_empty, PINNED

def T(f, reps=5):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); r = f(); call("qr_stream_synchronize", None); ts.append(time.perf_counter() - t0); del r
    return min(ts) * 1e3, float(np.median(ts)) * 1e3

labels, coeffs = H.xxz_chain(20, 1.0, 0.7)
op = Q.SparsePauliOp([Q.Pauli(l) for l in labels], coeffs)
terms = op.terms()
plan = op.plan()
nb = 352 << 20
print("cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
try:
    import subprocess
    print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:1500])
    print(subprocess.run(["bash", "-c", "lscpu | grep -i -E 'numa|socket|model name'"], capture_output=True, text=True).stdout)
except Exception as e:
    print(e)
print("plan create      min/med ms", T(lambda: Q.SparsePauliOp.from_terms(20, terms).plan()))
print("cudaMalloc+Free 352MB     ", T(lambda: DeviceBuffer(nb)))
d = DeviceBuffer(nb)
print("pinned alloc 352MB (cold) ", T(lambda: pinned_empty(nb // 16, np.complex128), reps=1))
print("pinned alloc 352MB (pool) ", T(lambda: pinned_empty(nb // 16, np.complex128)))
h = pinned_empty(nb // 16, np.complex128)
ms = T(lambda: d.download(h))
print("D2H 352MB pinned          ", ms, "GB/s", nb / ms[0] / 1e6)
hp = np.empty(nb // 16, np.complex128)
ms = T(lambda: d.download(hp))
print("D2H 352MB pageable        ", ms, "GB/s", nb / ms[0] / 1e6)
ms = T(lambda: d.upload(h))
print("H2D 352MB pinned          ", ms, "GB/s", nb / ms[0] / 1e6)
print("build device-resident     ", T(lambda: op.to_matrix_mode("Cuda")))
def full():
    return Q.SparsePauliOp.from_terms(20, terms).to_matrix_mode("Cuda").export()
print("full e2e step             ", T(full))
m = op.to_matrix_mode("Cuda")
t0 = time.perf_counter(); out = m.export(); print("export only ms", (time.perf_counter() - t0) * 1e3)
# windowed qr_build_host into pinned buffers
G, dim = plan.n_groups, plan.dim
ip = pinned_empty(dim + 1, np.uint64); ix = pinned_empty(dim * G, np.uint64); dt = pinned_empty(dim * G, np.complex128)
print("qr_build_host pinned      ", T(lambda: call("qr_build_host", plan.handle, 0, dim, ip.ctypes.data, ix.ctypes.data, dt.ctypes.data, 0)))
