#!/usr/bin/env python
"""BASELINE configs 4/5 on N GPUs (one process per GPU, torchrun): row-sharded CSR build left in
HBM, sampled rows verified against the oracle, then distributed matrix-free H.v (NCCL all-gather
inside the library) verified on sampled rows.  Prints one JSON line from rank 0.
  torchrun --nproc-per-node 8 tools/multi_gpu_config.py C5
"""
import ctypes as C, json, os, sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch, torch.distributed as dist
import qrusty_b200 as Q
from qrusty_b200 import _ffi, hamiltonians as H, dist as qd
from qrusty_b200._ffi import call
from qrusty_b200._runtime import DeviceBuffer
from oracle import oracle as O

cfg = sys.argv[1] if len(sys.argv) > 1 else "C5"
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
call("qr_set_device", local)

def barrier():
    call("qr_stream_synchronize", None)
    if world > 1:
        dist.barrier(); torch.cuda.synchronize()

def maxr(x):
    return qd.max_over_ranks(dist, x, "cuda") if world > 1 else x

sys.path.insert(0, str(ROOT / "tools"))
from fill_sweep import get_workload
labels, coeffs = get_workload(cfg)
n, params = O.make_params(labels, coeffs)
op = Q.SparsePauliOp([Q.Pauli(l) for l in labels], coeffs)
plan = op.plan(local)
G, dim = plan.n_groups, plan.dim
lo, hi = qd.row_block(rank, world, dim)
rows = hi - lo
st = C.c_void_p(); call("qr_stream_create", C.byref(st))
def ev():
    e = C.c_void_p(); call("qr_event_create", C.byref(e)); return e
def elapsed(a, b):
    ms = C.c_float(); call("qr_event_elapsed_ms", a, b, C.byref(ms)); return ms.value

out = {"config": cfg, "workload": H.CONFIGS[cfg][0] if cfg in H.CONFIGS else cfg, "n_gpus": world, "n_qubits": n, "n_terms": len(labels), "n_groups": G,
       "nnz": G * dim, "rows_per_gpu": rows, "csr_bytes_per_gpu": rows * G * 24 + (rows + 1) * 8}
# ---- build the shard in HBM ----
d_ip, d_ix, d_dt = DeviceBuffer((rows + 1) * 8, local), DeviceBuffer(rows * G * 8, local), DeviceBuffer(rows * G * 16, local)
for _ in range(2):
    call("qr_build_rows_device", plan.handle, lo, hi, d_ip.ptr, d_ix.ptr, d_dt.ptr, _ffi.QR_INDPTR_GLOBAL, st)
call("qr_stream_synchronize", st)
e0, e1 = ev(), ev(); reps = 5
barrier()
call("qr_event_record", e0, st)
for _ in range(reps):
    call("qr_build_rows_device", plan.handle, lo, hi, d_ip.ptr, d_ix.ptr, d_dt.ptr, _ffi.QR_INDPTR_GLOBAL, st)
call("qr_event_record", e1, st); call("qr_stream_synchronize", st)
t_build = maxr(elapsed(e0, e1) / reps)
out.update(build_ms=t_build, build_nnz_per_s=G * dim / (t_build * 1e-3),
           build_GBps_per_gpu=(rows * G * 24 + (rows + 1) * 8) / t_build / 1e6)
# verify 1024 sampled rows of this shard against the oracle
rng = np.random.default_rng(100 + rank)
sample = np.unique(np.r_[lo, hi - 1, rng.integers(lo, hi, min(1022, rows))])
bad = 0
row_ix, row_dt, ipv = np.empty(G, np.uint64), np.empty(G, np.complex128), np.empty(2, np.uint64)
for r in sample:
    o = (int(r) - lo) * G
    d_ix.download(row_ix, offset=o * 8); d_dt.download(row_dt, offset=o * 16); d_ip.download(ipv, offset=(int(r) - lo) * 8)
    cols, vals = O.make_row(params, int(r))
    ok = np.array_equal(cols, row_ix) and np.array_equal(vals.view(np.uint64), row_dt.view(np.uint64)) \
        and ipv[0] == int(r) * G and ipv[1] == (int(r) + 1) * G
    bad += 0 if ok else 1
out["rows_verified_per_gpu"] = int(len(sample)); out["rows_bad"] = int(maxr(bad))
del d_ip, d_ix, d_dt
# ---- distributed matrix-free H.v ----
d_vs, d_vf, d_y = DeviceBuffer(rows * 16, local), DeviceBuffer(dim * 16, local), DeviceBuffer(rows * 16, local)
for c0 in range(lo, hi, 1 << 22):
    v = H.lanczos_start_vector(c0, min(hi, c0 + (1 << 22)))
    call("qr_memcpy_h2d", d_vs.ptr + (c0 - lo) * 16, v.ctypes.data, v.nbytes, None)
comm = qd.create_comm(dist, local) if world > 1 else None
def hv():
    if comm is None: call("qr_apply_device", plan.handle, lo, hi, d_vs.ptr, d_y.ptr, st)
    else: call("qr_apply_distributed", plan.handle, comm, d_vs.ptr, d_vf.ptr, d_y.ptr, st)
for _ in range(3): hv()
call("qr_stream_synchronize", st)
barrier(); reps = 10
call("qr_event_record", e0, st)
for _ in range(reps): hv()
call("qr_event_record", e1, st); call("qr_stream_synchronize", st)
t_hv = maxr(elapsed(e0, e1) / reps)
out.update(hv_ms=t_hv, hv_GBps_compulsory=32.0 * dim / t_hv / 1e6,
           hv_nvlink_GBps_in_per_gpu=(16.0 * dim * (world - 1) / world / t_hv / 1e6) if world > 1 else None)
# fused variant: peers' shards read in place over NVLink (no all-gather, no v_full)
if comm is not None:
    ptrs, opened = qd.share_shards(dist, d_vs.ptr)
    parr = qd.pointer_array(ptrs)
    d_y2 = DeviceBuffer(rows * 16, local)
    for _ in range(3): call("qr_apply_p2p", plan.handle, comm, parr, d_y2.ptr, st)
    call("qr_stream_synchronize", st)
    barrier()
    call("qr_event_record", e0, st)
    for _ in range(reps): call("qr_apply_p2p", plan.handle, comm, parr, d_y2.ptr, st)
    call("qr_event_record", e1, st); call("qr_stream_synchronize", st)
    t_p2p = maxr(elapsed(e0, e1) / reps)
    ya, yb = np.empty(min(rows, 1 << 16), np.complex128), np.empty(min(rows, 1 << 16), np.complex128)
    d_y.download(ya); d_y2.download(yb)
    diff = np.abs(ya - yb)
    out.update(hv_p2p_ms=t_p2p, hv_p2p_GBps_compulsory=32.0 * dim / t_p2p / 1e6, hv_p2p_max_abs_diff=maxr(float(diff.max())),
               hv_p2p_n_diff=int(maxr(float(np.count_nonzero(diff)))), hv_p2p_first_diff=int(np.flatnonzero(diff)[0]) if diff.any() else -1,
               hv_p2p_equals_allgather=bool(maxr(0.0 if np.array_equal(ya.view(np.uint64), yb.view(np.uint64)) else 1.0) == 0.0))
    barrier()
    qd.close_shards(opened)
ys = rng.integers(lo, hi, 4096)
yv = np.empty(1, np.complex128); worst = 0.0
absH = float(np.abs(params["re"] + 1j * params["im"]).sum())
for r in ys[:512]:
    d_y.download(yv, offset=(int(r) - lo) * 16)
    cols, vals = O.make_row(params, int(r))
    ref = np.sum(vals * H.lanczos_start_at(cols))
    worst = max(worst, abs(yv[0] - ref) / (absH * 1.5))
out["hv_rows_verified_per_gpu"] = 512; out["hv_max_rel_err"] = maxr(worst)
if comm is not None: call("qr_comm_destroy", comm)
barrier()
if rank == 0: print(json.dumps(out), flush=True)
if world > 1: dist.destroy_process_group()
