#!/usr/bin/env python
"""bench.py -- SparsePauliOp -> CSR throughput (nnz/s) on N B200s, plus the roofline of the
fill kernel, the matrix-free H.v figure, the end-to-end number through the public API with host
buffers, and the CPU port of the reference algorithm timed beside it.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config C2]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one operator already resident in HBM: the
canonicalisation kernel (K1) followed by the fill kernel(s) (K3, which also writes indptr),
writing a device-resident CSR shard.  N=1: BASELINE config 2 (XXZ periodic chain n=20).
N>1: the same operator with log2(N) spectator qubits (H (x) I), row-block sharded, 2^20 rows of 21
entries per GPU at every N (weak scaling; the build needs no collective).  The multi-GPU BASELINE
configs are measured AND verified in the same run (`extras.c4`: TFIM 5x5, n=25, strong-scaled build,
fused distributed H.v, 50 Lanczos iterations; `extras.c5`: Heisenberg n=28 where it fits, N >= 2):
sampled rows of every shard against the oracle's make_row, sampled H.v elements against the oracle's
dot, the fused peer-memory H.v against the all-gather form bit for bit.  torch is used only for the
rendezvous, the barrier and the max-over-ranks reduction; the reference arm imports neither torch
nor the CUDA library.
"""
import argparse
import ctypes as C
import hashlib
import importlib.util
import json
import math
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC, UNIT = "csr_build_nnz_per_s", "nnz/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="auto", help="auto | C2 | C4 | xxz<n>")
    ap.add_argument("--replays", type=int, default=10, help="extra timed replays of the K-step graph (median + spread)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-hv", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-c5", action="store_true", help="skip extras.c5 (Heisenberg n=28) at N >= 2")
    ap.add_argument("--no-graph", action="store_true", help="time K eager step launches instead of one CUDA graph of K steps")
    return ap.parse_args()


def load_hamiltonians():
    """qrusty_b200/hamiltonians.py loaded by path: the generators need numpy only, and importing the
    package would dlopen the CUDA library -- which the reference arm must not do."""
    spec = importlib.util.spec_from_file_location("qr_hamiltonians", ROOT / "qrusty_b200" / "hamiltonians.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def workload(args, world):
    """-> (name, labels, coeffs).  N=1: BASELINE config 2.  N>1 (weak scaling): the same operator with
    log2(N) spectator qubits on top, H (x) I -- 2^20 rows of 21 entries per GPU at every N, so the
    per-GPU work is exactly that of the N=1 line."""
    H = load_hamiltonians()
    cfg = args.config
    spect = int(math.log2(world)) if cfg == "auto" else 0
    if cfg in ("auto", "C2"):
        cfg = "xxz20"
    if cfg == "C4":
        return "tfim_5x5_n25", *H.tfim_lattice(5, 5, 1.0, 3.0)
    if cfg.startswith("xxz"):
        n = int(cfg[3:])
        labels, coeffs = H.xxz_chain(n, 1.0, 0.7)
        name = "xxz_periodic_n%d_J1_delta0.7" % n
        if spect:
            labels = ["I" * spect + l for l in labels]
            name += " (x) I^%d (%d spectator qubits on top: %d qubits, rows sharded over %d GPUs)" % (spect, spect, n + spect, world)
        return name, labels, coeffs
    raise SystemExit("unknown --config " + cfg)


def config_dict(name, n, n_terms, G, world):
    """The `config` object of the JSON line -- built by this one function for BOTH arms."""
    dim = 1 << n
    rows = dim // world
    bytes_local = rows * G * 24 + (rows + 1) * 8               # B_csr
    return {"workload": name, "n_qubits": n, "n_terms": n_terms, "n_groups": G, "nnz": G * dim,
            "rows_per_gpu": rows, "bytes_per_gpu": bytes_local, "parallelism": "row-block x%d, no collective" % world,
            "step": "canonicalise kernel + fill kernel(s), outputs device-resident",
            "l2": "each step writes %.0f MB per GPU (> 126 MB L2), no flush needed" % (bytes_local / 1e6)}


# ------------------------------------------------------------------------------------------
# clocks: sampled DURING the timed region with NVML (nvidia-smi as a fallback)
# ------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown",
               0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}

    def __init__(self, local_rank, interval=0.01, enabled=True):
        # NVML queries contend with CUDA driver calls (a 2 ms poll slowed cudaMemcpy 4x), so poll
        # gently and stop before the host-facing e2e loop
        self.interval = interval
        self.samples, self.reasons, self.max_mhz, self.ok = [], set(), None, False
        self._stop = threading.Event()
        if not enabled:
            self.err = "disabled on this rank"
            return
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            idx = local_rank
            if vis and all(p.strip().isdigit() for p in vis.split(",")):
                idx = int(vis.split(",")[local_rank])
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:                       # pragma: no cover - depends on the box
            self.err = repr(e)

    def _loop(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.interval)

    def start(self):
        if self.ok:
            self.t = threading.Thread(target=self._loop, daemon=True)
            self.t.start()

    def stop(self):
        if self.ok:
            self._stop.set()
            self.t.join()

    def summary(self, window):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "window": window,
                    "note": "no NVML samples" + (": " + getattr(self, "err", "") if not self.ok else "")}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples), "window": window}


# ------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle's port of accel.rs:267-336 on the host cores
# ------------------------------------------------------------------------------------------
def cpu_build_rate(labels, coeffs, budget_s, max_rows=None):
    """Times oracle.build_csr (chunk step 1000 as perf_giant.py:34, all host threads) on the whole
    matrix, or on a leading row window if one build would exceed the budget.  -> dict."""
    from oracle import oracle as O
    n, params = O.make_params(labels, coeffs)
    G = len(np.unique(params["x"]))
    threads = O.hardware_threads()
    dim = 1 << n
    probe = min(dim, 1 << 14)
    t0 = time.perf_counter(); O.build_csr(params, n, 0, probe, step=1000, groups=G); t_probe = time.perf_counter() - t0
    rows = dim
    est = t_probe * dim / probe
    if est > budget_s / 2:                                        # bounded sample
        rows = max(probe, int(dim * (budget_s / 2) / est) // 1024 * 1024)
    if max_rows:
        rows = min(rows, max_rows)
    times = []
    t_start = time.perf_counter()
    while not times or (time.perf_counter() - t_start < budget_s and len(times) < 5):
        t0 = time.perf_counter(); O.build_csr(params, n, 0, rows, step=1000, groups=G); times.append(time.perf_counter() - t0)
    t = float(np.median(times))
    sample = ("full matrix" if rows == dim else "rows [0,%d) of %d" % (rows, dim)) + ", %d runs, median" % len(times)
    return {"value": rows * G / t, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": sample, "seconds_per_run": t}


def cpu_tuned_rate(labels, coeffs, runs=3):
    """A tuned CPU variant beside the port, NOT the reference's algorithm: terms grouped by X-mask once, every entry written
    straight to its closed-form slot, no per-row sort, no concat passes (oracle_build_grouped) -- what the host cores do with
    the same observations the CUDA path rests on.  Whole matrix, all host threads, median of `runs`."""
    from oracle import oracle as O
    n, params = O.make_params(labels, coeffs)
    G = len(np.unique(params["x"]))
    times = []
    for _ in range(runs):
        t0 = time.perf_counter(); O.build_csr_grouped(params, n, groups=G); times.append(time.perf_counter() - t0)
    t = float(np.median(times))
    return {"value": (1 << n) * G / t, "unit": UNIT, "cores": O.hardware_threads(), "kind": "tuned variant, not the reference's algorithm",
            "sample": "full matrix, %d runs, median (includes allocating the %d MB of output, as the port's timing does)" % (runs, ((1 << n) * G * 24) >> 20),
            "seconds_per_run": t}


def run_reference(args, rank, world):
    """The reference's CPU path (its port, oracle/qrusty_oracle.c) on the host cores.  Imports nothing of
    qrusty_b200 -- neither the package nor libqrusty_cuda.so is loaded in this process."""
    if rank != 0:
        return
    name, labels, coeffs = workload(args, world)
    from oracle import oracle as O
    n, params = O.make_params(labels, coeffs)
    G = len(np.unique(params["x"]))
    dim = 1 << n
    # per-step sample: whole matrix when one build is ~1 s or less, else a leading row window
    t0 = time.perf_counter(); O.build_csr(params, n, 0, 1 << 14, step=1000, groups=G); tp = time.perf_counter() - t0
    rows = dim if tp * dim / (1 << 14) <= 3.0 else max(1 << 14, int((1 << 14) * 3.0 / tp) // 1024 * 1024)
    for _ in range(args.warmup):
        O.build_csr(params, n, 0, rows, step=1000, groups=G)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.build_csr(params, n, 0, rows, step=1000, groups=G)
    t = (time.perf_counter() - t0) / args.steps
    value = rows * G / t
    sample = "full matrix per step" if rows == dim else "rows [0,%d) of %d per step" % (rows, dim)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(name, n, len(labels), G, world),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": O.hardware_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "native_loaded": sorted(m for m in sys.modules if m.startswith("qrusty_b200")),
            "note": "CPU port (oracle/qrusty_oracle.c) of qrusty accel.rs:267-336 on all host threads; the Rust "
                    "reference cannot be built here (no cargo; un-vendored git deps)"}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def peak_hbm():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def fill_source_hash():
    """Identifies the fill kernels' source: an ncu traffic figure is only quoted for the code it was taken from."""
    h = hashlib.sha256()
    for f in ("fill.cuh", "plan.cuh", "scan.cuh"):
        h.update((ROOT / "qrusty_b200" / "csrc" / f).read_bytes())
    return h.hexdigest()[:16]


def traffic_from_profile(name, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture
    (profiles/fill_traffic.json, written by tools/ncu_traffic.py), refused when it was taken on another
    workload, another kernel or another version of the fill source."""
    p = ROOT / "profiles" / "fill_traffic.json"
    if not p.exists():
        return None, "no profiles/fill_traffic.json"
    try:
        d = json.loads(p.read_text())
    except Exception as exc:                                       # noqa: BLE001
        return None, "unreadable: %r" % exc
    if d.get("workload") != name:
        return None, "capture is of another workload"
    if not str(d.get("kernel", "")).startswith(kernel):
        return None, "capture is of another kernel (%s)" % d.get("kernel")
    if d.get("fill_source_sha256_16") != fill_source_hash():
        return None, "stale: capture predates the current fill source (%s != %s)" % (d.get("fill_source_sha256_16"), fill_source_hash())
    return d.get("dram_bytes_per_launch"), d.get("source")


class Rig:
    """One rank's handles: device, stream, events, rendezvous helpers."""

    def __init__(self, rank, local_rank, world):
        from qrusty_b200 import _ffi
        from qrusty_b200._ffi import call
        self.rank, self.device, self.world, self.call, self.ffi = rank, local_rank, world, call, _ffi
        self.dist = None
        if world > 1:
            import torch
            import torch.distributed as dist
            torch.cuda.set_device(local_rank)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            self.dist, self.torch = dist, torch
        if _ffi.device_count() <= local_rank:
            raise SystemExit("bench.py: no CUDA device %d -- there is no CPU fallback" % local_rank)
        call("qr_set_device", local_rank)
        self.stream = C.c_void_p(); call("qr_stream_create", C.byref(self.stream))

    def barrier(self):
        self.call("qr_stream_synchronize", None)
        if self.dist is not None:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        """float or 1-D list -> the element-wise maximum over the ranks."""
        if self.dist is None:
            return x
        scalar = not isinstance(x, (list, tuple, np.ndarray))
        t = self.torch.tensor([x] if scalar else list(x), dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t[0].item()) if scalar else [float(v) for v in t.tolist()]

    def ev(self):
        e = C.c_void_p(); self.call("qr_event_create", C.byref(e)); return e

    def elapsed(self, a, b):
        ms = C.c_float(); self.call("qr_event_elapsed_ms", a, b, C.byref(ms)); return ms.value

    def timed(self, fn, reps, warm=3, sync_ranks=False):
        """mean ms per call of fn over `reps` back-to-back calls on self.stream, after `warm` untimed ones."""
        for _ in range(warm):
            fn()
        self.call("qr_stream_synchronize", self.stream)
        a, b = self.ev(), self.ev()
        if sync_ranks:
            self.barrier()
        self.call("qr_event_record", a, self.stream)
        for _ in range(reps):
            fn()
        self.call("qr_event_record", b, self.stream)
        self.call("qr_stream_synchronize", self.stream)
        t = self.elapsed(a, b) / reps
        return self.max_over_ranks(t) if sync_ranks else t


def upload_start_vector(rig, H, d_buf, lo, hi, seed=25):
    chunk = 1 << 22
    for c0 in range(lo, hi, chunk):
        v = H.lanczos_start_vector(c0, min(hi, c0 + chunk), seed)
        rig.call("qr_memcpy_h2d", d_buf.ptr + (c0 - lo) * 16, v.ctypes.data, v.nbytes, None)


def verify_rows(rig, params, G, lo, hi, d_ip, d_ix, d_dt, n_sample, seed):
    """Sampled rows of a device-resident shard (global indptr) against the oracle's make_row, bit for bit."""
    from oracle import oracle as O
    rng = np.random.default_rng(seed)
    rows = hi - lo
    sample = np.unique(np.r_[lo, hi - 1, rng.integers(lo, hi, max(0, min(n_sample - 2, rows)))])
    bad = 0
    row_ix, row_dt, ipv = np.empty(G, np.uint64), np.empty(G, np.complex128), np.empty(2, np.uint64)
    for r in sample:
        r = int(r)
        o = (r - lo) * G
        d_ix.download(row_ix, offset=o * 8); d_dt.download(row_dt, offset=o * 16); d_ip.download(ipv, offset=(r - lo) * 8)
        cols, vals = O.make_row(params, r)
        ok = (np.array_equal(cols, row_ix) and np.array_equal(vals.view(np.uint64), row_dt.view(np.uint64))
              and ipv[0] == r * G and ipv[1] == (r + 1) * G)
        bad += 0 if ok else 1
    return int(len(sample)), bad


def verify_hv(rig, H, params, lo, hi, d_y, n_sample, seed):
    """Sampled elements of y = H v0 (v0 = the regenerable Lanczos start vector) against the oracle's row dot.
    -> max |y - ref| / (1.5 * sum_t |c'_t|)   (|v0_i| <= sqrt(2))."""
    from oracle import oracle as O
    rng = np.random.default_rng(seed)
    yv = np.empty(1, np.complex128)
    absH = float(np.abs(params["re"] + 1j * params["im"]).sum())
    worst = 0.0
    for r in rng.integers(lo, hi, n_sample):
        r = int(r)
        d_y.download(yv, offset=(r - lo) * 16)
        cols, vals = O.make_row(params, r)
        ref = np.sum(vals * H.lanczos_start_at(cols))
        worst = max(worst, abs(yv[0] - ref) / (absH * 1.5))
    return worst


def run_baseline_config(rig, cfg, build_reps=5, hv_reps=20, lanczos_iters=50):
    """One BASELINE multi-GPU config (C4 / C5), row-sharded over the ranks: CSR shard built in HBM and verified on
    sampled rows, distributed matrix-free H.v (all-gather form and fused peer-memory form) verified on sampled
    elements, Lanczos iterations.  Every rank returns the same dict (the timings are maxima over the ranks)."""
    import qrusty_b200 as Q
    from qrusty_b200 import hamiltonians as H, dist as qd, lanczos as qlz
    from qrusty_b200._runtime import DeviceBuffer
    from oracle import oracle as O
    call, _ffi, st, world, rank, device = rig.call, rig.ffi, rig.stream, rig.world, rig.rank, rig.device
    peak, _ = peak_hbm()
    name, gen = H.CONFIGS[cfg]
    labels, coeffs = gen()
    n, params = O.make_params(labels, coeffs)
    op = Q.SparsePauliOp([Q.Pauli(l) for l in labels], coeffs)
    plan = op.plan(device)
    G, dim = plan.n_groups, plan.dim
    lo, hi = qd.row_block(rank, world, dim)
    rows = hi - lo
    csr_bytes = rows * G * 24 + (rows + 1) * 8
    out = {"workload": name, "n_qubits": n, "n_terms": len(labels), "n_groups": G, "nnz": G * dim, "n_gpus": world,
           "rows_per_gpu": rows, "csr_bytes_per_gpu": csr_bytes, "fill_kernel": plan.fill_kernel}
    # ---- build the shard in HBM, verify sampled rows ----
    d_ip, d_ix, d_dt = DeviceBuffer((rows + 1) * 8, device), DeviceBuffer(rows * G * 8, device), DeviceBuffer(rows * G * 16, device)
    t_build = rig.timed(lambda: call("qr_build_rows_device", plan.handle, lo, hi, d_ip.ptr, d_ix.ptr, d_dt.ptr, _ffi.QR_INDPTR_GLOBAL, st),
                        build_reps, warm=2, sync_ranks=True)
    n_ver, bad = verify_rows(rig, params, G, lo, hi, d_ip, d_ix, d_dt, 1024, 100 + rank)
    out.update(build_ms=t_build, build_nnz_per_s=G * dim / (t_build * 1e-3), GBps_per_gpu=csr_bytes / t_build / 1e6,
               frac_of_peak=csr_bytes / t_build / 1e6 / peak, rows_verified=n_ver * world, rows_bad=int(rig.max_over_ranks(float(bad))))
    del d_ip, d_ix, d_dt
    # ---- matrix-free H.v on the row-sharded start vector ----
    d_vs, d_y = DeviceBuffer(rows * 16, device), DeviceBuffer(rows * 16, device)
    upload_start_vector(rig, H, d_vs, lo, hi)
    x, _, _ = plan.groups()
    n_remote = int(np.count_nonzero(x >= np.uint64(rows))) if world > 1 else 0
    comm = qd.create_comm(rig.dist, device) if world > 1 else None
    if comm is None:
        hv_ms = rig.timed(lambda: call("qr_apply_device", plan.handle, lo, hi, d_vs.ptr, d_y.ptr, st), hv_reps)
        out.update(hv_ms=hv_ms, hv_form="local matrix-free apply (gather kernel, diag(H) cached)")
    else:
        d_vf, d_y2 = DeviceBuffer(dim * 16, device), DeviceBuffer(rows * 16, device)
        ag_ms = rig.timed(lambda: call("qr_apply_distributed", plan.handle, comm, d_vs.ptr, d_vf.ptr, d_y2.ptr, st), hv_reps, sync_ranks=True)
        ptrs, opened = qd.share_shards(rig.dist, d_vs.ptr)
        parr = qd.pointer_array(ptrs)
        hv_ms = rig.timed(lambda: call("qr_apply_p2p", plan.handle, comm, parr, d_y.ptr, st), hv_reps, sync_ranks=True)
        ya, yb = np.empty(min(rows, 1 << 18), np.complex128), np.empty(min(rows, 1 << 18), np.complex128)
        d_y2.download(ya); d_y.download(yb)
        differ = 0.0 if np.array_equal(ya.view(np.uint64), yb.view(np.uint64)) else 1.0
        same = rig.max_over_ranks(differ) == 0.0
        max_diff = rig.max_over_ranks(float(np.abs(ya - yb).max()))
        rig.barrier()
        qd.close_shards(opened)
        nv_bytes = 16.0 * rows * n_remote
        out.update(hv_ms=hv_ms, hv_form="fused peer-memory apply (qr_apply_p2p): peers' shards read in place over NVLink inside the gather kernel, device-side epoch flags (no NCCL call)",
                   hv_allgather_ms=ag_ms, hv_remote_groups=n_remote, hv_nvlink_bytes_in_per_gpu=nv_bytes,
                   hv_nvlink_GBps_in=nv_bytes / hv_ms / 1e6, hv_allgather_nvlink_GBps_in=16.0 * dim * (world - 1) / world / ag_ms / 1e6,
                   hv_p2p_equals_allgather=bool(same), hv_p2p_vs_allgather_max_abs_diff=max_diff)
        del d_vf, d_y2
    out["hv_GBps_compulsory"] = 32.0 * dim / out["hv_ms"] / 1e6
    out["hv_rows_verified"] = 512 * world
    out["hv_max_rel_err"] = rig.max_over_ranks(verify_hv(rig, H, params, lo, hi, d_y, 512, 200 + rank))
    del d_vs, d_y
    # ---- Lanczos iterations (H.v, <v,w>, three-term update, norm; everything device-resident) ----
    if lanczos_iters:
        res = qlz.lanczos(op, n_iter=lanczos_iters, device=device, dist=rig.dist, comm=comm)
        out.update(lanczos_iterations=int(res["iterations"]), lanczos_iter_ms=rig.max_over_ranks(float(res["iter_ms"])),
                   lanczos_hv_ms=rig.max_over_ranks(float(res["hv_ms"])),
                   lanczos_ritz_min=float(qlz.ritz_values(res["alphas"], res["betas"])[0]),
                   lanczos_form="2 passes per iteration: apply with <u, H u> folded in its epilogue (qr_apply_dot_device / qr_apply_p2p_dot), "
                                "update fused with the norm on unnormalised vectors; scalars stay on the device (no host read inside an iteration)")
        if lanczos_iters >= 20:                                        # the round-1 loop beside it: 4 passes, 2 host reads per iteration
            old = qlz.lanczos(op, n_iter=lanczos_iters, device=device, dist=rig.dist, comm=comm, device_scalars=False)
            out.update(lanczos_iter_ms_host_scalars=rig.max_over_ranks(float(old["iter_ms"])),
                       lanczos_max_alpha_diff=float(np.abs(res["alphas"] - old["alphas"][:len(res["alphas"])]).max()))
    if comm is not None:
        call("qr_comm_destroy", comm)
    rig.barrier()
    return out


def run_b200(args, rank, local_rank, world):
    import qrusty_b200 as Q
    from qrusty_b200 import _ffi, hamiltonians as H
    from qrusty_b200._runtime import DeviceBuffer
    from oracle import oracle as O

    rig = Rig(rank, local_rank, world)
    call, device, stream = rig.call, rig.device, rig.stream
    ev, elapsed, barrier, max_over_ranks = rig.ev, rig.elapsed, rig.barrier, rig.max_over_ranks

    name, labels, coeffs = workload(args, world)
    op = Q.SparsePauliOp([Q.Pauli(l) for l in labels], coeffs)
    plan = op.plan(device)
    n, G, dim = plan.n_qubits, plan.n_groups, plan.dim
    rows = dim // world
    lo, hi = rank * rows, (rank + 1) * rows
    nnz_local, nnz_total = rows * G, dim * G
    bytes_local = nnz_local * 24 + (rows + 1) * 8          # B_csr

    d_ip, d_ix, d_dt = DeviceBuffer((rows + 1) * 8, device), DeviceBuffer(nnz_local * 8, device), DeviceBuffer(nnz_local * 16, device)

    def step(e_fill0=None, e_fill1=None):
        call("qr_plan_canonicalise_async", plan.handle, stream)
        if e_fill0 is not None:
            call("qr_event_record", e_fill0, stream)
        call("qr_build_rows_device", plan.handle, lo, hi, d_ip.ptr, d_ix.ptr, d_dt.ptr, _ffi.QR_INDPTR_GLOBAL, stream)
        if e_fill1 is not None:
            call("qr_event_record", e_fill1, stream)

    K, W = args.steps, max(args.warmup, 3)
    sampler = ClockSampler(local_rank, enabled=(rank == 0))     # 8 processes polling NVML stall launches
    for _ in range(W):
        step()
    call("qr_stream_synchronize", stream)
    # The K timed steps are recorded once into a CUDA graph and replayed with one launch: with 8 ranks
    # sharing one host the Python launch rate (4 ctypes calls per 95 us step) is otherwise what gets timed.
    # Two graphs of the same K steps: the TIMED one holds nothing but the kernels (canonicalise and fill are chained by
    # programmatic dependent launch, which an event record between them would break); the second one brackets every fill
    # launch with events on the same stream and is replayed right after each timed region for the roofline figure.
    fill_ev = [(ev(), ev()) for _ in range(K)]                  # around every fill launch of the instrumented steps
    graph, graph_ev, launches0 = None, None, _ffi.kernel_launches()
    if not args.no_graph:
        try:
            call("qr_graph_begin_capture", stream)
            for i in range(K):
                step()
            g = C.c_void_p()
            call("qr_graph_end_capture", stream, C.byref(g))
            graph = g
            call("qr_graph_launch", graph, stream)               # untimed: uploads the graph
            call("qr_stream_synchronize", stream)
        except _ffi.QrustyCudaError as exc:                       # capture unsupported: eager launches
            sys.stderr.write("bench.py: CUDA graph capture failed (%s); timing eager launches\n" % exc)
            graph = None
    launches_per_k = _ffi.kernel_launches() - launches0
    if graph is not None:
        call("qr_graph_begin_capture", stream)
        for i in range(K):
            step(*fill_ev[i])
        g = C.c_void_p()
        call("qr_graph_end_capture", stream, C.byref(g))
        graph_ev = g
        call("qr_graph_launch", graph_ev, stream)
        call("qr_stream_synchronize", stream)

    def timed_k_steps():
        """exactly K steps between two events on `stream`, bracketed by barrier + synchronize -> (ms, mean fill ms, launches)"""
        e0, e1 = ev(), ev()
        barrier()
        l0 = _ffi.kernel_launches()
        call("qr_event_record", e0, stream)
        if graph is not None:
            call("qr_graph_launch", graph, stream)
        else:
            for i in range(K):
                step()
        call("qr_event_record", e1, stream)
        call("qr_stream_synchronize", stream)
        n_l = launches_per_k if graph is not None else _ffi.kernel_launches() - l0
        barrier()
        # the same K steps once more, instrumented: per-launch duration of the fill kernel
        if graph_ev is not None:
            call("qr_graph_launch", graph_ev, stream)
        else:
            for i in range(K):
                step(*fill_ev[i])
        call("qr_stream_synchronize", stream)
        return elapsed(e0, e1), float(np.mean([elapsed(a, b) for a, b in fill_ev])), n_l

    sampler.start()
    runs = [timed_k_steps() for _ in range(1 + max(0, args.replays))]
    launches = runs[0][2]
    step_ms_runs = [v / K for v in max_over_ranks([r[0] for r in runs])]   # per replay: max over ranks
    fill_ms_runs = max_over_ranks([r[1] for r in runs])
    t_ms = float(np.median(step_ms_runs)) * K
    fill_ms = float(np.median(fill_ms_runs))

    # ---- matrix-free H.v: BASELINE config 4 (TFIM 5x5, n=25), rows sharded over the ranks: extras.c4 carries the
    # verified numbers; `hv` repeats the headline ones
    extras = {}
    hv = None
    if not args.no_hv:
        c4 = run_baseline_config(rig, "C4")
        extras["c4"] = c4
        hv = {"workload": c4["workload"], "n_groups": c4["n_groups"], "ms": c4["hv_ms"], "gbs_compulsory": c4["hv_GBps_compulsory"],
              "gbs_gather_effective": 16.0 * (c4["n_groups"] + 1) * (1 << c4["n_qubits"]) / c4["hv_ms"] / 1e6,
              "allgather_variant_ms": c4.get("hv_allgather_ms"), "max_rel_err": c4["hv_max_rel_err"],
              "p2p_equals_allgather": c4.get("hv_p2p_equals_allgather"), "nvlink_GBps_in": c4.get("hv_nvlink_GBps_in"),
              "note": c4["hv_form"] + "; compulsory bytes = read v once + write y once (32 B/row); verified in extras.c4"}
    sampler.stop()

    # ---- e2e: the public API with host buffers, copies inside the timed region -------------------
    e2e = None
    if not args.no_e2e:
        terms = op.terms()

        def e2e_step():
            o = Q.SparsePauliOp.from_terms(n, terms)              # fresh plan: H2D of the term table + K1
            m = o.to_matrix_rows(lo, hi, device) if world > 1 else o.to_matrix_mode("Cuda")   # K3, fresh buffers
            return m.export()                                     # D2H into pinned host memory
        for _ in range(2):
            out = e2e_step()
        ref_ok = None
        if rank == 0 and world == 1:                              # the exported arrays ARE the reference's, bit for bit
            _, e_data, e_indices, e_indptr = out
            n_o, params = O.make_params(labels, coeffs)
            probe = np.random.default_rng(9).integers(0, dim, 256)
            ref_ok = True
            for r in probe:
                cols, vals = O.make_row(params, int(r))
                a, b = int(e_indptr[int(r)]), int(e_indptr[int(r) + 1])
                ref_ok &= (b - a == G and np.array_equal(e_indices[a:b], cols) and
                           np.array_equal(e_data[a:b].view(np.uint64), vals.view(np.uint64)))
        del out
        barrier()
        reps = max(3, min(K, 10))
        t0 = time.perf_counter()
        for _ in range(reps):
            out = e2e_step()
            del out
        call("qr_stream_synchronize", None)
        t_e2e = max_over_ranks((time.perf_counter() - t0) / reps)
        e2e = {"value": nnz_total / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int(len(terms) * 32),
               "d2h_bytes_per_step": int(_ffi.last_d2h_bytes()),
               "host_bytes_produced_per_step": int(bytes_local), "ms_per_step": t_e2e * 1e3,
               "sampled_rows_equal_oracle": ref_ok,
               "wire": "data as stored + one group id per entry (u8/u16); indptr and the u64 columns are rebuilt by host threads during the DMA",
               "path": "SparsePauliOp.from_terms(terms).to_matrix_mode('Cuda').export(): plan (H2D + K1), K3, D2H of the CSR into pinned host arrays (per rank: its row block)"}
    barrier()

    # ---- extras: the multi-GPU BASELINE config 5 where it fits (N >= 2: 94.5 GB of CSR per GPU at N = 2) ----
    if world > 1 and not args.no_extras and not args.no_c5:
        try:
            extras["c5"] = run_baseline_config(rig, "C5", build_reps=3, hv_reps=10, lanczos_iters=10)
        except _ffi.QrustyCudaError as exc:                       # e.g. out of memory on a smaller part: say so, keep the line
            extras["c5"] = {"error": str(exc)}
            barrier()

    # ---- extras (N=1 only): the other regimes of the build, each timed with events on `stream` --------
    if world == 1 and not args.no_extras:
        peak, _ = peak_hbm()
        timed = rig.timed

        # (1) BASELINE config 3: random 2000-term sum, G = 1500 -> the large-G kernel (whole rows through shared memory) on a row window
        cl, cc = H.random_pauli_sum(24, 2000, 1500, 100, 24)
        cplan = Q.SparsePauliOp([Q.Pauli(l) for l in cl], cc).plan(device)
        crow = 1 << 18                                                  # 9.4 GB per window (2^18-row windows)
        cG = cplan.n_groups
        c_ip, c_ix, c_dt = DeviceBuffer((crow + 1) * 8, device), DeviceBuffer(crow * cG * 8, device), DeviceBuffer(crow * cG * 16, device)
        clo = (cplan.dim // 2)
        ms = timed(lambda: call("qr_build_rows_device", cplan.handle, clo, clo + crow, c_ip.ptr, c_ix.ptr, c_dt.ptr, 0, stream), 5)
        cbytes = crow * cG * 24 + (crow + 1) * 8
        extras["large_g"] = {"workload": "random_T2000_n24 (BASELINE config 3), rows [2^23, 2^23 + 2^18)", "n_groups": cG,
                             "kernel": cplan.fill_kernel, "ms": ms, "nnz_per_s": crow * cG / ms * 1e3,
                             "achieved_GBps": cbytes / ms / 1e6, "frac_of_peak": cbytes / ms / 1e6 / peak}
        del c_ip, c_ix, c_dt, cplan

        # (1b) molecular Hamiltonians from the reference's own fixtures (H8: 16 qubits, T = 5793, G = 981; H12: 24 qubits,
        #      T = 4497, G = 811, one group of 301 terms): the term-rich side of the large-G path, build and H.v
        fx_path = ROOT / "tests" / "golden" / "h_fixtures.json.gz"
        if fx_path.exists():
            import gzip
            fxs = json.load(gzip.open(fx_path))
            for key, log2_rows in (("H8", 16), ("H10", 17), ("H12", 18)):
                if key not in fxs:
                    continue
                fx = fxs[key]
                mop = Q.SparsePauliOp([Q.Pauli(l) for l in fx["labels"]], [complex(a, b) for a, b in fx["coeffs"]])
                mplan = mop.plan(device)
                mrow, mG = min(1 << log2_rows, mplan.dim), mplan.n_groups
                m_ip, m_ix, m_dt = DeviceBuffer((mrow + 1) * 8, device), DeviceBuffer(mrow * mG * 8, device), DeviceBuffer(mrow * mG * 16, device)
                mlo = 0 if mrow == mplan.dim else mplan.dim // 2
                ms = timed(lambda: call("qr_build_rows_device", mplan.handle, mlo, mlo + mrow, m_ip.ptr, m_ix.ptr, m_dt.ptr, 0, stream), 5)
                mbytes = mrow * mG * 24 + (mrow + 1) * 8
                entry = {"workload": "%s fixture (qrusty H_fixtures.py), rows [%d, %d)" % (key, mlo, mlo + mrow), "n_terms": mplan.n_terms,
                         "n_groups": mG, "kernel": mplan.fill_kernel, "ms": ms, "nnz_per_s": mrow * mG / ms * 1e3,
                         "achieved_GBps": mbytes / ms / 1e6, "frac_of_peak": mbytes / ms / 1e6 / peak}
                del m_ip, m_ix, m_dt
                extras["molecular" if key == "H12" else "molecular_" + key] = entry
                if key in ("H10", "H12"):
                    # matrix-free H.v on the whole vector (H10: 2^20, the 20-qubit operator of main10.rs; H12: 2^24): every
                    # group re-evaluates its terms per row (compute-bound) -- apply_fold_kernel, with the gather kernel
                    # (QR_APPLY_FOLD=0, a fresh plan) timed beside it
                    mdim = mplan.dim
                    d_v, d_y = DeviceBuffer(mdim * 16, device), DeviceBuffer(mdim * 16, device)
                    upload_start_vector(rig, H, d_v, 0, mdim)
                    ms = timed(lambda: call("qr_apply_device", mplan.handle, 0, mdim, d_v.ptr, d_y.ptr, stream), 3, warm=1)
                    n_m, mparams = O.make_params(fx["labels"], [complex(a, b) for a, b in fx["coeffs"]])
                    err = verify_hv(rig, H, mparams, 0, mdim, d_y, 64, 77)
                    os.environ["QR_APPLY_FOLD"] = "0"
                    try:
                        gplan = Q.SparsePauliOp([Q.Pauli(l) for l in fx["labels"]], [complex(a, b) for a, b in fx["coeffs"]]).plan(device)
                        ms_gather = timed(lambda: call("qr_apply_device", gplan.handle, 0, mdim, d_v.ptr, d_y.ptr, stream), 3, warm=1)
                        del gplan
                    finally:
                        del os.environ["QR_APPLY_FOLD"]
                    extras["hv_molecular" if key == "H12" else "hv_molecular_" + key] = {
                        "workload": "%s fixture, full vector (2^%d)" % (key, n_m), "n_terms": mplan.n_terms, "n_groups": mG,
                        "kernel": mplan.apply_kernel(), "ms": ms, "gather_kernel_ms": ms_gather,
                        "gbs_compulsory": 32.0 * mdim / ms / 1e6, "term_row_evaluations_per_s": mplan.n_terms * mdim / ms * 1e3,
                        "max_rel_err": err}
                    del d_v, d_y
                del mplan, mop

        # (1b) the reference's own H.v: CSR SpMV (accel.rs:338-370) over the bench operator's built matrix, resident in HBM
        #      (d_ip / d_ix / d_dt hold the last timed build); sampled rows against the oracle's sequential sums, bit for bit
        if n <= 22:
            sv_h = H.lanczos_start_vector(0, dim)
            s_v, s_y = DeviceBuffer(dim * 16, device), DeviceBuffer(rows * 16, device)
            s_v.upload(sv_h)
            ms = timed(lambda: call("qr_spmv_device", rows, d_ip.ptr, d_ix.ptr, d_dt.ptr, s_v.ptr, s_y.ptr, stream), 20)
            y_h = np.empty(rows, np.complex128); s_y.download(y_h)
            _, prm = O.make_params(labels, coeffs)
            r_lo = lo + rows // 2
            ref = O.build_csr(prm, n, r_lo, r_lo + 2048)
            same = bool(np.array_equal(y_h[r_lo - lo:r_lo - lo + 2048].view(np.uint64), O.spmv(*ref, sv_h).view(np.uint64)))
            extras["spmv_csr"] = {"workload": name, "kernel": "spmv_csr_unrolled_kernel<4>", "ms": ms, "nnz_per_s": nnz_local / ms * 1e3,
                                  "GBps_matrix": nnz_local * 24 / ms / 1e6, "frac_of_peak_matrix_bytes": nnz_local * 24 / ms / 1e6 / peak,
                                  "rows_verified": 2048, "rows_equal_oracle_bitwise": same,
                                  "note": "spmat_dot_densevec on the device-resident CSR, thread per row, reference summation order; "
                                          "the matrix-free apply of the same operator is extras-independent (hv)"}
            del s_v, s_y

        # (2) fused drop-zeros build of the bench operator: count_rows + scan + fill_compact
        kept = C.c_uint64()
        z_ip = DeviceBuffer((rows + 1) * 8, device)
        call("qr_build_compact_count", plan.handle, lo, hi, 1e-7, z_ip.ptr, C.byref(kept), stream)
        z_ix, z_dt = DeviceBuffer(max(kept.value * 8, 16), device), DeviceBuffer(max(kept.value * 16, 16), device)

        def drop_zeros():
            k2 = C.c_uint64()
            call("qr_build_compact_count", plan.handle, lo, hi, 1e-7, z_ip.ptr, C.byref(k2), stream)
            call("qr_build_compact_fill", plan.handle, lo, hi, 1e-7, z_ip.ptr, z_ix.ptr, z_dt.ptr, stream)
        ms = timed(drop_zeros, 10)
        extras["drop_zeros"] = {"workload": name, "tolerance": 1e-7, "stored_entries": int(nnz_local), "kept_entries": int(kept.value),
                                "ms": ms, "kept_nnz_per_s": kept.value / ms * 1e3, "entries_evaluated_per_s": 2 * nnz_local / ms * 1e3,
                                "bytes_written": int(kept.value * 24 + (rows + 1) * 8),
                                "note": "count_rows_kernel + 3 scan kernels + fill_compact_kernel; includes one 8-byte D2H of nnz"}
        del z_ip, z_ix, z_dt

        # (3) e2e with ORDINARY (pageable) numpy arrays as the destination -- what a Rust Vec is
        p_ip, p_ix, p_dt = np.empty(rows + 1, np.uint64), np.empty(nnz_local, np.uint64), np.empty(nnz_local, np.complex128)
        p_ix[:] = 0; p_dt[:] = 0; p_ip[:] = 0                            # fault the pages in once, as a reused buffer would be

        def e2e_pageable(flags):
            t0 = time.perf_counter()
            o = Q.SparsePauliOp.from_terms(n, op.terms())
            call("qr_build_host", o.plan(device).handle, lo, hi, p_ip.ctypes.data, p_ix.ctypes.data, p_dt.ctypes.data, flags)
            return time.perf_counter() - t0
        for fl, key in ((0, "staged"), (_ffi.QR_HOST_NO_STAGING, "plain_cudaMemcpy")):
            e2e_pageable(fl)
            t = float(np.median([e2e_pageable(fl) for _ in range(5)]))
            extras.setdefault("e2e_pageable", {})[key] = {"ms": t * 1e3, "nnz_per_s": nnz_local / t}
        extras["e2e_pageable"]["note"] = ("qr_build_host into plain numpy arrays: pinned staging windows + host copy threads vs "
                                          "cudaMemcpy straight into pageable memory")

    if rank == 0:
        peak, peak_src = peak_hbm()
        achieved = bytes_local / (fill_ms * 1e-3) / 1e9
        kernel = plan.fill_kernel
        traffic, traffic_src = traffic_from_profile(name, kernel)
        line = {
            "metric": METRIC, "value": nnz_total / (t_ms * 1e-3 / K), "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": t_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": config_dict(name, n, len(labels), G, world),
            "run": {"launch": ("one CUDA graph holding the K steps, canonicalise -> fill chained by programmatic dependent launch" if graph is not None else "K eager step launches"),
                    "fill_timing": "events around every fill launch of an instrumented replay of the same K steps, right after each timed region",
                    "timed_regions": len(step_ms_runs),
                    "value_is": "median over the timed regions (each exactly K steps, max over ranks)",
                    "ms_per_step_runs": [round(v, 5) for v in step_ms_runs],
                    "ms_per_step_min": min(step_ms_runs), "ms_per_step_max": max(step_ms_runs),
                    "fill_ms_runs": [round(v, 5) for v in fill_ms_runs]},
            "roofline": {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": bytes_local, "kernel_ms": fill_ms,
                         "bytes_per_nnz": 24 + 8.0 / G},
            "gpu_launches": int(launches),
            "clocks": sampler.summary("timed regions + H.v / Lanczos loops"),
        }
        if hv:
            line["hv"] = hv
        if e2e:
            line["e2e"] = e2e
        if extras:
            line["extras"] = extras
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_build_rate(labels, coeffs, budget_s=12.0)
            try:                                                  # fairness line; never allowed to break the bench line
                line["cpu_tuned"] = cpu_tuned_rate(labels, coeffs)
            except Exception as exc:                              # noqa: BLE001
                line["cpu_tuned"] = {"error": repr(exc)}
        print(json.dumps(line), flush=True)
    if rig.dist is not None:
        rig.dist.destroy_process_group()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and world & (world - 1):
        raise SystemExit("bench.py: the number of ranks must be a power of two")
    if args.impl == "reference":
        from oracle import oracle
        if rank == 0:
            oracle.build()
        run_reference(args, rank, world)
        return
    if rank == 0 and not (ROOT / "qrusty_b200" / "lib" / "libqrusty_cuda.so").exists():
        import __graft_entry__
        __graft_entry__.build()
    run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
