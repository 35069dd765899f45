"""Lanczos-style iteration on device-resident, row-sharded vectors -- the loop BASELINE config 4 is
defined by: w = H v; alpha = <v,w>; w -= alpha v + beta v_prev; beta = ||w||.

Everything stays in HBM.  Default (device_scalars=True): two passes per iteration and no host round trip --
  1. y = H u and <u, y> in ONE kernel (qr_apply_dot_device, or the fused peer-memory qr_apply_p2p_dot when
     sharded: the apply kernels fold the Rayleigh quotient per CTA in their epilogue);
  2. w = (H u)/b - (alpha/b) u - (b/b') u_prev fused with ||w||^2 (qr_lanczos_update_dev), the coefficients read
     from a device-resident state that two one-thread kernels maintain (qr_lanczos_coef_device); the vectors stay
     unnormalised (u_{k+1} = w_k), so no pass is spent on a rescale.
Sharded runs all-reduce the two scalars in place on the device (qr_allreduce_sum_f64, stream-ordered).  The host
reads alphas and betas once, after the last iteration.
device_scalars=False is the round-1 loop: apply, <v,w> (qr_dotc_device), update fused with the norm
(qr_lanczos_update_device), rescale (qr_ax_device) -- four passes and two host reads per iteration.

The reference leaves this loop to scipy/PRIMME calling spmat_dot_densevec + axpby/axpy/ax
(pyqrusty/sandbox/test1.py:60-87, qrusty/src/accel.rs:338-393).
"""
import ctypes as C
import math
import time

import numpy as np

from . import _ffi, hamiltonians
from ._ffi import call
from ._runtime import DeviceBuffer


def _c2(a):
    a = complex(a)
    return (C.c_double * 2)(a.real, a.imag)


def lanczos(op, n_iter=50, device=0, seed=25, dist=None, comm=None, fused_p2p=True, v0=None, device_scalars=True):
    """-> dict(alphas, betas, hv_ms, iter_ms).  `dist`/`comm` given: this rank owns rows
    dist.row_block(rank, world, dim) and `comm` is a qr_comm (dist.create_comm)."""
    from . import dist as qd
    plan = op.plan(device)
    dim = plan.dim
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist is not None else (0, 1)
    lo, hi = qd.row_block(rank, world, dim)
    rows = hi - lo
    call("qr_set_device", device)
    stream = C.c_void_p(); call("qr_stream_create", C.byref(stream))
    # three rotating shard buffers: v_prev, v, w.  With the fused path peers read v in place, so
    # all three are shared once and rotate together on every rank.
    bufs = [DeviceBuffer(rows * 16, device) for _ in range(3)]
    d_full = DeviceBuffer(dim * 16, device) if (world > 1 and not fused_p2p) else None
    d_s = DeviceBuffer(32, device)                         # [<v,w> (2 doubles), ||w||^2, pad]
    scal = np.zeros(4, np.float64)
    shared = None
    if world > 1 and fused_p2p:
        shared = [qd.share_shards(dist, b.ptr) for b in bufs]

    def allreduce(ptr, count):
        if world > 1:
            call("qr_allreduce_sum_f64", comm, ptr, count, stream)

    def read_scalars():
        call("qr_memcpy_d2h", scal.ctypes.data, d_s.ptr, 32, stream)
        call("qr_stream_synchronize", stream)

    # start vector v0[i] = u1 + i u2 (regenerable anywhere), normalised
    chunk = 1 << 22
    for c0 in range(lo, hi, chunk):
        v = hamiltonians.lanczos_start_vector(c0, min(hi, c0 + chunk), seed) if v0 is None else np.ascontiguousarray(v0[c0:min(hi, c0 + chunk)], np.complex128)
        call("qr_memcpy_h2d", bufs[1].ptr + (c0 - lo) * 16, v.ctypes.data, v.nbytes, None)
    call("qr_dotc_device", rows, bufs[1].ptr, bufs[1].ptr, d_s.ptr, stream)
    allreduce(d_s.ptr, 2)
    read_scalars()
    call("qr_ax_device", rows, _c2(1.0 / math.sqrt(scal[0])), bufs[1].ptr, bufs[1].ptr, stream)

    def ev():
        e = C.c_void_p(); call("qr_event_create", C.byref(e)); return e

    if device_scalars and (world == 1 or fused_p2p):
        res = _lanczos_device_scalars(plan, n_iter, lo, hi, rows, bufs, shared, world, dist, comm, stream, allreduce, ev, qd)
        if shared:
            for _, opened in shared:
                qd.close_shards(opened)
        call("qr_stream_destroy", stream)
        return res
    e0, e1 = ev(), ev()
    alphas, betas, hv_ms = [], [], 0.0
    beta = 0.0
    i_prev, i_v, i_w = 0, 1, 2
    if dist is not None:
        dist.barrier()
    call("qr_stream_synchronize", stream)
    t_start = time.perf_counter()
    for it in range(n_iter):
        call("qr_event_record", e0, stream)
        if world == 1:
            call("qr_apply_device", plan.handle, lo, hi, bufs[i_v].ptr, bufs[i_w].ptr, stream)
        elif fused_p2p:
            call("qr_apply_p2p", plan.handle, comm, qd.pointer_array(shared[i_v][0]), bufs[i_w].ptr, stream)
        else:
            call("qr_apply_distributed", plan.handle, comm, bufs[i_v].ptr, d_full.ptr, bufs[i_w].ptr, stream)
        call("qr_event_record", e1, stream)
        call("qr_dotc_device", rows, bufs[i_v].ptr, bufs[i_w].ptr, d_s.ptr, stream)
        allreduce(d_s.ptr, 2)
        read_scalars()
        ms = C.c_float(); call("qr_event_elapsed_ms", e0, e1, C.byref(ms)); hv_ms += ms.value
        alpha = scal[0]                                    # H Hermitian: <v,Hv> is real
        call("qr_lanczos_update_device", rows, _c2(alpha), _c2(beta), bufs[i_w].ptr, bufs[i_v].ptr,
             bufs[i_prev].ptr if it > 0 else None, bufs[i_w].ptr, d_s.ptr + 16, stream)
        allreduce(d_s.ptr + 16, 1)
        read_scalars()
        beta = math.sqrt(scal[2])
        alphas.append(alpha); betas.append(beta)
        if beta == 0.0:
            break
        call("qr_ax_device", rows, _c2(1.0 / beta), bufs[i_w].ptr, bufs[i_w].ptr, stream)
        i_prev, i_v, i_w = i_v, i_w, i_prev
    call("qr_stream_synchronize", stream)
    total_ms = (time.perf_counter() - t_start) * 1e3
    if dist is not None:
        dist.barrier()
    if shared:
        for _, opened in shared:
            qd.close_shards(opened)
    call("qr_stream_destroy", stream)
    n_done = len(alphas)
    return {"alphas": np.array(alphas), "betas": np.array(betas), "iterations": n_done,
            "hv_ms": hv_ms / n_done, "iter_ms": total_ms / n_done}


def _lanczos_device_scalars(plan, K, lo, hi, rows, bufs, shared, world, dist, comm, stream, allreduce, ev, qd):
    """The two-pass loop with the scalars on the device (module docstring).  bufs[1] holds the normalised start vector."""
    head = 16
    state = np.zeros(head + 2 * K, np.float64)
    state[5] = 1.0                                          # ||u_0||
    d_state = DeviceBuffer(state.nbytes, plan.device)
    d_state.upload(state)
    events = [(ev(), ev()) for _ in range(K)]
    i_prev, i_u, i_y = 0, 1, 2
    if dist is not None:
        dist.barrier()
    call("qr_stream_synchronize", stream)
    t_start = time.perf_counter()
    for k in range(K):
        e0, e1 = events[k]
        call("qr_event_record", e0, stream)
        if world == 1:
            call("qr_apply_dot_device", plan.handle, lo, hi, bufs[i_u].ptr, bufs[i_y].ptr, d_state.ptr, stream)
        else:
            call("qr_apply_p2p_dot", plan.handle, comm, qd.pointer_array(shared[i_u][0]), bufs[i_y].ptr, d_state.ptr, stream)
        call("qr_event_record", e1, stream)
        allreduce(d_state.ptr, 2)
        call("qr_lanczos_coef_device", d_state.ptr, k, K, 0, stream)
        # w overwrites y in place (element i is read before it is written by the same thread)
        call("qr_lanczos_update_dev", rows, d_state.ptr, bufs[i_y].ptr, bufs[i_u].ptr, bufs[i_prev].ptr if k > 0 else None,
             bufs[i_y].ptr, stream)
        allreduce(d_state.ptr + 16, 1)
        call("qr_lanczos_coef_device", d_state.ptr, k, K, 1, stream)
        i_prev, i_u, i_y = i_u, i_y, i_prev
    call("qr_stream_synchronize", stream)
    total_ms = (time.perf_counter() - t_start) * 1e3
    if dist is not None:
        dist.barrier()
    d_state.download(state)
    hv_ms = 0.0
    for e0, e1 in events:
        ms = C.c_float(); call("qr_event_elapsed_ms", e0, e1, C.byref(ms)); hv_ms += ms.value
    alphas, betas = state[head:head + K].copy(), state[head + K:head + 2 * K].copy()
    n_done = K
    dead = np.flatnonzero(~(betas > 0.0))                   # breakdown (beta = 0) or what follows it (nan): cut there
    if len(dead):
        n_done = int(dead[0]) + 1
        alphas, betas = alphas[:n_done], betas[:n_done]
    return {"alphas": alphas, "betas": betas, "iterations": n_done, "hv_ms": hv_ms / K, "iter_ms": total_ms / K,
            "passes_per_iteration": 2, "host_reads_per_iteration": 0}


def ritz_values(alphas, betas):
    """Eigenvalues of the Lanczos tridiagonal (host, k x k)."""
    k = len(alphas)
    t = np.diag(alphas) + np.diag(betas[:k - 1], 1) + np.diag(betas[:k - 1], -1)
    return np.linalg.eigvalsh(t)
