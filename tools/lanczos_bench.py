#!/usr/bin/env python
"""BASELINE config 4's loop: 50 Lanczos iterations on the TFIM 5x5 lattice (n=25), vectors
device-resident and row-sharded over the ranks.  Reports H.v time and full-iteration time.
  python tools/lanczos_bench.py [C4|xxz20|...]            (1 GPU)
  torchrun --nproc-per-node N tools/lanczos_bench.py C4   (N GPUs: fused peer-memory H.v and all-gather H.v)
"""
import json, os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tools"))
import numpy as np
import qrusty_b200 as Q
from qrusty_b200 import lanczos as L, dist as qd
from qrusty_b200._ffi import call
from fill_sweep import get_workload

cfg = sys.argv[1] if len(sys.argv) > 1 else "C4"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 50
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dist = comm = None
if world > 1:
    import torch, torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    call("qr_set_device", local)
    comm = qd.create_comm(dist, local)
labels, coeffs = get_workload(cfg)
op = Q.SparsePauliOp([Q.Pauli(l) for l in labels], coeffs)
out = {"workload": cfg, "n_gpus": world, "n_qubits": len(labels[0]), "iterations": iters}
L.lanczos(op, 3, device=local, dist=dist, comm=comm)                       # warm-up (diag cache, IPC)
res = L.lanczos(op, iters, device=local, dist=dist, comm=comm, fused_p2p=True)
out.update(hv_ms=res["hv_ms"], iter_ms=res["iter_ms"], ritz_min=float(L.ritz_values(res["alphas"], res["betas"])[0]),
           alpha0=float(res["alphas"][0]), beta0=float(res["betas"][0]))
if world > 1:
    res2 = L.lanczos(op, iters, device=local, dist=dist, comm=comm, fused_p2p=False)
    out.update(allgather_hv_ms=res2["hv_ms"], allgather_iter_ms=res2["iter_ms"],
               max_alpha_diff=float(np.abs(res["alphas"] - res2["alphas"]).max()))
    call("qr_comm_destroy", comm)
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    dist.destroy_process_group()
