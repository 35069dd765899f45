"""world_size-2 gloo test (CPU) of the host-side logic of the row-sharded path: row blocks
partition the matrix, shards built independently concatenate to the full CSR (checked with the
oracle, since there is no GPU here), the unique-id exchange and the max-over-ranks reduction work."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    from oracle import oracle as O
    from qrusty_b200 import dist as qd, hamiltonians as H
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        uid = qd.exchange_unique_id(dist, lambda: bytes(range(128)))
        assert uid == bytes(range(128))
        assert qd.max_over_ranks(dist, 1.0 + rank) == float(world)
        labels, coeffs = H.xxz_chain(10, 1.0, 0.7)
        n, params = O.make_params(labels, coeffs)
        G = len(np.unique(params["x"]))
        lo, hi = qd.row_block(rank, world, 1 << n)
        indptr, indices, data = O.build_csr(params, n, lo, hi, step=100, n_threads=1, groups=G)
        np.savez(Path(out_dir) / f"shard{rank}.npz", lo=lo, hi=hi, indptr=indptr, indices=indices, data=data)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_row_sharding_world2(tmp_path):
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, str(ROOT))
    from oracle import oracle as O
    from qrusty_b200 import hamiltonians as H
    labels, coeffs = H.xxz_chain(10, 1.0, 0.7)
    n, params = O.make_params(labels, coeffs)
    full = O.build_csr(params, n)
    shards = [np.load(tmp_path / f"shard{r}.npz") for r in range(world)]
    assert [int(s["lo"]) for s in shards] == [0, 512] and [int(s["hi"]) for s in shards] == [512, 1024]
    assert np.array_equal(np.concatenate([s["indices"] for s in shards]), full[1])
    assert np.array_equal(np.concatenate([s["data"] for s in shards]).view(np.uint64), full[2].view(np.uint64))
    G = len(full[1]) >> n
    # local indptr of shard p + p*rows*G == the global indptr slice
    for p, s in enumerate(shards):
        assert np.array_equal(s["indptr"] + np.uint64(p * 512 * G), full[0][p * 512:(p + 1) * 512 + 1])


def test_row_block_properties():
    sys.path.insert(0, str(ROOT))
    from qrusty_b200 import dist as qd
    dim = 1 << 12
    for world in (1, 2, 4, 8):
        blocks = [qd.row_block(r, world, dim) for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == dim
        assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
        assert all(qd.owner_of_row(lo, world, dim) == r and qd.owner_of_row(hi - 1, world, dim) == r
                   for r, (lo, hi) in enumerate(blocks))
    with pytest.raises(ValueError):
        qd.row_block(0, 3, dim)
    with pytest.raises(ValueError):
        qd.row_block(2, 2, dim)
