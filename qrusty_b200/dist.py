"""One-process-per-GPU plumbing.  The CSR build shards by contiguous row blocks and needs no
collective; the matrix-free H.v all-gathers the row-sharded vector with NCCL (qr_comm_*, inside
the C library).  The package imports no framework: every function takes the caller's rendezvous
object `dist` -- anything with get_rank(), get_world_size(), broadcast_object_list(),
all_gather_object() and barrier() (torch.distributed with any backend, "gloo" in the CPU tests, or a
ten-line MPI shim) -- and uses it only to carry the NCCL unique id, the CUDA-IPC handles and a few
Python floats.
"""
import ctypes as C


def row_block(rank, world, dim):
    """Rows [lo,hi) owned by `rank`: contiguous blocks = fixed top log2(world) row bits
    (SURVEY.md 8(e)).  world must divide dim."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("row_block: bad rank/world")
    if dim % world:
        raise ValueError("row_block: %d ranks do not divide %d rows" % (world, dim))
    rows = dim // world
    return rank * rows, (rank + 1) * rows


def owner_of_row(row, world, dim):
    return row // (dim // world)


def exchange_unique_id(dist, make_id, src=0):
    """Rank `src` creates the 128-byte NCCL unique id (make_id() -> bytes); everyone gets it."""
    box = [make_id() if dist.get_rank() == src else None]
    dist.broadcast_object_list(box, src=src)
    if not isinstance(box[0], (bytes, bytearray)) or len(box[0]) != 128:
        raise RuntimeError("exchange_unique_id: malformed id")
    return bytes(box[0])


def max_over_ranks(dist, value, device=None):
    """max of one Python float over the ranks (timings: device-timed on every rank, reduced here)."""
    vals = [None] * dist.get_world_size()
    dist.all_gather_object(vals, float(value))
    return max(vals)


def create_comm(dist, device):
    """-> qr_comm handle (ctypes void*) spanning the default process group."""
    from . import _ffi

    def make_id():
        buf = C.create_string_buffer(_ffi.QR_UNIQUE_ID_BYTES)
        _ffi.call("qr_comm_unique_id", buf)
        return buf.raw

    uid = exchange_unique_id(dist, make_id)
    comm = C.c_void_p()
    _ffi.call("qr_comm_create", uid, dist.get_world_size(), dist.get_rank(), device, C.byref(comm))
    return comm


def share_shards(dist, local_ptr):
    """CUDA-IPC exchange of one device buffer per rank (e.g. the v shards of a row-sharded H.v):
    returns (pointers[world], opened) where pointers[rank] is local_ptr and the others are peer
    mappings usable in kernels; pass `opened` to close_shards when done."""
    from . import _ffi
    buf = C.create_string_buffer(_ffi.QR_IPC_HANDLE_BYTES)
    _ffi.call("qr_ipc_get_handle", local_ptr, buf)
    handles = [None] * dist.get_world_size()
    dist.all_gather_object(handles, buf.raw)
    ptrs, opened = [], []
    for r, h in enumerate(handles):
        if r == dist.get_rank():
            ptrs.append(local_ptr)
        else:
            p = C.c_void_p()
            _ffi.call("qr_ipc_open_handle", h, C.byref(p))
            ptrs.append(p.value)
            opened.append(p.value)
    return ptrs, opened


def close_shards(opened):
    from . import _ffi
    for p in opened:
        _ffi.call("qr_ipc_close_handle", p)


def pointer_array(ptrs):
    return (C.c_void_p * len(ptrs))(*ptrs)
