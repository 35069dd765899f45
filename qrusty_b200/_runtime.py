"""Device / pinned-host memory owned through the C ABI (no torch, no cupy)."""
import ctypes as C
import threading

import numpy as np

from . import _ffi
from ._ffi import call, lib


class DeviceBuffer:
    """cudaMalloc'ed block on one device, freed with the object."""

    def __init__(self, nbytes, device=0):
        self.nbytes = int(nbytes)
        self.device = device
        call("qr_set_device", device)
        p = C.c_void_p()
        call("qr_malloc_device", C.byref(p), self.nbytes)
        self.ptr = p.value

    def free(self):
        if getattr(self, "ptr", None) and lib is not None:
            lib.qr_set_device(self.device)
            lib.qr_free_device(self.ptr)
            self.ptr = None

    __del__ = free

    def upload(self, arr, stream=None):
        arr = np.ascontiguousarray(arr)
        assert arr.nbytes <= self.nbytes
        call("qr_set_device", self.device)
        call("qr_memcpy_h2d", self.ptr, arr.ctypes.data, arr.nbytes, stream)

    def download(self, out, nbytes=None, offset=0, stream=None):
        """Copy into a C-contiguous numpy array or a HostBuffer-backed array."""
        nbytes = out.nbytes if nbytes is None else nbytes
        call("qr_set_device", self.device)
        call("qr_memcpy_d2h", out.ctypes.data, self.ptr + offset, nbytes, stream)
        return out


class _PinnedPool:
    """Page-locked host blocks are slow to allocate (~0.2 ms/MB), so exported CSR
    arrays recycle them by size.  A block returns to the pool when the numpy array
    that wraps it is garbage collected."""

    def __init__(self, max_cached_bytes=8 << 30):
        self.free = {}
        self.cached = 0
        self.max_cached = max_cached_bytes
        self.lock = threading.Lock()

    def take(self, nbytes):
        with self.lock:
            lst = self.free.get(nbytes)
            if lst:
                self.cached -= nbytes
                return lst.pop()
        p = C.c_void_p()
        call("qr_malloc_host", C.byref(p), nbytes)
        return p.value

    def give(self, ptr, nbytes):
        with self.lock:
            if self.cached + nbytes <= self.max_cached:
                self.free.setdefault(nbytes, []).append(ptr)
                self.cached += nbytes
                return
        lib.qr_free_host(ptr)

    def clear(self):
        with self.lock:
            for lst in self.free.values():
                for p in lst:
                    lib.qr_free_host(p)
            self.free.clear()
            self.cached = 0


PINNED = _PinnedPool()


class HostBuffer:
    """A pinned host block exposed to numpy through __array_interface__; the ndarray
    made by np.asarray(buf) keeps `buf` alive, and `buf` hands the block back to the
    pool when it dies -- the moral equivalent of the reference moving its Vecs into
    numpy on export (pyqrusty/src/lib.rs:199-209)."""

    def __init__(self, count, dtype):
        self.dtype = np.dtype(dtype)
        self.count = int(count)
        self.nbytes = max(16, self.count * self.dtype.itemsize)
        self.ptr = PINNED.take(self.nbytes)
        self.__array_interface__ = {
            "version": 3, "shape": (self.count,), "typestr": self.dtype.str,
            "data": (self.ptr, False),
        }

    def array(self):
        return np.asarray(self)

    def __del__(self):
        if getattr(self, "ptr", None) and PINNED is not None and lib is not None:
            PINNED.give(self.ptr, self.nbytes)
            self.ptr = None


def pinned_empty(count, dtype):
    return HostBuffer(count, dtype).array()


def synchronize(stream=None):
    call("qr_stream_synchronize", stream)


def device_name(device=0):
    buf = C.create_string_buffer(128)
    call("qr_device_name", device, buf, 128)
    return buf.value.decode()
