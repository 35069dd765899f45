"""qrusty_b200 -- host-side mirror of the pyqrusty interface over the sm_100a library.

Same names, argument meaning and error behaviour as the reference's Python module
(pyqrusty/src/lib.rs, pyqrusty/python/pyqrusty/__init__.py) for the hot path:

    Pauli, SparsePauliOp, SpMat, csr_matrix, spmat_dot_densevec, axpby, axpy, ax, precond, precond2

plus what the CUDA path adds: to_matrix_mode("Cuda") / ("Cuda/<ngpus>"), and the
matrix-free SparsePauliOp.apply(v).  Every matrix is built by the CUDA kernels in
csrc/ through the C ABI of include/qrusty_cuda.h; there is no CPU implementation
here, so the reference's CPU-strategy mode strings ("", "Rowwise", "RowwiseUnsafeChunked/n",
...) are accepted and executed by the same CUDA build -- they all denote the same matrix.

Out of scope: SpMat +/-, a_spmat_p_b_spmat, the CPU build strategies (their entry
points exist and run the CUDA build).
"""
import ctypes as C
import re
import sys

import numpy as np

from . import _ffi
from ._ffi import QrustyCudaError, call, lib
from ._runtime import DeviceBuffer, HostBuffer, PINNED, device_name, pinned_empty, synchronize

__all__ = ["Pauli", "SparsePauliOp", "SpMat", "csr_matrix", "spmat_dot_densevec",
           "axpby", "axpy", "ax", "precond", "precond2", "QrustyCudaError"]

TERM_DTYPE = np.dtype([("z", "<u8"), ("x", "<u8"), ("re", "<f8"), ("im", "<f8")])   # == qr_term

_LABEL_RE = re.compile(r"^([+-]?)1?([ij]?)([IXYZ]+)$")            # lib.rs:127
# AccelMode::try_from (lib.rs:293-331); the two chunked forms are unanchored is_match.
_MODES_EXACT = {"Binary", "Accel", "Rowwise", "RowwiseUnsafe", "Reduce", "Rayon"}
_MODES_RE = [re.compile(r"RowwiseUnsafeChunked/(\d+)"), re.compile(r"RayonChunked/(\d+)")]
_CUDA_RE = re.compile(r"^Cuda(?:/(\d+))?$")


def _rust_f64(x):
    """Rust `{}` of an f64 (util.rs:140-142 complex64_to_string): shortest round-trip,
    never scientific, no trailing '.0'."""
    if x != x:
        return "NaN"
    if x in (float("inf"), float("-inf")):
        return "inf" if x > 0 else "-inf"
    s = repr(float(x))
    if "e" in s or "E" in s:
        from decimal import Decimal
        s = format(Decimal(s), "f")
    return s[:-2] if s.endswith(".0") else s


class Pauli:
    """pyqrusty.Pauli (pyqrusty/src/lib.rs:262-303) over qrusty::Pauli (lib.rs:117-267)."""

    def __init__(self, label):
        m = _LABEL_RE.match(label) if isinstance(label, str) else None
        if not m:
            raise Exception("error: malformed label")                  # lib.rs:129
        sign, imag, body = m.groups()
        self.base_phase = (1 if imag else 0) + (2 if sign == "-" else 0)  # lib.rs:135-141
        self._body = body
        x = z = ny = 0
        for k, ch in enumerate(reversed(body)):                        # qubit k = k-th char from the right
            if ch in "XY":
                x |= 1 << k
            if ch in "YZ":
                z |= 1 << k
            if ch == "Y":
                ny += 1
        self._x, self._z, self._ny = x, z, ny

    def num_qubits(self):
        return len(self._body)

    def label(self):
        return self._body                                               # lib.rs:150-154 (prefix dropped)

    def x_indices(self):
        return self._x

    def z_indices(self):
        return self._z

    def phase(self):
        return (self.base_phase + self._ny) % 4                        # lib.rs:179-181

    def __repr__(self):
        return "Pauli('%s')" % self._body

    def __str__(self):
        return self._body

    def to_matrix(self):
        return SparsePauliOp([self], [1.0 + 0.0j]).to_matrix()


def _parse_mode(mode):
    """-> ("cuda", ngpus).  Unknown strings raise like pyqrusty (pyqrusty/src/lib.rs:406-415)."""
    if mode == "":
        return 1
    m = _CUDA_RE.match(mode)
    if m:
        n = int(m.group(1)) if m.group(1) else 1
        if n < 1:
            raise Exception("to_matrix_mode: unrecognized mode %s" % mode)
        return n
    if mode in _MODES_EXACT or any(r.search(mode) for r in _MODES_RE):
        return 1
    raise Exception("to_matrix_mode: unrecognized mode %s" % mode)


class _Plan:
    """Owns a qr_plan handle."""

    def __init__(self, n_qubits, terms, device, flags=0):
        self.terms = np.ascontiguousarray(terms, dtype=TERM_DTYPE)
        h = C.c_void_p()
        call("qr_plan_create", n_qubits, self.terms.ctypes.data, len(self.terms), device, flags, C.byref(h))
        self.handle = h.value
        info = _ffi.PlanInfo()
        call("qr_plan_info", self.handle, C.byref(info))
        self.n_qubits, self.device = info.n_qubits, info.device
        self.dim, self.n_terms, self.n_groups, self.nnz = info.dim, info.n_terms, info.n_groups, info.nnz
        c = C.c_uint64()
        call("qr_plan_canonical_terms", self.handle, C.byref(c))
        self.n_terms_canonical = c.value

    @property
    def fill_kernel(self):
        """Name of the fill kernel an aligned row window of this plan is built with."""
        return (lib.qr_plan_fill_kernel(self.handle) or b"").decode()

    def apply_kernel(self, row_lo=0, row_hi=None):
        """Name of the kernel a matrix-free H.v on rows [row_lo, row_hi) of this plan runs."""
        name = (lib.qr_plan_apply_kernel(self.handle, row_lo, self.dim if row_hi is None else row_hi) or b"").decode()
        if not name:
            raise _ffi.QrustyCudaError(_ffi.QR_ERR_INVALID, (lib.qr_last_error() or b"").decode())
        return name

    def groups(self):
        x = np.zeros(self.n_groups, np.uint64)
        off = np.zeros(self.n_groups + 1, np.uint32)
        order = np.zeros(self.n_terms, np.uint32)
        call("qr_plan_groups", self.handle, x.ctypes.data, off.ctypes.data, order.ctypes.data)
        return x, off, order

    def __del__(self):
        if getattr(self, "handle", None) and lib is not None:      # lib is None at interpreter exit
            lib.qr_plan_destroy(self.handle)
            self.handle = None


class SparsePauliOp:
    """pyqrusty.SparsePauliOp (pyqrusty/src/lib.rs:305-433) over qrusty::SparsePauliOp
    (lib.rs:345-586)."""

    def __init__(self, paulis, coeffs):
        paulis, coeffs = list(paulis), list(coeffs)
        if len(paulis) != len(coeffs):
            raise Exception("SparsePauliOp::new: paulis and coeffs must have same length")
        if len(paulis) == 0:
            raise Exception("SparsePauliOp::new: at least one pauli must be supplied")
        n = paulis[0].num_qubits()
        if any(p.num_qubits() != n for p in paulis):
            raise Exception("SparsePauliOp::new: all supplied paulis must have the same #qubits")
        self._members = [(p, complex(c)) for p, c in zip(paulis, coeffs)]
        self._terms = None
        self._plans = {}

    @classmethod
    def from_terms(cls, n_qubits, terms):
        """Builds an operator straight from make_params-style tuples (z, x, c') --
        the synthetic-Hamiltonian generators use this; labels are not materialised."""
        self = cls.__new__(cls)
        self._members = None
        self._n_qubits = int(n_qubits)
        self._terms = np.ascontiguousarray(terms, dtype=TERM_DTYPE)
        if len(self._terms) == 0:
            raise Exception("SparsePauliOp::new: at least one pauli must be supplied")
        self._plans = {}
        return self

    # -- container protocol --------------------------------------------------------
    def __len__(self):
        return len(self._terms) if self._members is None else len(self._members)

    def __getitem__(self, idx):
        if self._members is None:
            raise Exception("__getitem__: operator was built from raw terms")
        if isinstance(idx, slice):
            start, stop, step = idx.indices(len(self._members))
            return self._members[start:stop][::step]                   # pyqrusty/src/lib.rs:357-366
        if not (0 <= idx < len(self._members)):
            raise Exception("__getitem__ called on invalid index %d" % idx)
        return self._members[idx]

    def num_qubits(self):
        return self._n_qubits if self._members is None else self._members[0][0].num_qubits()

    def __repr__(self):
        labels = "','".join(p.label() for p, _ in self._members)
        coeffs = ", ".join("%s+%sj" % (_rust_f64(c.real), _rust_f64(c.imag)) for _, c in self._members)
        return "SparsePauliOp('%s', [%s])" % (labels, coeffs)

    __str__ = __repr__

    def __add__(self, other):                                          # lib.rs:588-593
        return SparsePauliOp([p for p, _ in self._members + other._members],
                             [c for _, c in self._members + other._members])

    # -- make_params (accel.rs:141-157) --------------------------------------------
    def terms(self):
        """qr_term[T]: (z, x, c') with c' = coeff * (+i)^base_phase * (-i)^#Y.

        The (-i)^#Y factor is accel.rs:147-153.  For the base phase this follows the
        reference's default to_matrix, base_coeff = (+i)^base_phase (lib.rs:182-190,211);
        the row-wise code uses (-i)^base_phase instead, which differs only for labels with
        an i/j prefix -- parity is defined against to_matrix."""
        if self._terms is None:
            unit = [(1.0, 0.0), (0.0, -1.0), (-1.0, 0.0), (0.0, 1.0)]
            t = np.zeros(len(self._members), TERM_DTYPE)
            for i, (p, c) in enumerate(self._members):
                ur, ui = unit[(p._ny - p.base_phase) % 4]
                t[i] = (p._z, p._x, ur * c.real - ui * c.imag, ur * c.imag + ui * c.real)
            self._terms = t
        return self._terms

    merge_duplicates = False     # opt-in: merge identical (x, z) terms on the GPU (data to 1e-12, not bit-exact)

    def plan(self, device=0):
        if device not in self._plans:
            n = self.num_qubits()
            if n > 32:
                raise QrustyCudaError(_ffi.QR_ERR_UNSUPPORTED, "n_qubits > 32 is not supported")
            flags = _ffi.QR_PLAN_MERGE_DUPLICATES if self.merge_duplicates else 0
            self._plans[device] = _Plan(n, self.terms(), device, flags)
        return self._plans[device]

    # -- matrix build ----------------------------------------------------------------
    def to_matrix(self):
        return self.to_matrix_mode("Cuda")

    def to_matrix_mode(self, mode=""):
        """Mode strings of AccelMode::try_from (lib.rs:293-331) plus "Cuda" and
        "Cuda/<ngpus>" (row blocks over the first <ngpus> devices of this process)."""
        ngpus = _parse_mode(mode)
        dim = 1 << self.num_qubits()
        if ngpus > dim or dim % ngpus:
            raise Exception("to_matrix_mode: %d GPUs do not divide %d rows" % (ngpus, dim))
        shards = []
        for d in range(ngpus):
            plan = self.plan(d)
            lo, hi = dim // ngpus * d, dim // ngpus * (d + 1)
            shards.append(_Shard.build(plan, lo, hi))
        return SpMat._from_shards((dim, dim), shards)

    # the reference's per-strategy entry points (pyqrusty/src/lib.rs:386-404): one matrix, one CUDA build
    def to_matrix_binary(self):
        return self.to_matrix_mode("Cuda")

    def to_matrix_accel(self):
        return self.to_matrix_mode("Cuda")

    def to_matrix_reduce(self):
        return self.to_matrix_mode("Cuda")

    def to_matrix_rayon(self):
        return self.to_matrix_mode("Cuda")

    def to_matrix_rayon_chunked(self, step):
        if int(step) <= 0:
            raise Exception("to_matrix_rayon_chunked: step must be positive")
        return self.to_matrix_mode("Cuda")

    def to_matrix_rows(self, row_lo, row_hi, device=0):
        """Rows [row_lo,row_hi) as a self-contained (row_hi-row_lo) x 2^n CSR shard on `device`
        (local indptr, global column ids) -- what one rank of a row-sharded build owns."""
        plan = self.plan(device)
        if not (0 <= row_lo < row_hi <= plan.dim):
            raise Exception("to_matrix_rows: bad row range")
        sh = _Shard.build(plan, row_lo, row_hi, local=True)
        return SpMat._from_shards((row_hi - row_lo, plan.dim), [sh])

    # -- matrix-free H.v -------------------------------------------------------------
    def apply(self, v, device=0):
        """y = H v without building H (host vectors in, host vector out)."""
        plan = self.plan(device)
        v = np.ascontiguousarray(v, dtype=np.complex128)
        if v.shape != (plan.dim,):
            raise Exception("apply: vector has the wrong length")
        y = np.empty(plan.dim, np.complex128)
        call("qr_apply_host", plan.handle, v.ctypes.data, y.ctypes.data)
        return y

    def diagonal(self, device=0):
        plan = self.plan(device)
        d = DeviceBuffer(plan.dim * 16, device)
        call("qr_diagonal_device", plan.handle, 0, plan.dim, d.ptr, None)
        return d.download(np.empty(plan.dim, np.complex128))


class _Shard:
    """Rows [lo,hi) of the CSR on one device.  Plan-backed shards are LAZY: nothing is built
    until the CSR is needed on the device (SpMV, copy) -- export() streams fill -> PCIe windows
    straight into pinned host arrays (qr_build_host) without ever holding the shard in HBM."""

    def __init__(self, plan, lo, hi, indptr=None, indices=None, data=None, off=None, local=False):
        self.plan, self.lo, self.hi = plan, lo, hi
        self.off = lo if off is None else off        # first row of this shard inside its SpMat
        self.local = local                           # indptr relative to the shard (else global)
        self.device = plan.device if plan is not None else indptr.device
        self.indptr, self.indices, self.data = indptr, indices, data
        self.nnz = (hi - lo) * plan.n_groups if plan is not None else None

    @classmethod
    def build(cls, plan, lo, hi, local=False):
        return cls(plan, lo, hi, off=0 if local else lo, local=local)

    def materialise(self, stream=None):
        """Build the shard in HBM (asynchronous on `stream`)."""
        if self.data is None:
            plan, rows, G = self.plan, self.hi - self.lo, self.plan.n_groups
            self.indptr = DeviceBuffer((rows + 1) * 8, plan.device)
            self.indices = DeviceBuffer(rows * G * 8, plan.device)
            self.data = DeviceBuffer(rows * G * 16, plan.device)
            call("qr_build_rows_device", plan.handle, self.lo, self.hi, self.indptr.ptr, self.indices.ptr,
                 self.data.ptr, _ffi.QR_INDPTR_LOCAL if self.local else _ffi.QR_INDPTR_GLOBAL, stream)
        return self


class SpMat:
    """pyqrusty.SpMat (pyqrusty/src/lib.rs:33-260): a boxed CSR that export() moves out.
    Here the CSR lives in HBM (one shard per GPU) until export() copies it to the host."""

    def __init__(self):
        self._shape = None
        self._shards = None

    @classmethod
    def _from_shards(cls, shape, shards):
        m = cls()
        m._shape, m._shards = shape, shards
        return m

    @staticmethod
    def new_unchecked(shape, data, indices, indptr, device=0):
        """pyqrusty/src/lib.rs:104-116: wraps caller arrays without validation (uploads them)."""
        data = np.ascontiguousarray(data, dtype=np.complex128)
        indices = np.ascontiguousarray(indices, dtype=np.uint64)
        indptr = np.ascontiguousarray(indptr, dtype=np.uint64)
        bufs = []
        for a in (indptr, indices, data):
            b = DeviceBuffer(max(a.nbytes, 16), device)
            b.upload(a)
            bufs.append(b)
        sh = _Shard(None, 0, int(shape[0]), *bufs, local=True)
        sh.nnz = len(data)
        return SpMat._from_shards((int(shape[0]), int(shape[1])), [sh])

    def _live(self, what):
        if self._shards is None:
            raise Exception(what)
        return self._shards

    def shape(self):
        self._live("SpMat.shape(): matrix is already dropped")
        return self._shape

    def nnz(self):
        return sum(s.nnz for s in self._live("cannot get NNZ of an exported sparse matrix"))

    def __repr__(self):
        if self._shards is None:
            return "<already-dropped sparse matrix of type Complex64>"
        return ("<%dx%d sparse matrix of type Complex64\n\twith %d stored elements in Compressed Sparse Row format>"
                % (self._shape[0], self._shape[1], self.nnz()))

    __str__ = __repr__

    def __copy__(self):
        shards = []
        for s in self._live("cannot copy an exported sparse matrix"):
            if s.data is None:                       # lazy: nothing to duplicate yet
                shards.append(_Shard(s.plan, s.lo, s.hi, off=s.off, local=s.local))
                continue
            bufs = []
            for b in (s.indptr, s.indices, s.data):
                nb = DeviceBuffer(b.nbytes, b.device)
                tmp = np.empty(b.nbytes, np.uint8)
                b.download(tmp)
                nb.upload(tmp)
                bufs.append(nb)
            c = _Shard(s.plan, s.lo, s.hi, *bufs, off=s.off, local=s.local)
            c.nnz = s.nnz
            if getattr(s, "compacted", False):
                c.compacted = True
            shards.append(c)
        return SpMat._from_shards(self._shape, shards)

    def diagonal(self):
        """pyqrusty/src/lib.rs:118-125: the diagonal of the STORED matrix.  A shard that has not been built yet (lazy, square
        placement) takes it straight from the plan -- the mask-0 group, no matrix needed; a shard that is resident (and may
        have been scaled or compacted since) is searched on the device."""
        shards = self._live("SpMat.diagonal(): already-exported sparse matrix")
        out = np.zeros(min(self._shape), np.complex128)
        for s in shards:
            rows = s.hi - s.lo
            n = min(rows, max(0, len(out) - s.off))
            if n == 0:
                continue
            call("qr_set_device", s.device)
            d = DeviceBuffer(rows * 16, s.device)
            if s.plan is not None and s.data is None and s.off == s.lo:
                call("qr_diagonal_device", s.plan.handle, s.lo, s.hi, d.ptr, None)
            else:
                s.materialise()
                call("qr_csr_diagonal_device", rows, s.off, s.indptr.ptr, s.indices.ptr, s.data.ptr, d.ptr, None)
            tmp = d.download(np.empty(rows, np.complex128))
            out[s.off:s.off + n] = tmp[:n]
        return out

    def scale(self, factor):
        """In place: every stored value times `factor` (pyqrusty/src/lib.rs:158-164), on the device with the
        `ax` kernel (num-complex's multiply, so bit-identical to the CPU)."""
        if self._shards is None:
            raise Exception("cannot scale already-exported sparse matrix")
        for s in self.to_device()._shards:
            if s.nnz:
                call("qr_set_device", s.device)
                call("qr_ax_device", s.nnz, _c2(factor), s.data.ptr, s.data.ptr, None)
        for s in self._shards:
            call("qr_set_device", s.device)
            synchronize()

    # -- zero elimination (pyqrusty/src/lib.rs:170-183 -> util.rs:144-171) ------------------------
    def _kept(self, tolerance, compact):
        """Per shard: count + scan (and optionally compact) on the device.  -> (kept, new shards).
        A shard that has not been built yet takes the fused path (qr_build_compact_*): its values
        are counted in registers and only the kept entries are ever written; a shard already
        resident in HBM is counted and compacted in place (qr_count_kept / qr_compact_rows)."""
        shards = self._live("cannot %s zeroes of an exported sparse matrix"
                            % ("eliminate" if compact else "count"))
        kept_total, out = 0, []
        for s in shards:
            call("qr_set_device", s.device)
            rows = s.hi - s.lo
            indptr = DeviceBuffer((rows + 1) * 8, s.device)
            kept = C.c_uint64()
            if s.plan is None or getattr(s, "compacted", False):
                # any CSR (new_unchecked, read back from disk, already compacted): driven by the stored indptr
                call("qr_csr_count_kept_device", rows, s.indptr.ptr, s.data.ptr, float(tolerance), indptr.ptr, C.byref(kept), None)
                kept_total += kept.value
                if compact:
                    indices = DeviceBuffer(max(kept.value * 8, 16), s.device)
                    data = DeviceBuffer(max(kept.value * 16, 16), s.device)
                    call("qr_csr_compact_device", rows, s.indptr.ptr, s.indices.ptr, s.data.ptr, float(tolerance), indptr.ptr,
                         indices.ptr, data.ptr, None)
                    c = _Shard(s.plan, s.lo, s.hi, indptr, indices, data, off=s.off, local=True)
                    c.nnz, c.compacted = kept.value, True
                    out.append(c)
                continue
            G = s.plan.n_groups
            fused = s.data is None
            if fused:
                call("qr_build_compact_count", s.plan.handle, s.lo, s.hi, float(tolerance), indptr.ptr, C.byref(kept), None)
            else:
                call("qr_count_kept_device", rows, G, s.data.ptr, float(tolerance), indptr.ptr, C.byref(kept), None)
            kept_total += kept.value
            if compact:
                indices = DeviceBuffer(max(kept.value * 8, 16), s.device)
                data = DeviceBuffer(max(kept.value * 16, 16), s.device)
                if fused:
                    call("qr_build_compact_fill", s.plan.handle, s.lo, s.hi, float(tolerance), indptr.ptr,
                         indices.ptr, data.ptr, None)
                else:
                    call("qr_compact_rows_device", rows, G, s.indices.ptr, s.data.ptr, float(tolerance), indptr.ptr,
                         indices.ptr, data.ptr, None)
                c = _Shard(s.plan, s.lo, s.hi, indptr, indices, data, off=s.off, local=True)
                c.nnz, c.compacted = kept.value, True
                out.append(c)
        return kept_total, out

    def count_zeros(self, tolerance=1e-7):
        """Entries with norm <= tolerance (util.rs:144-152)."""
        self._live("cannot count zeroes of an exported sparse matrix")
        kept, _ = self._kept(tolerance, compact=False)
        return self.nnz() - kept

    def eliminate_zeros(self, tolerance=1e-7):
        """New SpMat holding only entries with norm > tolerance (util.rs:154-171)."""
        _, shards = self._kept(tolerance, compact=True)
        for s in shards:
            call("qr_set_device", s.device)
            synchronize()
        return SpMat._from_shards(self._shape, shards)

    # -- on-disk formats (rawio.rs:128-179; pyqrusty/src/lib.rs:216-259) -----------------------------
    def rawio_write(self, path):
        """qrusty::rawio::write (rawio.rs:128-148): "MI" mark, storage tag, shape, then the three
        length-prefixed arrays.  A single plan-backed shard that was never built is streamed from the
        GPU in row windows (qr_write_rawio); anything else is exported to the host first."""
        shards = self._live("cannot write an already-destroyed sparse matrix")
        if len(shards) == 1 and shards[0].data is None and shards[0].plan is not None:
            s = shards[0]
            call("qr_write_rawio", s.plan.handle, s.lo, s.hi, str(path).encode())
            return
        shape, data, indices, indptr = self.__copy__().export()
        with open(path, "wb") as f:
            f.write(b"MI" if sys.byteorder == "little" else b"IM")
            np.array([0, shape[0], shape[1], len(indptr)], np.uint64).tofile(f)
            indptr.tofile(f)
            np.array([len(indices)], np.uint64).tofile(f); indices.tofile(f)
            np.array([len(data)], np.uint64).tofile(f); data.tofile(f)

    @staticmethod
    def rawio_read(path, device=0):
        """qrusty::rawio::read (rawio.rs:150-179), byte-swapping when the mark says so."""
        with open(path, "rb") as f:
            mark = f.read(2)
            native = b"MI" if sys.byteorder == "little" else b"IM"
            if mark not in (b"MI", b"IM"):
                raise Exception("rawio_read: bad endian mark")
            u8 = np.dtype(np.uint64) if mark == native else np.dtype(np.uint64).newbyteorder()
            c16 = np.dtype(np.complex128) if mark == native else np.dtype(np.complex128).newbyteorder()
            storage, rows, cols, n_ptr = (int(v) for v in np.fromfile(f, u8, 4))
            if storage != 0:
                raise Exception("rawio_read: only CSR storage is supported")      # rawio.rs:153-157 also has CSC
            indptr = np.fromfile(f, u8, n_ptr).astype(np.uint64)
            indices = np.fromfile(f, u8, int(np.fromfile(f, u8, 1)[0])).astype(np.uint64)
            data = np.fromfile(f, c16, int(np.fromfile(f, u8, 1)[0])).astype(np.complex128)
        return SpMat.new_unchecked((rows, cols), data, indices, indptr, device)

    def matrixmarket_write(self, path):
        """pyqrusty/src/lib.rs:216-230 (sprs::io::write_matrix_market): coordinate / complex / general,
        one-based "row col re im" lines in storage order.  Text formatting is host work."""
        shape, data, indices, indptr = self.__copy__().export()
        rows = np.repeat(np.arange(shape[0], dtype=np.uint64), np.diff(indptr.astype(np.int64)))
        with open(path, "w") as f:
            f.write("%%MatrixMarket matrix coordinate complex general\n")
            f.write("% written by qrusty_b200\n")
            f.write("%d %d %d\n" % (shape[0], shape[1], len(data)))
            for r, c, v in zip(rows.tolist(), indices.tolist(), data.tolist()):
                f.write("%d %d %r %r\n" % (r + 1, c + 1, v.real, v.imag))

    @staticmethod
    def matrixmarket_read(path, device=0):
        """pyqrusty/src/lib.rs:232-247: triplets -> CSR (duplicates summed, as TriMat::to_csr)."""
        import scipy.io
        import scipy.sparse
        m = scipy.sparse.csr_matrix(scipy.io.mmread(str(path)), dtype=np.complex128)
        m.sum_duplicates(); m.sort_indices()
        return SpMat.new_unchecked(m.shape, m.data, m.indices.astype(np.uint64), m.indptr.astype(np.uint64), device)

    def to_device(self):
        """Materialise every shard in HBM (needed for SpMV on the stored matrix); returns self."""
        shards = self._live("cannot use an exported sparse matrix")
        for s in shards:
            s.materialise()
        for s in shards:
            call("qr_set_device", s.device)
            synchronize()
        return self

    def export(self):
        """-> ((rows, cols), data, indices, indptr), then the matrix is gone
        (pyqrusty/src/lib.rs:190-214).  Arrays live in pinned host memory.  Shards not yet
        built are produced by qr_build_host: row windows filled on the GPU and copied over PCIe
        on two streams, never resident in HBM as a whole."""
        shards = self._live("cannot export from an already-exported sparse matrix")
        nnz = self.nnz()
        data = pinned_empty(nnz, np.complex128)
        indices = pinned_empty(nnz, np.uint64)
        indptr = pinned_empty(self._shape[0] + 1, np.uint64)
        off = 0
        streams, lazy = [], []
        for s in shards:
            rows = s.hi - s.lo
            if s.data is None:
                lazy.append((s, off))
            else:
                call("qr_set_device", s.device)
                st = C.c_void_p()
                call("qr_stream_create", C.byref(st))
                streams.append((s.device, st))
                s.data.download(data[off:off + s.nnz], stream=st)
                s.indices.download(indices[off:off + s.nnz], stream=st)
                s.indptr.download(indptr[s.off:s.off + rows + 1], stream=st)
            off += s.nnz

        def build_host(item):
            s, o = item
            call("qr_build_host", s.plan.handle, s.lo, s.hi, indptr[s.off:].ctypes.data,
                 indices[o:].ctypes.data, data[o:].ctypes.data,
                 _ffi.QR_INDPTR_LOCAL if s.local else _ffi.QR_INDPTR_GLOBAL)
        if len(lazy) > 1:
            # one host thread per GPU: every shard streams over its own PCIe link at the same time
            # (ctypes drops the GIL for the duration of the call).  Shards overlap by one indptr entry
            # (a shard's last = the next one's first, the same value), so concurrent writes agree.
            from concurrent.futures import ThreadPoolExecutor
            with ThreadPoolExecutor(len(lazy)) as pool:
                list(pool.map(build_host, lazy))
        else:
            for item in lazy:
                build_host(item)
        for dev, st in streams:
            call("qr_set_device", dev)
            call("qr_stream_synchronize", st)
            call("qr_stream_destroy", st)
        base = 0
        for s in shards:                             # compacted shards carry shard-local indptr
            if getattr(s, "compacted", False) and base:
                indptr[s.off + 1:s.off + (s.hi - s.lo) + 1] += np.uint64(base)
                indptr[s.off] = base                 # overwritten by this shard's local 0
            base += s.nnz
        shape = self._shape
        self._shards = None
        return shape, data, indices, indptr


def csr_matrix(m):
    """pyqrusty/python/pyqrusty/__init__.py:14-17."""
    (shape, data, indices, indptr) = m.export()
    from scipy.sparse import csr_matrix as _csr
    return _csr((data, indices, indptr), shape=shape, dtype=complex)


def spmat_dot_densevec(spmat, x):
    """pyqrusty spmat_dot_densevec (pyqrusty/src/lib.rs:476-491) -> accel.rs:338-370:
    CSR SpMV over the device-resident matrix, sequential per row in stored order."""
    shards = spmat.to_device()._live("cannot multiply with an exported sparse matrix")
    x = np.ascontiguousarray(x, dtype=np.complex128)
    if x.shape != (spmat._shape[1],):
        raise Exception("spmat_dot_densevec: vector of length %d against a matrix with %d columns" % (x.size, spmat._shape[1]))
    y = np.empty(spmat._shape[0], np.complex128)
    for s in shards:
        rows = s.hi - s.lo
        dv = DeviceBuffer(max(x.nbytes, 16), s.device)
        dv.upload(x)
        dy = DeviceBuffer(max(rows * 16, 16), s.device)
        call("qr_spmv_device", rows, s.indptr.ptr, s.indices.ptr, s.data.ptr, dv.ptr, dy.ptr, None)
        dy.download(y[s.off:s.off + rows])
    return y


def _c2(a):
    a = complex(a)
    return (C.c_double * 2)(a.real, a.imag)


def _vec_op(name, n, scalars, inputs):
    bufs = []
    for v in inputs:
        b = DeviceBuffer(max(n * 16, 16))
        b.upload(v)
        bufs.append(b)
    z = DeviceBuffer(max(n * 16, 16))
    args = [n]
    it_s, it_b = iter(scalars), iter(bufs)
    for kind in {"qr_axpby_device": "svsv", "qr_axpy_device": "svv", "qr_ax_device": "sv"}[name]:
        args.append(next(it_s) if kind == "s" else next(it_b).ptr)
    call(name, *args, z.ptr, None)
    return z.download(np.empty(n, np.complex128))


def _vec(x):
    return np.ascontiguousarray(x, dtype=np.complex128)


def axpby(a, x, b, y):
    """z = a*x + b*y (accel.rs:374-379)."""
    x, y = _vec(x), _vec(y)
    if x.shape != y.shape:
        raise Exception("axpby: x and y differ in length")
    return _vec_op("qr_axpby_device", len(x), [_c2(a), _c2(b)], [x, y])


def axpy(a, x, y):
    """z = a*x + y (accel.rs:381-386)."""
    x, y = _vec(x), _vec(y)
    if x.shape != y.shape:
        raise Exception("axpy: x and y differ in length")
    return _vec_op("qr_axpy_device", len(x), [_c2(a)], [x, y])


def ax(a, x):
    """z = a*x (accel.rs:388-393)."""
    x = _vec(x)
    return _vec_op("qr_ax_device", len(x), [_c2(a)], [x])


def precond2(diag, dx, e, tol):
    """dx / reg(diag - e, tol) (pyqrusty/src/lib.rs:457-468, 579-591)."""
    diag, dx = _vec(diag), _vec(dx)
    if len(diag) != len(dx):
        raise ValueError("precond2: diag and dx differ in length")
    n = len(dx)
    bd, bx, bz = DeviceBuffer(max(n * 16, 16)), DeviceBuffer(max(n * 16, 16)), DeviceBuffer(max(n * 16, 16))
    bd.upload(diag); bx.upload(dx)
    call("qr_precond2_device", n, bd.ptr, bx.ptr, _c2(e), float(tol), bz.ptr, None)
    return bz.download(np.empty(n, np.complex128))


def precond(spmat, dx, e, tol):
    """dx / reg(spmat.diagonal() - e, tol) (pyqrusty/src/lib.rs:442-455, 564-577)."""
    if spmat._shards is None:
        raise Exception("cannot call precond with an exported sparse matrix")
    return precond2(spmat.diagonal(), dx, e, tol)
