"""ctypes front end of the CPU oracle (oracle/qrusty_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py -- never by qrusty_b200.

Parity status: pinned (see the header of qrusty_oracle.c and tests/test_oracle.py).
"""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "_build" / "libqrusty_oracle.so"

PARAM_DTYPE = np.dtype([("z", "<u8"), ("x", "<u8"), ("re", "<f8"), ("im", "<f8")])


def build(force=False):
    """Compile the oracle with gcc (oracle/Makefile).  Building the checker is not using it."""
    src = _HERE / "qrusty_oracle.c"
    if force or not _SO.exists() or _SO.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE)], check=True, capture_output=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(_SO))
        u64p, vp = C.POINTER(C.c_uint64), C.c_void_p
        L.oracle_parse_label.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), u64p, u64p, C.POINTER(C.c_int)]
        L.oracle_parse_label.restype = C.c_int
        L.oracle_make_params.argtypes = [vp, vp, vp, vp, vp, C.c_size_t, C.c_int, vp]
        L.oracle_make_params.restype = None
        L.oracle_make_row.argtypes = [vp, C.c_size_t, C.c_uint64, vp, vp]
        L.oracle_make_row.restype = C.c_size_t
        L.oracle_build_chunked.argtypes = [vp, C.c_size_t, C.c_uint64, C.c_uint64, C.c_size_t, C.c_int, vp, vp, vp, u64p]
        L.oracle_build_chunked.restype = C.c_int
        L.oracle_build_grouped.argtypes = [vp, C.c_size_t, C.c_int, C.c_uint64, C.c_uint64, C.c_int, vp, vp, vp, u64p]
        L.oracle_build_grouped.restype = C.c_int
        L.oracle_single_pauli.argtypes = [C.c_uint64, C.c_uint64, C.c_double, C.c_double, C.c_int, C.c_int, vp, vp, vp]
        L.oracle_single_pauli.restype = None
        L.oracle_spmv.argtypes = [vp, vp, vp, C.c_uint64, vp, vp, C.c_int]
        L.oracle_spmv.restype = None
        L.oracle_apply_rows.argtypes = [vp, C.c_size_t, vp, C.c_size_t, vp, vp]
        L.oracle_apply_rows.restype = None
        L.oracle_hardware_threads.restype = C.c_int
        d = C.c_double
        L.oracle_axpby.argtypes = [d, d, vp, d, d, vp, vp, C.c_size_t]
        L.oracle_axpy.argtypes = [d, d, vp, vp, vp, C.c_size_t]
        L.oracle_ax.argtypes = [d, d, vp, vp, C.c_size_t]
        L.oracle_precond2.argtypes = [vp, vp, d, d, d, vp, C.c_size_t]
        L.oracle_precond2.restype = None
        for f in (L.oracle_axpby, L.oracle_axpy, L.oracle_ax):
            f.restype = None
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def hardware_threads():
    return int(lib().oracle_hardware_threads())


def parse_label(label):
    """-> (base_phase, n_qubits, x, z, n_y); raises ValueError like Pauli::new (lib.rs:156-159)."""
    bp, nq, ny = C.c_int(), C.c_int(), C.c_int()
    x, z = C.c_uint64(), C.c_uint64()
    rc = lib().oracle_parse_label(label.encode(), C.byref(bp), C.byref(nq), C.byref(x), C.byref(z), C.byref(ny))
    if rc != 0:
        raise ValueError("error: malformed label")
    return bp.value, nq.value, x.value, z.value, ny.value


def make_params(labels, coeffs, convention="to_matrix"):
    """labels + complex coeffs -> (n_qubits, params[T]) as make_params does (accel.rs:141-157).

    convention "rowwise" = accel.rs verbatim; "to_matrix" = the default to_matrix's
    (+i)^base_phase (lib.rs:182-190,211).  Identical unless a label has an i/j prefix."""
    if len(labels) != len(coeffs):
        raise ValueError("SparsePauliOp::new: paulis and coeffs must have same length")      # lib.rs:355-357
    if len(labels) == 0:
        raise ValueError("SparsePauliOp::new: at least one pauli must be supplied")          # lib.rs:358-360
    T = len(labels)
    bp = np.zeros(T, np.int32); ny = np.zeros(T, np.int32)
    x = np.zeros(T, np.uint64); z = np.zeros(T, np.uint64)
    nq0 = None
    for t, lab in enumerate(labels):
        b, nq, xm, zm, y = parse_label(lab)
        if nq0 is None:
            nq0 = nq
        elif nq != nq0:
            raise ValueError("SparsePauliOp::new: all supplied paulis must have the same #qubits")  # lib.rs:362-365
        bp[t], ny[t], x[t], z[t] = b, y, xm, zm
    c = np.ascontiguousarray(np.asarray(coeffs, dtype=np.complex128))
    out = np.zeros(T, PARAM_DTYPE)
    lib().oracle_make_params(_p(bp), _p(ny), _p(x), _p(z), _p(c), T, 0 if convention == "rowwise" else 1, _p(out))
    return nq0, out


def make_row(params, row):
    """accel.rs:171-210 -> (cols u64[k], vals c128[k])."""
    T = len(params)
    cols = np.zeros(T, np.uint64); vals = np.zeros(T, np.complex128)
    k = lib().oracle_make_row(_p(params), T, int(row), _p(cols), _p(vals))
    return cols[:k].copy(), vals[:k].copy()


def build_csr(params, n_qubits, row_lo=0, row_hi=None, step=1000, n_threads=None, groups=None):
    """accel.rs:267-336 on rows [row_lo,row_hi) -> (indptr, indices, data).  indptr starts at 0."""
    params = np.ascontiguousarray(params)
    T = len(params)
    dim = 1 << n_qubits
    row_hi = dim if row_hi is None else row_hi
    rows = row_hi - row_lo
    if groups is None:
        groups = len(np.unique(params["x"]))
    cap = rows * groups
    indptr = np.zeros(rows + 1, np.uint64)
    indices = np.zeros(cap, np.uint64)
    data = np.zeros(cap, np.complex128)
    nnz = C.c_uint64()
    nt = n_threads or hardware_threads()
    rc = lib().oracle_build_chunked(_p(params), T, row_lo, row_hi, step, nt, _p(indptr), _p(indices), _p(data), C.byref(nnz))
    if rc != 0:
        raise MemoryError("oracle_build_chunked")
    assert nnz.value == cap, (nnz.value, cap)
    return indptr, indices, data


def build_csr_grouped(params, n_qubits, row_lo=0, row_hi=None, n_threads=None, groups=None):
    """Tuned CPU variant (NOT the reference's algorithm; see oracle_build_grouped): same bytes as build_csr."""
    params = np.ascontiguousarray(params)
    dim = 1 << n_qubits
    row_hi = dim if row_hi is None else row_hi
    rows = row_hi - row_lo
    if groups is None:
        groups = len(np.unique(params["x"]))
    indptr = np.zeros(rows + 1, np.uint64)
    indices = np.zeros(rows * groups, np.uint64)
    data = np.zeros(rows * groups, np.complex128)
    nnz = C.c_uint64()
    rc = lib().oracle_build_grouped(_p(params), len(params), n_qubits, row_lo, row_hi, n_threads or hardware_threads(),
                                    _p(indptr), _p(indices), _p(data), C.byref(nnz))
    if rc != 0:
        raise MemoryError("oracle_build_grouped")
    assert nnz.value == rows * groups
    return indptr, indices, data


def single_pauli(z, x, coeff, phase, n_qubits):
    """accel.rs:22-122 (Pauli::to_unsafe_vectors, lib.rs:215-223)."""
    dim = 1 << n_qubits
    indptr = np.zeros(dim + 1, np.uint64); indices = np.zeros(dim, np.uint64); data = np.zeros(dim, np.complex128)
    coeff = complex(coeff)
    lib().oracle_single_pauli(z, x, coeff.real, coeff.imag, phase, n_qubits, _p(indptr), _p(indices), _p(data))
    return indptr, indices, data


def spmv(indptr, indices, data, v, n_threads=None):
    """accel.rs:338-370."""
    rows = len(indptr) - 1
    v = np.ascontiguousarray(v, dtype=np.complex128)
    y = np.zeros(rows, np.complex128)
    lib().oracle_spmv(_p(indptr), _p(indices), _p(data), rows, _p(v), _p(y), n_threads or hardware_threads())
    return y


def apply_rows(params, rows, v):
    """y[i] = (row rows[i] of the reference-built CSR) . v, in stored order, without the CSR."""
    params = np.ascontiguousarray(params)
    rows = np.ascontiguousarray(rows, dtype=np.uint64)
    v = np.ascontiguousarray(v, dtype=np.complex128)
    y = np.zeros(len(rows), np.complex128)
    lib().oracle_apply_rows(_p(params), len(params), _p(rows), len(rows), _p(v), _p(y))
    return y


def _cv(x):
    return np.ascontiguousarray(x, dtype=np.complex128)


def axpby(a, x, b, y):
    """accel.rs:374-379."""
    x, y = _cv(x), _cv(y); z = np.empty_like(x); a, b = complex(a), complex(b)
    lib().oracle_axpby(a.real, a.imag, _p(x), b.real, b.imag, _p(y), _p(z), len(x))
    return z


def axpy(a, x, y):
    """accel.rs:381-386."""
    x, y = _cv(x), _cv(y); z = np.empty_like(x); a = complex(a)
    lib().oracle_axpy(a.real, a.imag, _p(x), _p(y), _p(z), len(x))
    return z


def ax(a, x):
    """accel.rs:388-393."""
    x = _cv(x); z = np.empty_like(x); a = complex(a)
    lib().oracle_ax(a.real, a.imag, _p(x), _p(z), len(x))
    return z


def precond2(diag, dx, e, tol):
    """pyqrusty/src/lib.rs:436-468: dx / reg(diag - e, tol)."""
    diag, dx = _cv(diag), _cv(dx); out = np.empty_like(dx); e = complex(e)
    lib().oracle_precond2(_p(diag), _p(dx), e.real, e.imag, float(tol), _p(out), len(dx))
    return out


def eliminate_zeros(indptr, indices, data, tolerance=1e-7):
    """util::csmatrix_eliminate_zeroes (qrusty/src/util.rs:154-171): keep entries with
    norm() > tolerance (norm = hypot), row-major order preserved; -> (indptr, indices, data)."""
    keep = np.hypot(data.real, data.imag) > tolerance
    rows = np.repeat(np.arange(len(indptr) - 1), np.diff(indptr.astype(np.int64)))
    counts = np.bincount(rows[keep], minlength=len(indptr) - 1).astype(np.uint64)
    new_indptr = np.concatenate([[np.uint64(0)], np.cumsum(counts, dtype=np.uint64)])
    return new_indptr, indices[keep].copy(), data[keep].copy()


def count_zeros(data, tolerance=1e-7):
    """util::csmatrix_nz (util.rs:144-152)."""
    return int(np.count_nonzero(np.hypot(data.real, data.imag) <= tolerance))


def rawio_bytes(shape, indptr, indices, data, byteorder="="):
    """qrusty::rawio::write (rawio.rs:128-148) as bytes: native-endian "MI" mark (rawio.rs:59-68), u64
    storage tag (0 = CSR), rows, cols, then indptr / indices / data each prefixed by its u64 length.
    byteorder ">" / "<" forces the producer's endianness (to exercise the reader's swab path)."""
    import sys
    little = sys.byteorder == "little" if byteorder == "=" else byteorder == "<"
    u8 = np.dtype("<u8" if little else ">u8")
    c16 = np.dtype("<c16" if little else ">c16")
    out = [b"MI" if little else b"IM"]              # u16::from_ne_bytes(['M','I']) written with to_ne_bytes
    out.append(np.array([0, shape[0], shape[1], len(indptr)], u8).tobytes())
    out.append(np.asarray(indptr).astype(u8).tobytes())
    out.append(np.array([len(indices)], u8).tobytes()); out.append(np.asarray(indices).astype(u8).tobytes())
    out.append(np.array([len(data)], u8).tobytes()); out.append(np.asarray(data).astype(c16).tobytes())
    return b"".join(out)
