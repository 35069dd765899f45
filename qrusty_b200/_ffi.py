"""ctypes binding of include/qrusty_cuda.h -- the same C ABI the Rust `qrusty::cuda`
module binds.  There is no fallback: if the shared library is
missing this module raises at import, and every compute call raises QrustyCudaError
when CUDA fails.
"""
import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("QRUSTY_CUDA_LIB", _PKG / "lib" / "libqrusty_cuda.so"))

QR_OK, QR_ERR_INVALID, QR_ERR_CUDA, QR_ERR_NCCL, QR_ERR_OOM, QR_ERR_UNSUPPORTED = range(6)
QR_INDPTR_LOCAL, QR_INDPTR_GLOBAL, QR_FILL_DIRECT, QR_HOST_NO_STAGING, QR_HOST_WIDE = 0, 1, 2, 4, 8
QR_PLAN_MERGE_DUPLICATES = 1
QR_UNIQUE_ID_BYTES = 128
QR_IPC_HANDLE_BYTES = 64


class QrustyCudaError(Exception):
    """A non-zero status from the C ABI (the reference raises PyException from QrustyErr,
    pyqrusty/src/lib.rs:48-52)."""

    def __init__(self, code, message):
        super().__init__(message)
        self.code = code


class Term(C.Structure):
    """qr_term == one tuple of rowwise::make_params (qrusty/src/accel.rs:141-157)."""
    _fields_ = [("z", C.c_uint64), ("x", C.c_uint64), ("re", C.c_double), ("im", C.c_double)]


class PlanInfo(C.Structure):
    _fields_ = [("n_qubits", C.c_int32), ("device", C.c_int32), ("dim", C.c_uint64),
                ("n_terms", C.c_uint64), ("n_groups", C.c_uint64), ("nnz", C.c_uint64)]


if not LIB_PATH.exists():
    raise ImportError(
        f"qrusty_b200: {LIB_PATH} not found.  Build it with `python qrusty_b200/build.py` "
        "(nvcc, sm_100a).  There is no CPU fallback.")

lib = C.CDLL(str(LIB_PATH))

_vp, _u64, _u32, _sz, _int = C.c_void_p, C.c_uint64, C.c_uint32, C.c_size_t, C.c_int
_dp = C.POINTER(C.c_double)

# name -> (argtypes, restype is int unless given)
SIGNATURES = {
    "qr_plan_create": [_int, _vp, _sz, _int, _u32, C.POINTER(_vp)],
    "qr_plan_destroy": [_vp],
    "qr_plan_info": [_vp, C.POINTER(PlanInfo)],
    "qr_plan_groups": [_vp, _vp, _vp, _vp],
    "qr_plan_canonical_terms": [_vp, C.POINTER(_u64)],
    "qr_plan_canonicalise_async": [_vp, _vp],
    "qr_build_rows_device": [_vp, _u64, _u64, _vp, _vp, _vp, _u32, _vp],
    "qr_build_host": [_vp, _u64, _u64, _vp, _vp, _vp, _u32],
    "qr_write_rawio": [_vp, _u64, _u64, C.c_char_p],
    "qr_csr_count_kept_device": [_u64, _vp, _vp, C.c_double, _vp, C.POINTER(_u64), _vp],
    "qr_csr_compact_device": [_u64, _vp, _vp, _vp, C.c_double, _vp, _vp, _vp, _vp],
    "qr_apply_device": [_vp, _u64, _u64, _vp, _vp, _vp],
    "qr_apply_host": [_vp, _vp, _vp],
    "qr_diagonal_device": [_vp, _u64, _u64, _vp, _vp],
    "qr_spmv_device": [_u64, _vp, _vp, _vp, _vp, _vp, _vp],
    "qr_csr_diagonal_device": [_u64, _u64, _vp, _vp, _vp, _vp, _vp],
    "qr_count_kept_device": [_u64, _u64, _vp, C.c_double, _vp, C.POINTER(_u64), _vp],
    "qr_compact_rows_device": [_u64, _u64, _vp, _vp, C.c_double, _vp, _vp, _vp, _vp],
    "qr_build_compact_count": [_vp, _u64, _u64, C.c_double, _vp, C.POINTER(_u64), _vp],
    "qr_build_compact_fill": [_vp, _u64, _u64, C.c_double, _vp, _vp, _vp, _vp],
    "qr_axpby_device": [_u64, _dp, _vp, _dp, _vp, _vp, _vp],
    "qr_axpy_device": [_u64, _dp, _vp, _vp, _vp, _vp],
    "qr_ax_device": [_u64, _dp, _vp, _vp, _vp],
    "qr_precond2_device": [_u64, _vp, _vp, _dp, C.c_double, _vp, _vp],
    "qr_dotc_device": [_u64, _vp, _vp, _vp, _vp],
    "qr_lanczos_update_device": [_u64, _dp, _dp, _vp, _vp, _vp, _vp, _vp, _vp],
    "qr_comm_unique_id": [_vp],
    "qr_comm_create": [_vp, _int, _int, _int, C.POINTER(_vp)],
    "qr_comm_destroy": [_vp],
    "qr_apply_distributed": [_vp, _vp, _vp, _vp, _vp, _vp],
    "qr_allreduce_sum_f64": [_vp, _vp, _sz, _vp],
    "qr_apply_dot_device": [_vp, _u64, _u64, _vp, _vp, _vp, _vp],
    "qr_apply_p2p_dot": [_vp, _vp, _vp, _vp, _vp, _vp],
    "qr_lanczos_coef_device": [_vp, _u32, _u32, _u32, _vp],
    "qr_lanczos_update_dev": [_u64, _vp, _vp, _vp, _vp, _vp, _vp],
    "qr_apply_p2p": [_vp, _vp, _vp, _vp, _vp],
    "qr_ipc_get_handle": [_vp, _vp],
    "qr_ipc_open_handle": [_vp, C.POINTER(_vp)],
    "qr_ipc_close_handle": [_vp],
    "qr_release_scratch": [],
    "qr_device_count": [C.POINTER(_int)],
    "qr_device_name": [_int, C.c_char_p, _sz],
    "qr_set_device": [_int],
    "qr_malloc_device": [C.POINTER(_vp), _sz],
    "qr_free_device": [_vp],
    "qr_malloc_host": [C.POINTER(_vp), _sz],
    "qr_free_host": [_vp],
    "qr_memcpy_h2d": [_vp, _vp, _sz, _vp],
    "qr_memcpy_d2h": [_vp, _vp, _sz, _vp],
    "qr_memset_device": [_vp, _int, _sz, _vp],
    "qr_stream_create": [C.POINTER(_vp)],
    "qr_stream_destroy": [_vp],
    "qr_stream_synchronize": [_vp],
    "qr_graph_begin_capture": [_vp],
    "qr_graph_end_capture": [_vp, C.POINTER(_vp)],
    "qr_graph_launch": [_vp, _vp],
    "qr_graph_destroy": [_vp],
    "qr_event_create": [C.POINTER(_vp)],
    "qr_event_destroy": [_vp],
    "qr_event_record": [_vp, _vp],
    "qr_event_elapsed_ms": [_vp, _vp, C.POINTER(C.c_float)],
}
for _name, _args in SIGNATURES.items():
    _fn = getattr(lib, _name)          # AttributeError here == header and library disagree
    _fn.argtypes = _args
    _fn.restype = C.c_int
lib.qr_last_error.restype = C.c_char_p
lib.qr_last_error.argtypes = []
lib.qr_version.restype = C.c_int
lib.qr_kernel_launches.restype = C.c_uint64
lib.qr_last_d2h_bytes.restype = C.c_uint64
lib.qr_last_d2h_bytes.argtypes = []
lib.qr_plan_fill_kernel.restype = C.c_char_p
lib.qr_plan_fill_kernel.argtypes = [_vp]
lib.qr_plan_apply_kernel.restype = C.c_char_p
lib.qr_plan_apply_kernel.argtypes = [_vp, _u64, _u64]
EXPORTS = sorted(list(SIGNATURES) + ["qr_last_error", "qr_version", "qr_kernel_launches", "qr_plan_fill_kernel", "qr_plan_apply_kernel", "qr_last_d2h_bytes"])


def check(rc):
    if rc != QR_OK:
        raise QrustyCudaError(rc, (lib.qr_last_error() or b"").decode(errors="replace"))


def call(name, *args):
    check(getattr(lib, name)(*args))


def last_d2h_bytes():
    return int(lib.qr_last_d2h_bytes())


def kernel_launches():
    return int(lib.qr_kernel_launches())


def device_count():
    n = C.c_int(0)
    rc = lib.qr_device_count(C.byref(n))
    return n.value if rc == QR_OK else 0
