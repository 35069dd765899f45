"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

indptr / indices: bit-exact.  data: bit-exact (same summation order, signed zeros included);
the 1e-12 tolerance of the north-star is therefore never needed for the build.  H.v
(matrix-free) sums the groups in a fixed mask order, not the per-row column order of the CSR
SpMV, so it is checked to 1e-12 * sum_k |a_k||v_k|.

Run on the B200 box:  python -m pytest tests -m gpu -x -q
"""
import ctypes as C
import hashlib

import numpy as np
import pytest

from oracle import oracle as O, oracle_np as N
import qrusty_b200 as Q
from qrusty_b200 import _ffi, hamiltonians as H
from qrusty_b200._runtime import DeviceBuffer

pytestmark = pytest.mark.gpu


def u64(a):
    return np.ascontiguousarray(a).view(np.uint64)


def make_op(labels, coeffs):
    return Q.SparsePauliOp([Q.Pauli(l) for l in labels], coeffs)


def device_build(plan, lo, hi, flags=0, with_indptr=True):
    """qr_build_rows_device into fresh buffers pre-filled with 0xFF (every slot must be written)."""
    rows, G = hi - lo, plan.n_groups
    ip = DeviceBuffer((rows + 1) * 8); ix = DeviceBuffer(rows * G * 8); dt = DeviceBuffer(rows * G * 16)
    for b in (ip, ix, dt):
        _ffi.call("qr_memset_device", b.ptr, 0xFF, b.nbytes, None)
    _ffi.call("qr_build_rows_device", plan.handle, lo, hi, ip.ptr if with_indptr else None, ix.ptr, dt.ptr, flags, None)
    _ffi.call("qr_stream_synchronize", None)
    return (ip.download(np.empty(rows + 1, np.uint64)), ix.download(np.empty(rows * G, np.uint64)),
            dt.download(np.empty(rows * G, np.complex128)))


def assert_same(got, ref, what=""):
    for name, a, b in zip(("indptr", "indices", "data"), got, ref):
        assert a.shape == b.shape, (what, name, a.shape, b.shape)
        if not np.array_equal(u64(a), u64(b)):
            bad = np.flatnonzero(u64(a).reshape(len(a), -1) != u64(b).reshape(len(b), -1))[:5]
            raise AssertionError(f"{what}: {name} differs at {bad}: {a[bad]} vs {b[bad]}")


SMALL = {
    "C1": lambda fx: H.tfim_chain(12),
    "xxz_n10": lambda fx: H.xxz_chain(10, 1.0, 0.7),
    "tfim_3x3": lambda fx: H.tfim_lattice(3, 3, 1.0, 3.0),
    "random_n10": lambda fx: H.random_pauli_sum(10, 300, 200, 30, 7),
    "H2": lambda fx: fx["H2"], "H2_rs": lambda fx: fx["H2_rs"],
    "H4": lambda fx: fx["H4"], "H4_rs": lambda fx: fx["H4_rs"], "H6": lambda fx: fx["H6"],
}


# ---- the reference's cross-mode tests (test_H.py:32-52, lib.rs:815-885) with the Cuda mode ----
@pytest.mark.parametrize("name", list(SMALL))
def test_build_matches_oracle_and_goldens(fixtures, golden_sums, name):
    labels, coeffs = SMALL[name](fixtures)
    n, params = O.make_params(labels, coeffs)
    ref = O.build_csr(params, n, step=100)
    op = make_op(labels, coeffs)
    m = op.to_matrix_mode("Cuda")
    assert m.shape() == (1 << n, 1 << n)
    assert m.nnz() == len(ref[2])
    shape, data, indices, indptr = m.export()
    assert_same((indptr, indices, data), ref, name)
    g = golden_sums[name]
    for key, arr in (("indptr", indptr), ("indices", indices), ("data", data)):
        assert hashlib.sha256(np.ascontiguousarray(arr).tobytes()).hexdigest() == g[key], key


@pytest.mark.parametrize("name", ["H2", "H4", "H6"])
def test_reference_test_H(fixtures, name):
    """pyqrusty/tests/test_H.py:32-48 verbatim in spirit: every mode gives the same dense matrix."""
    labels, coeffs = fixtures[name]
    op = make_op(labels, coeffs)
    m1 = Q.csr_matrix(op.to_matrix())
    assert np.array_equal(m1.todense(), Q.csr_matrix(op.to_matrix_mode(mode="")).todense())
    m2 = op.to_matrix_mode(mode="RowwiseUnsafeChunked/100")
    assert np.array_equal(m1.todense(), Q.csr_matrix(m2).todense())
    assert np.array_equal(m1.todense(), Q.csr_matrix(op.to_matrix_mode("Cuda")).todense())
    if name != "H6":
        assert np.array_equal(np.asarray(m1.todense()), N.spop_dense(labels, coeffs))


def test_mode_and_export_errors(fixtures):
    op = make_op(*fixtures["H2"])
    with pytest.raises(Exception):
        op.to_matrix_mode("foo")                                   # test_H.py:28-30
    m = op.to_matrix()
    assert repr(m) == "<16x16 sparse matrix of type Complex64\n\twith 64 stored elements in Compressed Sparse Row format>"
    m.export()
    with pytest.raises(Exception, match="already-exported"):
        m.export()                                                 # pyqrusty/src/lib.rs:195-197
    with pytest.raises(Exception):
        m.shape()
    assert repr(m) == "<already-dropped sparse matrix of type Complex64>"


# ---- lib.rs:721-768: single Pauli strings (G = 1) ----------------------------------------
@pytest.mark.parametrize("label", ["IX", "XI", "I", "Y", "YY", "ZXYI", "-YZYX", "X", "Z"])
def test_single_pauli(label):
    bp, nq, x, z, ny = O.parse_label(label)
    ref = O.single_pauli(z, x, (-1 if bp == 2 else 1) + 0j, ny % 4, nq)
    shape, data, indices, indptr = Q.Pauli(label).to_matrix().export()
    assert np.array_equal(indptr, ref[0]) and np.array_equal(indices, ref[1])
    assert np.array_equal(data, ref[2])
    assert np.array_equal(np.asarray(Q.csr_matrix(Q.Pauli(label).to_matrix()).todense()), N.pauli_dense(label))


def test_pauli_known_answers():
    """pyqrusty/tests/test_it.py:60-101."""
    I = np.array([[1.0, 0.0], [0.0, 1.0]], dtype=complex); Z = np.array([[1.0, 0.0], [0.0, -1.0]], dtype=complex)
    X = np.array([[0.0, 1.0], [1.0, 0.0]], dtype=complex); Y = np.array([[0.0, -1.0j], [1.0j, 0.0]], dtype=complex)
    for lab, mat in (("I", I), ("X", X), ("Z", Z), ("Y", Y)):
        assert np.array_equal(mat, Q.csr_matrix(Q.Pauli(lab).to_matrix()).todense())
    spop = Q.SparsePauliOp([Q.Pauli("I"), Q.Pauli("Y")], [1.0 + 0.0j, 2.0 + 0.0j])
    assert np.array_equal(I + 2.0 * Y, Q.csr_matrix(spop.to_matrix()).todense())
    spop = Q.SparsePauliOp([Q.Pauli("I"), Q.Pauli("X")], [1.0 + 0.0j, 2.0 + 0.0j])
    assert np.array_equal([[1, 2], [2, 1]], Q.csr_matrix(spop.to_matrix()).todense())   # lib.rs:786-793


def test_prefixed_labels_follow_to_matrix():
    """i/j prefixes use the default to_matrix's (+i)^base_phase."""
    labels, coeffs = ["iXY", "-jZI", "-XX", "YZ", "+1ZZ"], [0.5 + 0.25j, 1.5 + 0j, -2 + 1j, 0.75 + 0j, 1j]
    got = np.asarray(Q.csr_matrix(make_op(labels, coeffs).to_matrix()).todense())
    assert np.array_equal(got, N.spop_dense(labels, coeffs))


# ---- tiny dimensions (dim < one warp) and ragged / unaligned row windows ---------------------
@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 6, 7])
def test_tiny_dims(n):
    labels, coeffs = H.random_pauli_sum(n, 40, min(1 << n, 12), 5, 100 + n)
    nq, params = O.make_params(labels, coeffs)
    ref = O.build_csr(params, nq)
    shape, data, indices, indptr = make_op(labels, coeffs).to_matrix().export()
    assert_same((indptr, indices, data), ref, f"n={n}")


@pytest.mark.parametrize("flags", [0, _ffi.QR_FILL_DIRECT])
def test_row_windows(fixtures, flags):
    labels, coeffs = fixtures["H4"]                                  # n=8, G=51
    n, params = O.make_params(labels, coeffs)
    plan = make_op(labels, coeffs).plan()
    G = plan.n_groups
    full = O.build_csr(params, n)
    for lo, hi in [(0, 256), (37, 201), (0, 1), (255, 256), (31, 33), (64, 128), (1, 255), (100, 101), (96, 160)]:
        ip, ix, dt = device_build(plan, lo, hi, flags)
        assert np.array_equal(ip, np.arange(hi - lo + 1, dtype=np.uint64) * G), (lo, hi)
        assert np.array_equal(ix, full[1][lo * G:hi * G]), (lo, hi)
        assert np.array_equal(u64(dt), u64(full[2][lo * G:hi * G])), (lo, hi)
        ipg, _, _ = device_build(plan, lo, hi, flags | _ffi.QR_INDPTR_GLOBAL)
        assert np.array_equal(ipg, (np.arange(hi - lo + 1, dtype=np.uint64) + lo) * G), (lo, hi)
        _, ix2, dt2 = device_build(plan, lo, hi, flags, with_indptr=False)
        assert np.array_equal(ix2, ix) and np.array_equal(u64(dt2), u64(dt))


def test_direct_equals_staged(fixtures):
    labels, coeffs = H.xxz_chain(14, 1.0, 0.7)
    plan = make_op(labels, coeffs).plan()
    a = device_build(plan, 0, plan.dim, 0)
    b = device_build(plan, 0, plan.dim, _ffi.QR_FILL_DIRECT)
    assert_same(a, b, "staged vs direct")


@pytest.mark.parametrize("name,force", [("xxz11", False), ("xxz15", False), ("xxz7", False), ("tfim_3x3", True), ("H2", False)])
def test_staged_padded_tiles(fixtures, monkeypatch, name, force):
    """fill_staged_kernel<.,.,PAD>: G a multiple of 4 (xxz n=11/15/7: G = 12/16/8, H2: 4) takes the padded
    shared-memory pitch and one TMA copy per row automatically; QR_FILL_PAD=1 forces it for any even G."""
    monkeypatch.setenv("QR_FILL_ROWS", "0")              # this test is about the staged kernel
    if force:
        monkeypatch.setenv("QR_FILL_PAD", "1")
    labels, coeffs = H.xxz_chain(int(name[3:]), 1.0, 0.7) if name.startswith("xxz") else SMALL[name](fixtures)
    n, params = O.make_params(labels, coeffs)
    ref = O.build_csr(params, n)
    plan = make_op(labels, coeffs).plan()
    G, dim = plan.n_groups, 1 << n
    assert G % 2 == 0
    assert_same(device_build(plan, 0, dim), ref, name)
    if dim >= 256:
        for lo, hi in [(5, dim - 3), (64, 192), (dim // 2 - 1, dim // 2 + 130)]:
            ip, ix, dt = device_build(plan, lo, hi, flags=_ffi.QR_INDPTR_GLOBAL)
            assert np.array_equal(ix, ref[1][lo * G:hi * G]) and np.array_equal(u64(dt), u64(ref[2][lo * G:hi * G])), (lo, hi)
            assert np.array_equal(ip, np.arange(lo, hi + 1, dtype=np.uint64) * G)


@pytest.mark.parametrize("cfg", ["2,8", "1,4", "4,16"])
@pytest.mark.parametrize("name", ["xxz11", "xxz15", "xxz7", "xxz19", "xxz13", "tfim_3x3", "H2", "H4_rs"])
def test_staged_swizzled_tiles(fixtures, monkeypatch, name, cfg):
    """fill_staged_swz_kernel: the tile in the TMA's 128-byte swizzle, boxes of 8 entries x R rows leaving through 2-D
    tensor maps of the output arrays (last box of a row clipped at column G).  Every even G (12, 16, 8, 20, 14, 10, 4, ...),
    whole matrices and ragged row windows (edge rows by the direct kernel), local and global indptr."""
    monkeypatch.setenv("QR_FILL_ROWS", "0")
    monkeypatch.setenv("QR_FILL_SWZ", "1")
    monkeypatch.setenv("QR_FILL_CFG", cfg)
    labels, coeffs = H.xxz_chain(int(name[3:]), 1.0, 0.7) if name.startswith("xxz") else SMALL[name](fixtures)
    n, params = O.make_params(labels, coeffs)
    plan = make_op(labels, coeffs).plan()
    G, dim = plan.n_groups, 1 << n
    if G % 2 or dim < 32 * int(cfg[0]):
        pytest.skip("odd G or fewer rows than a tile")
    R = 32 * int(cfg[0])
    lo_hi = [(0, dim)] + ([(5, dim - 3), (R, 3 * R), (dim // 2 - 1, dim // 2 + R + 2)] if dim >= 4 * R else [])
    if n > 16:                                                      # xxz19: windows, not 10^7 rows of oracle
        lo_hi = [(dim // 2 + 5, dim // 2 + 4096 + 82), (dim // 2, dim // 2 + 8192), (dim - 4096, dim)]
    for lo, hi in lo_hi:
        ref = O.build_csr(params, n, lo, hi)
        before = _ffi.kernel_launches()
        ip, ix, dt = device_build(plan, lo, hi, flags=_ffi.QR_INDPTR_GLOBAL)
        assert _ffi.kernel_launches() - before <= 3
        assert np.array_equal(ix, ref[1]) and np.array_equal(u64(dt), u64(ref[2])), (lo, hi)
        assert np.array_equal(ip, np.arange(lo, hi + 1, dtype=np.uint64) * G)
        assert_same(device_build(plan, lo, hi), ref, f"{name} [{lo},{hi}) local indptr")


@pytest.mark.parametrize("E", [1, 2])
@pytest.mark.parametrize("S", [32, 48, 64, 128])
@pytest.mark.parametrize("name", ["H4", "H6", "random_n10", "xxz_n10", "C1", "H2"])
def test_blocked_kernel(fixtures, monkeypatch, name, S, E):
    """Large-G path (partition_kernel + fill_blocked_kernel) forced on every case: subtree blocks
    of at most S groups, whole matrix and ragged windows."""
    monkeypatch.setenv("QR_FILL_BLOCK", str(S))
    monkeypatch.setenv("QR_FILL_BLOCK_E", str(E))
    labels, coeffs = SMALL[name](fixtures)
    n, params = O.make_params(labels, coeffs)
    ref = O.build_csr(params, n)
    plan = make_op(labels, coeffs).plan()
    G, dim = plan.n_groups, 1 << n
    assert_same(device_build(plan, 0, dim), ref, f"{name} S={S}")
    if dim >= 128:
        for lo, hi in [(5, dim - 3), (32, 64), (dim // 2 - 1, dim // 2 + 33)]:
            ip, ix, dt = device_build(plan, lo, hi)
            assert np.array_equal(ix, ref[1][lo * G:hi * G]) and np.array_equal(u64(dt), u64(ref[2][lo * G:hi * G])), (lo, hi)
            assert np.array_equal(ip, np.arange(hi - lo + 1, dtype=np.uint64) * G)


@pytest.mark.parametrize("W,sync", [(8, 0), (8, 2), (16, 1), (16, 2), (32, 2)])
@pytest.mark.parametrize("R", [5, 6, 8])
@pytest.mark.parametrize("name", ["H4", "H6", "random_n10", "xxz_n10", "C1", "H2", "tfim_3x3"])
def test_lanes_kernel(fixtures, monkeypatch, name, R, W, sync):
    """Large-G path (fill_lanes_kernel: lane <-> group, rows in Gray-code order, heavy groups
    lane <-> row at the start of every 32-row strip) forced on every case: whole matrix and ragged windows."""
    monkeypatch.setenv("QR_FILL_LANES", "1")
    monkeypatch.setenv("QR_FILL_LANES_R", str(R))
    monkeypatch.setenv("QR_FILL_LANES_W", str(W))
    monkeypatch.setenv("QR_FILL_LANES_SYNC", str(sync))
    labels, coeffs = SMALL[name](fixtures)
    n, params = O.make_params(labels, coeffs)
    ref = O.build_csr(params, n)
    plan = make_op(labels, coeffs).plan()
    G, dim = plan.n_groups, 1 << n
    assert_same(device_build(plan, 0, dim), ref, f"{name} R={R} W={W} sync={sync}")
    if dim >= 128:
        for lo, hi in [(5, dim - 3), (32, 64), (dim // 2 - 1, dim // 2 + 33)]:
            ip, ix, dt = device_build(plan, lo, hi, flags=_ffi.QR_INDPTR_GLOBAL)
            assert np.array_equal(ix, ref[1][lo * G:hi * G]) and np.array_equal(u64(dt), u64(ref[2][lo * G:hi * G])), (lo, hi)
            assert np.array_equal(ip, np.arange(lo, hi + 1, dtype=np.uint64) * G)


@pytest.mark.parametrize("W", [8, 16, 32])
def test_lanes_kernel_clusters(monkeypatch, W):
    """G = 1100: 5 / 3 / 2 sibling CTAs per row run, launched as one thread-block cluster that
    re-aligns every 32 rows (odd cluster sizes included); rows in runs of 32 and 256."""
    labels, coeffs = H.random_pauli_sum(12, 1500, 1100, 50, 5)
    n, params = O.make_params(labels, coeffs)
    ref = O.build_csr(params, n)
    monkeypatch.setenv("QR_FILL_ROWS", "0")
    for R in (5, 8):
        monkeypatch.setenv("QR_FILL_LANES_R", str(R))
        monkeypatch.setenv("QR_FILL_LANES_W", str(W))
        plan = make_op(labels, coeffs).plan()
        assert plan.n_groups == 1100 and plan.fill_kernel == "fill_lanes_kernel"
        assert_same(device_build(plan, 0, 1 << n), ref, f"G=1100 W={W} R={R}")
        lo, hi = 100, 4000
        ip, ix, dt = device_build(plan, lo, hi)
        assert np.array_equal(ix, ref[1][lo * 1100:hi * 1100]) and np.array_equal(u64(dt), u64(ref[2][lo * 1100:hi * 1100]))


@pytest.mark.parametrize("REGT,Q,R,SL", [(1, 1, 1, 0), (1, 2, 5, 0), (0, 1, 7, 0), (0, 2, 3, 0), (1, 1, 4, 0), (0, 2, 8, 0), (1, 1, 6, 0), (1, 2, 7, 0),
                                         (0, 1, 5, 0), (1, 3, 7, 0), (1, 3, 3, 0), (1, 3, 7, None), (1, 1, 6, 2), (1, 2, 7, 1), (1, 3, 8, 4), (1, 1, 3, 1)])
@pytest.mark.parametrize("HV", [None, "0", "2"])
@pytest.mark.parametrize("name", ["H4", "H6", "random_n10", "xxz_n10", "C1", "H2", "tfim_3x3"])
def test_rows_kernel(fixtures, monkeypatch, name, REGT, Q, R, SL, HV):
    """Rows kernel (fill_rows_kernel: persistent CTAs own whole rows, thread <-> group, batches of 2^Q rows
    in Gray-code order through two shared-memory buffers and the TMA) forced on every case: whole matrix
    and ragged windows (edge rows go through the direct kernel, misaligned windows through the default path).
    REGT: 1 = up to 6 terms of a group in registers (512 threads), 0 = first term in registers, the others in shared
    memory (1024 threads).  SL: log2 of the sub-batches the threads split into when G <= 256 (None: as many as fit; the
    library lowers it when 512 >> SL < G).  HV: threshold of the CTA-wide heavy-group path (default: more than 6 terms; "0": off;
    "2": nearly every group is heavy; both force the shared-memory variant)."""
    monkeypatch.setenv("QR_FILL_ROWS", "1")
    if HV is not None:
        monkeypatch.setenv("QR_FILL_ROWS_HV", HV)
    monkeypatch.setenv("QR_FILL_ROWS_REGT", str(REGT))
    monkeypatch.setenv("QR_FILL_ROWS_Q", str(Q))
    monkeypatch.setenv("QR_FILL_ROWS_R", str(R))
    if SL is not None:
        monkeypatch.setenv("QR_FILL_ROWS_SL", str(SL))
    labels, coeffs = SMALL[name](fixtures)
    n, params = O.make_params(labels, coeffs)
    ref = O.build_csr(params, n)
    plan = make_op(labels, coeffs).plan()
    assert plan.fill_kernel == "fill_rows_kernel"
    G, dim = plan.n_groups, 1 << n
    assert_same(device_build(plan, 0, dim), ref, f"{name} REGT={REGT} Q={Q} R={R} SL={SL} HV={HV}")
    if dim >= 128:
        for lo, hi in [(5, dim - 3), (32, 64), (dim // 2 - 1, dim // 2 + 33), (6, dim - 2)]:
            ip, ix, dt = device_build(plan, lo, hi, flags=_ffi.QR_INDPTR_GLOBAL)
            assert np.array_equal(ix, ref[1][lo * G:hi * G]) and np.array_equal(u64(dt), u64(ref[2][lo * G:hi * G])), (lo, hi)
            assert np.array_equal(ip, np.arange(lo, hi + 1, dtype=np.uint64) * G)


@pytest.mark.parametrize("dec,exthv", [("1", "0"), ("2", "0"), ("0", "1"), ("1", "1")])
@pytest.mark.parametrize("Q,R,HV", [(1, 1, None), (2, 5, None), (1, 4, "2"), (2, 7, "2"), (2, 3, "0"), (1, 6, None)])
@pytest.mark.parametrize("name", ["H4", "H6", "random_n10", "xxz_n10", "C1", "H2", "tfim_3x3"])
def test_rows_kernel_decoupled_whole_rows(fixtures, monkeypatch, name, Q, R, HV, dec, exthv):
    """Two options of the whole-row register variant forced on small cases.  dec: the decoupled hand-over (DEC, see
    test_rows_kernel_split) where the default keeps the batch barrier -- in-CTA heavy phase between two named barriers,
    first-arriver hand-over of whole batches.  exthv = 1: the heavy groups' values come from heavy_values_kernel instead of the
    CTA's own heavy phase (the default when that lets a batch hold more rows: H8)."""
    monkeypatch.setenv("QR_FILL_ROWS", "1")
    monkeypatch.setenv("QR_FILL_ROWS_DEC", dec)
    monkeypatch.setenv("QR_FILL_ROWS_EXTHV", exthv)
    monkeypatch.setenv("QR_FILL_ROWS_REGT", "1")
    monkeypatch.setenv("QR_FILL_ROWS_SL", "0")
    monkeypatch.setenv("QR_FILL_ROWS_Q", str(Q))
    monkeypatch.setenv("QR_FILL_ROWS_R", str(R))
    if HV is not None:
        monkeypatch.setenv("QR_FILL_ROWS_HV", HV)
    labels, coeffs = SMALL[name](fixtures)
    n, params = O.make_params(labels, coeffs)
    ref = O.build_csr(params, n)
    plan = make_op(labels, coeffs).plan()
    assert plan.fill_kernel == "fill_rows_kernel"
    G, dim = plan.n_groups, 1 << n
    assert_same(device_build(plan, 0, dim), ref, f"{name} Q={Q} R={R} HV={HV} dec={dec} exthv={exthv}")
    if dim >= 128:
        for lo, hi in [(5, dim - 3), (32, 64), (dim // 2 - 1, dim // 2 + 33)]:
            ip, ix, dt = device_build(plan, lo, hi, flags=_ffi.QR_INDPTR_GLOBAL)
            assert np.array_equal(ix, ref[1][lo * G:hi * G]) and np.array_equal(u64(dt), u64(ref[2][lo * G:hi * G])), (lo, hi)
            assert np.array_equal(ip, np.arange(lo, hi + 1, dtype=np.uint64) * G)


@pytest.mark.parametrize("G,T,REGT,Q", [(1100, 1500, 0, 1), (1100, 1500, 0, 2), (2200, 2600, 0, 1), (600, 2400, 0, 2), (600, 2400, 1, 2),
                                        (1000, 3000, 1, 1), (1000, 3000, 0, 1), (450, 500, None, 0), (700, 3000, None, 0), (400, 900, 1, 3),
                                        (200, 300, None, 0), (160, 700, None, 0), (100, 250, 1, 3), (60, 200, 1, 2)])
def test_rows_kernel_large_G(monkeypatch, G, T, REGT, Q):
    """The shapes the rows kernel is chosen for by default (G >= 400): 1..3 groups per thread, terms in registers
    or in shared memory (T - G extras, up to 4 per group on average, so some groups are heavy), 2- and 4-row
    batches, runs of 8 and 128 rows, several runs per persistent CTA (a 2^12-row matrix on 148 CTAs has 512 runs
    of 8 rows).  REGT None: the library's own choice (registers for the term-rich case)."""
    labels, coeffs = H.random_pauli_sum(12, T, G, 50, 5)
    n, params = O.make_params(labels, coeffs)
    ref = O.build_csr(params, n)
    if REGT is not None:
        monkeypatch.setenv("QR_FILL_ROWS_REGT", str(REGT))
        monkeypatch.setenv("QR_FILL_ROWS_Q", str(Q))
    if G <= 150:
        monkeypatch.setenv("QR_FILL_ROWS", "1")          # the staged kernel's range: force
    for R in (3, 7):
        monkeypatch.setenv("QR_FILL_ROWS_R", str(R))
        plan = make_op(labels, coeffs).plan()
        assert plan.n_groups == G and plan.fill_kernel == "fill_rows_kernel"
        assert_same(device_build(plan, 0, 1 << n), ref, f"G={G} REGT={REGT} Q={Q} R={R}")
        lo, hi = 100, 4000
        ip, ix, dt = device_build(plan, lo, hi)
        assert np.array_equal(ix, ref[1][lo * G:hi * G]) and np.array_equal(u64(dt), u64(ref[2][lo * G:hi * G]))
        assert np.array_equal(ip, np.arange(hi - lo + 1, dtype=np.uint64) * G)


@pytest.mark.parametrize("CL", [2, 4])
@pytest.mark.parametrize("R", [2, 7])
@pytest.mark.parametrize("name", ["random_n10", "H6", "tfim_3x3", "H2", "H4", "C1"])
def test_rows_kernel_clusters(fixtures, monkeypatch, name, R, CL):
    """Cluster variant of the rows kernel forced on small cases: the CTAs of a cluster split the groups and the rows of a
    4-row batch and exchange entries through distributed shared memory.  (Odd G with a cluster of 4 needs misaligned
    single-row copies: the library falls back to the one-CTA variant there -- H4, C1 -- which must still be right.)"""
    monkeypatch.setenv("QR_FILL_ROWS", "1")
    monkeypatch.setenv("QR_FILL_ROWS_CL", str(CL))
    monkeypatch.setenv("QR_FILL_ROWS_R", str(R))
    labels, coeffs = SMALL[name](fixtures)
    n, params = O.make_params(labels, coeffs)
    ref = O.build_csr(params, n)
    plan = make_op(labels, coeffs).plan()
    assert plan.fill_kernel == "fill_rows_kernel"
    G, dim = plan.n_groups, 1 << n
    assert_same(device_build(plan, 0, dim), ref, f"{name} CL={CL} R={R}")
    if dim >= 128:
        for lo, hi in [(6, dim - 2), (32, 64), (dim // 2 - 2, dim // 2 + 34)]:
            ip, ix, dt = device_build(plan, lo, hi, flags=_ffi.QR_INDPTR_GLOBAL)
            assert np.array_equal(ix, ref[1][lo * G:hi * G]) and np.array_equal(u64(dt), u64(ref[2][lo * G:hi * G])), (lo, hi)
            assert np.array_equal(ip, np.arange(lo, hi + 1, dtype=np.uint64) * G)


@pytest.mark.parametrize("exthv,dec", [("1", "1"), ("0", "1"), ("1", "0"), ("0", "0"), ("1", "2"), ("0", "2")])
@pytest.mark.parametrize("S", [32, 64, 1024])
@pytest.mark.parametrize("R", [1, 2, 7])
@pytest.mark.parametrize("name", ["random_n10", "H6", "tfim_3x3", "H2", "H4", "C1", "xxz_n10"])
def test_rows_kernel_split(fixtures, monkeypatch, name, R, S, exthv, dec):
    """Split mode of the rows kernel forced on small cases: the sorted masks are cut into trie subtrees of <= S groups, a
    CTA owns one subtree and writes its segment of every row (odd segment starts: shifted id buffer + edge stores).
    exthv = 1: the heavy groups' values come from heavy_values_kernel (the default), 0: folded inside the owning CTA.
    dec = 1: decoupled warps (the batch buffers handed over through mbarriers, the first warp to arrive gives the batch to the
    TMA; the split-mode default), 2: the same with the hand-over waiting for the TMA's read at once, 0: one CTA barrier per batch."""
    monkeypatch.setenv("QR_FILL_ROWS", "1")
    monkeypatch.setenv("QR_FILL_ROWS_EXTHV", exthv)
    monkeypatch.setenv("QR_FILL_ROWS_DEC", dec)
    monkeypatch.setenv("QR_FILL_ROWS_SPLIT", str(S))
    monkeypatch.setenv("QR_FILL_ROWS_R", str(R))
    labels, coeffs = SMALL[name](fixtures)
    n, params = O.make_params(labels, coeffs)
    ref = O.build_csr(params, n)
    plan = make_op(labels, coeffs).plan()
    assert plan.fill_kernel == "fill_rows_kernel"
    G, dim = plan.n_groups, 1 << n
    assert_same(device_build(plan, 0, dim), ref, f"{name} S={S} R={R}")
    if dim >= 128:
        for lo, hi in [(5, dim - 3), (32, 64), (dim // 2 - 1, dim // 2 + 34)]:
            ip, ix, dt = device_build(plan, lo, hi, flags=_ffi.QR_INDPTR_GLOBAL)
            assert np.array_equal(ix, ref[1][lo * G:hi * G]) and np.array_equal(u64(dt), u64(ref[2][lo * G:hi * G])), (lo, hi)
            assert np.array_equal(ip, np.arange(lo, hi + 1, dtype=np.uint64) * G)


@pytest.mark.parametrize("G,T,n", [(2600, 3900, 12), (3001, 3500, 12), (1500, 3600, 12), (2048, 4000, 12), (3998, 4090, 12), (5001, 9000, 13)])
def test_rows_kernel_split_large_G(G, T, n):
    """The shapes split mode is chosen for: more than 1024 groups with >= 2 terms per group on average, or rows too long
    for one CTA's shared memory (any G, odd ones included).  n = 12 / 13: 4096 / 8192 rows of up to 5001 entries."""
    labels, coeffs = H.random_pauli_sum(n, T, G, 50, 5)
    n, params = O.make_params(labels, coeffs)
    ref = O.build_csr(params, n)
    plan = make_op(labels, coeffs).plan()
    assert plan.n_groups == G and plan.fill_kernel == "fill_rows_kernel"
    assert_same(device_build(plan, 0, 1 << n), ref, f"G={G} T={T}")
    lo, hi = 101, 4000
    ip, ix, dt = device_build(plan, lo, hi)
    assert np.array_equal(ix, ref[1][lo * G:hi * G]) and np.array_equal(u64(dt), u64(ref[2][lo * G:hi * G]))
    assert np.array_equal(ip, np.arange(hi - lo + 1, dtype=np.uint64) * G)


@pytest.mark.parametrize("name,lo_hi", [("H10", [(1 << 19, (1 << 19) + 1024), ((1 << 19) + 77, (1 << 19) + 1500)]),
                                        ("H11", [(3 << 20, (3 << 20) + 512)])])
def test_split_external_heavy_on_molecular_operators(fixtures, name, lo_hi):
    """H10 / H11 (20 / 22 qubits, G = 2 536 / 3 792, the reference's perf_giant.py operators): split mode with the heavy
    groups (91 / 111 of them, up to 254 terms) folded by heavy_values_kernel -- row windows against the oracle, bit for bit."""
    if name not in fixtures:
        pytest.skip("fixture not in the golden file")
    labels, coeffs = fixtures[name]
    n, params = O.make_params(labels, coeffs)
    plan = make_op(labels, coeffs).plan()
    assert plan.fill_kernel == "fill_rows_kernel"
    G = plan.n_groups
    for lo, hi in lo_hi:
        ref = O.build_csr(params, n, lo, hi)
        ip, ix, dt = device_build(plan, lo, hi)
        assert np.array_equal(ix, ref[1]) and np.array_equal(u64(dt), u64(ref[2])), (lo, hi)
        assert np.array_equal(ip, np.arange(hi - lo + 1, dtype=np.uint64) * G)


def test_rows_kernel_selection(fixtures, monkeypatch):
    """Default choice: staged for rows that fit a 32-row tile, lanes for rows too long for two shared-memory
    batches, rows in between; QR_FILL_ROWS=0 restores the lanes kernel."""
    def kernel_of(labels, coeffs):
        return make_op(labels, coeffs).plan().fill_kernel
    assert kernel_of(*H.xxz_chain(10, 1.0, 0.7)) == "fill_staged_kernel"
    assert kernel_of(*H.tfim_lattice(5, 6, 1.0, 3.0)) == "fill_staged_kernel"      # G = 31
    assert kernel_of(*H.xxz_chain(23, 1.0, 0.7)) == "fill_staged_swz_kernel"       # G = 24: bank conflicts in the plain tile -> swizzled tile
    assert kernel_of(*H.xxz_chain(19, 1.0, 0.7)) == "fill_staged_swz_kernel"       # G = 20
    assert kernel_of(*H.xxz_chain(27, 1.0, 0.7)) == "fill_rows_kernel"             # G = 28: the rows kernel is ahead of both tiles
    assert kernel_of(*H.random_pauli_sum(12, 60, 40, 5, 5)) == "fill_rows_kernel"  # G = 40, no long group
    assert kernel_of(*H.random_pauli_sum(12, 30, 20, 5, 5)) == "fill_staged_swz_kernel"   # G = 20
    assert kernel_of(*H.random_pauli_sum(12, 30, 21, 5, 5)) == "fill_staged_kernel"
    assert kernel_of(*fixtures["H4"]) == "fill_staged_kernel"                      # G = 51 with heavy groups
    assert kernel_of(*fixtures["H6"]) == "fill_rows_kernel"                        # G = 286
    big = H.random_pauli_sum(12, 1500, 1100, 50, 5)
    assert kernel_of(*big) == "fill_rows_kernel"
    assert kernel_of(*H.random_pauli_sum(13, 3500, 3000, 50, 5)) == "fill_rows_kernel"     # split mode: a CTA per subtree of <= 1024 masks
    assert kernel_of(*H.random_pauli_sum(13, 5501, 5001, 50, 5)) == "fill_rows_kernel"
    monkeypatch.setenv("QR_FILL_ROWS_SPLIT", "0")
    assert kernel_of(*H.random_pauli_sum(13, 5501, 5001, 50, 5)) == "fill_lanes_kernel"    # 5001 * 96 B > 227 KB
    monkeypatch.delenv("QR_FILL_ROWS_SPLIT")
    monkeypatch.setenv("QR_FILL_ROWS", "0")
    assert kernel_of(*big) == "fill_lanes_kernel"


def test_large_G_default_is_one_launch(fixtures):
    """G = 286 rows do not fit the staged kernel's shared-memory tile: the default path must be one launch of a
    large-G kernel (rows; lanes with QR_FILL_ROWS=0), and it must agree with the oracle."""
    labels, coeffs = fixtures["H6"]
    n, params = O.make_params(labels, coeffs)
    ref = O.build_csr(params, n)
    plan = make_op(labels, coeffs).plan()
    before = _ffi.kernel_launches()
    assert_same(device_build(plan, 0, 1 << n), ref, "H6 default")
    assert _ffi.kernel_launches() - before == 1          # one fill kernel, no edge launches


def test_build_host_windows(fixtures):
    labels, coeffs = fixtures["H4"]
    n, params = O.make_params(labels, coeffs)
    full = O.build_csr(params, n)
    plan = make_op(labels, coeffs).plan()
    G = plan.n_groups
    for lo, hi in [(0, 256), (3, 250)]:
        rows = hi - lo
        ip = np.full(rows + 1, 0xFFFF, np.uint64); ix = np.zeros(rows * G, np.uint64); dt = np.zeros(rows * G, np.complex128)
        _ffi.call("qr_build_host", plan.handle, lo, hi, ip.ctypes.data, ix.ctypes.data, dt.ctypes.data, 0)
        assert np.array_equal(ip, np.arange(rows + 1, dtype=np.uint64) * G)
        assert np.array_equal(ix, full[1][lo * G:hi * G]) and np.array_equal(u64(dt), u64(full[2][lo * G:hi * G]))


@pytest.mark.parametrize("threads", ["1", "5"])
def test_build_host_pageable_destination(monkeypatch, threads):
    """Ordinary (pageable) numpy arrays as the destination -- what a Rust Vec is: several 32 MB windows
    through pinned staging and the host copy pool, against pinned destinations and the plain-cudaMemcpy
    path (QR_HOST_NO_STAGING), on ragged and global-indptr requests."""
    monkeypatch.setenv("QR_HOST_COPY_THREADS", threads)
    labels, coeffs = H.xxz_chain(18, 1.0, 0.7)                      # 113 MB: four staging windows
    n, params = O.make_params(labels, coeffs)
    plan = make_op(labels, coeffs).plan()
    G, dim = plan.n_groups, 1 << n
    for lo, hi, flags in [(0, dim, 0), (77, dim - 5, _ffi.QR_INDPTR_GLOBAL), (1000, 1032, 0)]:
        rows = hi - lo
        ref = O.build_csr(params, n, lo, hi)
        want_ip = ref[0] + (np.uint64(lo * G) if flags & _ffi.QR_INDPTR_GLOBAL else np.uint64(0))
        for extra in (0, _ffi.QR_HOST_NO_STAGING):
            ip = np.full(rows + 1, 0xFFFF, np.uint64); ix = np.full(rows * G, 0xFFFF, np.uint64)
            dt = np.full(rows * G, np.nan, np.complex128)
            _ffi.call("qr_build_host", plan.handle, lo, hi, ip.ctypes.data, ix.ctypes.data, dt.ctypes.data, flags | extra)
            assert_same((ip, ix, dt), (want_ip, ref[1], ref[2]), f"pageable [{lo},{hi}) flags={flags | extra}")


@pytest.mark.parametrize("name,wire_col,want_col", [("xxz16", None, 1), ("xxz16", "2", 2), ("xxz16", "4", 4), ("random_n14", None, 2),
                                                   ("random_n14", "4", 4), ("H4", None, 1)])
@pytest.mark.parametrize("pinned", [True, False])
def test_build_host_wire_forms(fixtures, monkeypatch, name, wire_col, want_col, pinned):
    """The compact wire form of qr_build_host (wire.cuh): group ids of 1 / 2 bytes or 32-bit columns cross PCIe, the host
    pool rebuilds the u64 columns and writes indptr itself.  The caller's arrays must equal the oracle's bit for bit, as they
    do with QR_HOST_WIDE (every array copied as stored), and the D2H byte count must be what the form promises."""
    labels, coeffs = {"xxz16": lambda: H.xxz_chain(16, 1.0, 0.7), "random_n14": lambda: H.random_pauli_sum(14, 400, 300, 30, 5),
                      "H4": lambda: fixtures["H4"]}[name]()
    n, params = O.make_params(labels, coeffs)
    plan = make_op(labels, coeffs).plan()
    G, dim = plan.n_groups, 1 << n
    if wire_col:
        monkeypatch.setenv("QR_HOST_WIRE_COL", wire_col)
    from qrusty_b200._runtime import pinned_empty
    for lo, hi, flags in [(0, dim, 0), (5, dim - 3, _ffi.QR_INDPTR_GLOBAL)]:
        rows = hi - lo
        ref = O.build_csr(params, n, lo, hi)
        want_ip = ref[0] + (np.uint64(lo * G) if flags & _ffi.QR_INDPTR_GLOBAL else np.uint64(0))
        for extra, col in ((0, want_col), (_ffi.QR_HOST_WIDE, 8)):
            alloc = pinned_empty if pinned else np.empty
            ip, ix, dt = alloc(rows + 1, np.uint64), alloc(rows * G, np.uint64), alloc(rows * G, np.complex128)
            ip[:] = 0xFFFF; ix[:] = 0xFFFF; dt[:] = np.nan
            _ffi.call("qr_build_host", plan.handle, lo, hi, ip.ctypes.data, ix.ctypes.data, dt.ctypes.data, flags | extra)
            assert_same((ip, ix, dt), (want_ip, ref[1], ref[2]), f"wire form {col} B/col [{lo},{hi}) pinned={pinned}")
            assert _ffi.last_d2h_bytes() == rows * G * (16 + col)


def test_abi_argument_errors(fixtures):
    plan = make_op(*fixtures["H2"]).plan()
    buf = DeviceBuffer(1 << 16)
    for lo, hi in [(5, 5), (7, 3), (0, 17)]:
        with pytest.raises(Q.QrustyCudaError) as e:
            _ffi.call("qr_build_rows_device", plan.handle, lo, hi, None, buf.ptr, buf.ptr, 0, None)
        assert e.value.code == _ffi.QR_ERR_INVALID
    with pytest.raises(Q.QrustyCudaError):
        _ffi.call("qr_build_rows_device", plan.handle, 0, 16, None, buf.ptr + 8, buf.ptr, 0, None)


# ---- K1: canonicalisation output against a stable sort on the host -----------------------------
@pytest.mark.parametrize("name", ["H2", "H6", "H8", "H10", "H11", "H12", "HJ", "C3"])
def test_canonicalise_groups(fixtures, name):
    labels, coeffs = H.CONFIGS["C3"][1]() if name == "C3" else fixtures[name]
    op = make_op(labels, coeffs)
    terms = op.terms()
    gx, goff, order = op.plan().groups()
    ref_order = np.argsort(terms["x"], kind="stable")
    assert np.array_equal(order, ref_order.astype(np.uint32))
    sx = terms["x"][ref_order]
    heads = np.flatnonzero(np.r_[True, sx[1:] != sx[:-1]])
    assert np.array_equal(gx, sx[heads])
    assert np.array_equal(goff, np.r_[heads, len(sx)].astype(np.uint32))


def test_merge_duplicates_opt_in(fixtures):
    """QR_PLAN_MERGE_DUPLICATES: identical (x, z) terms are merged by the (x,z)-keyed sort +
    segmented reduce; structure stays bit-exact, data agrees to 1e-12 (summation order changes)."""
    for labels, coeffs in (H.random_pauli_sum(10, 300, 200, 30, 7), H.random_pauli_sum(12, 900, 120, 400, 3),
                           fixtures["H4"], H.xxz_chain(10, 1.0, 0.7)):
        n, params = O.make_params(labels, coeffs)
        ref = O.build_csr(params, n)
        op = make_op(labels, coeffs)
        op.merge_duplicates = True
        plan = op.plan()
        keys = np.unique(np.stack([params["x"], params["z"]], axis=1), axis=0)
        assert plan.n_terms_canonical == len(keys) <= plan.n_terms
        gx, goff, order = plan.groups()
        assert np.array_equal(gx, np.unique(params["x"])) and goff[-1] == len(keys)
        shape, data, indices, indptr = op.to_matrix().export()
        assert np.array_equal(indptr, ref[0]) and np.array_equal(indices, ref[1])
        scale = np.abs(params["re"] + 1j * params["im"]).sum()
        assert np.abs(data - ref[2]).max() <= 1e-12 * scale
        v = H.lanczos_start_vector(0, 1 << n, seed=9)
        assert np.abs(op.apply(v) - O.spmv(*ref, v)).max() <= 1e-12 * scale * 2


# ---- full-size configs -------------------------------------------------------------------------
def test_C2_full_matrix():
    """BASELINE config 2 (XXZ periodic n=20, nnz = 22 020 096) compared in full."""
    labels, coeffs = H.xxz_chain(20, 1.0, 0.7)
    n, params = O.make_params(labels, coeffs)
    ref = O.build_csr(params, n, step=1000)
    shape, data, indices, indptr = make_op(labels, coeffs).to_matrix_mode("Cuda").export()
    assert (len(data), shape) == (22020096, (1 << 20, 1 << 20))
    assert_same((indptr, indices, data), ref, "C2")
    assert np.mean(data == 0) > 0.4            # XX+YY cancellation leaves explicit zeros, kept


def _sorted_cols(rows, gx):
    return np.sort(rows[:, None] ^ gx[None, :], axis=1).ravel()


@pytest.mark.parametrize("cfg,win", [("C3", 1 << 10), ("C4", 1 << 12), ("C5", 1 << 12)])
def test_big_config_windows(cfg, win):
    """C3/C4/C5 do not fit the host (or the GPU): row windows at sampled positions, indices also
    against the closed form sort_g(r ^ x_g)."""
    labels, coeffs = H.CONFIGS[cfg][1]()
    n, params = O.make_params(labels, coeffs)
    plan = make_op(labels, coeffs).plan()
    G = plan.n_groups
    gx = np.unique(params["x"])
    dim = 1 << n
    rng = np.random.default_rng(5)
    starts = [0, dim - win, dim // 2 - win // 2] + [int(s) for s in rng.integers(0, dim - win, 3)]
    for lo in starts:
        hi = lo + win
        ip, ix, dt = device_build(plan, lo, hi, _ffi.QR_INDPTR_GLOBAL)
        assert np.array_equal(ip, (np.arange(win + 1, dtype=np.uint64) + np.uint64(lo)) * np.uint64(G))
        assert np.array_equal(ix, _sorted_cols(np.arange(lo, hi, dtype=np.uint64), gx)), (cfg, lo)
        ref = O.build_csr(params, n, row_lo=lo, row_hi=hi, groups=G)
        assert np.array_equal(ix, ref[1]) and np.array_equal(u64(dt), u64(ref[2])), (cfg, lo)


def test_C4_full_size_properties():
    """All 2^25 rows of C4 (872 415 232 entries, 21 GB) built in 2^21-row windows; every window's
    indices equal the closed form, and a checksum of data is compared with the oracle on 8 windows."""
    labels, coeffs = H.tfim_lattice(5, 5, 1.0, 3.0)
    n, params = O.make_params(labels, coeffs)
    plan = make_op(labels, coeffs).plan()
    G, dim, win = plan.n_groups, 1 << n, 1 << 21
    gx = np.unique(params["x"])
    ip = DeviceBuffer((win + 1) * 8); ix = DeviceBuffer(win * G * 8); dt = DeviceBuffer(win * G * 16)
    hix = np.empty(win * G, np.uint64); hdt = np.empty(win * G, np.complex128)
    check_data = set(range(0, dim // win, 2))
    for w in range(dim // win):
        lo, hi = w * win, (w + 1) * win
        _ffi.call("qr_build_rows_device", plan.handle, lo, hi, ip.ptr, ix.ptr, dt.ptr, 0, None)
        ix.download(hix)
        r = np.arange(lo, hi, dtype=np.uint64)
        # columns of row r are {r ^ x_g}: XOR-sum and sum are order-free invariants; sortedness pins order
        cols = hix.reshape(win, G)
        assert (cols[:, 1:] > cols[:, :-1]).all()
        assert np.array_equal(np.bitwise_xor.reduce(cols, axis=1), np.bitwise_xor.reduce(r[:, None] ^ gx[None, :], axis=1))
        assert np.array_equal(cols.sum(axis=1, dtype=np.uint64), (r[:, None] ^ gx[None, :]).sum(axis=1, dtype=np.uint64))
        if w in check_data:
            dt.download(hdt)
            sub = slice(0, 4096 * G)
            ref = O.build_csr(params, n, row_lo=lo, row_hi=lo + 4096, groups=G)
            assert np.array_equal(hix[sub], ref[1]) and np.array_equal(u64(hdt[sub]), u64(ref[2]))
            # TFIM: diagonal (slot of mask 0) carries -J * sum of bond signs, off-diagonals are -h exactly
            vals = hdt.reshape(win, G)
            diag_slot = (cols == r[:, None])
            assert diag_slot.sum() == win
            assert np.all(vals[~diag_slot] == -3.0)


# ---- H.v -------------------------------------------------------------------------------------
@pytest.mark.parametrize("unroll", ["4", "1", "2", "8"])
@pytest.mark.parametrize("name", ["H2", "H4", "H6", "C1", "random_n10", "xxz_n10", "tfim_3x3"])
def test_spmv_ragged_rows(fixtures, monkeypatch, name, unroll):
    """spmat_dot_densevec on a compacted matrix (eliminate_zeros: ragged and empty rows) and on a row shard -- the
    thread-per-row kernel with 1 / 2 / 4 (default) / 8 entries of a row in flight, every form bit for bit the reference's
    sequential sums (accel.rs:338-370)."""
    monkeypatch.setenv("QR_SPMV_UNROLL", unroll)
    labels, coeffs = SMALL[name](fixtures)
    n, params = O.make_params(labels, coeffs)
    rng = np.random.default_rng(11)
    v = rng.uniform(-3, 3, 1 << n) + 1j * rng.uniform(-5, 5, 1 << n)
    tol = float(np.median(np.abs(O.build_csr(params, n)[2])))       # drops about half the entries
    want = O.eliminate_zeros(*O.build_csr(params, n), tolerance=tol)
    assert len(want[2]) < (1 << n) * len(set(params["x"].tolist())) and len(want[2]) > 0
    m = make_op(labels, coeffs).to_matrix_mode("Cuda").eliminate_zeros(tol)
    y = Q.spmat_dot_densevec(m, v)
    assert np.array_equal(u64(y), u64(O.spmv(*want, v)))
    dim = 1 << n
    if dim >= 512:
        lo, hi = 100, dim - 37                                      # a ragged shard: 256-row tiles with a short last one
        ref = O.build_csr(params, n, lo, hi)
        ms = make_op(labels, coeffs).to_matrix_rows(lo, hi)
        assert np.array_equal(u64(Q.spmat_dot_densevec(ms, v)), u64(O.spmv(*ref, v)))


@pytest.mark.parametrize("name", ["H2", "H6", "C1", "random_n10"])
def test_spmv_bit_exact(fixtures, name):
    """spmat_dot_densevec on the device CSR == accel.rs:338-370 bit for bit (lib.rs:887-919)."""
    labels, coeffs = SMALL[name](fixtures)
    n, params = O.make_params(labels, coeffs)
    ref = O.build_csr(params, n)
    rng = np.random.default_rng(3)
    v = rng.uniform(0, 10, 1 << n) + 1j * rng.uniform(-5, 5, 1 << n)
    m = make_op(labels, coeffs).to_matrix()
    y = Q.spmat_dot_densevec(m, v)
    assert np.array_equal(u64(y), u64(O.spmv(*ref, v)))
    csr = Q.csr_matrix(m)
    assert np.allclose(y, csr.dot(v))                              # test_it.py:218-230


@pytest.mark.parametrize("name", ["H2", "H6", "C1", "random_n10", "xxz_n10"])
def test_apply_matrix_free(fixtures, name):
    labels, coeffs = SMALL[name](fixtures)
    n, params = O.make_params(labels, coeffs)
    indptr, indices, data = O.build_csr(params, n)
    rng = np.random.default_rng(4)
    v = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    ref = O.spmv(indptr, indices, data, v)
    # north_star's contract for a value: |d - d_ref| <= 1e-12 * sum_{t in group} |c'_t| (the term-rich kernels fold a group's
    # terms in bucket order, so a value that cancels is accurate relative to the sum of its terms, not to itself)
    pa = params.copy(); pa["re"] = np.hypot(params["re"], params["im"]); pa["im"] = 0.0; pa["z"] = 0
    bound = O.spmv(*O.build_csr(pa, n), np.abs(v).astype(np.complex128)).real
    op = make_op(labels, coeffs)
    y = op.apply(v)
    assert np.all(np.abs(y - ref) <= 1e-12 * bound + 1e-300)
    assert np.array_equal(op.diagonal(), np.asarray(N.csr_to_dense(indptr, indices, data, 1 << n).diagonal())) if n <= 10 else True
    assert np.array_equal(op.to_matrix().diagonal(), op.diagonal())


def test_apply_C2_and_linearity():
    labels, coeffs = H.xxz_chain(20, 1.0, 0.7)
    n, params = O.make_params(labels, coeffs)
    op = make_op(labels, coeffs)
    dim = 1 << n
    v1 = H.lanczos_start_vector(0, dim, seed=25)
    v2 = H.lanczos_start_vector(0, dim, seed=26)
    y1, y2 = op.apply(v1), op.apply(v2)
    rows = np.random.default_rng(6).integers(0, dim, 4096).astype(np.uint64)
    ref = O.apply_rows(params, rows, v1)
    assert np.all(np.abs(y1[rows.astype(np.int64)] - ref) <= 1e-12 * 64 * np.abs(v1).max())
    a, b = 0.5 - 1.25j, -2.0 + 0.75j
    y12 = op.apply(a * v1 + b * v2)
    assert np.abs(y12 - (a * y1 + b * y2)).max() <= 1e-11 * np.abs(y12).max()
    # H is Hermitian: <v2, H v1> == conj(<v1, H v2>)
    assert abs(np.vdot(v2, y1) - np.conj(np.vdot(v1, y2))) <= 1e-9 * abs(np.vdot(v2, y1))


def _apply_block(plan, lo, hi, v_dev):
    dy = DeviceBuffer((hi - lo) * 16)
    _ffi.call("qr_memset_device", dy.ptr, 0xFF, dy.nbytes, None)
    _ffi.call("qr_apply_device", plan.handle, lo, hi, v_dev.ptr, dy.ptr, None)
    return dy.download(np.empty(hi - lo, np.complex128))


@pytest.mark.parametrize("name", ["xxz16", "tfim_4x4", "H8", "random_n14", "heis_n18"])
def test_apply_tiled_passes(fixtures, monkeypatch, name):
    """Multi-pass shared-memory H.v (apply_pass_kernel) against the v0 gather kernel and the oracle,
    on the whole vector and on row-sharded blocks (groups touching bits above the block go direct)."""
    labels, coeffs = {"xxz16": lambda: H.xxz_chain(16, 1.0, 0.7), "tfim_4x4": lambda: H.tfim_lattice(4, 4, 1.0, 3.0),
                      "H8": lambda: fixtures["H8"], "random_n14": lambda: H.random_pauli_sum(14, 300, 200, 30, 11),
                      "heis_n18": lambda: H.heisenberg_chain(18)}[name]()
    n, params = O.make_params(labels, coeffs)
    op = make_op(labels, coeffs)
    plan = op.plan()
    dim = 1 << n
    v = H.lanczos_start_vector(0, dim, seed=31)
    dv = DeviceBuffer(dim * 16); dv.upload(v)
    monkeypatch.setenv("QR_APPLY_V0", "1")
    y0 = _apply_block(plan, 0, dim, dv)
    monkeypatch.setenv("QR_APPLY_V0", "0")
    y1 = _apply_block(plan, 0, dim, dv)
    absH = np.abs(params["re"] + 1j * params["im"]).sum()
    tol = 1e-12 * absH * np.abs(v).max()
    assert np.abs(y1 - y0).max() <= tol
    rows = np.random.default_rng(12).integers(0, dim, 512).astype(np.uint64)
    ref = O.apply_rows(params, rows, v)
    assert np.abs(y1[rows.astype(np.int64)] - ref).max() <= tol
    for parts in (2, 4):
        blk = dim // parts
        ys = np.concatenate([_apply_block(plan, p * blk, (p + 1) * blk, dv) for p in range(parts)])
        assert np.abs(ys - y0).max() <= tol, parts


@pytest.mark.parametrize("name", ["xxz16", "tfim_4x4", "H8", "random_n14", "heis_n18"])
@pytest.mark.parametrize("cut", ["0", "10", "12"])
def test_apply_two_pass(fixtures, monkeypatch, name, cut):
    """Tiled H.v (apply_tile.cuh): a TMA-fed run of v per CTA serves the masks inside it, the rest is gathered (cut 0:
    that pass alone); with a cut bit a second pass adds the masks from that bit up out of shared-memory tiles spanning
    the top row bits.  Against the gather kernel and the oracle, on the whole vector and on aligned row blocks of a full
    vector (groups reaching outside the block pull their runs from the neighbouring blocks)."""
    labels, coeffs = {"xxz16": lambda: H.xxz_chain(16, 1.0, 0.7), "tfim_4x4": lambda: H.tfim_lattice(4, 4, 1.0, 3.0),
                      "H8": lambda: fixtures["H8"], "random_n14": lambda: H.random_pauli_sum(14, 300, 200, 30, 11),
                      "heis_n18": lambda: H.heisenberg_chain(18)}[name]()
    n, params = O.make_params(labels, coeffs)
    dim = 1 << n
    v = H.lanczos_start_vector(0, dim, seed=33)
    dv = DeviceBuffer(dim * 16); dv.upload(v)
    monkeypatch.setenv("QR_APPLY_TILE", "0")
    y0 = _apply_block(make_op(labels, coeffs).plan(), 0, dim, dv)
    monkeypatch.setenv("QR_APPLY_TILE", "1")
    monkeypatch.setenv("QR_APPLY_TILE_MIN", "9")
    monkeypatch.setenv("QR_APPLY_D", cut)
    plan = make_op(labels, coeffs).plan()
    _apply_block(plan, 0, dim, dv)                                   # builds the diag cache
    before = _ffi.kernel_launches()
    y1 = _apply_block(plan, 0, dim, dv)
    if name != "H8":                                                 # H8 (981 groups) is beyond the tile kernel's descriptor budget: gather
        assert _ffi.kernel_launches() - before == (1 if cut == "0" else 2)
    absH = np.abs(params["re"] + 1j * params["im"]).sum()
    tol = 1e-12 * absH * np.abs(v).max()
    assert np.abs(y1 - y0).max() <= tol
    rows = np.random.default_rng(13).integers(0, dim, 512).astype(np.uint64)
    ref = O.apply_rows(params, rows, v)
    assert np.abs(y1[rows.astype(np.int64)] - ref).max() <= tol
    for parts in (2, 4):
        blk = dim // parts
        ys = np.concatenate([_apply_block(plan, p * blk, (p + 1) * blk, dv) for p in range(parts)])
        assert np.abs(ys - y0).max() <= tol, parts


FOLD_OPS = {
    "H6": lambda fx: fx["H6"], "H8": lambda fx: fx["H8"],
    "random_n14": lambda fx: H.random_pauli_sum(14, 300, 200, 30, 11),           # complex coefficients, long groups
    "random_n13_rich": lambda fx: H.random_pauli_sum(13, 4000, 150, 40, 5),       # ~27 terms per group
    "xxz16": lambda fx: H.xxz_chain(16, 1.0, 0.7), "tfim_4x4": lambda fx: H.tfim_lattice(4, 4, 1.0, 3.0),
}


@pytest.mark.parametrize("name", list(FOLD_OPS))
def test_apply_fold(fixtures, monkeypatch, name):
    """Term-rich H.v (apply_fold.cuh: terms bucketed by the thread's three row bits, in-register Walsh-Hadamard
    fold) against the gather kernel and the oracle, forced onto every operator shape: whole vector, aligned row
    blocks of a full vector, and a window that is not a whole number of CTAs (must fall back to the gather kernel)."""
    labels, coeffs = FOLD_OPS[name](fixtures)
    n, params = O.make_params(labels, coeffs)
    dim = 1 << n
    v = H.lanczos_start_vector(0, dim, seed=41)
    dv = DeviceBuffer(dim * 16); dv.upload(v)
    monkeypatch.setenv("QR_APPLY_PTILE", "0")
    monkeypatch.setenv("QR_APPLY_FOLD", "0")
    plan0 = make_op(labels, coeffs).plan()
    assert plan0.apply_kernel() == "apply_direct_kernel"
    y0 = _apply_block(plan0, 0, dim, dv)
    monkeypatch.setenv("QR_APPLY_FOLD", "1")
    plan = make_op(labels, coeffs).plan()
    assert plan.apply_kernel() == "apply_fold_kernel"
    assert plan.apply_kernel(0, dim - 512) == "apply_direct_kernel"
    y1 = _apply_block(plan, 0, dim, dv)
    absH = np.abs(params["re"] + 1j * params["im"]).sum()
    tol = 1e-12 * absH * np.abs(v).max()
    assert np.abs(y1 - y0).max() <= tol
    rows = np.random.default_rng(14).integers(0, dim, 512).astype(np.uint64)
    ref = O.apply_rows(params, rows, v)
    assert np.abs(y1[rows.astype(np.int64)] - ref).max() <= tol
    assert np.array_equal(u64(_apply_block(plan, 0, dim, dv)), u64(y1))            # deterministic
    for parts in (2, 4):
        blk = dim // parts
        ys = np.concatenate([_apply_block(plan, p * blk, (p + 1) * blk, dv) for p in range(parts)])
        assert np.array_equal(u64(ys), u64(y1)), parts                              # a row's value does not depend on the block it is in
    lo, hi = 1024, dim - 512
    assert np.abs(_apply_block(plan, lo, hi, dv) - y0[lo:hi]).max() <= tol


@pytest.mark.parametrize("name", list(FOLD_OPS))
@pytest.mark.parametrize("K,nbuf", [("10", "2"), ("10", "4"), ("11", "3"), ("12", "3"), ("12", "2")])
def test_apply_ptile(fixtures, monkeypatch, name, K, nbuf):
    """Partner-tile H.v (apply_fold.cuh, K4c): one TMA bulk load of the partner tile per segment of groups sharing
    x >> K, values as the fold kernel's -- bit-identical to it, within tolerance of the gather kernel and the oracle;
    whole vector and aligned row blocks of a full vector."""
    labels, coeffs = FOLD_OPS[name](fixtures)
    n, params = O.make_params(labels, coeffs)
    dim = 1 << n
    v = H.lanczos_start_vector(0, dim, seed=43)
    dv = DeviceBuffer(dim * 16); dv.upload(v)
    monkeypatch.setenv("QR_APPLY_PTILE", "0")
    monkeypatch.setenv("QR_APPLY_FOLD", "0")
    y0 = _apply_block(make_op(labels, coeffs).plan(), 0, dim, dv)
    monkeypatch.setenv("QR_APPLY_FOLD", "1")
    yf = _apply_block(make_op(labels, coeffs).plan(), 0, dim, dv)
    monkeypatch.setenv("QR_APPLY_PTILE", "1")
    monkeypatch.setenv("QR_APPLY_PTILE_K", K)
    monkeypatch.setenv("QR_APPLY_PTILE_NBUF", nbuf)
    plan = make_op(labels, coeffs).plan()
    assert plan.apply_kernel() == "apply_ptile_kernel"
    assert plan.apply_kernel(0, dim - 512) != "apply_ptile_kernel"
    y1 = _apply_block(plan, 0, dim, dv)
    assert np.array_equal(u64(y1), u64(yf))                                        # same arithmetic as the fold kernel
    absH = np.abs(params["re"] + 1j * params["im"]).sum()
    tol = 1e-12 * absH * np.abs(v).max()
    assert np.abs(y1 - y0).max() <= tol
    rows = np.random.default_rng(15).integers(0, dim, 512).astype(np.uint64)
    assert np.abs(y1[rows.astype(np.int64)] - O.apply_rows(params, rows, v)).max() <= tol
    for parts in (2, 4):
        blk = dim // parts
        ys = np.concatenate([_apply_block(plan, p * blk, (p + 1) * blk, dv) for p in range(parts)])
        assert np.array_equal(u64(ys), u64(y1)), parts


def test_apply_fold_is_the_default_for_term_rich_operators(fixtures, monkeypatch):
    monkeypatch.delenv("QR_APPLY_FOLD", raising=False)
    assert make_op(*fixtures["H8"]).plan().apply_kernel() == "apply_fold_kernel"
    assert make_op(*H.xxz_chain(16, 1.0, 0.7)).plan().apply_kernel() == "apply_direct_kernel"
    assert make_op(*H.tfim_lattice(4, 4, 1.0, 3.0)).plan().apply_kernel() == "apply_direct_kernel"
    assert make_op(*H.heisenberg_chain(18)).plan().apply_kernel() == "apply_direct_kernel"
    assert make_op(*fixtures["H2"]).plan().apply_kernel() == "apply_direct_kernel"   # 16 rows: less than one CTA of the fold kernel


# ---- accel.rs:374-393 (test_it.py:232-269, lib.rs:921-990) ------------------------------------
def test_vector_ops_bit_exact():
    rng = np.random.default_rng(8)
    for rows in (4, 1 << 20):
        x = rng.random(rows) + rng.random(rows) * 1j
        y = rng.random(rows) + rng.random(rows) * 1j
        a, b = 1.0 + 2.0j, 3.0 + 4.0j
        assert np.array_equal(u64(Q.axpby(a, x, b, y)), u64(O.axpby(a, x, b, y)))
        assert np.array_equal(u64(Q.axpy(a, x, y)), u64(O.axpy(a, x, y)))
        assert np.array_equal(u64(Q.ax(a, x)), u64(O.ax(a, x)))
        assert np.allclose(Q.axpby(a, x, b, y), a * x + b * y)
        assert np.allclose(Q.axpy(1j, x, y), 1j * x + y) and np.allclose(Q.ax(1j, x), 1j * x)


def test_precond_bit_exact(fixtures, precond_vectors):
    """precond / precond2 (pyqrusty/src/lib.rs:436-468; test_it.py:303-313) against the oracle."""
    PRECOND_DX, PRECOND_E, PRECOND_RV = precond_vectors
    labels, coeffs = fixtures["H2"]
    m = make_op(labels, coeffs).to_matrix_mode("Cuda")
    diag = m.diagonal()
    got = Q.precond2(diag, PRECOND_DX, PRECOND_E, 1e-14)
    assert np.array_equal(u64(got), u64(O.precond2(diag, PRECOND_DX, PRECOND_E, 1e-14)))
    assert np.allclose(got, PRECOND_RV, rtol=1e-8, atol=0)
    assert np.array_equal(u64(Q.precond(m, PRECOND_DX, PRECOND_E, 1e-14)), u64(got))
    rng = np.random.default_rng(3)
    n = (1 << 20) + 5
    d = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    dx = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    d[::7] = 0.25 + 0j                                    # denominators that hit reg()
    e = 0.25 + 1e-9j
    assert np.array_equal(u64(Q.precond2(d, dx, e, 1e-7)), u64(O.precond2(d, dx, e, 1e-7)))
    m.export()
    with pytest.raises(Exception, match="exported"):
        Q.precond(m, PRECOND_DX, PRECOND_E, 1e-14)


def test_lanczos_matches_numpy(fixtures):
    """The device Lanczos loop (BASELINE config 4's measurement loop) against numpy on the oracle's CSR."""
    import scipy.sparse as sps
    from qrusty_b200 import lanczos as L
    labels, coeffs = H.tfim_lattice(3, 4, 1.0, 3.0)                  # 12 qubits, Hermitian
    n, params = O.make_params(labels, coeffs)
    indptr, indices, data = O.build_csr(params, n)
    A = sps.csr_matrix((data, indices.astype(np.int64), indptr.astype(np.int64)), shape=(1 << n, 1 << n))
    res = L.lanczos(make_op(labels, coeffs), n_iter=30)                  # two passes, scalars on the device
    old = L.lanczos(make_op(labels, coeffs), n_iter=30, device_scalars=False)   # round-1 loop: four passes, host scalars
    assert res["passes_per_iteration"] == 2
    assert np.allclose(res["alphas"], old["alphas"], rtol=1e-10, atol=1e-10) and np.allclose(res["betas"], old["betas"], rtol=1e-10, atol=1e-10)
    v = H.lanczos_start_vector(0, 1 << n); v /= np.linalg.norm(v)
    v_prev, beta, al, be = np.zeros_like(v), 0.0, [], []
    for _ in range(30):
        w = A @ v
        a = np.vdot(v, w).real
        w = w - a * v - beta * v_prev
        beta = np.linalg.norm(w)
        al.append(a); be.append(beta)
        v_prev, v = v, w / beta
    assert np.allclose(res["alphas"], al, rtol=1e-9, atol=1e-9) and np.allclose(res["betas"], be, rtol=1e-9, atol=1e-9)
    exact = np.linalg.eigvalsh(A.toarray())[0]
    assert L.ritz_values(res["alphas"], res["betas"])[0] >= exact - 1e-9
    assert res["hv_ms"] > 0 and res["iterations"] == 30


@pytest.mark.parametrize("name", ["C1", "H6", "xxz16", "random_n10"])
def test_apply_dot(fixtures, name):
    """qr_apply_dot_device: y is the plain apply's bit for bit, the fused <v, H v> equals the vdot of the downloaded
    vectors to rounding (both apply kernels: gather and fold; whole vector and a row block)."""
    labels, coeffs = H.xxz_chain(16, 1.0, 0.7) if name == "xxz16" else SMALL[name](fixtures)
    plan = make_op(labels, coeffs).plan()
    dim = plan.dim
    v = H.lanczos_start_vector(0, dim, seed=51)
    dv = DeviceBuffer(dim * 16); dv.upload(v)
    for lo, hi in [(0, dim), (dim // 4, dim // 2), (3, dim - 5)]:
        y0 = _apply_block(plan, lo, hi, dv)
        dy = DeviceBuffer((hi - lo) * 16); dd = DeviceBuffer(16)
        _ffi.call("qr_apply_dot_device", plan.handle, lo, hi, dv.ptr, dy.ptr, dd.ptr, None)
        y1 = dy.download(np.empty(hi - lo, np.complex128))
        dot = dd.download(np.empty(1, np.complex128))[0]
        assert np.array_equal(u64(y1), u64(y0)), (lo, hi)
        want = np.vdot(v[lo:hi], y0)
        assert abs(dot - want) <= 1e-12 * np.abs(v[lo:hi]).dot(np.abs(y0)) + 1e-300, (lo, hi, dot, want)
        dot2 = DeviceBuffer(16)
        _ffi.call("qr_apply_dot_device", plan.handle, lo, hi, dv.ptr, dy.ptr, dot2.ptr, None)
        assert np.array_equal(u64(dot2.download(np.empty(1, np.complex128))), u64(np.array([dot])))   # deterministic


def test_dotc():
    rng = np.random.default_rng(9)
    n = (1 << 18) + 17
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    y = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    dx, dy, out = DeviceBuffer(n * 16), DeviceBuffer(n * 16), DeviceBuffer(16)
    dx.upload(x); dy.upload(y)
    _ffi.call("qr_dotc_device", n, dx.ptr, dy.ptr, out.ptr, None)
    got = out.download(np.empty(1, np.complex128))[0]
    assert abs(got - np.vdot(x, y)) <= 1e-10 * np.abs(x).sum()


# ---- K2: count + scan + compaction (test_it.py:141-148, util.rs:144-171) -----------------------
@pytest.mark.parametrize("resident", [False, True])
@pytest.mark.parametrize("name", ["H2", "H4", "H6", "C1", "random_n10", "xxz_n10"])
def test_eliminate_zeros(fixtures, name, resident):
    """resident=False: the fused drop-zeros build (count_rows_kernel + scan + fill_compact_kernel) on a
    matrix that was never written; resident=True: count + scan + compaction of the built shard."""
    labels, coeffs = SMALL[name](fixtures)
    n, params = O.make_params(labels, coeffs)
    ref = O.build_csr(params, n)
    for tol in (1e-7, 0.0, 0.05):
        want = O.eliminate_zeros(*ref, tolerance=tol)
        m = make_op(labels, coeffs).to_matrix()
        if resident:
            m.to_device()
        assert m.count_zeros(tol) == O.count_zeros(ref[2], tol)
        m2 = m.eliminate_zeros(tol)
        assert m2.nnz() == len(want[2]) and m2.shape() == (1 << n, 1 << n)
        assert m2.count_zeros(tol) == 0
        v = H.lanczos_start_vector(0, 1 << n, seed=3)
        y2 = Q.spmat_dot_densevec(m2, v)
        assert np.array_equal(u64(y2), u64(O.spmv(*want, v)))
        shape, data, indices, indptr = m2.export()
        assert_same((indptr, indices, data), want, f"{name} tol={tol}")
        assert m.nnz() == len(ref[2])                      # the original is untouched


@pytest.mark.parametrize("count_rows", [False, True, "3"])
@pytest.mark.parametrize("win_mb", [None, "1"])
@pytest.mark.parametrize("n,T,G", [(10, 900, 400), (11, 1500, 700), (11, 2400, 1100)])
def test_eliminate_zeros_long_rows(monkeypatch, n, T, G, win_mb, count_rows):
    """Fused drop-zeros build when a row is too long for a shared-memory tile (G > 302): row windows built by the fill
    kernels into the per-device scratch and compacted by compact_rows_kernel; counted by count_rows_kernel (default) or,
    window by window, by count_kept_kernel -- one window and many (1 MB windows: the counts of neighbouring windows must not clobber each other)."""
    if win_mb:
        monkeypatch.setenv("QR_COMPACT_WIN_MB", win_mb)
    if not count_rows:
        monkeypatch.setenv("QR_COMPACT_COUNT_WINDOWED", "1")
    elif count_rows == "3":                                         # count_rows_kernel with the groups cut into 3 slices (atomic counts)
        monkeypatch.setenv("QR_COUNT_ROWS_SLICES", "3")
    labels, coeffs = H.random_pauli_sum(n, T, G, 50, 9)
    nq, params = O.make_params(labels, coeffs)
    ref = O.build_csr(params, nq)
    tol = float(np.median(np.abs(ref[2])))
    want = O.eliminate_zeros(*ref, tolerance=tol)
    assert 0 < len(want[2]) < len(ref[2])
    m = make_op(labels, coeffs).to_matrix()
    assert m.count_zeros(tol) == O.count_zeros(ref[2], tol)
    m2 = m.eliminate_zeros(tol)
    shape, data, indices, indptr = m2.export()
    assert_same((indptr, indices, data), want, f"G={G} win={win_mb} count_rows={count_rows}")


def test_eliminate_zeros_tiny_values():
    """test_it.py:141-148: a matrix of 1e-8 entries has 2 'zeros'; eliminating them leaves nothing."""
    m = Q.SparsePauliOp([Q.Pauli("I")], [1e-8 + 0j]).to_matrix()
    assert m.count_zeros() == 2
    m2 = m.eliminate_zeros()
    assert m2.count_zeros() == 0 and m2.nnz() == 0
    shape, data, indices, indptr = m2.export()
    assert len(data) == 0 and np.array_equal(indptr, [0, 0, 0])


def test_eliminate_zeros_any_csr(fixtures):
    """util.rs:144-171 take any CsMat: a matrix wrapped by new_unchecked (ragged rows, empty rows, a global indptr
    offset is not involved) and a matrix that is already compacted go through the indptr-driven kernels."""
    rng = np.random.default_rng(21)
    n_rows, n_cols = 300, 517
    lens = rng.integers(0, 70, n_rows); lens[[0, 17, 299]] = 0; lens[5] = 200
    indptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    indices = np.concatenate([np.sort(rng.choice(n_cols, int(k), replace=False)) for k in lens]).astype(np.uint64)
    data = rng.standard_normal(len(indices)) + 1j * rng.standard_normal(len(indices))
    small = rng.random(len(indices)) < 0.3
    data[small] *= 1e-9
    data[rng.random(len(indices)) < 0.05] = 0.0
    m = Q.SpMat.new_unchecked((n_rows, n_cols), data, indices, indptr)
    for tol in (1e-7, 0.0, 0.5):
        want = O.eliminate_zeros(indptr, indices, data, tolerance=tol)
        assert m.count_zeros(tol) == O.count_zeros(data, tol)
        m2 = m.eliminate_zeros(tol)
        assert m2.nnz() == len(want[2]) and m2.count_zeros(tol) == 0
        # ... and once more on the compacted matrix, with a larger tolerance
        want3 = O.eliminate_zeros(*want, tolerance=1.0)
        m3 = m2.eliminate_zeros(1.0)
        assert m2.count_zeros(1.0) == O.count_zeros(want[2], 1.0)
        assert_same(m3.export()[:0:-1], want3, f"twice tol={tol}")
        assert_same(m2.export()[:0:-1], want, f"tol={tol}")
    # a plan-built matrix, compacted, then compacted again
    labels, coeffs = SMALL["random_n10"](fixtures)
    n, params = O.make_params(labels, coeffs)
    ref = O.eliminate_zeros(*O.build_csr(params, n), tolerance=0.3)
    mm = make_op(labels, coeffs).to_matrix().eliminate_zeros(0.3)
    assert mm.count_zeros(0.9) == O.count_zeros(ref[2], 0.9)
    assert_same(mm.eliminate_zeros(0.9).export()[:0:-1], O.eliminate_zeros(*ref, tolerance=0.9), "plan-built, twice")


@pytest.mark.parametrize("name,lo,hi", [("H4", 3, 250), ("H4", 32, 96), ("xxz_n10", 1, 1023), ("H6", 100, 3001),
                                        ("H8", 4096 + 7, 8192 + 100)])
def test_fused_drop_zeros_row_windows(fixtures, name, lo, hi):
    """qr_build_compact_* on ragged row windows (first/last tile partly outside the window); H8's rows
    (G = 981) do not fit a shared-memory tile and take the windowed build + compaction inside the library."""
    labels, coeffs = fixtures[name] if name in fixtures else SMALL[name](fixtures)
    n, params = O.make_params(labels, coeffs)
    G = len(np.unique(params["x"]))
    ref = O.build_csr(params, n, lo, hi)
    ref = (ref[0] - ref[0][0], ref[1], ref[2])
    for tol in (1e-7, 0.02):
        want = O.eliminate_zeros(*ref, tolerance=tol)
        m = make_op(labels, coeffs).to_matrix_rows(lo, hi)
        assert m.count_zeros(tol) == O.count_zeros(ref[2], tol)
        shape, data, indices, indptr = m.eliminate_zeros(tol).export()
        assert shape == (hi - lo, 1 << n)
        assert_same((indptr, indices, data), want, f"{name}[{lo},{hi}) tol={tol}")


def test_eliminate_zeros_C2_full():
    labels, coeffs = H.xxz_chain(20, 1.0, 0.7)
    n, params = O.make_params(labels, coeffs)
    want = O.eliminate_zeros(*O.build_csr(params, n))
    shape, data, indices, indptr = make_op(labels, coeffs).to_matrix().eliminate_zeros().export()
    assert_same((indptr, indices, data), want, "C2 compacted")
    assert len(data) < 22020096 * 0.6


# ---- on-disk formats (rawio.rs:128-179, pyqrusty/src/lib.rs:216-259) ----------------------------
@pytest.mark.parametrize("name,lo,hi", [("H4", 0, 256), ("H6", 0, 4096), ("xxz_n10", 100, 901), ("C1", 0, 4096)])
def test_rawio_streamed_write(fixtures, tmp_path, name, lo, hi):
    """qr_write_rawio (fill -> pinned staging -> pwrite, row windows) byte for byte against the
    restatement of rawio::write applied to the oracle's CSR; then read back."""
    labels, coeffs = fixtures[name] if name in fixtures else SMALL[name](fixtures)
    n, params = O.make_params(labels, coeffs)
    ref = O.build_csr(params, n, lo, hi)
    op = make_op(labels, coeffs)
    m = op.to_matrix() if (lo, hi) == (0, 1 << n) else op.to_matrix_rows(lo, hi)
    path = tmp_path / "m.rawio"
    m.rawio_write(path)
    want = O.rawio_bytes((hi - lo, 1 << n), *ref)
    assert path.read_bytes() == want
    back = Q.SpMat.rawio_read(path)
    shape, data, indices, indptr = back.export()
    assert shape == (hi - lo, 1 << n)
    assert_same((indptr, indices, data), ref, name + " rawio round trip")
    # a resident (already built) matrix takes the export path: same bytes
    m2 = op.to_matrix_rows(lo, hi).to_device()
    m2.rawio_write(path)
    assert path.read_bytes() == want
    # a file written on a machine of the other endianness (rawio.rs:150-152: need_swab)
    import sys
    path.write_bytes(O.rawio_bytes((hi - lo, 1 << n), *ref, byteorder=">" if sys.byteorder == "little" else "<"))
    shape, data, indices, indptr = Q.SpMat.rawio_read(path).export()
    assert_same((indptr, indices, data), ref, name + " rawio swabbed")


def test_rawio_many_windows(tmp_path):
    """A matrix larger than the writer's 64 MB staging window: several windows, two streams."""
    labels, coeffs = H.xxz_chain(18, 1.0, 0.7)
    n, params = O.make_params(labels, coeffs)
    ref = O.build_csr(params, n)
    path = tmp_path / "xxz18.rawio"
    make_op(labels, coeffs).to_matrix().rawio_write(path)
    assert hashlib.sha256(path.read_bytes()).hexdigest() == hashlib.sha256(O.rawio_bytes((1 << n, 1 << n), *ref)).hexdigest()


def test_matrixmarket_round_trip(fixtures, tmp_path):
    labels, coeffs = fixtures["H4"]
    n, params = O.make_params(labels, coeffs)
    want = O.eliminate_zeros(*O.build_csr(params, n), tolerance=0.0)       # a reader drops nothing, but
    m = make_op(labels, coeffs).to_matrix().eliminate_zeros(0.0)           # scipy's mmread sums/drops zeros
    path = tmp_path / "h4.mtx"
    m.matrixmarket_write(path)
    assert path.read_text().splitlines()[0] == "%%MatrixMarket matrix coordinate complex general"
    shape, data, indices, indptr = Q.SpMat.matrixmarket_read(path).export()
    assert_same((indptr, indices, data), want, "H4 MatrixMarket round trip")


def test_multi_gpu_single_process(fixtures):
    if _ffi.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    labels, coeffs = fixtures["H6"]
    n, params = O.make_params(labels, coeffs)
    ref = O.build_csr(params, n)
    shape, data, indices, indptr = make_op(labels, coeffs).to_matrix_mode("Cuda/2").export()
    assert_same((indptr, indices, data), ref, "Cuda/2")


@pytest.mark.parametrize("n,n_masks", [(32, 20), (32, 300), (31, 24), (30, 1100)])
def test_row_windows_at_the_top_of_32_bit_row_space(n, n_masks):
    """n = 30..32 qubits: windows that end at row 2^n, straddle 2^31, or sit in the middle -- row ids and
    column ids use every bit of a u32, offsets are 64-bit.  Staged (G = 20, 24), lanes (G = 300), rows (G = 1100;
    its misaligned windows fall through to the lanes kernel)."""
    labels, coeffs = H.random_pauli_sum(n, n_masks + n_masks // 2, n_masks, 5, 1000 + n_masks)
    nq, params = O.make_params(labels, coeffs)
    plan = make_op(labels, coeffs).plan()
    G, dim = plan.n_groups, 1 << n
    assert G == n_masks
    for lo, hi in [(dim - 4096, dim), (dim // 2 - 100, dim // 2 + 229), (dim - 33, dim), (3 * (dim // 4) + 5, 3 * (dim // 4) + 2053)]:
        ref = O.build_csr(params, nq, lo, hi)
        ip, ix, dt = device_build(plan, lo, hi, flags=_ffi.QR_INDPTR_GLOBAL)
        assert np.array_equal(ix, ref[1]) and np.array_equal(u64(dt), u64(ref[2])), (lo, hi)
        assert np.array_equal(ip, np.arange(lo, hi + 1, dtype=np.uint64) * np.uint64(G))
    # the fused drop-zeros build on the last rows
    lo, hi = dim - 2048, dim
    ref = O.build_csr(params, nq, lo, hi)
    want = O.eliminate_zeros(ref[0] - ref[0][0], ref[1], ref[2], tolerance=0.3)
    shape, data, indices, indptr = make_op(labels, coeffs).to_matrix_rows(lo, hi).eliminate_zeros(0.3).export()
    assert_same((indptr, indices, data), want, "drop-zeros at the top rows")


def test_spmat_scale_and_strategy_aliases(fixtures):
    """SpMat.scale (pyqrusty/src/lib.rs:158-164) in place on the device, and the per-strategy to_matrix_*
    entry points (lib.rs:386-404), which all denote the same matrix."""
    labels, coeffs = fixtures["H4"]
    n, params = O.make_params(labels, coeffs)
    ref = O.build_csr(params, n)
    op = make_op(labels, coeffs)
    m = op.to_matrix()
    m.scale(0.5 - 2j)
    shape, data, indices, indptr = m.export()
    assert np.array_equal(indices, ref[1]) and np.array_equal(indptr, ref[0])
    assert np.array_equal(u64(data), u64(O.ax(0.5 - 2j, ref[2])))
    with pytest.raises(Exception, match="exported"):
        m.scale(2.0)
    for build in (op.to_matrix_binary, op.to_matrix_accel, op.to_matrix_reduce, op.to_matrix_rayon,
                  lambda: op.to_matrix_rayon_chunked(100)):
        shape, data, indices, indptr = build().export()
        assert_same((indptr, indices, data), ref, "strategy alias")


@pytest.mark.parametrize("name", ["H4", "xxz_n10", "C1"])
def test_diagonal_follows_the_stored_matrix(fixtures, name):
    """SpMat.diagonal / precond read the STORED matrix (pyqrusty/src/lib.rs:118-125, 442-455): after scale() the scaled
    diagonal, after eliminate_zeros() zero where the entry was dropped, on a row-window shard M[i,i] = H[lo+i, i]."""
    labels, coeffs = SMALL[name](fixtures)
    n, params = O.make_params(labels, coeffs)
    indptr, indices, data = O.build_csr(params, n)
    dim = 1 << n
    import scipy.sparse
    dense_diag = scipy.sparse.csr_matrix((data, indices.astype(np.int64), indptr.astype(np.int64)), shape=(dim, dim)).diagonal()
    op = make_op(labels, coeffs)
    m = op.to_matrix()
    assert np.array_equal(m.diagonal(), np.asarray(dense_diag))                             # lazy shard: from the plan (scipy drops the sign of a stored -0.0)
    f = 0.5 - 2j
    m.scale(f)
    want = O.ax(f, np.asarray(dense_diag, dtype=np.complex128))
    assert np.array_equal(m.diagonal(), want)                                               # resident, scaled
    dx = np.random.default_rng(3).standard_normal(dim) + 0j
    assert np.array_equal(Q.precond(m, dx, 0.25 + 0j, 1e-8), O.precond2(want, dx, 0.25 + 0j, 1e-8))
    tol = float(np.median(np.abs(dense_diag))) if np.any(dense_diag) else 1e-7             # drops about half of the diagonal
    z = op.to_matrix().eliminate_zeros(tol)
    kept = np.where(np.hypot(dense_diag.real, dense_diag.imag) > tol, dense_diag, 0)
    assert np.array_equal(z.diagonal(), kept)
    lo, hi = dim // 4, dim // 2                                                             # rows [lo,hi): M[i,i] = H[lo+i, i]
    sh = op.to_matrix_rows(lo, hi)
    ref_rows = scipy.sparse.csr_matrix((data, indices.astype(np.int64), indptr.astype(np.int64)), shape=(dim, dim))[lo:hi]
    assert np.array_equal(sh.diagonal(), ref_rows.diagonal())


def test_vector_length_checks(fixtures):
    labels, coeffs = fixtures["H2"]
    m = make_op(labels, coeffs).to_matrix()
    with pytest.raises(Exception, match="columns"):
        Q.spmat_dot_densevec(m, np.zeros(15, complex))
    with pytest.raises(Exception, match="differ in length"):
        Q.axpby(1.0, np.zeros(4, complex), 2.0, np.zeros(5, complex))
    with pytest.raises(Exception, match="differ in length"):
        Q.axpy(1.0, np.zeros(4, complex), np.zeros(5, complex))
    with pytest.raises(Exception, match="wrong length"):
        make_op(labels, coeffs).apply(np.zeros(15, complex))


def test_graph_capture_replay(fixtures):
    """qr_graph_*: the canonicalise -> fill sequence recorded once, replayed into zeroed buffers; event
    records inside the capture become graph nodes whose timestamps are readable after the replay."""
    labels, coeffs = fixtures["H4"]
    n, params = O.make_params(labels, coeffs)
    ref = O.build_csr(params, n)
    plan = make_op(labels, coeffs).plan()
    G, dim = plan.n_groups, 1 << n
    ip, ix, dt = DeviceBuffer((dim + 1) * 8), DeviceBuffer(dim * G * 8), DeviceBuffer(dim * G * 16)
    st, g = C.c_void_p(), C.c_void_p()
    e0, e1 = C.c_void_p(), C.c_void_p()
    _ffi.call("qr_stream_create", C.byref(st))
    _ffi.call("qr_event_create", C.byref(e0)); _ffi.call("qr_event_create", C.byref(e1))
    _ffi.call("qr_graph_begin_capture", st)
    _ffi.call("qr_plan_canonicalise_async", plan.handle, st)
    _ffi.call("qr_event_record", e0, st)
    _ffi.call("qr_build_rows_device", plan.handle, 0, dim, ip.ptr, ix.ptr, dt.ptr, 0, st)
    _ffi.call("qr_event_record", e1, st)
    _ffi.call("qr_graph_end_capture", st, C.byref(g))
    for b in (ip, ix, dt):
        _ffi.call("qr_memset_device", b.ptr, 0xFF, b.nbytes, st)       # nothing ran during the capture
    _ffi.call("qr_stream_synchronize", st)
    assert ix.download(np.empty(4, np.uint64))[0] == 0xFFFFFFFFFFFFFFFF
    for _ in range(2):
        _ffi.call("qr_graph_launch", g, st)
    _ffi.call("qr_stream_synchronize", st)
    got = (ip.download(np.empty(dim + 1, np.uint64)), ix.download(np.empty(dim * G, np.uint64)),
           dt.download(np.empty(dim * G, np.complex128)))
    assert_same(got, ref, "graph replay")
    ms = C.c_float()
    _ffi.call("qr_event_elapsed_ms", e0, e1, C.byref(ms))
    assert 0.0 < ms.value < 50.0
    _ffi.call("qr_graph_destroy", g)
    with pytest.raises(_ffi.QrustyCudaError):
        _ffi.call("qr_graph_begin_capture", None)


def test_kernels_were_launched():
    before = _ffi.kernel_launches()
    make_op(*H.tfim_chain(8)).to_matrix().export()
    assert _ffi.kernel_launches() >= before + 2
