"""Pins the CPU oracle (oracle/qrusty_oracle.c) against every known answer the
reference's own tests hold for the hot path, and against an independent numpy
restatement of the reference's default kron-and-add to_matrix.  CPU only.

Reference tests mirrored here (paths into /root/reference):
  qrusty/src/lib.rs:608-693   parse_labels          -> test_parse_labels*
  qrusty/src/lib.rs:695-719   simple_pauli_matrices -> test_single_pauli_dense
  qrusty/src/lib.rs:721-768   pauli_matrices, accel -> test_accel_equals_rowwise
  qrusty/src/lib.rs:775-813   sparse_pauli_op       -> test_I_plus_2X, test_chunk_invariance
  qrusty/src/lib.rs:815-885   h2 / h4 / h6          -> test_fixture_matches_kron
  qrusty/src/lib.rs:887-919   spmat_dot_densevec    -> test_spmv
  pyqrusty/tests/test_it.py:60-101, test_H.py:32-52 -> same functions
"""
import hashlib

import numpy as np
import pytest

from oracle import oracle as O, oracle_np as N
from qrusty_b200 import hamiltonians as H

ONE, TWO, ZERO, I1 = 1 + 0j, 2 + 0j, 0j, 1j


def dense(labels, coeffs, **kw):
    n, params = O.make_params(labels, coeffs, **kw)
    indptr, indices, data = O.build_csr(params, n, step=100)
    return N.csr_to_dense(indptr, indices, data, 1 << n)


# ---- lib.rs:608-693 ----------------------------------------------------------------
@pytest.mark.parametrize("bad", ["W", "", "+i", "IXW", "2I", "i", "-"])
def test_parse_labels_errors(bad):
    with pytest.raises(ValueError):
        O.parse_label(bad)
    with pytest.raises(ValueError):
        N.parse(bad)


@pytest.mark.parametrize("label,phase", [("I", 0), ("+I", 0), ("+iI", 1), ("+jI", 1), ("-1jI", 3),
                                         ("-1I", 2), ("-1jIX", 3), ("IXYZ", 0)])
def test_parse_labels_phase(label, phase):
    bp, nq, x, z, ny = O.parse_label(label)
    assert bp == phase
    assert N.parse(label)[0] == phase


def test_parse_labels_qiskit_masks():
    # lib.rs:645-660: Pauli("IXYZ").x = [F,T,T,F], .z = [T,T,F,F] (index 0 = right-most char)
    bp, nq, x, z, ny = O.parse_label("IXYZ")
    assert nq == 4
    assert [(x >> k) & 1 for k in range(4)] == [0, 1, 1, 0]
    assert [(z >> k) & 1 for k in range(4)] == [1, 1, 0, 0]
    assert ny == 1
    # lib.rs:663-691: 22 qubits, base_phase 2, phase 0
    bp, nq, x, z, ny = O.parse_label("-IIIIIIIIIIIIIIIIIIYXXY")
    assert (bp, nq, (bp + ny) % 4) == (2, 22, 0)
    assert [(z >> k) & 1 for k in range(22)] == [1, 0, 0, 1] + [0] * 18
    assert [(x >> k) & 1 for k in range(22)] == [1, 1, 1, 1] + [0] * 18
    # IX -> [X, I] reversed (lib.rs:619)
    assert O.parse_label("-1jIX")[2] == 0b01


# ---- lib.rs:695-719, test_it.py:60-83 -------------------------------------------------
def test_single_pauli_dense():
    assert np.array_equal(dense(["I"], [ONE]), [[ONE, ZERO], [ZERO, ONE]])
    assert np.array_equal(dense(["X"], [ONE]), [[ZERO, ONE], [ONE, ZERO]])
    assert np.array_equal(dense(["Y"], [ONE]), [[ZERO, -I1], [I1, ZERO]])
    assert np.array_equal(dense(["Z"], [ONE]), [[ONE, ZERO], [ZERO, -ONE]])
    assert np.array_equal(dense(["IX"], [ONE]), [[0, 1, 0, 0], [1, 0, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]])
    assert np.array_equal(dense(["XI"], [ONE]), np.kron([[0, 1], [1, 0]], np.eye(2)))


# ---- lib.rs:775-813, test_it.py:85-101 ------------------------------------------------
def test_I_plus_2X_and_2Y():
    assert np.array_equal(dense(["I", "X"], [ONE, TWO]), [[ONE, TWO], [TWO, ONE]])
    assert np.array_equal(dense(["I", "Y"], [ONE, TWO]), np.eye(2) + 2.0 * np.array([[0, -1j], [1j, 0]]))


def test_constructor_validation():
    # lib.rs:354-376 / :783-784 / test_it.py:34-38
    with pytest.raises(ValueError):
        O.make_params(["I", "II"], [ONE, ONE])
    with pytest.raises(ValueError):
        O.make_params([], [])
    with pytest.raises(ValueError):
        O.make_params(["I"], [ONE, ONE])


# ---- lib.rs:721-768: single-Pauli fast path == row-wise path, exact struct equality ----
@pytest.mark.parametrize("label", ["IX", "XI", "I", "Y", "YY", "ZXYI", "YZYX"])
def test_accel_equals_rowwise(label):
    bp, nq, x, z, ny = O.parse_label(label)
    a = O.single_pauli(z, x, ONE, (bp + ny) % 4, nq)
    n, params = O.make_params([label], [ONE], convention="rowwise")
    b = O.build_csr(params, n)
    for u, v in zip(a, b):
        assert np.array_equal(u.view(np.uint64), v.view(np.uint64))
    assert np.array_equal(N.csr_to_dense(*b, 1 << n), N.pauli_dense(label))


# ---- lib.rs:815-885, test_H.py:32-48 --------------------------------------------------
@pytest.mark.parametrize("name", ["H2", "H2_rs", "H4", "H4_rs", "H6", "H6_rs"])
def test_fixture_matches_kron(fixtures, golden_sums, name):
    labels, coeffs = fixtures[name]
    n, params = O.make_params(labels, coeffs)
    indptr, indices, data = O.build_csr(params, n, step=100)
    G = len(np.unique(params["x"]))
    # uniform rows, sorted columns, explicit zeros kept
    assert np.array_equal(indptr, np.arange((1 << n) + 1, dtype=np.uint64) * G)
    assert (np.diff(indices.reshape(-1, G).astype(np.int64), axis=1) > 0).all()
    if n <= 8:
        assert np.array_equal(N.spop_dense(labels, coeffs), N.csr_to_dense(indptr, indices, data, 1 << n))
    elif name == "H6":      # 12 qubits: sparse kron fold, compared by value (H6_rs: checksum only)
        import scipy.sparse as sps
        m = sps.csr_matrix((data, indices.astype(np.int64), indptr.astype(np.int64)), shape=(1 << n, 1 << n))
        assert (N.spop_sparse(labels, coeffs) != m).nnz == 0
    g = golden_sums[name]
    assert (g["n_groups"], g["nnz"]) == (G, len(data))
    for key, arr in (("indptr", indptr), ("indices", indices), ("data", data)):
        assert hashlib.sha256(arr.tobytes()).hexdigest() == g[key], key


def test_golden_h2_full(fixtures, golden_sums):
    labels, coeffs = fixtures["H2"]
    n, params = O.make_params(labels, coeffs)
    indptr, indices, data = O.build_csr(params, n)
    full = golden_sums["H2"]["full"]
    assert indptr.tolist() == full["indptr"] and indices.tolist() == full["indices"]
    assert [[float(v.real).hex(), float(v.imag).hex()] for v in data] == full["data_hex"]


def test_shape_h6(fixtures):
    labels, _ = fixtures["H6"]                                  # test_H.py:50-52
    assert O.make_params(labels, [ONE] * len(labels))[0] == 12


@pytest.mark.parametrize("name", ["C1", "xxz_n10", "tfim_3x3", "random_n10"])
def test_synthetic_matches_kron(golden_sums, name):
    gen = {"C1": lambda: H.tfim_chain(12), "xxz_n10": lambda: H.xxz_chain(10, 1.0, 0.7),
           "tfim_3x3": lambda: H.tfim_lattice(3, 3, 1.0, 3.0),
           "random_n10": lambda: H.random_pauli_sum(10, 300, 200, 30, 7)}[name]
    labels, coeffs = gen()
    n, params = O.make_params(labels, coeffs)
    indptr, indices, data = O.build_csr(params, n)
    if n <= 10 and len(labels) <= 64:
        assert np.array_equal(N.spop_dense(labels, coeffs), N.csr_to_dense(indptr, indices, data, 1 << n))
    else:
        import scipy.sparse as sps
        m = sps.csr_matrix((data, indices.astype(np.int64), indptr.astype(np.int64)), shape=(1 << n, 1 << n))
        assert (N.spop_sparse(labels, coeffs) != m).nnz == 0
    g = golden_sums[name]
    for key, arr in (("indptr", indptr), ("indices", indices), ("data", data)):
        assert hashlib.sha256(arr.tobytes()).hexdigest() == g[key], key


def test_config_sizes():
    # (n, T, G) of the five BASELINE configs
    expect = {"C1": (12, 23, 13), "C2": (20, 60, 21), "C3": (24, 2000, 1500), "C4": (25, 65, 26), "C5": (28, 84, 29)}
    for k, (_, gen) in H.CONFIGS.items():
        labels, coeffs = gen()
        n, params = O.make_params(labels, coeffs)
        assert (n, len(labels), len(np.unique(params["x"]))) == expect[k]


# ---- lib.rs:812: chunk size and thread count do not change the result ------------------
def test_chunk_invariance(fixtures):
    labels, coeffs = fixtures["H4"]
    n, params = O.make_params(labels, coeffs)
    base = O.build_csr(params, n, step=500, n_threads=1)
    for step, nt in [(1, 3), (100, 2), (1000, 8), (37, 5)]:
        other = O.build_csr(params, n, step=step, n_threads=nt)
        for u, v in zip(base, other):
            assert np.array_equal(u.view(np.uint64), v.view(np.uint64))


def test_row_window_equals_slice(fixtures):
    labels, coeffs = fixtures["H4"]
    n, params = O.make_params(labels, coeffs)
    G = len(np.unique(params["x"]))
    indptr, indices, data = O.build_csr(params, n)
    ip, ix, dt = O.build_csr(params, n, row_lo=37, row_hi=201, step=10)
    assert np.array_equal(ip, np.arange(201 - 37 + 1, dtype=np.uint64) * G)
    assert np.array_equal(ix, indices[37 * G:201 * G]) and np.array_equal(dt.view(np.uint64), data[37 * G:201 * G].view(np.uint64))
    cols, vals = O.make_row(params, 123)
    assert np.array_equal(cols, indices[123 * G:124 * G]) and np.array_equal(vals, data[123 * G:124 * G])


# ---- the i/j-prefix quirk --------------------------------------------------
def test_base_phase_conventions():
    labels, coeffs = ["iXY", "-jZI", "-XX", "YZ"], [0.5 + 0.25j, 1.5 + 0j, -2 + 1j, 0.75 + 0j]
    ref = N.spop_dense(labels, coeffs)                            # reference default to_matrix
    assert np.array_equal(ref, dense(labels, coeffs, convention="to_matrix"))
    assert not np.array_equal(ref, dense(labels, coeffs, convention="rowwise"))
    # without i/j prefixes both conventions coincide bit for bit
    labels = ["-XY", "ZI", "+XX", "-1YZ"]
    a = O.make_params(labels, coeffs, convention="to_matrix")[1]
    b = O.make_params(labels, coeffs, convention="rowwise")[1]
    assert a.tobytes() == b.tobytes()


# ---- lib.rs:887-919, test_it.py:218-230 -------------------------------------------------
@pytest.mark.parametrize("name", ["H2", "H6"])
def test_spmv(fixtures, name):
    labels, coeffs = fixtures[name]
    n, params = O.make_params(labels, coeffs)
    indptr, indices, data = O.build_csr(params, n, step=500)
    rng = np.random.default_rng(1)
    v = rng.uniform(0, 10, 1 << n) + 1j * rng.uniform(0, 10, 1 << n)
    y = O.spmv(indptr, indices, data, v)
    import scipy.sparse as sps
    ref = sps.csr_matrix((data, indices.astype(np.int64), indptr.astype(np.int64)), shape=(1 << n, 1 << n)) @ v
    assert np.allclose(y, ref, rtol=0, atol=1e-7)                   # the reference's epsilon
    assert np.abs(y - ref).max() <= 1e-12 * np.abs(ref).max()
    rows = np.array([0, 5, (1 << n) - 1], dtype=np.uint64)
    assert np.array_equal(O.apply_rows(params, rows, v), y[rows.astype(np.int64)])


def h2_diagonal(fixtures):
    labels, coeffs = fixtures["H2"]
    n, params = O.make_params(labels, coeffs)
    indptr, indices, data = O.build_csr(params, n)
    G = len(indices) // (1 << n)
    rows = np.repeat(np.arange(1 << n, dtype=np.uint64), G)
    diag = np.zeros(1 << n, complex)
    diag[rows[indices == rows].astype(np.int64)] = data[indices == rows]
    return diag


def test_precond2_known_answer(fixtures, precond_vectors):
    """test_it.py:303-313 with the vectors stored there (rv printed to 9 digits)."""
    PRECOND_DX, PRECOND_E, PRECOND_RV = precond_vectors
    diag = h2_diagonal(fixtures)
    got = O.precond2(diag, PRECOND_DX, PRECOND_E, 1e-14)
    assert np.allclose(got, PRECOND_RV, rtol=1e-8, atol=0)
    # test_it.py:271-282: numpy restatement dx / reg(diag - e, tol)
    x = diag - PRECOND_E
    x = np.where(np.abs(x) < 1e-14, 1e-14, x)
    assert np.allclose(got, PRECOND_DX / x, rtol=1e-15, atol=0)
    # reg(): a denominator below tol becomes (tol, 0)
    got = O.precond2([PRECOND_E + 1e-16j], [2.0 + 4.0j], PRECOND_E, 1e-3)[0]
    assert got == complex(2.0 * 1e-3 / (1e-3 * 1e-3), 4.0 * 1e-3 / (1e-3 * 1e-3))        # num-complex's division


def test_rawio_layout():
    """rawio.rs:181-234: write_matrix (2x2 identity) and the swab constants, on the restatement."""
    indptr, indices, data = np.array([0, 1, 2], np.uint64), np.array([0, 1], np.uint64), np.array([1, 1], complex)
    b = O.rawio_bytes((2, 2), indptr, indices, data, byteorder="<")
    assert b[:2] == b"MI" and len(b) == 2 + 4 * 8 + 3 * 8 + 8 + 2 * 8 + 8 + 2 * 16
    head = np.frombuffer(b, "<u8", 4, 2)
    assert head.tolist() == [0, 2, 2, 3]                               # CSR, rows, cols, len(indptr)
    assert np.frombuffer(b, "<u8", 3, 34).tolist() == [0, 1, 2]
    assert np.frombuffer(b, "<u8", 1, 58)[0] == 2 and np.frombuffer(b, "<u8", 2, 66).tolist() == [0, 1]
    assert np.frombuffer(b, "<u8", 1, 82)[0] == 2 and np.frombuffer(b, "<c16", 2, 90).tolist() == [1, 1]
    big = O.rawio_bytes((2, 2), indptr, indices, data, byteorder=">")
    assert big[:2] == b"IM"                                            # the reader sees a foreign mark -> swab
    assert np.frombuffer(big, ">u8", 4, 2).tolist() == [0, 2, 2, 3]
    assert np.array([0x0123456789abcdef], "<u8").byteswap()[0] == 0xefcdab8967452301    # rawio.rs:188-200


def test_pair_t_layout():
    assert O.PARAM_DTYPE.itemsize == 32


def test_tuned_cpu_variant_matches_the_restatement(fixtures):
    """oracle_build_grouped (the fairness line of bench.py: group-first, closed-form slots, no per-row sort) is not the
    reference's algorithm, but it must produce the same bytes as the restatement of accel.rs:267-336."""
    from qrusty_b200 import hamiltonians as H
    cases = [fixtures["H2"], fixtures["H4"], fixtures["H6"], H.tfim_chain(12), H.xxz_chain(10, 1.0, 0.7),
             H.random_pauli_sum(10, 300, 200, 30, 7)]
    for labels, coeffs in cases:
        n, params = O.make_params(labels, coeffs)
        ref = O.build_csr(params, n)
        for threads in (1, 3):
            got = O.build_csr_grouped(params, n, n_threads=threads)
            for a, b in zip(got, ref):
                assert np.array_equal(np.ascontiguousarray(a).view(np.uint64), np.ascontiguousarray(b).view(np.uint64))
        dim = 1 << n
        if dim >= 64:
            lo, hi = 5, dim - 3
            got = O.build_csr_grouped(params, n, lo, hi)
            want = O.build_csr(params, n, lo, hi)
            for a, b in zip(got, want):
                assert np.array_equal(np.ascontiguousarray(a).view(np.uint64), np.ascontiguousarray(b).view(np.uint64))


def test_random_operators_three_ways():
    """Property test (hypothesis): for random small Pauli sums -- repeated X-masks, exact (x, z) duplicates, complex
    coefficients -- the restatement of accel.rs:267-336, the tuned CPU variant and the dense kron-and-add restatement of
    the reference's default to_matrix (oracle_np) agree: the first two bit for bit, the dense one to 1e-12."""
    from hypothesis import given, settings, strategies as st

    letters = st.sampled_from("IXYZ")

    @settings(max_examples=60, deadline=None)
    @given(st.integers(1, 6).flatmap(lambda n: st.tuples(
        st.just(n),
        st.lists(st.tuples(st.text(letters, min_size=n, max_size=n),
                           st.complex_numbers(max_magnitude=4.0, allow_nan=False, allow_infinity=False)),
                 min_size=1, max_size=24))))
    def check(case):
        n, terms = case
        labels = [t[0] for t in terms]
        coeffs = [t[1] for t in terms]
        nq, params = O.make_params(labels, coeffs)
        assert nq == n
        a = O.build_csr(params, n)
        b = O.build_csr_grouped(params, n, n_threads=2)
        for u, v in zip(a, b):
            assert np.array_equal(np.ascontiguousarray(u).view(np.uint64), np.ascontiguousarray(v).view(np.uint64))
        G = len(np.unique(params["x"]))
        assert np.array_equal(a[0], np.arange((1 << n) + 1, dtype=np.uint64) * G)
        dense = N.csr_to_dense(*a, 1 << n)
        want = N.spop_dense(labels, coeffs)
        scale = max(1.0, float(np.abs(np.asarray(coeffs)).sum()))
        assert np.abs(dense - want).max() <= 1e-12 * scale

    check()
