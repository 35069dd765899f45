"""CPU-only tests of the host side: the pyqrusty-compatible operator model, the mode-string
grammar, make_params parity with the oracle, and that the C-ABI library loads and exports every
symbol include/qrusty_cuda.h declares.  No compute call is made unless it is expected to fail."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle as O
import qrusty_b200 as Q
from qrusty_b200 import _ffi, hamiltonians as H

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "qrusty_cuda.h").read_text()
    declared = set(re.findall(r"^QR_API\s+[\w\s\*]+?\b(qr_\w+)\s*\(", header, re.M))
    assert len(declared) >= 35
    assert declared == set(_ffi.EXPORTS)
    for name in declared:
        assert hasattr(_ffi.lib, name), name
    assert _ffi.lib.qr_version() == 100
    assert C.sizeof(_ffi.Term) == 32 and Q.TERM_DTYPE.itemsize == 32
    assert C.sizeof(_ffi.PlanInfo) == 40


def _c_prototypes(header):
    """name -> list of parameter type strings, from the QR_API declarations of the header."""
    protos = {}
    for m in re.finditer(r"^QR_API\s+[\w\s\*]+?\b(qr_\w+)\s*\(([^;]*?)\)\s*;", re.sub(r"/\*.*?\*/", "", header, flags=re.S), re.M | re.S):
        args = [a.strip() for a in m.group(2).split(",")]
        protos[m.group(1)] = [] if args in ([""], ["void"]) else args
    return protos


def test_rust_shim_in_lock_step_with_the_header():
    """rust/cuda.rs cannot be compiled here (no cargo in the image), so its extern block is checked against the header the
    other way: every function it binds exists with the same number of parameters, pointer-ness and integer width per
    position, and the constants it copies have the header's values."""
    header = (ROOT / "include" / "qrusty_cuda.h").read_text()
    rust = (ROOT / "rust" / "cuda.rs").read_text()
    protos = _c_prototypes(header)
    block = re.search(r'extern "C" \{(.*?)\n\}', rust, re.S).group(1)
    bound = re.findall(r"fn (qr_\w+)\s*\((.*?)\)\s*(?:->\s*([^;]+))?;", block, re.S)
    assert len(bound) >= 15
    width = {"c_int": "int", "u32": "uint32_t", "u64": "uint64_t", "usize": "size_t", "f64": "double"}
    for name, args, ret in bound:
        assert name in protos, "rust/cuda.rs binds %s, which the header does not declare" % name
        r_args = [a.strip() for a in args.split(",") if a.strip()]
        c_args = protos[name]
        assert len(r_args) == len(c_args), (name, r_args, c_args)
        for ra, ca in zip(r_args, c_args):
            r_type = ra.split(":", 1)[1].strip()
            assert r_type.startswith("*") == ("*" in ca or "[" in ca), (name, ra, ca)
            if not r_type.startswith("*"):
                assert width[r_type] in ca, (name, ra, ca)
    for const, value in re.findall(r"pub const (QR_\w+): \w+ = (\d+);", rust):
        m = re.search(r"#define\s+%s\s+(\d+)" % const, header)
        if m is None:                                                # status codes are an enum in the header
            m = re.search(r"\b%s\s*=\s*(\d+)" % const, header)
        assert m is not None and int(m.group(1)) == int(value), const


def test_no_oracle_in_product():
    """The product must never import, link or call the oracle."""
    for f in (ROOT / "qrusty_b200").rglob("*"):
        if f.suffix in (".py", ".cu", ".cuh", ".h") and f.is_file():
            assert "oracle" not in f.read_text().lower().replace("no cpu", ""), f


def test_argument_validation_without_gpu():
    t = (_ffi.Term * 2)()
    t[0].x, t[0].z = 1, 0
    h = C.c_void_p()
    assert _ffi.lib.qr_plan_create(33, t, 2, 0, 0, C.byref(h)) == _ffi.QR_ERR_UNSUPPORTED
    assert _ffi.lib.qr_plan_create(0, t, 2, 0, 0, C.byref(h)) == _ffi.QR_ERR_INVALID
    assert _ffi.lib.qr_plan_create(4, t, 0, 0, 0, C.byref(h)) == _ffi.QR_ERR_INVALID
    assert b"at least one" in _ffi.lib.qr_last_error()
    t[1].x = 1 << 5
    assert _ffi.lib.qr_plan_create(4, t, 2, 0, 0, C.byref(h)) == _ffi.QR_ERR_INVALID
    assert b"outside n_qubits" in _ffi.lib.qr_last_error()
    assert _ffi.lib.qr_plan_info(None, None) == _ffi.QR_ERR_INVALID
    assert _ffi.lib.qr_plan_destroy(None) == _ffi.QR_OK


def test_fails_loudly_without_gpu():
    if _ffi.device_count() > 0:
        pytest.skip("a GPU is present")
    op = Q.SparsePauliOp([Q.Pauli("IX")], [1.0])
    with pytest.raises(Q.QrustyCudaError):
        op.to_matrix()
    with pytest.raises(Q.QrustyCudaError):
        op.apply(np.zeros(4, complex))
    with pytest.raises(Q.QrustyCudaError):
        Q.axpy(1.0, np.zeros(4, complex), np.zeros(4, complex))


# ---- operator model: lib.rs:608-693, test_it.py:28-58 --------------------------------------
@pytest.mark.parametrize("bad", ["W", "", "+i", "IXW", "2I", "i", "-", "ix"])
def test_malformed_labels(bad):
    with pytest.raises(Exception):
        Q.Pauli(bad)


@pytest.mark.parametrize("label", ["I", "+I", "+iI", "+jI", "-1jI", "-1I", "-1jIX", "IXYZ",
                                   "-IIIIIIIIIIIIIIIIIIYXXY", "iYYZX", "-jYXZ"])
def test_pauli_matches_oracle(label):
    p = Q.Pauli(label)
    bp, nq, x, z, ny = O.parse_label(label)
    assert (p.base_phase, p.num_qubits(), p.x_indices(), p.z_indices(), p.phase()) == (bp, nq, x, z, (bp + ny) % 4)


def test_pauli_api():
    p = Q.Pauli("IXYZ")                                   # test_it.py:28-32
    assert p.num_qubits() == 4 and p.label() == "IXYZ" and repr(p) == "Pauli('IXYZ')" and str(p) == "IXYZ"
    assert Q.Pauli("-1jIX").label() == "IX"


def test_spop_api(fixtures):
    with pytest.raises(Exception):                        # test_it.py:34-38
        Q.SparsePauliOp([Q.Pauli("I"), Q.Pauli("IXYZ")], [1.0 + 0.0j, 1.0 + 0.0j])
    with pytest.raises(Exception):
        Q.SparsePauliOp([], [])
    with pytest.raises(Exception):
        Q.SparsePauliOp([Q.Pauli("I")], [1.0, 2.0])
    spop = Q.SparsePauliOp([Q.Pauli("IIII"), Q.Pauli("IXYZ")], [1.0 + 0.0j, 1.0 + 0.0j])
    assert repr(spop) == "SparsePauliOp('IIII','IXYZ', [1+0j, 1+0j])"      # test_it.py:40-44
    assert len(spop) == 2 and spop.num_qubits() == 4
    labels, coeffs = fixtures["H2"]
    h2 = Q.SparsePauliOp([Q.Pauli(l) for l in labels], coeffs)
    assert repr(h2[3]) == "(Pauli('IIXX'), (-0.00170526+0j))"              # test_it.py:211-216
    assert repr(h2[3:8:2]) == ("[(Pauli('IIXX'), (-0.00170526+0j)), (Pauli('IZII'), (0.18388659+0j)), "
                               "(Pauli('XXII'), (-0.00170526+0j))]")
    with pytest.raises(Exception):
        h2[27]
    assert len(h2 + h2) == 54
    assert Q._rust_f64(1e-7) == "0.0000001" and Q._rust_f64(-2.5) == "-2.5" and Q._rust_f64(1e21) == "1000000000000000000000"


@pytest.mark.parametrize("mode,n", [("", 1), ("Binary", 1), ("Accel", 1), ("Rowwise", 1), ("RowwiseUnsafe", 1),
                                    ("RowwiseUnsafeChunked/100", 1), ("Reduce", 1), ("Rayon", 1),
                                    ("RayonChunked/1000", 1), ("Cuda", 1), ("Cuda/8", 8)])
def test_mode_strings(mode, n):
    assert Q._parse_mode(mode) == n                       # lib.rs:293-331 + the new Cuda arms


@pytest.mark.parametrize("mode", ["foo", "cuda", "Cuda/", "Cuda/0", "Cuda/x", "RowwiseUnsafeChunked/", "Rowwise "])
def test_bad_mode_strings(mode):
    with pytest.raises(Exception, match="unrecognized mode"):    # test_H.py:28-30
        Q._parse_mode(mode)


# ---- make_params parity (accel.rs:141-157) --------------------------------------------------
@pytest.mark.parametrize("name", ["H2", "H4", "H6", "H2_rs"])
def test_terms_match_oracle(fixtures, name):
    labels, coeffs = fixtures[name]
    op = Q.SparsePauliOp([Q.Pauli(l) for l in labels], coeffs)
    n, params = O.make_params(labels, coeffs)
    assert op.num_qubits() == n
    assert op.terms().tobytes() == params.tobytes()


def test_terms_with_prefixes_and_complex_coeffs():
    labels = ["iXY", "-jZI", "-XX", "YZ", "+1ZZ", "-1jYY"]
    coeffs = [0.5 + 0.25j, 1.5 - 0.0j, -2 + 1j, 0.75 + 0j, 1j, -0.0 + 3j]
    op = Q.SparsePauliOp([Q.Pauli(l) for l in labels], coeffs)
    assert op.terms().tobytes() == O.make_params(labels, coeffs, convention="to_matrix")[1].tobytes()
    labels, coeffs = H.random_pauli_sum(10, 300, 200, 30, 7)
    op = Q.SparsePauliOp([Q.Pauli(l) for l in labels], coeffs)
    assert op.terms().tobytes() == O.make_params(labels, coeffs)[1].tobytes()
    op2 = Q.SparsePauliOp.from_terms(10, op.terms())
    assert len(op2) == 300 and op2.num_qubits() == 10


def test_generators_roundtrip_masks():
    for n, x, z in [(5, 0b10110, 0b00111), (1, 1, 1), (8, 0, 255)]:
        p = Q.Pauli(H.label_from_masks(n, x, z))
        assert (p.x_indices(), p.z_indices(), p.num_qubits()) == (x, z, n)
    v = H.lanczos_start_vector(0, 64)
    assert np.array_equal(v[10:20], H.lanczos_start_vector(10, 20)) and np.all(np.abs(v.real) <= 1) and len(set(v)) == 64


# ---- bench.py contract pieces that need no GPU ---------------------------------------------------
def test_bench_reference_arm_and_workloads():
    """`bench.py --impl reference` (the CPU port of accel.rs:267-336 on the host threads) prints one JSON
    line with the contract's keys; the weak-scaling family keeps T = 60 and G = 21 at every N."""
    import json
    import subprocess
    import sys
    import types
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    out = subprocess.run([sys.executable, str(root / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--config", "xxz12"], capture_output=True, text=True, timeout=300, cwd=str(root))
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "csr_build_nnz_per_s" and line["unit"] == "nnz/s"
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["config"]["nnz"] == 13 * 4096 and line["value"] > 0

    sys.path.insert(0, str(root))
    import bench
    import numpy as np
    from oracle import oracle as O
    for world in (1, 2, 4, 8):
        name, labels, coeffs = bench.workload(types.SimpleNamespace(config="auto"), world)
        n, params = O.make_params(labels, coeffs)
        assert n == 20 + int(np.log2(world)) and len(labels) == 60 and len(np.unique(params["x"])) == 21
        assert int(params["x"].max()) < 1 << 20 and int(params["z"].max()) < 1 << 20       # spectators untouched
    # the fairness line: tuned CPU variant, labelled as not the reference's algorithm
    tuned = bench.cpu_tuned_rate(*H.xxz_chain(12, 1.0, 0.7), runs=2)
    assert tuned["value"] > 0 and tuned["unit"] == "nnz/s" and "not the reference" in tuned["kind"] and tuned["cores"] >= 1
