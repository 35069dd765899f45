#!/usr/bin/env python
"""Extracts the reference's Hamiltonian fixture DATA (labels + coefficients) into
tests/golden/h_fixtures.json.gz.  Runs only in the authoring container, where
/root/reference exists; the GPU box uses the committed JSON.

Sources (inputs only -- the reference stores no expected outputs):
  pyqrusty/tests/H_fixtures.py:21-95518   H2 H4 H6 H8 H10 H11 H12 HJ
  qrusty/src/fixtures.rs:27-217           H2 H4 H6 (Rust ordering / full-precision coeffs)

H_fixtures.py is executed with a stub `pyqrusty` module whose Pauli /
SparsePauliOp only record their arguments.
"""
import gzip, json, re, sys, types
from pathlib import Path

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent / "h_fixtures.json.gz"


def from_python():
    stub = types.ModuleType("pyqrusty")

    class Pauli:
        def __init__(self, label): self.label = label

    class SparsePauliOp:
        def __init__(self, paulis, coeffs):
            self.labels = [p.label for p in paulis]
            self.coeffs = [complex(c) for c in coeffs]

    stub.Pauli, stub.SparsePauliOp = Pauli, SparsePauliOp
    stub.__all__ = ["Pauli", "SparsePauliOp"]
    sys.modules["pyqrusty"] = stub
    ns = {}
    src = (REF / "pyqrusty/tests/H_fixtures.py").read_text()
    exec(compile(src, "H_fixtures.py", "exec"), ns)
    out = {}
    for name in ["H2", "H4", "H6", "H8", "H10", "H11", "H12", "HJ"]:
        op = ns[name]
        out[name] = {"labels": op.labels, "coeffs": [[c.real, c.imag] for c in op.coeffs]}
    return out


def from_rust():
    src = (REF / "qrusty/src/fixtures.rs").read_text()
    out = {}
    for name in ["H2", "H4", "H6"]:
        m = re.search(r"pub static ref %s : TestCase = \{(.*?)TestCase \{" % name, src, re.S)
        body = m.group(1)
        lab = re.search(r"let labels = vec!\[(.*?)\] ;", body, re.S).group(1)
        cof = re.search(r"let coeffs = vec!\[(.*?)\] ;", body, re.S).group(1)
        labels = re.findall(r'"([^"]+)"', lab)
        coeffs = [complex(s.replace(" ", "")) for s in re.findall(r'"([^"]+)"', cof)]
        assert len(labels) == len(coeffs) and labels
        out[name + "_rs"] = {"labels": labels, "coeffs": [[c.real, c.imag] for c in coeffs]}
    return out


if __name__ == "__main__":
    data = from_python()
    data.update(from_rust())
    for k, v in data.items():
        print(k, len(v["labels"]), "terms", len(v["labels"][0]), "qubits")
    with gzip.GzipFile(OUT, "wb", mtime=0) as f:
        f.write(json.dumps(data, separators=(",", ":")).encode())
    print("wrote", OUT, OUT.stat().st_size, "bytes")
