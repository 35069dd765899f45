"""numpy restatement of the reference's DEFAULT to_matrix: per-qubit 2x2 matrices,
Kronecker fold, scale by base_coeff, then a sequential left-to-right sum.

TEST INFRASTRUCTURE ONLY (dense, n <= 12).  It is independent of the row-wise
algorithm in qrusty_oracle.c, so agreement between the two pins both, exactly as
the reference's own tests do (lib.rs:806-812, :815-885; test_H.py:32-48).

  SimplePauli::to_matrix       qrusty/src/lib.rs:93-114
  Pauli::to_matrix             qrusty/src/lib.rs:203-213   (fold kron(x, acc), scale(base_coeff))
  Pauli::base_coeff            qrusty/src/lib.rs:182-190   ((+i)^base_phase)
  SparsePauliOp::to_matrix     qrusty/src/lib.rs:401-412   (sum = c0*P0; sum = sum + ci*Pi)
"""
import re

import numpy as np

_P = {
    "I": np.array([[1, 0], [0, 1]], dtype=np.complex128),
    "X": np.array([[0, 1], [1, 0]], dtype=np.complex128),
    "Y": np.array([[0, -1j], [1j, 0]], dtype=np.complex128),
    "Z": np.array([[1, 0], [0, -1]], dtype=np.complex128),
}
_RE = re.compile(r"^([+-]?)1?([ij]?)([IXYZ]+)$")          # lib.rs:127
_BASE = [1 + 0j, 1j, -1 + 0j, -1j]                         # lib.rs:183-188


def parse(label):
    m = _RE.match(label)
    if not m:
        raise ValueError("error: malformed label")
    sign, imag, body = m.groups()
    base_phase = (1 if imag else 0) + (2 if sign == "-" else 0)
    return base_phase, body[::-1]                          # index 0 = right-most char (lib.rs:144)


def pauli_dense(label):
    base_phase, l = parse(label)
    acc = _P[l[0]]
    for ch in l[1:]:
        acc = np.kron(_P[ch], acc)                         # lib.rs:209
    return acc * _BASE[base_phase]                         # lib.rs:211


def spop_dense(labels, coeffs):
    total = pauli_dense(labels[0]) * complex(coeffs[0])    # lib.rs:403-404
    for lab, c in zip(labels[1:], coeffs[1:]):
        total = total + complex(c) * pauli_dense(lab)      # lib.rs:409
    return total


def spop_sparse(labels, coeffs):
    """Same fold as spop_dense with scipy.sparse matrices (for 12+ qubits, where dense 4096^2
    adds per term are too slow).  scipy drops entries that cancel, so compare by VALUE with
    `(a != b).nnz == 0`; structure with explicit zeros is pinned by the dense cases."""
    import scipy.sparse as sps
    P = {k: sps.csr_matrix(v) for k, v in _P.items()}

    def pauli(label):
        base_phase, l = parse(label)
        acc = P[l[0]]
        for ch in l[1:]:
            acc = sps.kron(P[ch], acc, format="csr")
        return acc * _BASE[base_phase]

    total = pauli(labels[0]) * complex(coeffs[0])
    for lab, c in zip(labels[1:], coeffs[1:]):
        total = total + complex(c) * pauli(lab)
    return total


def csr_to_dense(indptr, indices, data, dim):
    """Dense view of a CSR triple (duplicates do not occur on this path)."""
    out = np.zeros((dim, dim), np.complex128)
    rows = np.repeat(np.arange(len(indptr) - 1), np.diff(indptr.astype(np.int64)))
    out[rows, indices.astype(np.int64)] = data
    return out
