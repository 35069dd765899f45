#!/usr/bin/env python
"""Writes tests/golden/csr_checksums.json: sha256 of the oracle's (indptr, indices, data)
for the small cases, after cross-checking each against the numpy restatement of the
reference's default kron-and-add to_matrix (oracle/oracle_np.py).  Also stores the full
CSR of H2 (64 entries) so one golden is human-readable.  The reference stores no expected
outputs, so these pin the oracle against regressions."""
import gzip, hashlib, json, sys
from pathlib import Path
import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import oracle as O, oracle_np as N
from qrusty_b200 import hamiltonians as H


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    fx = json.load(gzip.open(ROOT / "tests/golden/h_fixtures.json.gz"))
    cases = {k: (fx[k]["labels"], [complex(a, b) for a, b in fx[k]["coeffs"]]) for k in ["H2", "H2_rs", "H4", "H4_rs", "H6", "H6_rs"]}
    cases["C1"] = H.tfim_chain(12)
    cases["xxz_n10"] = H.xxz_chain(10, 1.0, 0.7)
    cases["tfim_3x3"] = H.tfim_lattice(3, 3, 1.0, 3.0)
    cases["random_n10"] = H.random_pauli_sum(10, 300, 200, 30, 7)
    out = {}
    for name, (labels, coeffs) in cases.items():
        n, params = O.make_params(labels, coeffs)
        indptr, indices, data = O.build_csr(params, n, step=100)
        if n <= 10 and len(labels) <= 400:
            dense = N.spop_dense(labels, coeffs)
            assert np.array_equal(dense, N.csr_to_dense(indptr, indices, data, 1 << n)), name
        else:
            import scipy.sparse as sps
            m = sps.csr_matrix((data, indices.astype(np.int64), indptr.astype(np.int64)), shape=(1 << n, 1 << n))
            assert (N.spop_sparse(labels, coeffs) != m).nnz == 0, name
        out[name] = {"n_qubits": n, "n_terms": len(labels), "n_groups": int(len(np.unique(params["x"]))),
                     "nnz": int(len(data)), "indptr": sha(indptr), "indices": sha(indices), "data": sha(data)}
        if name == "H2":
            out[name]["full"] = {"indptr": indptr.tolist(), "indices": indices.tolist(),
                                 "data_hex": [[float(v.real).hex(), float(v.imag).hex()] for v in data]}
        print(name, out[name]["n_groups"], out[name]["nnz"])
    (ROOT / "tests/golden/csr_checksums.json").write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
