"""CPU model of apply_fold_kernel's arithmetic (qrusty_b200/csrc/apply_fold.cuh) against the oracle: a thread owns the 8
rows r0 ^ (e << 7); popc(r & z) = popc(r0 & z) + popc(e & q) with q = (z >> 7) & 7; groups of <= 12 terms add every term
into the 8 row values with the sign pattern of its bucket, longer groups sum per bucket and finish with a 3-stage
Walsh-Hadamard butterfly.  No GPU: this pins the index arithmetic and the identity the kernel relies on."""
import numpy as np

from oracle import oracle as O
from qrusty_b200 import hamiltonians as H

B0, WHT_MIN = 7, 12


def popc(a):
    return np.array([bin(int(v)).count("1") for v in np.atleast_1d(a)])


def wht8(S):
    S = S.copy()
    b = 1
    while b < 8:
        for i in range(8):
            if not i & b:
                a, c = S[i], S[i | b]
                S[i], S[i | b] = a + c, a - c
        b <<= 1
    return S


def fold_apply(params, n, v):
    """y = H v the way the kernel computes it: per block of 1024 rows, per thread (128 of them), per group."""
    dim = 1 << n
    x = params["x"].astype(np.uint64); z = params["z"].astype(np.uint64); c = params["re"] + 1j * params["im"]
    order = np.argsort(x, kind="stable")
    gx, start = np.unique(x[order], return_index=True)
    bounds = list(start) + [len(x)]
    y = np.zeros(dim, np.complex128)
    for base in range(0, dim, 1024):
        for tid in range(128):
            r0 = base + tid
            acc = np.zeros(8, np.complex128)
            for g, xm in enumerate(gx):
                t = order[bounds[g]:bounds[g + 1]]
                zq = (z[t] >> np.uint64(B0)) & np.uint64(7)
                sign0 = np.where(popc(np.uint64(r0) & z[t]) & 1, -1.0, 1.0)
                if len(t) <= WHT_MIN:
                    val = np.array([np.sum(c[t] * sign0 * np.where(popc(np.uint64(e) & zq) & 1, -1.0, 1.0)) for e in range(8)])
                else:
                    S = np.array([np.sum((c[t] * sign0)[zq == q]) for q in range(8)])
                    val = wht8(S)
                for e in range(8):
                    r = r0 ^ (e << B0)
                    acc[e] += val[e] * v[r ^ int(xm)]
            for e in range(8):
                y[r0 ^ (e << B0)] = acc[e]
    return y


def test_fold_arithmetic_matches_the_oracle():
    n = 11
    labels, coeffs = H.random_pauli_sum(n, 90, 12, 10, 3)          # 12 masks, ~7 terms each, one group well above 12
    labels += labels[:20]; coeffs += [c * 0.37 for c in coeffs[:20]]
    n_o, params = O.make_params(labels, coeffs)
    assert n_o == n
    rng = np.random.default_rng(2)
    v = rng.standard_normal(1 << n) + 1j * rng.standard_normal(1 << n)
    rows = np.arange(1 << n, dtype=np.uint64)
    ref = O.apply_rows(params, rows, v)
    y = fold_apply(params, n, v)
    absH = np.abs(params["re"] + 1j * params["im"]).sum()
    assert np.abs(y - ref).max() <= 1e-12 * absH * np.abs(v).max()
    sizes = np.unique(params["x"], return_counts=True)[1]
    assert sizes.max() > WHT_MIN and sizes.min() <= WHT_MIN            # both paths of the kernel were exercised
