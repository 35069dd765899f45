"""Synthetic Hamiltonians of BASELINE.json's five configs, as (labels, coeffs).

Conventions: qubit k is the k-th label character from the
right (qrusty/src/lib.rs:143-146); coefficients are real unless stated; randomness is
numpy default_rng(seed).  Term order is part of the definition: it is the reference's
left-to-right summation order inside a group (accel.rs:191-205).
"""
import numpy as np


def label_from_masks(n, x, z):
    chars = []
    for k in range(n - 1, -1, -1):
        xb, zb = (x >> k) & 1, (z >> k) & 1
        chars.append("IXZY"[xb + 2 * zb])
    return "".join(chars)


def _two_site(n, i, j, ch):
    s = ["I"] * n
    s[n - 1 - i] = ch
    s[n - 1 - j] = ch
    return "".join(s)


def _one_site(n, i, ch):
    s = ["I"] * n
    s[n - 1 - i] = ch
    return "".join(s)


def tfim_chain(n=12, J=1.0, h=0.5):
    """C1: open 1D transverse-field Ising chain, H = -J sum Z_i Z_{i+1} - h sum X_i.
    n=12: T=23, G=13, nnz=53 248."""
    labels = [_two_site(n, i, i + 1, "Z") for i in range(n - 1)] + [_one_site(n, i, "X") for i in range(n)]
    coeffs = [-J] * (n - 1) + [-h] * n
    return labels, [complex(c) for c in coeffs]


def xxz_chain(n=20, J=1.0, delta=0.7):
    """C2 (delta=0.7, n=20: T=60, G=21, nnz=22 020 096) and C5 (delta=1, n=28: T=84, G=29):
    periodic chain, sum_i [J (X_i X_{i+1} + Y_i Y_{i+1}) + delta Z_i Z_{i+1}], indices mod n."""
    labels, coeffs = [], []
    for i in range(n):
        j = (i + 1) % n
        labels += [_two_site(n, i, j, "X"), _two_site(n, i, j, "Y"), _two_site(n, i, j, "Z")]
        coeffs += [J, J, delta]
    return labels, [complex(c) for c in coeffs]


def heisenberg_chain(n=28):
    """C5: periodic Heisenberg XXX chain (J = delta = 1)."""
    return xxz_chain(n, 1.0, 1.0)


def tfim_lattice(rows=5, cols=5, J=1.0, h=3.0):
    """C4: open 2D transverse-field Ising lattice, site q = cols*row + col:
    nearest-neighbour -J Z_q Z_q' bonds (horizontal first, then vertical, row-major) and -h X_q.
    5x5: T=65, G=26, nnz=872 415 232."""
    n = rows * cols
    labels, coeffs = [], []
    for r in range(rows):
        for c in range(cols):
            q = cols * r + c
            if c + 1 < cols:
                labels.append(_two_site(n, q, q + 1, "Z")); coeffs.append(-J)
            if r + 1 < rows:
                labels.append(_two_site(n, q, q + cols, "Z")); coeffs.append(-J)
    for q in range(n):
        labels.append(_one_site(n, q, "X")); coeffs.append(-h)
    return labels, [complex(c) for c in coeffs]


def random_pauli_sum(n=24, n_terms=2000, n_masks=1500, n_dup=100, seed=24):
    """C3: pool of n_masks distinct uniform n-bit X-masks; the first n_masks terms use each
    once, the rest re-use uniformly drawn pool masks, n_dup of those also copying the Z-mask of
    an earlier term with the same X-mask (exact (x,z) duplicates); Z-masks uniform; coefficients
    N(0,1) + i N(0,1); term order shuffled.  G = n_masks."""
    rng = np.random.default_rng(seed)
    pool = set()
    while len(pool) < n_masks:
        pool.update(int(v) for v in rng.integers(0, 1 << n, size=n_masks - len(pool)))
    pool = np.array(sorted(pool), dtype=np.int64)
    rng.shuffle(pool)
    x = np.empty(n_terms, np.int64)
    z = rng.integers(0, 1 << n, size=n_terms)
    x[:n_masks] = pool
    extra = n_terms - n_masks
    x[n_masks:] = pool[rng.integers(0, n_masks, size=extra)]
    for t in range(n_masks, n_masks + min(n_dup, extra)):
        earlier = np.flatnonzero(x[:t] == x[t])
        z[t] = z[earlier[0]]
    coeffs = rng.standard_normal(n_terms) + 1j * rng.standard_normal(n_terms)
    order = rng.permutation(n_terms)
    labels = [label_from_masks(n, int(x[t]), int(z[t])) for t in order]
    return labels, [complex(coeffs[t]) for t in order]


CONFIGS = {
    "C1": ("tfim_chain_n12", lambda: tfim_chain(12)),
    "C2": ("xxz_periodic_n20", lambda: xxz_chain(20, 1.0, 0.7)),
    "C3": ("random_T2000_n24", lambda: random_pauli_sum(24, 2000, 1500, 100, 24)),
    "C4": ("tfim_5x5_n25", lambda: tfim_lattice(5, 5, 1.0, 3.0)),
    "C5": ("heisenberg_periodic_n28", lambda: heisenberg_chain(28)),
}


def splitmix64(x):
    """Vectorised splitmix64 finaliser (uint64 in, uint64 out)."""
    x = (np.asarray(x, dtype=np.uint64) + np.uint64(0x9E3779B97F4A7C15))
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


def lanczos_start_at(indices, seed=25):
    """The same start vector evaluated at arbitrary row indices."""
    i = np.asarray(indices, dtype=np.uint64)
    base = np.uint64(seed) << np.uint64(40)
    with np.errstate(over="ignore"):
        a = splitmix64(i * np.uint64(2) + base)
        b = splitmix64(i * np.uint64(2) + np.uint64(1) + base)
    scale = 1.0 / float(1 << 53)
    u1 = (a >> np.uint64(11)).astype(np.float64) * scale * 2.0 - 1.0
    u2 = (b >> np.uint64(11)).astype(np.float64) * scale * 2.0 - 1.0
    return u1 + 1j * u2


def lanczos_start_vector(lo, hi, seed=25):
    """Elements [lo,hi) of the (un-normalised) start vector of SURVEY.md 8(d) C4:
    v[i] = u1 + i*u2 with u1,u2 uniform(-1,1) from splitmix64(2i + {0,1} + seed*2^40), so any
    shard or the CPU can regenerate any element."""
    i = np.arange(lo, hi, dtype=np.uint64)
    base = np.uint64(seed) << np.uint64(40)
    with np.errstate(over="ignore"):
        a = splitmix64(i * np.uint64(2) + base)
        b = splitmix64(i * np.uint64(2) + np.uint64(1) + base)
    scale = 1.0 / float(1 << 53)
    u1 = (a >> np.uint64(11)).astype(np.float64) * scale * 2.0 - 1.0
    u2 = (b >> np.uint64(11)).astype(np.float64) * scale * 2.0 - 1.0
    return u1 + 1j * u2
