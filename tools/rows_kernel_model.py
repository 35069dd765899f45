#!/usr/bin/env python
"""Python model of fill_rows_kernel's control flow (qrusty_b200/csrc/fill.cuh): thread <-> group state,
Gray-coded batches of 2^Q rows, the +-cnt slot steps, the extras table and the two batch buffers, run
thread by thread on the CPU and compared bit for bit with the oracle.  A design check that needs no GPU
(`python tools/rows_kernel_model.py`); the kernel itself is tested in tests/test_gpu_parity.py."""
import sys, gzip, json
import numpy as np
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import oracle as O
from qrusty_b200 import hamiltonians as H

def plan_tables(params, n):
    z = params["z"].astype(np.uint64); x = params["x"].astype(np.uint64)
    c = np.ascontiguousarray(np.stack([params["re"], params["im"]], axis=1)).view(np.complex128).ravel()
    order = np.argsort(x, kind="stable")
    xs, zs, cs = x[order], z[order], c[order]
    gx, first = np.unique(xs, return_index=True)
    G = len(gx); goff = np.append(first, len(xs))
    cnt = np.zeros((G, 32), dtype=np.int64)
    for g in range(G):
        for h in range(G):
            if h != g:
                cnt[g, int(gx[g] ^ gx[h]).bit_length() - 1] += 1
    return gx, goff, zs, cs, cnt

def popc(v): return bin(int(v)).count("1")

def emulate(params, n, row_lo, row_hi, NG, TH, Q, log2R):
    gx, goff, zs, cs, cnt = plan_tables(params, n)
    G = len(gx); T = len(zs)
    RT = 1 << Q; R = 1 << log2R
    s0 = (row_lo + R - 1) // R * R; s1 = row_hi // R * R
    assert s1 > s0
    n_runs = (s1 - s0) // R
    rows = row_hi - row_lo
    indices = np.full(rows * G, -1, dtype=np.int64); data = np.full(rows * G, np.nan, dtype=np.complex128)
    tile_n = RT * G
    # extras
    n_extra = T - G
    s_ez = np.zeros(n_extra, dtype=np.uint64); s_ec = np.zeros(n_extra, dtype=np.complex128)
    threads = []
    for tid in range(TH):
        st = []
        for k in range(NG):
            g = tid + k * TH; gg = min(g, G - 1)
            t0, t1 = int(goff[gg]), int(goff[gg + 1])
            if g < G:
                for t in range(t0 + 1, t1):
                    s_ez[t - gg - 1] = zs[t]; s_ec[t - gg - 1] = cs[t]
            sd = [(-int(cnt[gg, b]) if (int(gx[gg]) >> b) & 1 else int(cnt[gg, b])) for b in range(Q + 2)]
            st.append(dict(g=g, gg=gg, x=int(gx[gg]), z0=int(zs[t0]), c0=cs[t0], eb=t0 - gg, ee=t1 - gg - 1, sd=sd, off=0))
        threads.append(st)
    n_ctas = min(n_runs, 3)
    for cta in range(n_ctas):
        parity = 0
        bufs_d = [np.full(tile_n, np.nan, dtype=np.complex128) for _ in range(2)]
        bufs_i = [np.full(tile_n, -1, dtype=np.int64) for _ in range(2)]
        for run in range(cta, n_runs, n_ctas):
            r0 = s0 + (run << log2R)
            for st in threads:
                for s in st:
                    xr = s["x"] ^ r0
                    s["off"] = sum(int(cnt[s["gg"], b]) for b in range(n) if (xr >> b) & 1)
            rb = r0
            for i in range(R >> Q):
                if i != 0:
                    b = Q + ((i & -i).bit_length() - 1)
                    rb ^= 1 << b
                    up = (rb >> b) & 1
                    for st in threads:
                        for s in st:
                            if b < Q + 2: sdv = s["sd"][b]
                            else:
                                cb = int(cnt[s["gg"], b]); sdv = -cb if (s["x"] >> b) & 1 else cb
                            s["off"] += sdv if up else -sdv
                bd, bi = bufs_d[parity], bufs_i[parity]
                bd[:] = np.nan; bi[:] = -1
                for st in threads:
                    for s in st:
                        if s["g"] < G:
                            for j in range(RT):
                                r = rb + j
                                sg = -1.0 if popc(r & s["z0"]) & 1 else 1.0
                                vr, vi = sg * s["c0"].real, sg * s["c0"].imag
                                for e in range(s["eb"], s["ee"]):
                                    ce = s_ec[e]
                                    sg = -1.0 if popc(r & int(s_ez[e])) & 1 else 1.0
                                    vr, vi = vr + sg * ce.real, vi + sg * ce.imag
                                v = np.array([vr, vi]).view(np.complex128)[0]
                                o = j * G + s["off"] + sum(s["sd"][b] for b in range(Q) if (j >> b) & 1)
                                assert 0 <= o < tile_n and bi[o] == -1, (o, tile_n)
                                bd[o] = v; bi[o] = r ^ s["x"]
                o = (rb - row_lo) * G
                data[o:o + tile_n] = bd; indices[o:o + tile_n] = bi
                parity ^= 1
    return s0, s1, indices, data

def check(name, labels, coeffs, NG, TH, Q, log2R, lo=None, hi=None):
    n, params = O.make_params(labels, coeffs)
    ref = O.build_csr(params, n)
    dim = 1 << n
    lo = 0 if lo is None else lo; hi = dim if hi is None else hi
    s0, s1, ix, dt = emulate(params, n, lo, hi, NG, TH, Q, log2R)
    G = len(ix) // (hi - lo)
    a, b = (s0 - lo) * G, (s1 - lo) * G
    assert np.array_equal(ix[a:b], ref[1][s0 * G:s1 * G].astype(np.int64)), name
    assert np.array_equal(dt[a:b].view(np.uint64), ref[2][s0 * G:s1 * G].view(np.uint64)), name
    print("ok", name, "G", G, NG, TH, Q, log2R, (lo, hi))

if __name__ == "__main__" and "--no-run" not in sys.argv:
    fx = json.load(gzip.open(ROOT / "tests/golden/h_fixtures.json.gz"))
    def fxop(k): return fx[k]["labels"], [complex(a, b) for a, b in fx[k]["coeffs"]]
    check("C1", *H.tfim_chain(12)[:2], 1, 32, 1, 5)
    check("C1q2", *H.tfim_chain(12)[:2], 1, 32, 2, 6, 100, 4000)
    l, c = H.random_pauli_sum(8, 120, 70, 10, 7)
    check("rand", l, c, 3, 32, 1, 4)
    check("rand", l, c, 2, 64, 2, 6, 3, 250)
    check("H4", *fxop("H4"), 2, 32, 2, 5)
    check("H4", *fxop("H4"), 1, 64, 1, 3)
