"""CPU restatement of the Walsh form of the HEX8 pair block (kernel_mat2.cuh): the same
butterflies, spectrum accumulation and synthesis the kernel runs, in numpy, against the plain quadrature sum on the
host mirror's own dN table (Gauss-2 and GLL-2)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools", "walsh"))

NODE_OF_SIGN = [0, 1, 3, 2, 4, 5, 7, 6]   # walsh_node_of_sign


def fwd8(v):
    v = list(v)
    for k in range(3):
        for i in range(8):
            if not i & (1 << k):
                lo, hi = v[i], v[i | (1 << k)]
                v[i], v[i | (1 << k)] = hi + lo, hi - lo
    return v


def syn8_z(v):
    v = list(v)
    v[0] = -v[1]
    for i in range(2, 8, 2):
        lo, hi = v[i], v[i + 1]
        v[i], v[i + 1] = lo - hi, lo + hi
    for k in (1, 2):
        for i in range(8):
            if not i & (1 << k):
                lo, hi = v[i], v[i | (1 << k)]
                v[i], v[i | (1 << k)] = lo - hi, lo + hi
    return v


def popc3(m):
    return bin(m).count("1")


def walsh_block(B, c):
    """M[a][b] (local node order) from the pulled-back blocks B[q][k1][k2], the kernel's way."""
    wc = [c ** n / 64.0 for n in range(5)]
    Mh = np.zeros((8, 8))
    for k1 in range(3):
        for k2 in range(3):
            bq = fwd8([B[q][k1][k2] for q in range(8)])
            for s1 in range(8):
                for s2 in range(8):
                    if not s1 & (1 << k1) and not s2 & (1 << k2):
                        Mh[s1 | (1 << k1)][s2 | (1 << k2)] += wc[popc3(s1) + popc3(s2)] * bq[s1 ^ s2]
    M = Mh.copy()
    for al in range(1, 8):
        M[al] = syn8_z(M[al])
    for ib in range(8):
        M[:, ib] = syn8_z(M[:, ib])
    out = np.zeros((8, 8))
    for ia in range(8):
        for ib in range(8):
            out[NODE_OF_SIGN[ia]][NODE_OF_SIGN[ib]] = M[ia][ib]
    return out


@pytest.mark.parametrize("q_type", ["GaussLegendre", "GaussLobattoLegendre"])
def test_walsh_pair_block_equals_quadrature_sum(q_type):
    sys.path.insert(0, os.path.join(ROOT, "finiteelementcontainers.jl_b200"))
    try:   # the host mirror needs libfecb200.so to import (no CPU fallback)
        import fecb200  # noqa: F401
        from fecb200.reference_fe import ReferenceFE
        dN = ReferenceFE("HEX8", q_type, 2).dN
    except Exception:
        pytest.skip("libfecb200.so not built")
    c = np.sqrt(8.0 * abs(dN[0, 0, 0])) - 1.0
    assert abs(c - (1.0 / np.sqrt(3.0) if q_type == "GaussLegendre" else 1.0)) < 1e-14
    rng = np.random.default_rng(3)
    B = rng.standard_normal((8, 3, 3))
    direct = np.einsum("qak,qkl,qbl->ab", dN, B, dN)
    M = walsh_block(B, c)
    assert np.abs(M - direct).max() < 1e-13 * np.abs(direct).max()
    # fused residual row: rr[a] = sum_q sum_k dN[q][a][k] Ph[q][k]
    Ph = rng.standard_normal((8, 3))
    wr = [c ** n / 8.0 for n in range(3)]
    rh = np.zeros(8)
    for k in range(3):
        pq = fwd8(Ph[:, k])
        for s1 in range(8):
            if not s1 & (1 << k):
                rh[s1 | (1 << k)] += wr[popc3(s1)] * pq[s1]
    rr = syn8_z(rh)
    direct_r = np.einsum("qak,qk->a", dN, Ph)
    got = np.zeros(8)
    for ia in range(8):
        got[NODE_OF_SIGN[ia]] = rr[ia]
    assert np.abs(got - direct_r).max() < 1e-13 * np.abs(direct_r).max()


def test_walsh_coefficient_derivation():
    import derive
    rng = np.random.default_rng(0)
    for c in (1 / np.sqrt(3.0), 1.0, 0.5):
        W, H = derive.walsh_coeffs(c)
        assert np.count_nonzero(W) == 144
        assert derive.check(c, W, H, rng) < 1e-13
