"""Derivation + numerical check of the Walsh form of the hex8 / 2x2x2-rule tangent block.

For trilinear hexahedra with a symmetric 2-point rule per axis (points +-c), dN_a/dxi_k at point q is
    D_q[a][k] = s_a^k / 8 * prod_{k' != k} (1 + c s_a^{k'} sigma_q^{k'}),     s_a, sigma_q in {+-1}^3,
so the (d1, d2) block of the element tangent
    M[a][b] = sum_q sum_{k1,k2} D_q[a][k1] D_q[b][k2] B_q[k1][k2],     B_q = Jinv_q (JxW A9)_q Jinv_q^T
is multilinear in the sign bits of q, a and b.  Hence
    Bh_m[k1][k2] = sum_q sigma_q^m B_q[k1][k2]                 (Walsh transform over the 8 points, 24 add/sub per entry)
    Mh[alpha][beta] = sum_{k1,k2,m} W[alpha,beta,k1,k2,m] Bh_m[k1][k2]     (144 non-zero W)
    M[a][b] = sum_{alpha,beta} s_a^alpha s_b^beta Mh[alpha][beta]          (inverse Walsh over 6 sign bits, 384 add/sub)
= 744 FP64 instructions per block after the 432 of B, against 2112 for the quadrature loop.
This script computes W and checks the identity against the direct sum on random data; the kernel (csrc/kernel_mat2.cuh)
uses the closed form of W -- Mh[alpha][beta] += c^(|alpha|+|beta|-2)/64 Bh_{(alpha\\k1) xor (beta\\k2)}[k1][k2] for k1 in alpha,
k2 in beta -- written as compile-time-unrolled loops, and tests/test_walsh_cpu.py restates those loops in numpy."""
import itertools
import sys

import numpy as np

c_sym = None


def signs(i):  # index -> (+-1)^3, bit k set <=> +1
    return np.array([1 if (i >> k) & 1 else -1 for k in range(3)])


def D(c, ia, iq):
    s, sg = signs(ia), signs(iq)
    out = np.zeros(3)
    for k in range(3):
        v = s[k] / 8.0
        for kp in range(3):
            if kp != k:
                v *= 1 + c * s[kp] * sg[kp]
        out[k] = v
    return out


def walsh_coeffs(c):
    """W[alpha, beta, k1, k2, m]: Walsh analysis (over the 6 sign bits of a, b) of
    coef_m(a, b, k1, k2) = (1/8) sum_q sigma_q^m D_q[a][k1] D_q[b][k2]   (so that sum_m coef_m Bh_m = sum_q D D B_q)."""
    W = np.zeros((8, 8, 3, 3, 8))
    H = np.array([[np.prod(signs(i)[[k for k in range(3) if (mono >> k) & 1]]) if mono else 1 for i in range(8)] for mono in range(8)], dtype=float)
    # H[mono, i] = s_i^mono
    for k1 in range(3):
        for k2 in range(3):
            # f[q, a, b] = D_q[a][k1] D_q[b][k2]
            f = np.array([[[D(c, a, q)[k1] * D(c, b, q)[k2] for b in range(8)] for a in range(8)] for q in range(8)])
            # coef_m[a, b] = (1/8) sum_q sigma_q^m f[q, a, b]
            coef = np.einsum("mq,qab->mab", H, f) / 8.0
            # Walsh analysis over a and b: Wab[alpha, beta] = (1/64) sum_{a,b} s_a^alpha s_b^beta coef[a, b]
            W[:, :, k1, k2, :] = np.einsum("xa,yb,mab->xym", H, H, coef) / 64.0
    W[np.abs(W) < 1e-15] = 0.0
    return W, H


def check(c, W, H, rng):
    B = rng.standard_normal((8, 3, 3))
    M_direct = np.zeros((8, 8))
    for q in range(8):
        for a in range(8):
            for b in range(8):
                M_direct[a, b] += D(c, a, q) @ B[q] @ D(c, b, q)
    Bh = np.einsum("mq,qij->mij", H, B)                 # forward Walsh over q
    Mh = np.einsum("xyijm,mij->xy", W, Bh)
    M = np.einsum("xa,yb,xy->ab", H, H, Mh)             # synthesis
    return np.abs(M - M_direct).max() / np.abs(M_direct).max()


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for c in (1 / np.sqrt(3.0), 1.0, 0.5):
        W, H = walsh_coeffs(c)
        print(f"c = {c:.6f}: non-zero W = {np.count_nonzero(W)}, identity error = {check(c, W, H, rng):.2e}")
    W, H = walsh_coeffs(1 / np.sqrt(3.0))
    per = [(k1, k2, int(np.count_nonzero(W[:, :, k1, k2, :]))) for k1 in range(3) for k2 in range(3)]
    print("terms per (k1,k2):", per)
    vals = np.unique(np.round(np.abs(W[W != 0]), 14))
    print("distinct |W|:", vals)
