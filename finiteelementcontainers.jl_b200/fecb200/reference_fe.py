"""Reference-element tables: the product-side stand-in for `ReferenceFE(elem{Lagrange,p}(), QT(q))`
(src/FunctionSpaces.jl:91-121).  ReferenceFiniteElements.jl is an un-vendored dependency of the
reference, so a Julia host passes `ref_fe.cell_interps[q]` straight through the C ABI; this module
only exists so the Python host can run standalone.  Node ordering is Exodus'.
"""
from __future__ import annotations

import itertools

import numpy as np

from . import _lib

ELEM_IDS = {"QUAD4": _lib.QUAD4, "TRI3": _lib.TRI3, "HEX8": _lib.HEX8, "TETRA4": _lib.TET4, "TETRA10": _lib.TET10}
_SIGNS = {"QUAD4": [(-1, -1), (1, -1), (1, 1), (-1, 1)],
          "HEX8": [(-1, -1, -1), (1, -1, -1), (1, 1, -1), (-1, 1, -1), (-1, -1, 1), (1, -1, 1), (1, 1, 1), (-1, 1, 1)]}
_TET_EDGES = [(0, 1), (1, 2), (0, 2), (0, 3), (1, 3), (2, 3)]


def _tensor_shape(signs, xi):
    s = np.asarray(signs, dtype=float)                    # (NNPE, ND)
    f = 1.0 + s[None, :, :] * xi[:, None, :]              # (NQ, NNPE, ND)
    scale = 0.5 ** s.shape[1]
    N = scale * f.prod(axis=2)
    dN = np.empty(f.shape)
    for j in range(s.shape[1]):
        others = [k for k in range(s.shape[1]) if k != j]
        dN[:, :, j] = scale * s[None, :, j] * f[:, :, others].prod(axis=2)
    return N, dN


def _simplex_shape(nd, quadratic, xi):
    nq = xi.shape[0]
    L = np.concatenate([1.0 - xi.sum(axis=1, keepdims=True), xi], axis=1)      # barycentric (NQ, nd+1)
    dL = np.concatenate([-np.ones((1, nd)), np.eye(nd)], axis=0)               # (nd+1, nd)
    if not quadratic:
        return L.copy(), np.broadcast_to(dL, (nq, nd + 1, nd)).copy()
    nv = nd + 1
    edges = _TET_EDGES if nd == 3 else [(0, 1), (1, 2), (2, 0)]
    N = np.empty((nq, nv + len(edges)))
    dN = np.empty((nq, nv + len(edges), nd))
    for a in range(nv):
        N[:, a] = L[:, a] * (2 * L[:, a] - 1)
        dN[:, a, :] = (4 * L[:, a] - 1)[:, None] * dL[a][None, :]
    for m, (a, b) in enumerate(edges):
        N[:, nv + m] = 4 * L[:, a] * L[:, b]
        dN[:, nv + m, :] = 4 * (L[:, a][:, None] * dL[b][None, :] + L[:, b][:, None] * dL[a][None, :])
    return N, dN


def _line_rule(q_type, npts):
    if q_type == "GaussLegendre":
        return np.polynomial.legendre.leggauss(npts)
    if q_type == "GaussLobattoLegendre":
        if npts == 2:
            return np.array([-1.0, 1.0]), np.array([1.0, 1.0])
        if npts == 3:
            return np.array([-1.0, 0.0, 1.0]), np.array([1.0, 4.0, 1.0]) / 3.0
    raise ValueError(f"unsupported 1-D rule {q_type}({npts})")


class ReferenceFE:
    """Tables of one block: N[q,a], dN[q,a,j] = dN_a/dxi_j, w[q]."""

    def __init__(self, elem_type: str, q_type: str = "GaussLegendre", q_degree: int = 2):
        self.elem_type = elem_type
        self.elem_id = ELEM_IDS[elem_type]
        self.q_type = q_type
        if elem_type in ("QUAD4", "HEX8"):
            nd = 2 if elem_type == "QUAD4" else 3
            x, w = _line_rule(q_type, q_degree)
            # x fastest, like a tensor-product rule
            idx = list(itertools.product(range(len(x)), repeat=nd))
            xi = np.array([[x[i] for i in reversed(t)] for t in idx])
            wq = np.array([np.prod([w[i] for i in t]) for t in idx])
            self.N, self.dN = _tensor_shape(_SIGNS[elem_type], xi)
            self.w = wq
        elif elem_type == "TRI3":
            if q_degree <= 1:
                xi, self.w = np.array([[1 / 3, 1 / 3]]), np.array([0.5])
            else:
                xi, self.w = np.array([[1 / 6, 1 / 6], [2 / 3, 1 / 6], [1 / 6, 2 / 3]]), np.full(3, 1 / 6)
            self.N, self.dN = _simplex_shape(2, False, xi)
        elif elem_type in ("TETRA4", "TETRA10"):
            if q_degree <= 1:
                xi, self.w = np.array([[0.25, 0.25, 0.25]]), np.array([1 / 6])
            else:
                a, b = 0.5854101966249685, 0.1381966011250105
                xi, self.w = np.array([[b, b, b], [a, b, b], [b, a, b], [b, b, a]]), np.full(4, 1 / 24)
            self.N, self.dN = _simplex_shape(3, elem_type == "TETRA10", xi)
        else:
            raise ValueError(f"unsupported element type {elem_type}")
        self.N = np.ascontiguousarray(self.N)
        self.dN = np.ascontiguousarray(self.dN)
        self.w = np.ascontiguousarray(self.w, dtype=float)

    def surface_tables(self, q_type=None, q_degree=2):
        """(Ns[q,a], dNs[q,a,k], ws[q]) of one side of this element: the reduced shape functions
        `interps.N_reduced` and the surface rule of MappedH1OrL2SurfaceInterpolants (ReferenceFiniteElements.jl,
        un-vendored -- a Julia host passes its own tables).  2-node edges (QUAD4, TRI3), 4-node faces (HEX8),
        3- / 6-node triangles (TETRA4 / TETRA10), node order = Exodus side order (meshes._SIDE_NODES)."""
        q_type = q_type or getattr(self, "q_type", "GaussLegendre")
        t = self.elem_type
        if t in ("QUAD4", "TRI3"):
            x, w = _line_rule(q_type, q_degree)
            Ns = np.stack([0.5 * (1 - x), 0.5 * (1 + x)], axis=1)
            dNs = np.broadcast_to(np.array([[-0.5], [0.5]]), (len(x), 2, 1)).copy()
            return np.ascontiguousarray(Ns), dNs, np.ascontiguousarray(w, dtype=float)
        if t == "HEX8":
            f = ReferenceFE("QUAD4", q_type, q_degree)
            return f.N, f.dN, f.w
        xi, w = np.array([[1 / 6, 1 / 6], [2 / 3, 1 / 6], [1 / 6, 2 / 3]]), np.full(3, 1 / 6)
        Ns, dNs = _simplex_shape(2, t == "TETRA10", xi)
        return np.ascontiguousarray(Ns), np.ascontiguousarray(dNs), w

    @classmethod
    def from_tables(cls, elem_type, N, dN, w):
        """wrap tables produced elsewhere (what a Julia host does with ref_fe.cell_interps)"""
        self = cls.__new__(cls)
        self.elem_type, self.elem_id = elem_type, ELEM_IDS[elem_type]
        self.N, self.dN, self.w = (np.ascontiguousarray(N, dtype=float), np.ascontiguousarray(dN, dtype=float),
                                   np.ascontiguousarray(w, dtype=float))
        return self

    @property
    def num_quadrature_points(self):
        return len(self.w)

    @property
    def num_cell_dofs(self):
        return self.N.shape[1]
