"""
fec_oracle.py -- CPU restatement (numpy, float64) of the FiniteElementContainers.jl
assembly hot path.  THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / reference
arm may import this module.  The product package (`fecb200`) never does.

Parity status
-------------
The reference is Julia and cannot run in the build container (no julia binary, no
package depot), so the oracle is pinned against what the reference's own tests hold:

  * `test/poisson/poisson.gold`  (full assembly + solve pipeline, TestPoisson.jl:97)
  * `test/TestMesh.jl:99-128`    (StructuredMesh known connectivities / coordinates)
  * `test/TestFormulations.jl`   (extract_stiffness / discrete_gradient index maps)
  * analytic checks (u = x exactly, TestPoisson.jl:605-721)

Everything else (absolute R/K values for hex8, neo-Hookean, J2/tet10) is
**parity unpinned** in the reference itself; for those the
oracle is the definition, cross-checked here against finite differences of the
strain-energy function (standing in for Tensors.jl AD).

All indices in this file are 0-based internally; functions that export
reference-visible integer arrays (`Is`, `Js`, `rowptr`, `colval`, dof maps) say
explicitly whether they are 0- or 1-based.  Layout conventions follow the reference:

  dof(d, n) = NF*n + d                       (src/DofManagers.jl:41-58, Fields.jl:36-40)
  u_el[NF*a + d] = U[d, conn[a]]             (src/assemblers/Assemblers.jl:161-173)
  grad_u[d, j]   = sum_a u_el[d,a] dN_X[a,j] (src/Physics.jl:66-76)
"""
from __future__ import annotations

import math
import numpy as np

# --------------------------------------------------------------------------------------
# Meshes  (src/meshes/StructuredMesh.jl)
# --------------------------------------------------------------------------------------


def structured_mesh(el_type: str, mins, maxs, counts):
    """Restatement of StructuredMesh(el_type, mins, maxs, counts)
    (src/meshes/StructuredMesh.jl:28-83).  Returns dict with
    coords (ND, NN) float64, conn (NNPE, NE) int64 **1-based**, nodesets {name: 1-based ids}.
    """
    t = el_type.upper()
    if any(a >= b for a, b in zip(mins, maxs)):
        raise IndexError("Dimension has negative or zero length")  # BoundsError in the reference
    if t in "HEX":
        return _hex8_mesh(mins, maxs, counts)
    if t in "QUAD":
        return _quad4_mesh(mins, maxs, counts)
    if t in "TRI":
        return _tri3_mesh(mins, maxs, counts)
    if t in "TET":
        raise AssertionError("Implement tet case")  # StructuredMesh.jl:50-51
    raise ValueError(f"Unsupported element type {el_type}")


def _hex8_mesh(mins, maxs, counts):
    # src/meshes/StructuredMesh.jl:85-131, 433-470
    Nx, Ny, Nz = counts
    xs = np.linspace(mins[0], maxs[0], Nx)
    ys = np.linspace(mins[1], maxs[1], Ny)
    zs = np.linspace(mins[2], maxs[2], Nz)
    coords = np.empty((3, Nx * Ny * Nz))
    n = 0
    # note the reference loops counts[1] outer / counts[3] inner (quirk B7): cubes only
    for kz in range(Nz):
        for jy in range(Ny):
            coords[0, n:n + Nx] = xs
            coords[1, n:n + Nx] = ys[jy]
            coords[2, n:n + Nx] = zs[kz]
            n += Nx

    def node(i, j, k):  # 1-based args, 1-based result
        return i + Nx * (j - 1) + Nx * Ny * (k - 1)

    Ex, Ey, Ez = Nx - 1, Ny - 1, Nz - 1
    ex, ey, ez = np.meshgrid(np.arange(1, Ex + 1), np.arange(1, Ey + 1), np.arange(1, Ez + 1),
                             indexing="ij")  # ex outer, ez inner -> C-order ravel
    ex, ey, ez = ex.ravel(), ey.ravel(), ez.ravel()
    conn = np.stack([
        node(ex, ey, ez), node(ex + 1, ey, ez), node(ex + 1, ey + 1, ez), node(ex, ey + 1, ez),
        node(ex, ey, ez + 1), node(ex + 1, ey, ez + 1), node(ex + 1, ey + 1, ez + 1),
        node(ex, ey + 1, ez + 1)]).astype(np.int64)
    I, J, K = np.arange(1, Nx + 1), np.arange(1, Ny + 1), np.arange(1, Nz + 1)

    def vec(f, A, B):  # Julia comprehension [f(a,b) for a in A, b in B] |> vec : a fastest
        a, b = np.meshgrid(A, B, indexing="ij")
        return f(a, b).ravel(order="F").astype(np.int64)

    nsets = {
        "bottom": vec(lambda i, k: node(i, 1, k), I, K),
        "right": vec(lambda j, k: node(Nx, j, k), J, K),
        "front": vec(lambda i, j: node(i, j, Nz), I, J),
        "top": vec(lambda i, k: node(i, Ny, k), I, K),
        "left": vec(lambda j, k: node(1, j, k), J, K),
        "back": vec(lambda i, j: node(i, j, 1), I, J),
    }
    return dict(coords=coords, conn=conn, nodesets=nsets, el_type="HEX8")


def _quad4_mesh(mins, maxs, counts):
    # src/meshes/StructuredMesh.jl:232-255, 472-500
    Nx, Ny = counts
    xs = np.linspace(mins[0], maxs[0], Nx)
    ys = np.linspace(mins[1], maxs[1], Ny)
    coords = np.empty((2, Nx * Ny))
    n = 0
    for jy in range(Ny):
        coords[0, n:n + Nx] = xs
        coords[1, n:n + Nx] = ys[jy]
        n += Nx

    def node(i, j):
        return i + Nx * (j - 1)

    Ex, Ey = Nx - 1, Ny - 1
    conn = np.empty((4, Ex * Ey), dtype=np.int64)
    n = 0
    # the reference enumerates ex outer, ey inner (TestMesh.jl:99-104 pins this)
    for ex in range(1, Ex + 1):
        for ey in range(1, Ey + 1):
            conn[:, n] = (node(ex, ey), node(ex + 1, ey), node(ex + 1, ey + 1), node(ex, ey + 1))
            n += 1
    nsets = {
        "bottom": np.array([node(i, 1) for i in range(1, Nx + 1)], dtype=np.int64),
        "right": np.array([node(Nx, j) for j in range(1, Ny + 1)], dtype=np.int64),
        "top": np.array([node(i, Ny) for i in range(1, Nx + 1)], dtype=np.int64),
        "left": np.array([node(1, j) for j in range(1, Ny + 1)], dtype=np.int64),
    }
    return dict(coords=coords, conn=conn, nodesets=nsets, el_type="QUAD4")


def _tri3_mesh(mins, maxs, counts):
    # src/meshes/StructuredMesh.jl:329-352 ; pinned by TestMesh.jl:120-128
    q = _quad4_mesh(mins, maxs, counts)
    c = q["conn"]
    NEq = c.shape[1]
    conn = np.empty((3, 2 * NEq), dtype=np.int64)
    conn[:, 0::2] = c[[0, 1, 2], :]
    conn[:, 1::2] = c[[0, 2, 3], :]
    return dict(coords=q["coords"], conn=conn, nodesets=q["nodesets"], el_type="TRI3")


def kuhn_tet10_mesh(n: int, lo=0.0, hi=1.0):
    """Synthetic tet10 mesh (BASELINE.json config 4; the reference has NO tet generator,
    StructuredMesh.jl:50-51 asserts): every cell of an n^3 grid is split into 6 Kuhn
    tetrahedra around the (0,0,0)-(1,1,1) body diagonal; P2 nodes are the points of the
    (2n+1)^3 half lattice.  Exodus TETRA10 local ordering: 4 vertices, then mid-edge nodes
    (1-2),(2-3),(1-3),(1-4),(2-4),(3-4).  Oracle-defined (parity unpinned).
    Returns coords (3,NN), conn (10,NE) 1-based, nodesets.
    """
    M = 2 * n + 1
    g = np.linspace(lo, hi, M)
    coords = np.empty((3, M ** 3))
    kk, jj, ii = np.meshgrid(np.arange(M), np.arange(M), np.arange(M), indexing="ij")
    coords[0] = g[ii.ravel()]
    coords[1] = g[jj.ravel()]
    coords[2] = g[kk.ravel()]

    def nid(p):  # p: (..., 3) half-lattice integer coordinates -> 1-based node id
        return p[..., 0] + M * p[..., 1] + M * M * p[..., 2] + 1

    # Kuhn: permutations of axis order; tet vertices v0=0, v1=e_p0, v2=e_p0+e_p1, v3=(1,1,1)
    perms = [(0, 1, 2), (0, 2, 1), (1, 0, 2), (1, 2, 0), (2, 0, 1), (2, 1, 0)]
    cx, cy, cz = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    base = 2 * np.stack([cx.ravel(), cy.ravel(), cz.ravel()], axis=-1)  # (NC,3) ex outer, ez inner
    NC = base.shape[0]
    conn = np.empty((10, NC * 6), dtype=np.int64)
    E = np.eye(3, dtype=np.int64)
    for t, p in enumerate(perms):
        v = [np.zeros(3, dtype=np.int64), E[p[0]], E[p[0]] + E[p[1]], np.ones(3, dtype=np.int64)]
        # orientation: make the tet positively oriented
        a, b, c = v[1] - v[0], v[2] - v[0], v[3] - v[0]
        if np.dot(np.cross(a, b), c) < 0:
            v[1], v[2] = v[2], v[1]
        V = [base + 2 * vi for vi in v]  # half-lattice coordinates of the 4 vertices
        edges = [(0, 1), (1, 2), (0, 2), (0, 3), (1, 3), (2, 3)]
        cols = slice(t, NC * 6, 6)
        for a_ in range(4):
            conn[a_, cols] = nid(V[a_])
        for m, (a_, b_) in enumerate(edges):
            conn[4 + m, cols] = nid((V[a_] + V[b_]) // 2)
    idx = np.arange(M ** 3)
    i3, j3, k3 = idx % M, (idx // M) % M, idx // (M * M)
    nsets = {
        "bottom": idx[j3 == 0] + 1, "top": idx[j3 == M - 1] + 1,
        "left": idx[i3 == 0] + 1, "right": idx[i3 == M - 1] + 1,
        "back": idx[k3 == 0] + 1, "front": idx[k3 == M - 1] + 1,
    }
    return dict(coords=coords, conn=conn, nodesets=nsets, el_type="TETRA10")


# --------------------------------------------------------------------------------------
# Reference element tables.  The reference takes these from ReferenceFiniteElements.jl
# 0.14 (NOT vendored): `ref_fe.cell_interps[q]` = (N, grad_N_xi, w).  The
# library never hard-codes them: they cross the C ABI as arrays.  The oracle ships the
# standard Lagrange shape functions in Exodus node order and tensor-Gauss / GLL /
# simplex rules so standalone runs are possible.
# --------------------------------------------------------------------------------------

def _gauss_1d(npts):
    x, w = np.polynomial.legendre.leggauss(npts)
    return x, w


def _gll_1d(npts):
    if npts == 2:
        return np.array([-1.0, 1.0]), np.array([1.0, 1.0])
    if npts == 3:
        return np.array([-1.0, 0.0, 1.0]), np.array([1 / 3, 4 / 3, 1 / 3])
    raise ValueError("GLL rule with %d points not tabulated" % npts)


_QUAD4_XI = np.array([[-1, -1], [1, -1], [1, 1], [-1, 1]], dtype=float)
_HEX8_XI = np.array([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1],
                     [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]], dtype=float)


def _shape_quad4(xi):
    N = np.array([0.25 * (1 + a * xi[0]) * (1 + b * xi[1]) for a, b in _QUAD4_XI])
    dN = np.array([[0.25 * a * (1 + b * xi[1]), 0.25 * b * (1 + a * xi[0])] for a, b in _QUAD4_XI])
    return N, dN


def _shape_hex8(xi):
    N = np.array([0.125 * (1 + a * xi[0]) * (1 + b * xi[1]) * (1 + c * xi[2]) for a, b, c in _HEX8_XI])
    dN = np.array([[0.125 * a * (1 + b * xi[1]) * (1 + c * xi[2]),
                    0.125 * b * (1 + a * xi[0]) * (1 + c * xi[2]),
                    0.125 * c * (1 + a * xi[0]) * (1 + b * xi[1])] for a, b, c in _HEX8_XI])
    return N, dN


def _shape_tri3(xi):
    N = np.array([1 - xi[0] - xi[1], xi[0], xi[1]])
    dN = np.array([[-1.0, -1.0], [1.0, 0.0], [0.0, 1.0]])
    return N, dN


def _shape_tet4(xi):
    N = np.array([1 - xi[0] - xi[1] - xi[2], xi[0], xi[1], xi[2]])
    dN = np.array([[-1.0, -1, -1], [1, 0, 0], [0, 1, 0], [0, 0, 1.0]])
    return N, dN


def _shape_tet10(xi):
    L = np.array([1 - xi[0] - xi[1] - xi[2], xi[0], xi[1], xi[2]])
    dL = np.array([[-1.0, -1, -1], [1, 0, 0], [0, 1, 0], [0, 0, 1.0]])
    N = np.empty(10)
    dN = np.empty((10, 3))
    for a in range(4):
        N[a] = L[a] * (2 * L[a] - 1)
        dN[a] = (4 * L[a] - 1) * dL[a]
    for m, (a, b) in enumerate([(0, 1), (1, 2), (0, 2), (0, 3), (1, 3), (2, 3)]):
        N[4 + m] = 4 * L[a] * L[b]
        dN[4 + m] = 4 * (L[a] * dL[b] + L[b] * dL[a])
    return N, dN


def ref_fe_tables(el_type: str, rule: str = "gauss2"):
    """Tables (N[q,a], dN[q,a,j], w[q]) for an element type / quadrature rule.
    rule: 'gauss2' (tensor 2-pt Gauss), 'gll2' (2-pt Gauss-Lobatto = vertices), 'gll3',
          simplex rules 'tri1','tri3','tet1','tet4'.
    """
    t = el_type.upper()
    if t in ("QUAD4", "QUAD", "HEX8", "HEX"):
        nd = 2 if t.startswith("QUAD") else 3
        if rule.startswith("gauss"):
            x, w = _gauss_1d(int(rule[5:]))
        elif rule.startswith("gll"):
            x, w = _gll_1d(int(rule[3:]))
        else:
            raise ValueError(rule)
        pts, wts = [], []
        if nd == 2:
            for j in range(len(x)):
                for i in range(len(x)):
                    pts.append((x[i], x[j])); wts.append(w[i] * w[j])
            shp = _shape_quad4
        else:
            for k in range(len(x)):
                for j in range(len(x)):
                    for i in range(len(x)):
                        pts.append((x[i], x[j], x[k])); wts.append(w[i] * w[j] * w[k])
            shp = _shape_hex8
    elif t in ("TRI3", "TRI"):
        shp = _shape_tri3
        if rule == "tri1":
            pts, wts = [(1 / 3, 1 / 3)], [0.5]
        else:  # degree-2 3-point rule
            pts = [(1 / 6, 1 / 6), (2 / 3, 1 / 6), (1 / 6, 2 / 3)]
            wts = [1 / 6] * 3
    elif t in ("TET4", "TETRA4", "TETRA", "TET", "TETRA10", "TET10"):
        shp = _shape_tet10 if t.endswith("10") else _shape_tet4
        if rule == "tet1":
            pts, wts = [(0.25, 0.25, 0.25)], [1 / 6]
        else:  # degree-2 4-point rule
            a, b = 0.5854101966249685, 0.1381966011250105
            pts = [(b, b, b), (a, b, b), (b, a, b), (b, b, a)]
            wts = [1 / 24] * 4
    else:
        raise ValueError(el_type)
    N = np.array([shp(np.array(p))[0] for p in pts])
    dN = np.array([shp(np.array(p))[1] for p in pts])
    return N, dN, np.array(wts, dtype=float)


# --------------------------------------------------------------------------------------
# Side sets and surface tables (src/meshes/StructuredMesh.jl:133-230, 257-330, 355-431;
# surface_connectivity / MappedH1OrL2SurfaceInterpolants live in ReferenceFiniteElements.jl,
# un-vendored: Exodus side numbering, surface Jacobian |dx/dxi| (edges) or |t_0 x t_1| (faces))
# --------------------------------------------------------------------------------------

SIDE_NODES = {  # Exodus side -> local nodes (0-based)
    "QUAD4": [(0, 1), (1, 2), (2, 3), (3, 0)],
    "TRI3": [(0, 1), (1, 2), (2, 0)],
    "HEX8": [(0, 1, 5, 4), (1, 2, 6, 5), (2, 3, 7, 6), (0, 4, 7, 3), (0, 3, 2, 1), (4, 5, 6, 7)],
    "TETRA4": [(0, 1, 3), (1, 2, 3), (0, 3, 2), (0, 2, 1)],
    "TETRA10": [(0, 1, 3, 4, 8, 7), (1, 2, 3, 5, 9, 8), (0, 3, 2, 7, 9, 6), (0, 2, 1, 6, 5, 4)],
}


def structured_sidesets(el_type, counts):
    """{name: (elements, sides)} 1-based, for StructuredMesh.  QUAD4 / TRI3 follow the reference loops verbatim
    (StructuredMesh.jl:257-330, 355-431).  HEX8: the reference's `_hex8_ssets` (:133-230) is unfinished (its side
    numbers name faces that do not lie on the named boundary, its `left` loop reads an undefined k, and its
    side-node matrices are sized for one row of faces), so the geometrically correct faces are used instead:
    bottom / top = y-min / y-max, left / right = x-min / x-max, back / front = z-min / z-max like the node sets."""
    t = el_type.upper()
    if t.startswith("QUAD") or t.startswith("TRI"):
        Ex, Ey = counts[0] - 1, counts[1] - 1
        quad = lambda i, j: (i - 1) * Ey + j
        I, J = np.arange(1, Ex + 1), np.arange(1, Ey + 1)
        if t.startswith("QUAD"):
            return {"bottom": (quad(I, 1), np.full(Ex, 1)), "right": (quad(Ex, J), np.full(Ey, 2)),
                    "top": (quad(I, Ey), np.full(Ex, 3)), "left": (quad(1, J), np.full(Ey, 4))}
        a, b = (lambda q: 2 * q - 1), (lambda q: 2 * q)
        return {"bottom": (a(quad(I, 1)), np.full(Ex, 1)), "right": (a(quad(Ex, J)), np.full(Ey, 2)),
                "top": (b(quad(I, Ey)), np.full(Ex, 2)), "left": (b(quad(1, J)), np.full(Ey, 3))}
    Ex, Ey, Ez = counts[0] - 1, counts[1] - 1, counts[2] - 1
    elem = lambda i, j, k: (i - 1) * Ey * Ez + (j - 1) * Ez + k

    def plane(f, A, B):
        aa, bb = np.meshgrid(A, B, indexing="ij")
        return f(aa.ravel(), bb.ravel())
    I, J, K = np.arange(1, Ex + 1), np.arange(1, Ey + 1), np.arange(1, Ez + 1)
    return {"bottom": (plane(lambda i, k: elem(i, 1, k), I, K), np.full(Ex * Ez, 1)),
            "top": (plane(lambda i, k: elem(i, Ey, k), I, K), np.full(Ex * Ez, 3)),
            "left": (plane(lambda j, k: elem(1, j, k), J, K), np.full(Ey * Ez, 4)),
            "right": (plane(lambda j, k: elem(Ex, j, k), J, K), np.full(Ey * Ez, 2)),
            "back": (plane(lambda i, j: elem(i, j, 1), I, J), np.full(Ex * Ey, 5)),
            "front": (plane(lambda i, j: elem(i, j, Ez), I, J), np.full(Ex * Ey, 6))}


def side_nodes(el_type, conn, elements, sides):
    """surface_connectivity of every (element, side): (nnps, nsides) 1-based global node ids."""
    canon = {"QUAD": "QUAD4", "TRI": "TRI3", "HEX": "HEX8", "TET4": "TETRA4", "TET10": "TETRA10"}
    tab = SIDE_NODES[canon.get(el_type.upper(), el_type.upper())]
    return np.stack([conn[list(tab[s - 1]), e - 1] for e, s in zip(elements, sides)], axis=1).astype(np.int64)


def surface_tables(el_type: str, rule: str = "gauss2"):
    """(Ns[q,a], dNs[q,a,k], ws[q]) of the sides of an element type: 2-node edges (QUAD4, TRI3), 4-node faces
    (HEX8), 3- / 6-node triangles (TETRA4 / TETRA10)."""
    t = el_type.upper()
    if t in ("QUAD4", "QUAD", "TRI3", "TRI"):
        x, w = (_gauss_1d(int(rule[5:])) if rule.startswith("gauss") else _gll_1d(int(rule[3:]))) if rule[0] == "g" else _gauss_1d(2)
        N = np.stack([0.5 * (1 - x), 0.5 * (1 + x)], axis=1)
        dN = np.broadcast_to(np.array([[-0.5], [0.5]]), (len(x), 2, 1)).copy()
        return N, dN, np.asarray(w, dtype=float)
    if t in ("HEX8", "HEX"):
        return ref_fe_tables("QUAD4", rule)
    pts, wts = [(1 / 6, 1 / 6), (2 / 3, 1 / 6), (1 / 6, 2 / 3)], [1 / 6] * 3
    if t in ("TET4", "TETRA4", "TETRA", "TET"):
        return ref_fe_tables("TRI3", "tri3")
    N, dN = [], []
    for xi in pts:  # 6-node triangle, nodes (v0, v1, v2, m01, m12, m20)
        L = np.array([1 - xi[0] - xi[1], xi[0], xi[1]])
        dL = np.array([[-1.0, -1.0], [1.0, 0.0], [0.0, 1.0]])
        n = [L[a] * (2 * L[a] - 1) for a in range(3)] + [4 * L[a] * L[b] for a, b in ((0, 1), (1, 2), (2, 0))]
        d = [(4 * L[a] - 1) * dL[a] for a in range(3)] + [4 * (L[a] * dL[b] + L[b] * dL[a]) for a, b in ((0, 1), (1, 2), (2, 0))]
        N.append(n); dN.append(d)
    return np.array(N), np.array(dN), np.array(wts)


def surface_jxw(dNs_q, w_q, x_s):
    """x_s (nsides, nnps, ND) -> JxW (nsides,) of one surface quadrature point."""
    t = np.einsum("eai,ak->eki", x_s, dNs_q)       # tangents (nsides, ND-1, ND)
    if x_s.shape[2] == 2:
        return np.linalg.norm(t[:, 0, :], axis=1) * w_q
    return np.linalg.norm(np.cross(t[:, 0, :], t[:, 1, :]), axis=1) * w_q


def surface_quadrature_points(snodes, tables, X):
    """X_q of every (q, side): (nqs, nsides, ND)   (interps.X_q in _update_bc_values!, NeumannBCs.jl:60-71)"""
    x_s = np.transpose(X[:, snodes - 1], (2, 1, 0))
    return np.einsum("qa,eai->qei", tables[0], x_s)


def assemble_vector_neumann_bc(R, snodes, tables, vals, X, nf):
    """_assemble_block_vector_weakly_enforced_bc! (WeaklyEnforcedBCs.jl:61-83): R[(n,d)] += JxW Ns[n] vals[d,q,e].
    snodes (nnps, nsides) 1-based; vals (NF, nqs, nsides).  Adds in place (no zeroing, :33)."""
    Ns, dNs, ws = tables
    x_s = np.transpose(X[:, snodes - 1], (2, 1, 0))
    for q in range(len(ws)):
        JxW = surface_jxw(dNs[q], ws[q], x_s)
        contrib = JxW[:, None, None] * Ns[q][None, :, None] * vals[:, q, :].T[:, None, :]   # (nsides, nnps, NF)
        dofs = nf * (snodes.T - 1)[:, :, None] + np.arange(nf)[None, None, :]
        np.add.at(R, dofs.ravel(), contrib.ravel())
    return R


def robin_update_bc_values(snodes, tables, X, U, func, dfuncdu, t=0.0):
    """_update_bc_values! of the Robin container (src/bcs/RobinBCs.jl:77-86): u_q = u_el * interps.N at every surface
    point, vals[q, e] = func(X_q, t, u_q), dvalsdu[q, e] = dfuncdu(X_q, t, u_q) (ForwardDiff.jacobian in the reference,
    :72-75; an explicit callable here).  U (NF, NN).  Returns vals (NF, nqs, nsides), dvalsdu (NF, NF, nqs, nsides)."""
    Ns = tables[0]
    nf = U.shape[0]
    Xq = surface_quadrature_points(snodes, tables, X)                   # (nqs, nsides, ND)
    u_s = np.transpose(U[:, snodes - 1], (2, 1, 0))                     # (nsides, nnps, NF)
    uq = np.einsum("qa,eac->qec", Ns, u_s)                              # (nqs, nsides, NF)
    nq, ne = Xq.shape[0], Xq.shape[1]
    vals = np.zeros((nf, nq, ne))
    dvals = np.zeros((nf, nf, nq, ne))
    for q in range(nq):
        for e in range(ne):
            vals[:, q, e] = np.asarray(func(Xq[q, e], t, uq[q, e]), dtype=float).reshape(nf)
            dvals[:, :, q, e] = np.asarray(dfuncdu(Xq[q, e], t, uq[q, e]), dtype=float).reshape(nf, nf)
    return vals, dvals


def assemble_matrix_robin_bc(storage, conn, elements, snodes, tables, dvals, X, nf):
    """_assemble_block_matrix_weakly_enforced_bc! (src/assemblers/WeaklyEnforcedBCs.jl:118-153): per side
    K_el = sum_q _expand_face_block(Nvec, JxW, dval, node_to_face_idx) (:167-180), added to the COO storage of the parent
    element, column-major, at (el_id - 1) * NDOF^2 (:155-165).  conn (NNPE, NE) of the block, elements block-local
    1-based, snodes (nnps, nsides), dvals (NF, NF, nqs, nsides).  `storage` is the block view of the COO values."""
    Ns, dNs, ws = tables
    nnpe = conn.shape[0]
    ndofs = nf * nnpe
    x_s = np.transpose(X[:, snodes - 1], (2, 1, 0))
    jxw = np.stack([surface_jxw(dNs[q], ws[q], x_s) for q in range(len(ws))])      # (nqs, nsides)
    for e, el in enumerate(elements):
        c = conn[:, el - 1]
        face_idx = [int(np.nonzero(snodes[:, e] == n)[0][0]) + 1 if n in snodes[:, e] else 0 for n in c]   # node_to_face_idx
        K_el = np.zeros((ndofs, ndofs))
        for q in range(len(ws)):
            for row in range(ndofs):
                ni, di = row // nf, row % nf
                for col in range(ndofs):
                    nj, dj = col // nf, col % nf
                    ii, jj = face_idx[ni], face_idx[nj]
                    if ii and jj:
                        K_el[row, col] += jxw[q, e] * Ns[q, ii - 1] * Ns[q, jj - 1] * dvals[di, dj, q, e]
        s0 = (el - 1) * ndofs * ndofs
        storage[s0:s0 + ndofs * ndofs] += K_el.reshape(-1, order="F")      # K_el.data[i], column-major
    return storage


def cell_quadrature_points(block, X):
    """X_q of every (q, e): (NQ, NE, ND)   (_update_source_values!, Sources.jl:55-66)"""
    return np.einsum("qa,eai->qei", block.N, _gather(X, block.conn))


def assemble_vector_source(R, block, vals, X, nf):
    """_assemble_block_vector_source! (Source.jl:44-63): R[(n,d)] += -JxW N[n] vals[d,q,e]; vals (NF, NQ, NE)."""
    x_el = _gather(X, block.conn)
    for q in range(len(block.w)):
        _, _, JxW = map_interpolants(block.N[q], block.dN[q], block.w[q], x_el)
        contrib = -JxW[:, None, None] * block.N[q][None, :, None] * vals[:, q, :].T[:, None, :]
        dofs = nf * (block.conn.T - 1)[:, :, None] + np.arange(nf)[None, None, :]
        np.add.at(R, dofs.ravel(), contrib.ravel())
    return R


# --------------------------------------------------------------------------------------
# DofManager maps  (src/DofManagers.jl:227-298)
# --------------------------------------------------------------------------------------

DIRICHLET_DOF = -1
PERIODIC_SIDE_B_DOF = -2


def update_dofs(nf: int, nn: int, dirichlet_dofs, per_a=(), per_b=()):
    """All arrays 1-based like the reference.  Returns dict with dirichlet_dofs,
    unknown_dofs, dof_to_unknown, periodic_side_a_dofs, periodic_side_b_dofs,
    periodic_side_b_to_side_a_unknown."""
    ndof = nf * nn
    dd = np.asarray(dirichlet_dofs, dtype=np.int64)
    pa = [int(v) for v in per_a]
    pb = [int(v) for v in per_b]
    # _resolve_periodic_chains (:300-325)
    b2a = {b: a for a, b in zip(pa, pb)}

    def canonical(d):
        while d in b2a:
            d = b2a[d]
        return d
    for b in list(b2a.keys()):
        b2a[b] = canonical(b2a[b])
    ra = [b2a[b] for b in pb]
    seen, pairs = set(), []
    for a, b in zip(ra, pb):
        if (a, b) not in seen:
            seen.add((a, b)); pairs.append((a, b))
    ra = np.array([p[0] for p in pairs], dtype=np.int64)
    rb = np.array([p[1] for p in pairs], dtype=np.int64)
    assert np.all((dd >= 1) & (dd <= ndof))
    mask = np.ones(ndof + 1, dtype=bool)
    mask[0] = False
    mask[dd] = False
    mask[rb] = False
    unknown = np.nonzero(mask)[0].astype(np.int64)
    assert len(np.unique(dd)) + len(np.unique(rb)) + len(unknown) == ndof
    d2u = np.zeros(ndof + 1, dtype=np.int64)
    d2u[unknown] = np.arange(1, len(unknown) + 1)
    d2u[dd] = DIRICHLET_DOF
    d2u[rb] = PERIODIC_SIDE_B_DOF
    b2au = np.zeros(ndof + 1, dtype=np.int64)
    for a, b in zip(ra, rb):
        assert d2u[a] != 0
        b2au[b] = d2u[a]
    return dict(dirichlet_dofs=dd, unknown_dofs=unknown, dof_to_unknown=d2u[1:],
                periodic_side_a_dofs=ra, periodic_side_b_dofs=rb,
                periodic_side_b_to_side_a_unknown=b2au[1:])


def dof_to_unknown_index(dof, g):
    """src/DofManagers.jl:188-201 ; g is a 1-based array of dof ids."""
    g = np.asarray(g)
    dtu = dof["dof_to_unknown"][g - 1]
    side_a = dof["periodic_side_b_to_side_a_unknown"][g - 1]
    return np.where(dtu == PERIODIC_SIDE_B_DOF, side_a, dtu)


# --------------------------------------------------------------------------------------
# Sparsity pattern and sparse realisation (src/assemblers/SparsityPatterns.jl)
# --------------------------------------------------------------------------------------

def _dof_conn(conn, nf):
    """conn (NNPE, NE) 1-based -> dof_conn (NF*NNPE, NE) 1-based, d fastest
    (SparsityPatterns.jl:74-75)."""
    nnpe, ne = conn.shape
    d = np.arange(1, nf + 1)
    return (nf * (conn[:, None, :] - 1) + d[None, :, None]).reshape(nnpe * nf, ne)


def matrix_pattern(blocks_conn, nf: int, dof=None, condensed: bool = True):
    """SparseMatrixPattern(dof) (:53-117) and, if not condensed and `dof` given,
    _update_dofs! (:160-231).  Returns 1-based Is, Js, unknown_dofs (COO slot ids),
    block_start_indices and the stable permutation of (Is<<32)|Js (1-based)."""
    Is, Js, starts = [], [], []
    carry = 1
    for conn in blocks_conn:
        dc = _dof_conn(conn, nf)  # (NDOF, NE)
        ndofe, ne = dc.shape
        # for e: for i: for j:  -> index order (e, i, j)
        I = np.broadcast_to(dc.T[:, :, None], (ne, ndofe, ndofe)).reshape(-1)
        J = np.broadcast_to(dc.T[:, None, :], (ne, ndofe, ndofe)).reshape(-1)
        Is.append(I); Js.append(J)
        starts.append(carry)
        carry += ndofe * ndofe * ne
    Is = np.concatenate(Is).astype(np.int64)
    Js = np.concatenate(Js).astype(np.int64)
    slots = np.arange(1, len(Is) + 1, dtype=np.int64)
    if not condensed and dof is not None:
        ri = dof_to_unknown_index(dof, Is)
        rj = dof_to_unknown_index(dof, Js)
        keep = (ri > 0) & (rj > 0)
        Is, Js, slots = ri[keep], rj[keep], slots[keep]
    keys = (Is << 32) | Js
    perm = np.argsort(keys, kind="stable") + 1
    return dict(Is=Is, Js=Js, unknown_dofs=slots, block_start_indices=np.array(starts),
                permutation=perm)


def sparse_csc(Is, Js, vals, n):
    """SparseArrays.sparse!(I, J, V, n, n, +) semantics (SparsityPatterns.jl:301-308):
    CSC, rows sorted ascending inside each column, duplicates summed in COO order,
    explicitly stored zeros kept.  1-based colptr/rowval."""
    key = (Js.astype(np.int64) << 32) | Is.astype(np.int64)
    order = np.argsort(key, kind="stable")
    ks = key[order]
    first = np.ones(len(ks), dtype=bool)
    first[1:] = ks[1:] != ks[:-1]
    starts = np.nonzero(first)[0]
    seg = np.cumsum(first) - 1
    nz = np.zeros(len(starts))
    np.add.at(nz, seg, vals[order])  # sequential accumulation in COO order
    rowval = (ks[starts] & 0xFFFFFFFF).astype(np.int64)
    cols = (ks[starts] >> 32).astype(np.int64)
    counts = np.bincount(cols, minlength=n + 1)[1:]
    colptr = np.concatenate([[1], 1 + np.cumsum(counts)]).astype(np.int64)
    return colptr, rowval, nz


def csc_to_csr(colptr, rowval, nzval, n):
    """SparseMatrixCSR(csc) (SparsityPatterns.jl:326-329): 1-based rowptr/colval with
    columns sorted ascending in each row."""
    nnz = len(rowval)
    cols = np.repeat(np.arange(1, n + 1, dtype=np.int64), np.diff(colptr))
    order = np.argsort((rowval << 32) | cols, kind="stable")
    counts = np.bincount(rowval, minlength=n + 1)[1:]
    rowptr = np.concatenate([[1], 1 + np.cumsum(counts)]).astype(np.int64)
    assert rowptr[-1] == nnz + 1
    return rowptr, cols[order], nzval[order]


# --------------------------------------------------------------------------------------
# Physics at a quadrature point, vectorised over elements.
# Signature (x_el, u_el are (NE,NNPE,ND)/(NE,NNPE,NF)):
# --------------------------------------------------------------------------------------

def map_interpolants(N_q, dN_q, w_q, x_el):
    """MappedH1OrL2Interpolants(interps, x_el) (src/Physics.jl:86-92):
    J[i,j] = sum_a x[a,i] dN[a,j]; dN_X = dN J^-1 ; JxW = det(J) w ; X_q = sum N_a x_a."""
    J = np.einsum("eai,aj->eij", x_el, dN_q)
    detJ = np.linalg.det(J)
    Jinv = np.linalg.inv(J)
    dN_X = np.einsum("aj,ejk->eak", dN_q, Jinv)
    X_q = np.einsum("a,eai->ei", N_q, x_el)
    return X_q, dN_X, detJ * w_q


class Poisson:
    """test/poisson/TestPoissonCommon.jl:4-139 ; AbstractPhysics{1,0,0}.  `func(X) -> f`
    is evaluated at quadrature points (vectorised over elements)."""
    NF, NS = 1, 0

    def __init__(self, func):
        self.func = func

    def residual_q(self, N_q, X_q, dN_X, JxW, u_el, props, so, sn):
        gu = np.einsum("ead,eaj->edj", u_el, dN_X)  # (NE,1,ND)
        f = self.func(X_q)  # (NE,)
        R = np.einsum("ej,eaj->ea", gu[:, 0, :], dN_X) - N_q[None, :] * f[:, None]
        return JxW[:, None] * R  # (NE, NNPE*1)

    def stiffness_q(self, N_q, X_q, dN_X, JxW, u_el, props, so, sn):
        return JxW[:, None, None] * np.einsum("eaj,ebj->eab", dN_X, dN_X)

    def mass_q(self, N_q, X_q, dN_X, JxW, u_el, props, so, sn):
        return JxW[:, None, None] * (N_q[:, None] * N_q[None, :])[None]

    def lumped_mass_q(self, N_q, X_q, dN_X, JxW, u_el, props, so, sn):
        """row sum of mass_q (partition of unity: sum_b N_b = 1), density 1"""
        return JxW[:, None] * N_q[None, :]

    def energy_q(self, N_q, X_q, dN_X, JxW, u_el, props, so, sn):
        """energy (test/poisson/TestPoissonCommon.jl:8-16): JxW (1/2 grad u . grad u - u_q f(X_q))"""
        gu = np.einsum("ead,eaj->edj", u_el, dN_X)[:, 0, :]
        u_q = np.einsum("a,ea->e", N_q, u_el[:, :, 0])
        return JxW * (0.5 * np.einsum("ej,ej->e", gu, gu) - u_q * self.func(X_q))


def _sym(A):
    return 0.5 * (A + np.swapaxes(A, -1, -2))


class _Mechanics:
    """Shared machinery of the mechanics physics: P = dpsi/d(grad u), A = d2psi.
    R[NF a + d] = JxW sum_j dN_X[a,j] P[d,j]                (Formulations.jl:27-49,462-496)
    K[NF a+d1, NF b+d2] = JxW sum dN_X[a,j1] A[d1,j1,d2,j2] dN_X[b,j2]   (:89-126,574-576)
    2-D (PlaneStrain): grad u padded to 3x3 with zeros (Formulations.jl:421-427)."""
    NS = 0

    def __init__(self, nd=3):
        self.nd = nd
        self.NF = nd

    def _grad3(self, u_el, dN_X):
        gu = np.einsum("ead,eaj->edj", u_el, dN_X)
        if self.nd == 3:
            return gu
        g3 = np.zeros((gu.shape[0], 3, 3))
        g3[:, :2, :2] = gu
        return g3

    def residual_q(self, N_q, X_q, dN_X, JxW, u_el, props, so, sn):
        P, _ = self.stress_tangent(self._grad3(u_el, dN_X), props, so, sn, need_A=False)
        nd = self.nd
        R = np.einsum("eaj,edj->ead", dN_X, P[:, :nd, :nd])
        return JxW[:, None] * R.reshape(R.shape[0], -1)

    def stiffness_q(self, N_q, X_q, dN_X, JxW, u_el, props, so, sn):
        _, A = self.stress_tangent(self._grad3(u_el, dN_X), props, so, None, need_A=True)
        nd = self.nd
        A = A[:, :nd, :nd, :nd, :nd]
        K = np.einsum("eaj,edjfk,ebk->eadbf", dN_X, A, dN_X)
        n = K.shape[1] * nd
        return JxW[:, None, None] * K.reshape(K.shape[0], n, n)

    def mass_q(self, N_q, X_q, dN_X, JxW, u_el, props, so, sn):
        nd = self.nd
        NN = N_q[:, None] * N_q[None, :]
        M = np.einsum("ab,df->adbf", NN, np.eye(nd)).reshape(len(N_q) * nd, len(N_q) * nd)
        return (JxW * props[0])[:, None, None] * M[None]

    def energy_q(self, N_q, X_q, dN_X, JxW, u_el, props, so, sn):
        """energy(physics::Mechanics, ...) (test/mechanics/TestMechanicsCommon.jl:38-50): JxW psi(grad u)"""
        return JxW * self.energy(self._grad3(u_el, dN_X), props)

    def lumped_mass_q(self, N_q, X_q, dN_X, JxW, u_el, props, so, sn):
        """lumped_mass(physics::Mechanics, ...) (test/mechanics/TestMechanicsCommon.jl:98-125):
        m_el[NF a + d] = props[1] * JxW * N[a], identical in every direction."""
        m = (JxW * props[0])[:, None] * N_q[None, :]          # (NE, NNPE)
        return np.repeat(m, self.nd, axis=1)                    # interleaved dof order


class LinearElastic(_Mechanics):
    """test/mechanics/TestMechanicsCommon.jl:3-236: psi = 1/2 K tr(eps)^2 + G dev eps:dev eps,
    props = (rho, K, G)."""

    def energy(self, gu, props):
        """strain_energy (test/mechanics/TestMechanicsCommon.jl:14-20): 1/2 K tr(eps)^2 + G dev(eps):dev(eps)"""
        K, G = props[1], props[2]
        eps = _sym(gu)
        tr = np.trace(eps, axis1=1, axis2=2)
        dev = eps - tr[:, None, None] / 3 * np.eye(3)
        return 0.5 * K * tr ** 2 + G * np.einsum("eij,eij->e", dev, dev)

    def stress_tangent(self, gu, props, so, sn, need_A=True):
        K, G = props[1], props[2]
        eps = _sym(gu)
        tr = np.trace(eps, axis1=1, axis2=2)
        I = np.eye(3)
        dev = eps - tr[:, None, None] / 3 * I
        P = K * tr[:, None, None] * I + 2 * G * dev
        A = None
        if need_A:
            II = np.einsum("ij,kl->ijkl", I, I)
            Is4 = 0.5 * (np.einsum("ik,jl->ijkl", I, I) + np.einsum("il,jk->ijkl", I, I))
            A = np.broadcast_to(K * II + 2 * G * (Is4 - II / 3), (gu.shape[0], 3, 3, 3, 3))
        return P, A


class NeoHookean(_Mechanics):
    """test/mechanics/TestMechanicsLargeDeformation.jl:17-27 (stale script, parity unpinned):
    psi = 1/2 K U(J) + 1/2 G (J^-2/3 tr(F F^T) - 3), F = I + grad u.
    variant 'standard'  : U = 1/2 (J^2-1) - ln J   (stress free at F = I; examples/electromechanics/script.jl:40)
    variant 'as_written': U = 1/2 (J-1)^2 - ln J   (the script verbatim)
    props = (rho, K, G)."""

    def __init__(self, nd=3, variant="standard"):
        super().__init__(nd)
        self.variant = variant

    def energy(self, gu, props):
        K, G = props[1], props[2]
        F = gu + np.eye(3)
        J = np.linalg.det(F)
        I1 = np.einsum("eij,eij->e", F, F)
        U = 0.5 * (J * J - 1) - np.log(J) if self.variant == "standard" else 0.5 * (J - 1) ** 2 - np.log(J)
        return 0.5 * K * U + 0.5 * G * (J ** (-2.0 / 3.0) * I1 - 3.0)

    def stress_tangent(self, gu, props, so, sn, need_A=True):
        K, G = props[1], props[2]
        I = np.eye(3)
        F = gu + I
        J = np.linalg.det(F)
        H = np.swapaxes(np.linalg.inv(F), 1, 2)  # F^-T : H[i,J]
        I1 = np.einsum("eij,eij->e", F, F)
        m = J ** (-2.0 / 3.0)
        if self.variant == "standard":
            c = 0.5 * K * (J * J - 1.0)
            cp = K * J
        else:
            c = 0.5 * K * (J * J - J - 1.0)
            cp = 0.5 * K * (2.0 * J - 1.0)
        P = c[:, None, None] * H + (G * m)[:, None, None] * (F - (I1 / 3.0)[:, None, None] * H)
        A = None
        if need_A:
            HH = np.einsum("eij,ekl->eijkl", H, H)
            HxH = np.einsum("eil,ekj->eijkl", H, H)
            dev = F - (I1 / 3.0)[:, None, None] * H
            A = (cp * J)[:, None, None, None, None] * HH - c[:, None, None, None, None] * HxH
            A = A + G * (
                -(2.0 / 3.0) * m[:, None, None, None, None] * np.einsum("eij,ekl->eijkl", dev, H)
                + m[:, None, None, None, None] * (
                    np.einsum("ik,jl->ijkl", I, I)[None]
                    - (2.0 / 3.0) * np.einsum("eij,ekl->eijkl", H, F)
                    + (I1 / 3.0)[:, None, None, None, None] * HxH))
        return P, A


class NonSymmetricTest(_Mechanics):
    """TEST law, not in the reference: linear elasticity + beta * delta_ij T_kl (T fixed, non-symmetric), i.e. a
    tangent with A_ijkl != A_klij.  Every law the reference ships has a symmetric tangent, for which the transposed COO
    labelling of its pattern (Assemblers.jl:109-124 vs SparsityPatterns.jl:76-83) cannot be seen; with this one the
    assembled matrix is the TRANSPOSE of dR/dU, exactly what the reference's loops produce.  props = (rho, K, G, beta)."""
    T = np.array([[0.25, 1.0, 0.0], [-0.5, 0.5, 2.0], [3.0, 0.0, 0.75]])

    def stress_tangent(self, gu, props, so, sn, need_A=True):
        P, A = LinearElastic.stress_tangent(self, gu, props, so, sn, need_A)
        beta, I = props[3], np.eye(3)
        P = P + beta * np.einsum("kl,ekl->e", self.T, gu)[:, None, None] * I
        if need_A:
            A = A + beta * np.einsum("ij,kl->ijkl", I, self.T)[None]
        return P, A


class J2Plasticity(_Mechanics):
    """Small-strain J2 plasticity with linear isotropic hardening, radial return.
    The reference only has the hooks (AbstractPhysics{.,.,7}, state_old/state_new views,
    test/mechanics_with_state/TestMechanicsWithState.jl:15-67) -- no J2 law exists there:
    oracle-defined, parity unpinned.
    state (NS=7) = [ep_xx, ep_yy, ep_zz, ep_yz, ep_xz, ep_xy, eqps]; props = (rho,K,G,sigma_y,H)."""
    NS = 7

    @staticmethod
    def _ep_tensor(s):
        ep = np.zeros((s.shape[0], 3, 3))
        ep[:, 0, 0], ep[:, 1, 1], ep[:, 2, 2] = s[:, 0], s[:, 1], s[:, 2]
        ep[:, 1, 2] = ep[:, 2, 1] = s[:, 3]
        ep[:, 0, 2] = ep[:, 2, 0] = s[:, 4]
        ep[:, 0, 1] = ep[:, 1, 0] = s[:, 5]
        return ep

    def stress_tangent(self, gu, props, so, sn, need_A=True):
        K, G, sy, Hh = props[1], props[2], props[3], props[4]
        I = np.eye(3)
        eps = _sym(gu)
        tr = np.trace(eps, axis1=1, axis2=2)
        ep_old = self._ep_tensor(so)
        a_old = so[:, 6]
        e_tr = eps - tr[:, None, None] / 3 * I - ep_old
        s_tr = 2 * G * e_tr
        nrm = np.sqrt(np.einsum("eij,eij->e", s_tr, s_tr))
        q = math.sqrt(1.5) * nrm
        f = q - (sy + Hh * a_old)
        yld = f > 0
        dg = np.where(yld, f / (3 * G + Hh), 0.0)
        safe = np.where(nrm > 0, nrm, 1.0)
        n = s_tr / safe[:, None, None]
        s = s_tr - (2 * G * math.sqrt(1.5) * dg)[:, None, None] * n
        P = K * tr[:, None, None] * I + s
        if sn is not None:
            ep_new = ep_old + (math.sqrt(1.5) * dg)[:, None, None] * n
            sn[:, 0], sn[:, 1], sn[:, 2] = ep_new[:, 0, 0], ep_new[:, 1, 1], ep_new[:, 2, 2]
            sn[:, 3], sn[:, 4], sn[:, 5] = ep_new[:, 1, 2], ep_new[:, 0, 2], ep_new[:, 0, 1]
            sn[:, 6] = a_old + dg
        A = None
        if need_A:
            II = np.einsum("ij,kl->ijkl", I, I)
            Is4 = 0.5 * (np.einsum("ik,jl->ijkl", I, I) + np.einsum("il,jk->ijkl", I, I))
            Idev = Is4 - II / 3
            qs = np.where(q > 0, q, 1.0)
            theta = np.where(yld, 1.0 - 3 * G * dg / qs, 1.0)
            thbar = np.where(yld, 1.0 / (1.0 + Hh / (3 * G)) - (1.0 - theta), 0.0)
            nn = np.einsum("eij,ekl->eijkl", n, n)
            A = (K * II)[None] + (2 * G * theta)[:, None, None, None, None] * Idev[None] \
                - (2 * G * thbar)[:, None, None, None, None] * nn
        return P, A


# --------------------------------------------------------------------------------------
# Assembly (src/assemblers/{Assemblers,Vector,Matrix,MatrixAction}.jl)
# --------------------------------------------------------------------------------------

class Block:
    """One element block: conn (NNPE,NE) 1-based, reference tables, physics, props,
    state_old/new (NS,NQ,NE)."""

    def __init__(self, conn, tables, physics, props=(), state_old=None, state_new=None):
        self.conn = np.asarray(conn, dtype=np.int64)
        self.N, self.dN, self.w = tables
        self.physics = physics
        self.props = np.asarray(props, dtype=float)
        ne, nq, ns = self.conn.shape[1], len(self.w), physics.NS
        self.state_old = np.zeros((ns, nq, ne)) if state_old is None else state_old
        self.state_new = np.zeros((ns, nq, ne)) if state_new is None else state_new


def _gather(F, conn):
    """_element_level_fields_flat (Assemblers.jl:161-173): F (NF,NN) -> (NE,NNPE,NF)."""
    return np.transpose(F[:, conn - 1], (2, 1, 0))


def _loop_q(block, X, U, kind, write_state):
    x_el = _gather(X, block.conn)
    u_el = _gather(U, block.conn)
    ph = block.physics
    out = None
    for q in range(len(block.w)):
        X_q, dN_X, JxW = map_interpolants(block.N[q], block.dN[q], block.w[q], x_el)
        so = block.state_old[:, q, :].T if ph.NS else None
        sn = (block.state_new[:, q, :].T.copy() if (ph.NS and write_state) else None)
        v = getattr(ph, kind + "_q")(block.N[q], X_q, dN_X, JxW, u_el, block.props, so, sn)
        if sn is not None:
            block.state_new[:, q, :] = sn.T
        out = v if out is None else out + v
    return out


def assemble_vector(blocks, X, U, nf):
    """assemble_vector! with `residual` (Vector.jl:25-74 + Assemblers.jl:72-87,402-432).
    X (ND,NN), U (NF,NN) full fields.  Returns flat R (NF*NN) in dof order."""
    R = np.zeros(U.shape[0] * U.shape[1])
    for b in blocks:
        Re = _loop_q(b, X, U, "residual", True)  # (NE, NNPE*NF)
        dc = _dof_conn(b.conn, nf)  # (NDOF, NE) 1-based
        np.add.at(R, (dc.T - 1).ravel(), Re.ravel())
    return R


def assemble_scalar(blocks, X, U):
    """assemble_scalar!(asm, energy, ...) (QuadratureQuantity.jl:4-45): storage[1, q, e] = energy_q per block
    (Assemblers.jl:47-51).  Returns one (NQ, NE) array per block."""
    out = []
    for b in blocks:
        x_el, u_el = _gather(X, b.conn), _gather(U, b.conn)
        vals = np.zeros((len(b.w), b.conn.shape[1]))
        for q in range(len(b.w)):
            X_q, dN_X, JxW = map_interpolants(b.N[q], b.dN[q], b.w[q], x_el)
            so = b.state_old[:, q, :].T if b.physics.NS else None
            vals[q] = b.physics.energy_q(b.N[q], X_q, dN_X, JxW, u_el, b.props, so, None)
        out.append(vals)
    return out


def assemble_lumped_mass(blocks, X, U, nf):
    """assemble_lumped_mass! (LumpedMass.jl:32-60): the AssembledVector path with func = lumped_mass."""
    R = np.zeros(U.shape[0] * U.shape[1])
    for b in blocks:
        Me = _loop_q(b, X, U, "lumped_mass", False)
        dc = _dof_conn(b.conn, nf)
        np.add.at(R, (dc.T - 1).ravel(), Me.ravel())
    return R


def assemble_diagonal(blocks, X, U, nf, kind="stiffness"):
    """assemble_diagonal! (Diagonal.jl:16-74): per quadrature point only diag(K_q) is accumulated
    (Assemblers.jl:42-45), then the nodal scatter of a vector."""
    R = np.zeros(U.shape[0] * U.shape[1])
    for b in blocks:
        Ke = _loop_q(b, X, U, kind, False)                      # (NE, NDOF, NDOF)
        De = np.einsum("eii->ei", Ke)
        dc = _dof_conn(b.conn, nf)
        np.add.at(R, (dc.T - 1).ravel(), De.ravel())
    return R


def assemble_matrix_coo(blocks, X, U, nf, kind="stiffness"):
    """assemble_matrix! (Matrix.jl:35-75): COO storage, storage[(e)*NDOF^2 + k] = K_el.data[k]
    with K_el column-major (Assemblers.jl:109-124)."""
    out = []
    for b in blocks:
        Ke = _loop_q(b, X, U, kind, False)  # (NE, NDOF, NDOF) [r, c]
        out.append(np.transpose(Ke, (0, 2, 1)).reshape(-1))  # column-major flatten per element
    return np.concatenate(out)


def assemble_matrix_action(blocks, X, U, V, nf, kind="stiffness"):
    """assemble_matrix_action! (MatrixAction.jl:154-238): Kv_el = K_el * v_el, nodal scatter."""
    out = np.zeros(U.shape[0] * U.shape[1])
    for b in blocks:
        Ke = _loop_q(b, X, U, kind, False)
        v_el = _gather(V, b.conn).reshape(Ke.shape[0], -1)
        Kv = np.einsum("erc,ec->er", Ke, v_el)
        dc = _dof_conn(b.conn, nf)
        np.add.at(out, (dc.T - 1).ravel(), Kv.ravel())
    return out


# --------------------------------------------------------------------------------------
# A small "assembler" mirroring SparseMatrixAssembler + Parameters for the oracle side
# --------------------------------------------------------------------------------------

class OracleAssembler:
    """SparseMatrixAssembler + the bits of Parameters the hot path touches
    (SparseMatrixAssembler.jl:7-124, Parameters.jl:404-425)."""

    def __init__(self, coords, blocks, nf, condensed=False, matrix_type="csr"):
        self.X = np.asarray(coords, dtype=float)
        self.blocks = blocks
        self.nf = nf
        self.nn = self.X.shape[1]
        self.ndof = nf * self.nn
        self.condensed = condensed
        self.matrix_type = matrix_type
        self.field = np.zeros(self.ndof)
        self.hvp_scratch = np.zeros(self.ndof)
        self.bc_dofs = np.zeros(0, dtype=np.int64)
        self.bc_vals = np.zeros(0)
        self.update_dofs([], [], [])

    # update_dofs!(asm, dbcs, pbcs)  (SparseMatrixAssembler.jl:228-274)
    def update_dofs(self, dirichlet_dofs, per_a=(), per_b=()):
        dd = np.unique(np.asarray(dirichlet_dofs, dtype=np.int64))
        self.dof = update_dofs(self.nf, self.nn, dd, per_a, per_b)
        self.constraint = np.zeros(self.ndof)
        self.constraint[dd - 1] = 1.0
        self.pattern = matrix_pattern([b.conn for b in self.blocks], self.nf, self.dof,
                                      condensed=self.condensed)
        self.bc_dofs = dd
        self.bc_vals = np.zeros(len(dd))
        self.n = self.ndof if self.condensed else len(self.dof["unknown_dofs"])

    def create_unknowns(self):
        return np.zeros(self.n)

    # _update_for_assembly! (Parameters.jl:404-425)
    def _update_field(self, field, Uu, with_bcs=True):
        if with_bcs:
            field[self.bc_dofs - 1] = self.bc_vals
        ud = self.dof["unknown_dofs"] - 1
        field[ud] = Uu[ud] if self.condensed else Uu
        if with_bcs:
            for a, b in zip(self.dof["periodic_side_a_dofs"], self.dof["periodic_side_b_dofs"]):
                field[b - 1] = field[a - 1]

    def _U(self):
        return self.field.reshape(self.nn, self.nf).T

    def assemble_vector(self, Uu):
        self._update_field(self.field, Uu)
        self.residual_storage = assemble_vector(self.blocks, self.X, self._U(), self.nf)

    # external loads: p.neumann_bcs / p.sources (Parameters.jl:37-73); values are set by the caller like
    # update_bc_values! does (NeumannBCs.jl:157-171, Sources.jl:55-66)
    def add_neumann_bc(self, snodes, tables, vals):
        self.neumann = getattr(self, "neumann", []) + [(np.asarray(snodes, dtype=np.int64), tables, np.asarray(vals, dtype=float))]

    def add_source(self, block_index, vals):
        self.sources = getattr(self, "sources", []) + [(block_index, np.asarray(vals, dtype=float))]

    def assemble_vector_neumann_bc(self, Uu=None):
        """assemble_vector_neumann_bc!(asm, Uu, p) (WeaklyEnforcedBCs.jl:4-15): adds to the residual storage"""
        for snodes, tables, vals in getattr(self, "neumann", []):
            assemble_vector_neumann_bc(self.residual_storage, snodes, tables, vals, self.X, self.nf)

    def assemble_vector_source(self, Uu=None):
        """assemble_vector_source!(asm, Uu, p) (Source.jl:10-42): adds to the residual storage"""
        for b, vals in getattr(self, "sources", []):
            assemble_vector_source(self.residual_storage, self.blocks[b], vals, self.X, self.nf)

    def assemble_stiffness(self, Uu, kind="stiffness"):
        self._update_field(self.field, Uu)
        self.stiffness_storage = assemble_matrix_coo(self.blocks, self.X, self._U(), self.nf, kind)

    # assemble_lumped_mass! / assemble_diagonal! write the residual storage (LumpedMass.jl:36, Diagonal.jl:19)
    def assemble_lumped_mass(self, Uu):
        self._update_field(self.field, Uu)
        self.residual_storage = assemble_lumped_mass(self.blocks, self.X, self._U(), self.nf)

    def assemble_diagonal(self, Uu, kind="stiffness"):
        self._update_field(self.field, Uu)
        self.residual_storage = assemble_diagonal(self.blocks, self.X, self._U(), self.nf, kind)

    def assemble_scalar(self, Uu):
        self._update_field(self.field, Uu)
        self.scalar_quadrature_storage = assemble_scalar(self.blocks, self.X, self._U())

    def vector_values(self):
        """lumped_mass(asm) / diagonal(asm) (LumpedMass.jl:70-80, Diagonal.jl:76-89): no constraint scaling, no fold"""
        R = self.residual_storage
        return R if self.condensed else R[self.dof["unknown_dofs"] - 1]

    def assemble_matrix_action(self, Uu, Vu, kind="stiffness"):
        self._update_field(self.field, Uu)
        self._update_field(self.hvp_scratch, Vu, with_bcs=False)
        V = self.hvp_scratch.reshape(self.nn, self.nf).T
        self.action_storage = assemble_matrix_action(self.blocks, self.X, self._U(), V, self.nf, kind)

    # accessors (Assemblers.jl:310-388, assemblers/Utils.jl:53-167)
    def residual(self):
        R = self.residual_storage
        if self.condensed:
            R *= (1.0 - self.constraint)
            return R
        for a, b in zip(self.dof["periodic_side_a_dofs"], self.dof["periodic_side_b_dofs"]):
            R[a - 1] += R[b - 1]
        return R[self.dof["unknown_dofs"] - 1]

    def hvp(self, v):
        Av = self.action_storage
        if self.condensed:
            c = self.constraint
            Av[:] = (1.0 - c) * Av + c * v
            return Av
        return Av[self.dof["unknown_dofs"] - 1]

    def stiffness(self):
        """Returns (ptr, idx, nzval) 1-based: CSC (colptr,rowval) or CSR (rowptr,colval)."""
        p = self.pattern
        vals = self.stiffness_storage[p["unknown_dofs"] - 1]
        colptr, rowval, nz = sparse_csc(p["Is"], p["Js"], vals, self.n)
        if self.matrix_type == "csc":
            if self.condensed:
                cols = np.repeat(np.arange(1, self.n + 1), np.diff(colptr))
                tr = nz[rowval == cols].sum()
                pen = 1.0e6 * tr / self.n
                c = self.constraint[cols - 1]
                nz = (1.0 - c) * nz
                d = rowval == cols
                nz[d] += pen * c[d]
            return colptr, rowval, nz
        rowptr, colval, nzr = csc_to_csr(colptr, rowval, nz, self.n)
        if self.condensed:
            rows = np.repeat(np.arange(1, self.n + 1), np.diff(rowptr))
            tr = nzr[colval == rows].sum()
            pen = 1.0e6 * tr / self.n
            c = self.constraint[rows - 1]
            nzr = (1.0 - c) * nzr
            d = colval == rows
            nzr[d] += pen * c[d]
        return rowptr, colval, nzr

    def stiffness_scipy(self):
        import scipy.sparse as sp
        ptr, idx, nz = self.stiffness()
        if self.matrix_type == "csc":
            return sp.csc_matrix((nz, idx - 1, ptr - 1), shape=(self.n, self.n))
        return sp.csr_matrix((nz, idx - 1, ptr - 1), shape=(self.n, self.n))


# --------------------------------------------------------------------------------------
# Solver loop (src/Solvers.jl:128-220) -- the caller of the hot path
# --------------------------------------------------------------------------------------

def cg(A, b, atol=None, rtol=None, itmax=None):
    """Krylov.jl `cg` defaults: atol = rtol = sqrt(eps), itmax = 2n, x0 = 0,
    stop when ||r|| <= atol + rtol*||r0||.  Returns x, iterations."""
    n = len(b)
    eps = np.finfo(float).eps
    atol = math.sqrt(eps) if atol is None else atol
    rtol = math.sqrt(eps) if rtol is None else rtol
    itmax = 2 * n if itmax is None else itmax
    x = np.zeros(n)
    r = b.copy()
    p = r.copy()
    gamma = r @ r
    rn0 = math.sqrt(gamma)
    tol = atol + rtol * rn0
    it = 0
    if rn0 == 0:
        return x, 0
    while math.sqrt(gamma) > tol and it < itmax:
        Ap = A @ p
        alpha = gamma / (p @ Ap)
        x += alpha * p
        r -= alpha * Ap
        gnew = r @ r
        p = r + (gnew / gamma) * p
        gamma = gnew
        it += 1
    return x, it


def newton_solve(asm: OracleAssembler, Uu, max_iters=10, abs_tol=1e-12, rel_tol=1e-12, direct=False):
    """solve!(NewtonSolver, Uu, p) (Solvers.jl:193-220) with IterativeLinearSolver (:128-153).
    Returns (Uu, n_newton_iterations, [cg iterations], [residual norms])."""
    import scipy.sparse.linalg as spla
    R0 = None
    hist, cgits = [], []
    for it in range(1, max_iters + 1):
        asm.assemble_vector(Uu)
        asm.assemble_vector_source(Uu)       # Solvers.jl:135
        asm.assemble_vector_neumann_bc(Uu)   # Solvers.jl:136
        R = asm.residual().copy()
        asm.assemble_stiffness(Uu)
        K = asm.stiffness_scipy()
        if direct:
            dU = spla.spsolve(K.tocsc(), R); nit = 0
        else:
            dU, nit = cg(K, R)
        Uu = Uu - dU
        cgits.append(nit)
        nR, ndU = np.linalg.norm(R), np.linalg.norm(dU)
        if R0 is None:
            R0 = nR
        hist.append(nR)
        if ndU < abs_tol or nR < abs_tol or (R0 > 0 and nR / R0 < rel_tol):
            return Uu, it, cgits, hist
    return Uu, max_iters, cgits, hist
