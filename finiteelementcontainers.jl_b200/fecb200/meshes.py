"""Meshes: StructuredMesh / UnstructuredMesh (src/meshes/*.jl) -- the input side of the path.

Containers follow the reference: `nodal_coords` is an H1Field (ND, NN), `element_conns`
maps block name -> (NNPE, NE) Int64 **1-based** connectivity, `nodeset_nodes` /
`sideset_nodes` map names -> 1-based node ids.
"""
from __future__ import annotations

import numpy as np

from .fields import H1Field

# Exodus side -> local nodes (0-based) used to turn side sets into node lists
_SIDE_NODES = {
    "QUAD4": [(0, 1), (1, 2), (2, 3), (3, 0)],
    "TRI3": [(0, 1), (1, 2), (2, 0)],
    "HEX8": [(0, 1, 5, 4), (1, 2, 6, 5), (2, 3, 7, 6), (0, 4, 7, 3), (0, 3, 2, 1), (4, 5, 6, 7)],
    "TETRA4": [(0, 1, 3), (1, 2, 3), (0, 3, 2), (0, 2, 1)],
    "TETRA10": [(0, 1, 3, 4, 8, 7), (1, 2, 3, 5, 9, 8), (0, 3, 2, 7, 9, 6), (0, 2, 1, 6, 5, 4)],
}
_CANON = {"QUAD": "QUAD4", "QUAD4": "QUAD4", "TRI": "TRI3", "TRI3": "TRI3", "HEX": "HEX8", "HEX8": "HEX8",
          "TET": "TETRA4", "TET4": "TETRA4", "TETRA": "TETRA4", "TETRA4": "TETRA4", "TETRA10": "TETRA10",
          "TET10": "TETRA10"}


class AbstractMesh:
    nodal_coords: H1Field
    element_block_names: list
    element_types: dict
    element_conns: dict
    nodeset_nodes: dict
    sideset_nodes: dict

    # side sets as (elements, sides, side_nodes), the fields of the reference's mesh structs
    # (src/meshes/StructuredMesh.jl:21-25); elements are 1-based ids over the concatenated blocks
    sideset_elems: dict
    sideset_sides: dict
    sideset_side_nodes: dict

    def _set_sidesets(self, sets):
        """sets: {name: (elements, sides)} -> sideset_elems / sideset_sides / sideset_side_nodes"""
        self.sideset_elems, self.sideset_sides, self.sideset_side_nodes = {}, {}, {}
        offs = np.cumsum([0] + [self.element_conns[b].shape[1] for b in self.element_block_names])
        for name, (el, sd) in sets.items():
            el, sd = np.asarray(el, dtype=np.int64), np.asarray(sd, dtype=np.int64)
            self.sideset_elems[name], self.sideset_sides[name] = el, sd
            if not len(el):
                self.sideset_side_nodes[name] = np.zeros((0, 0), dtype=np.int64)
                continue
            # vectorised over the sides: one fancy-indexing gather per (block, side number) group
            blk = np.searchsorted(offs, el - 1, side="right") - 1
            out, nnps = None, None
            for b in np.unique(blk):
                bn = self.element_block_names[int(b)]
                conn, tab = self.element_conns[bn], _SIDE_NODES[self.element_types[bn]]
                for sn in np.unique(sd[blk == b]):
                    sel = np.nonzero((blk == b) & (sd == sn))[0]
                    loc = list(tab[int(sn) - 1])
                    if out is None:
                        nnps = len(loc)
                        out = np.empty((nnps, len(el)), dtype=np.int64)
                    out[:, sel] = conn[np.asarray(loc)[:, None], (el[sel] - 1 - offs[b])[None, :]]
            self.sideset_side_nodes[name] = out

    def _sidesets_from_nodesets(self, nodesets, blocks=None):
        """every element side whose nodes all belong to the node set (meshes without side-set records);
        `blocks` restricts the search (rank-local meshes: only the `owned` block carries surface loads)"""
        sets, off = {}, 0
        per_block = []
        for bn in self.element_block_names:
            if blocks is None or bn in blocks:
                per_block.append((off, self.element_conns[bn], _SIDE_NODES[self.element_types[bn]]))
            off += self.element_conns[bn].shape[1]
        for name, nodes in nodesets.items():
            mark = np.zeros(self.num_nodes() + 1, dtype=bool)
            mark[np.asarray(nodes, dtype=np.int64)] = True
            el, sd = [], []
            for o, conn, tab in per_block:
                for s, loc in enumerate(tab):
                    hit = np.nonzero(mark[conn[list(loc)]].all(axis=0))[0]
                    el.append(hit + 1 + o)
                    sd.append(np.full(len(hit), s + 1))
            el, sd = np.concatenate(el), np.concatenate(sd)
            order = np.argsort(el, kind="stable")
            sets[name] = (el[order], sd[order])
        return sets

    def num_dimensions(self):
        return self.nodal_coords.shape[0]

    def num_nodes(self):
        return self.nodal_coords.shape[1]

    def __repr__(self):
        ne = sum(c.shape[1] for c in self.element_conns.values())
        return (f"{type(self).__name__}: {self.num_dimensions()}-D, {self.num_nodes()} nodes, {ne} elements, "
                f"blocks {[(b, self.element_types[b]) for b in self.element_block_names]}")


class StructuredMesh(AbstractMesh):
    """StructuredMesh(element_type, mins, maxs, counts)  (src/meshes/StructuredMesh.jl:28-83).

    `counts` are NODE counts per axis.  hex8: node(i,j,k) = i + Nx(j-1) + NxNy(k-1), elements
    enumerated ex outer / ey / ez inner, Exodus local ordering (:85-131); quad4 ex outer / ey
    inner (:232-255); tri3 = each quad split (1,2,3),(1,3,4) (:329-352).
    """

    def __init__(self, element_type, mins, maxs, counts):
        if not (len(mins) == len(maxs) == len(counts)):
            raise AssertionError("mins, maxs and counts must have the same length")
        for m1, m2 in zip(mins, maxs):
            if m1 >= m2:
                raise IndexError("Dimension has negative or zero length")  # BoundsError
        t = element_type.upper()
        if t in "HEX":
            et, (coords, conn, nsets) = "HEX8", self._hex8(mins, maxs, counts)
        elif t in "QUAD":
            et, (coords, conn, nsets) = "QUAD4", self._quad4(mins, maxs, counts)
        elif t in "TET":
            raise AssertionError("Implement tet case")
        elif t in "TRI":
            et = "TRI3"
            coords, cq, nsets = self._quad4(mins, maxs, counts)
            conn = np.empty((3, 2 * cq.shape[1]), dtype=np.int64)
            conn[:, 0::2] = cq[[0, 1, 2]]
            conn[:, 1::2] = cq[[0, 2, 3]]
        else:
            raise ValueError(f"Unsupported element type {element_type}")
        self.nodal_coords = H1Field(coords)
        self.element_block_names = ["block_1"]
        self.element_types = {"block_1": et}
        self.element_conns = {"block_1": conn}
        self.nodeset_nodes = nsets
        self.sideset_nodes = dict(nsets)  # node lists of the boundary sides == node sets
        self._set_sidesets(self._structured_sidesets(et, counts))

    @staticmethod
    def _structured_sidesets(et, counts):
        """(elements, sides) per named boundary.  QUAD4 / TRI3: the reference's loops (src/meshes/StructuredMesh.jl:
        257-330, 355-431).  HEX8: the reference's `_hex8_ssets` (:133-230) is unfinished (side numbers that do not lie
        on the named boundary, a loop over an undefined k, side-node matrices sized for one row of faces); the faces
        that do lie on the boundary are used: bottom/top = y-min/max (sides 1/3), left/right = x-min/max (4/2),
        back/front = z-min/max (5/6), consistent with the node sets (:452-470)."""
        if et in ("QUAD4", "TRI3"):
            Ex, Ey = counts[0] - 1, counts[1] - 1
            quad = lambda i, j: (i - 1) * Ey + j
            I, J = np.arange(1, Ex + 1), np.arange(1, Ey + 1)
            if et == "QUAD4":
                return {"bottom": (quad(I, 1), np.full(Ex, 1)), "right": (quad(Ex, J), np.full(Ey, 2)),
                        "top": (quad(I, Ey), np.full(Ex, 3)), "left": (quad(1, J), np.full(Ey, 4))}
            return {"bottom": (2 * quad(I, 1) - 1, np.full(Ex, 1)), "right": (2 * quad(Ex, J) - 1, np.full(Ey, 2)),
                    "top": (2 * quad(I, Ey), np.full(Ex, 2)), "left": (2 * quad(1, J), np.full(Ey, 3))}
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

    @staticmethod
    def _quad4(mins, maxs, counts):
        Nx, Ny = counts
        xs, ys = np.linspace(mins[0], maxs[0], Nx), np.linspace(mins[1], maxs[1], Ny)
        coords = np.empty((2, Nx * Ny))
        coords[0] = np.tile(xs, Ny)
        coords[1] = np.repeat(ys, Nx)
        ex = np.repeat(np.arange(1, Nx), Ny - 1)   # ex outer
        ey = np.tile(np.arange(1, Ny), Nx - 1)     # ey inner
        n = lambda i, j: i + Nx * (j - 1)
        conn = np.stack([n(ex, ey), n(ex + 1, ey), n(ex + 1, ey + 1), n(ex, ey + 1)]).astype(np.int64)
        I, J = np.arange(1, Nx + 1), np.arange(1, Ny + 1)
        nsets = {"bottom": n(I, 1), "right": n(Nx, J), "top": n(I, Ny), "left": n(1, J)}
        return coords, conn, {k: v.astype(np.int64) for k, v in nsets.items()}

    @staticmethod
    def _hex8(mins, maxs, counts):
        Nx, Ny, Nz = counts
        xs = np.linspace(mins[0], maxs[0], Nx)
        ys = np.linspace(mins[1], maxs[1], Ny)
        zs = np.linspace(mins[2], maxs[2], Nz)
        coords = np.empty((3, Nx * Ny * Nz))
        coords[0] = np.tile(xs, Ny * Nz)
        coords[1] = np.tile(np.repeat(ys, Nx), Nz)
        coords[2] = np.repeat(zs, Nx * Ny)
        Ex, Ey, Ez = Nx - 1, Ny - 1, Nz - 1
        n = lambda i, j, k: i + Nx * (j - 1) + Nx * Ny * (k - 1)
        # elements ex outer / ey / ez inner (StructuredMesh.jl:113-127): first node of every element by broadcasting, the
        # other seven are fixed offsets from it (Exodus ordering, :116-123)
        n0 = (np.arange(1, Ex + 1, dtype=np.int64)[:, None, None] + Nx * np.arange(Ey, dtype=np.int64)[None, :, None]
              + Nx * Ny * np.arange(Ez, dtype=np.int64)[None, None, :]).reshape(-1)
        offs = np.array([0, 1, 1 + Nx, Nx, Nx * Ny, 1 + Nx * Ny, 1 + Nx + Nx * Ny, Nx + Nx * Ny], dtype=np.int64)
        conn = n0[None, :] + offs[:, None]
        I, J, K = np.arange(1, Nx + 1), np.arange(1, Ny + 1), np.arange(1, Nz + 1)

        def face(f, A, B):  # [f(a, b) for a in A, b in B] |> vec   (a fastest)
            bb, aa = np.meshgrid(B, A, indexing="ij")
            return f(aa.ravel(), bb.ravel()).astype(np.int64)

        nsets = {
            "bottom": face(lambda i, k: n(i, 1, k), I, K), "top": face(lambda i, k: n(i, Ny, k), I, K),
            "left": face(lambda j, k: n(1, j, k), J, K), "right": face(lambda j, k: n(Nx, j, k), J, K),
            "back": face(lambda i, j: n(i, j, 1), I, J), "front": face(lambda i, j: n(i, j, Nz), I, J),
        }
        return coords, conn, nsets


class KuhnTet10Mesh(AbstractMesh):
    """Synthetic TETRA10 mesh for the stateful-mechanics configuration (BASELINE.json config 4).
    The reference has no tet generator (StructuredMesh.jl:50-51); this one splits every cell of an
    n^3 grid into six Kuhn tetrahedra sharing the (0,0,0)-(1,1,1) diagonal; the P2 nodes are the
    (2n+1)^3 half-lattice points (x fastest).  Exodus TETRA10 ordering."""

    def __init__(self, n, lo=0.0, hi=1.0):
        M = 2 * n + 1
        g = np.linspace(lo, hi, M)
        coords = np.empty((3, M ** 3))
        coords[0] = np.tile(g, M * M)
        coords[1] = np.tile(np.repeat(g, M), M)
        coords[2] = np.repeat(g, M * M)
        cz = np.tile(np.arange(n), n * n)
        cy = np.tile(np.repeat(np.arange(n), n), n)
        cx = np.repeat(np.arange(n), n * n)
        base = 2 * np.stack([cx, cy, cz], axis=1)
        nid = lambda p: p[:, 0] + M * p[:, 1] + M * M * p[:, 2] + 1
        E = np.eye(3, dtype=np.int64)
        NC = n ** 3
        conn = np.empty((10, 6 * NC), dtype=np.int64)
        for t, p in enumerate([(0, 1, 2), (0, 2, 1), (1, 0, 2), (1, 2, 0), (2, 0, 1), (2, 1, 0)]):
            v = [np.zeros(3, dtype=np.int64), E[p[0]], E[p[0]] + E[p[1]], np.ones(3, dtype=np.int64)]
            if np.dot(np.cross(v[1] - v[0], v[2] - v[0]), v[3] - v[0]) < 0:
                v[1], v[2] = v[2], v[1]
            V = [base + 2 * vi for vi in v]
            for a in range(4):
                conn[a, t::6] = nid(V[a])
            for m, (a, b) in enumerate([(0, 1), (1, 2), (0, 2), (0, 3), (1, 3), (2, 3)]):
                conn[4 + m, t::6] = nid((V[a] + V[b]) // 2)
        idx = np.arange(M ** 3)
        i3, j3, k3 = idx % M, (idx // M) % M, idx // (M * M)
        self.nodal_coords = H1Field(coords)
        self.element_block_names = ["block_1"]
        self.element_types = {"block_1": "TETRA10"}
        self.element_conns = {"block_1": conn}
        self.nodeset_nodes = {"bottom": idx[j3 == 0] + 1, "top": idx[j3 == M - 1] + 1, "left": idx[i3 == 0] + 1,
                              "right": idx[i3 == M - 1] + 1, "back": idx[k3 == 0] + 1, "front": idx[k3 == M - 1] + 1}
        self.sideset_nodes = dict(self.nodeset_nodes)
        self._set_sidesets(self._sidesets_from_nodesets(self.nodeset_nodes))


class UnstructuredMesh(AbstractMesh):
    """UnstructuredMesh(file)  (src/meshes/UnstructuredMesh.jl, src/meshes/Exodus.jl).
    Reads Exodus II files (NetCDF classic) with scipy, or the .npz fixtures under tests/golden/."""

    def __init__(self, path=None, *, data=None):
        if data is None:
            data = self._read_npz(path) if str(path).endswith(".npz") else self._read_exodus(path)
        self.nodal_coords = H1Field(np.ascontiguousarray(data["coords"], dtype=float))
        self.element_block_names = list(data["block_names"])
        self.element_types = {b: _CANON[t.upper()] for b, t in zip(data["block_names"], data["types"])}
        self.element_conns = {b: np.ascontiguousarray(c, dtype=np.int64) for b, c in zip(data["block_names"], data["conns"])}
        self.nodeset_nodes = dict(data["nodesets"])
        self.sideset_nodes = dict(data["sidesets"])
        # Exodus files carry (elem_ss, side_ss); the .npz fixtures only node lists -> sides recovered from them
        self._set_sidesets(data["sideset_records"] if "sideset_records" in data
                           else self._sidesets_from_nodesets(self.sideset_nodes))

    @staticmethod
    def _read_npz(path):
        z = np.load(path, allow_pickle=False)
        nb = int(z["n_blocks"])
        d = dict(coords=z["coords"], block_names=[str(z[f"block_name_{b}"]) for b in range(nb)],
                 types=[str(z[f"type_{b}"]) for b in range(nb)], conns=[z[f"conn_{b}"] for b in range(nb)])
        d["nodesets"] = {str(n): z[f"nodeset_{i}"] for i, n in enumerate(z["nodeset_names"])}
        d["sidesets"] = {str(n): z[f"sideset_nodes_{i}"] for i, n in enumerate(z["sideset_names"])}
        return d

    @staticmethod
    def _read_exodus(path):
        from scipy.io import netcdf_file
        nc = netcdf_file(path, "r", mmap=False)

        def names(key, n, prefix, ids):
            out = []
            for i in range(n):
                s = b"".join(nc.variables[key].data[i]).split(b"\x00")[0].decode().strip() if key in nc.variables else ""
                out.append(s if s else f"{prefix}_{ids[i]}")
            return out

        nd = nc.dimensions["num_dim"]
        coords = np.stack([np.array(nc.variables["coord" + "xyz"[i]].data, dtype=float) for i in range(nd)])
        nb = nc.dimensions["num_el_blk"]
        conns = [np.array(nc.variables[f"connect{b+1}"].data, dtype=np.int64).T.copy() for b in range(nb)]
        types = [nc.variables[f"connect{b+1}"].elem_type.decode().strip().upper() for b in range(nb)]
        bnames = names("eb_names", nb, "block", np.array(nc.variables["eb_prop1"].data))
        nns = nc.dimensions.get("num_node_sets", 0) or 0
        nss = nc.dimensions.get("num_side_sets", 0) or 0
        nsn = names("ns_names", nns, "nset", np.array(nc.variables["ns_prop1"].data)) if nns else []
        ssn = names("ss_names", nss, "sset", np.array(nc.variables["ss_prop1"].data)) if nss else []
        nodesets = {nsn[i]: np.array(nc.variables[f"node_ns{i+1}"].data, dtype=np.int64) for i in range(nns)}
        offs = np.cumsum([0] + [c.shape[1] for c in conns])
        sidesets, records = {}, {}
        for i in range(nss):
            el = np.array(nc.variables[f"elem_ss{i+1}"].data, dtype=np.int64)
            sd = np.array(nc.variables[f"side_ss{i+1}"].data, dtype=np.int64)
            records[ssn[i]] = (el, sd)
            nodes = []
            for e, s in zip(el, sd):
                b = int(np.searchsorted(offs, e - 1, side="right") - 1)
                loc = _SIDE_NODES[_CANON[types[b]]][s - 1]
                nodes.extend(conns[b][list(loc), e - 1 - offs[b]].tolist())
            _, first = np.unique(nodes, return_index=True)
            sidesets[ssn[i]] = np.array(nodes, dtype=np.int64)[np.sort(first)]
        return dict(coords=coords, block_names=bnames, types=types, conns=conns, nodesets=nodesets, sidesets=sidesets,
                    sideset_records=records)
