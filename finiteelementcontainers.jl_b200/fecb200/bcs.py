"""Dirichlet boundary conditions (src/bcs/DirichletBCs.jl) -- the caller-side bookkeeping that
feeds `U[bc_dofs] = vals` before every assembly."""
from __future__ import annotations

import numpy as np


class DirichletBC:
    """DirichletBC(var_name, func; block_name | nodeset_name | sideset_name)  (:9-41).
    `func(X, t)` is called vectorised: X is (n, ND), returns (n,) (or a scalar)."""

    def __init__(self, var_name, func, *, block_name=None, nodeset_name=None, sideset_name=None):
        given = [x is not None for x in (block_name, nodeset_name, sideset_name)]
        if sum(given) == 0:
            raise ValueError("block_name, nodeset_name, or sideset_name required as input arguments in DirichletBC")
        if sum(given) != 1:
            raise ValueError("More than one entity type specificed in DirichletBC")
        self.var_name, self.func = var_name, func
        self.block_name, self.nset_name, self.sset_name = block_name, nodeset_name, sideset_name


class DirichletBCs:
    """Container (:43-120, :380-418): concatenated `dofs`, `nodes`, `vals` of all BCs, each BC's
    entries sorted and de-duplicated by dof like `_unique_sort_perm`."""

    def __init__(self, mesh, dof, dbcs):
        self.bc_funcs, self.bc_lengths = [], []
        dofs, nodes = [], []
        nf = dof.nf
        for bc in dbcs:
            d = dof.dof_index(bc.var_name)
            if bc.block_name is not None:
                n = np.unique(mesh.element_conns[bc.block_name])
            elif bc.nset_name is not None:
                n = np.asarray(mesh.nodeset_nodes[bc.nset_name], dtype=np.int64)
            else:
                n = np.asarray(mesh.sideset_nodes[bc.sset_name], dtype=np.int64)
            g = nf * (n - 1) + d + 1
            g, first = np.unique(g, return_index=True)
            dofs.append(g)
            nodes.append(n[first])
            self.bc_funcs.append(bc.func)
            self.bc_lengths.append(len(g))
        self.dofs = np.concatenate(dofs).astype(np.int64) if dofs else np.zeros(0, dtype=np.int64)
        self.nodes = np.concatenate(nodes).astype(np.int64) if nodes else np.zeros(0, dtype=np.int64)
        self.vals = np.zeros(len(self.dofs))

    def __len__(self):
        return len(self.bc_funcs)

    def dirichlet_dofs(self):
        return np.unique(self.dofs)  # unique(sort(...))  (:396-398)

    def update_bc_values(self, X, t):
        """update_bc_values!(bcs, X, t) (:400-409)"""
        off = 0
        for func, n in zip(self.bc_funcs, self.bc_lengths):
            nodes = self.nodes[off:off + n]
            v = func(np.asarray(X)[:, nodes - 1].T, t)
            self.vals[off:off + n] = np.broadcast_to(np.asarray(v, dtype=float), (n,))
            off += n


class TimeStepper:
    """TimeStepper(t0, t1, n) (src/TimeSteppers.jl)"""

    def __init__(self, t0=0.0, t1=0.0, n=1):
        self.time_start, self.time_end = float(t0), float(t1)
        self.time_current = float(t0)
        self.dt = (float(t1) - float(t0)) / n if n else 0.0


def _as_values(v, n, nf):
    """func(X, t) may return (n, NF), (NF,), (n,) for NF = 1, or a scalar; -> (n, NF)"""
    v = np.asarray(v, dtype=float)
    if v.ndim == 2:
        return np.ascontiguousarray(np.broadcast_to(v, (n, nf)))
    if v.ndim == 1 and v.shape[0] == nf and (nf > 1 or n == 1):
        return np.ascontiguousarray(np.broadcast_to(v[None, :], (n, nf)))
    return np.ascontiguousarray(np.broadcast_to(v.reshape(-1, 1) if v.ndim else v, (n, nf)))


class NeumannBC:
    """NeumannBC(var_name, func, sset_name)  (src/bcs/NeumannBCs.jl:7-20).  `func(X, t)` returns the flux vector
    (all NF components, `SVector{NF}` in the reference); it is called vectorised: X is (n, ND), result (n, NF).
    Sign convention: the assembler ADDS +int g N dGamma to the residual (test/laplace_with_source/TestLaplace.jl:438-440)."""

    def __init__(self, var_name, func, sset_name):
        self.var_name, self.func, self.sset_name = var_name, func, sset_name


class NeumannBCs:
    """NeumannBCs(mesh, dof, neumann_bcs) (src/bcs/NeumannBCs.jl:78-131) with `_setup_sideset`
    (src/bcs/BoundaryConditions.jl:337-411): one cache per BC holding the side set's elements (block-local), sides,
    surface connectivity, the block's surface tables and `vals[NF, nqs, nsides]`."""

    def __init__(self, mesh, dof, neumann_bcs):
        fspace = dof.var.fspace
        self.nf = dof.nf
        self.bc_funcs, self.bc_caches, self.block_ids, self.block_names = [], [], [], []
        offs = np.cumsum([0] + [mesh.element_conns[b].shape[1] for b in mesh.element_block_names])
        for bc in neumann_bcs:
            dof.dof_index(bc.var_name)  # ValueError if the variable does not exist (_dof_index_from_var_name)
            el = np.asarray(mesh.sideset_elems[bc.sset_name], dtype=np.int64)
            sd = np.asarray(mesh.sideset_sides[bc.sset_name], dtype=np.int64)
            blocks = np.searchsorted(offs, el - 1, side="right") - 1
            if len(np.unique(blocks)) > 1:
                raise AssertionError("Sidesets need to be in a single block")
            b = int(blocks[0]) if len(blocks) else 0
            Ns, dNs, ws = fspace.ref_fes[b].surface_tables()
            sn = np.ascontiguousarray(mesh.sideset_side_nodes[bc.sset_name], dtype=np.int64)
            if not len(sd):                       # e.g. a rank-local mesh that does not touch this boundary
                sn = np.zeros((Ns.shape[1], 0), dtype=np.int64)
            cache = dict(block=b, elements=el - offs[b], sides=sd, side_nodes=sn,
                         Ns=Ns, dNs=dNs, ws=ws, vals=np.zeros((self.nf, len(ws), len(sd)), order="F"))
            self.bc_caches.append(cache)
            self.bc_funcs.append(bc.func)
            self.block_ids.append(b)
            self.block_names.append(mesh.element_block_names[b])

    def __len__(self):
        return len(self.bc_caches)

    def update_bc_values(self, X, t):
        """update_bc_values!(bcs, asm, X, t) (:157-171, :60-71): vals[q, e] = func(X_q, t) at the surface points"""
        X = np.asarray(X)
        for func, c in zip(self.bc_funcs, self.bc_caches):
            if c["side_nodes"].shape[1] == 0:
                continue
            xs = X[:, c["side_nodes"] - 1]                               # (ND, nnps, nsides)
            Xq = np.einsum("qa,dae->eqd", c["Ns"], xs)                   # (nsides, nqs, ND)
            n = Xq.shape[0] * Xq.shape[1]
            v = _as_values(func(Xq.reshape(n, -1), t), n, self.nf)       # (nsides*nqs, NF), q fastest
            c["vals"] = np.asfortranarray(v.reshape(Xq.shape[0], Xq.shape[1], self.nf).transpose(2, 1, 0))


class RobinBC:
    """RobinBC(var_name, func, sset_name)  (src/bcs/RobinBCs.jl:7-20).  `func(X, t, u)` returns the flux vector (NF
    components) as a function of the field value at the surface point; called vectorised: X (n, ND), u (n, NF),
    result (n, NF).  The device path needs the law in AFFINE form g0(X, t) + D(X, t) u (the reference differentiates
    the closure with ForwardDiff, :72-75; a closure cannot cross the C ABI): g0 and D are recovered from func at u = 0
    and u = e_c, and a law that is not affine in u is rejected loudly."""

    def __init__(self, var_name, func, sset_name):
        self.var_name, self.func, self.sset_name = var_name, func, sset_name


class RobinBCs(NeumannBCs):
    """RobinBCs(mesh, dof, robin_bcs) (src/bcs/RobinBCs.jl:88-131): the Neumann side-set caches plus
    `g0[NF, nqs, nsides]`, `dvalsdu[NF, NF, nqs, nsides]` (vals = g0 + dvalsdu u_q is formed on the device)."""

    def update_bc_values(self, X, t):
        X = np.asarray(X)
        rng = np.random.default_rng(0)
        for func, c in zip(self.bc_funcs, self.bc_caches):
            ns = c["side_nodes"].shape[1]
            nq = len(c["ws"])
            c["g0"] = np.zeros((self.nf, nq, ns), order="F")
            c["dvalsdu"] = np.zeros((self.nf, self.nf, nq, ns), order="F")
            if ns == 0:
                continue
            xs = X[:, c["side_nodes"] - 1]
            Xq = np.einsum("qa,dae->eqd", c["Ns"], xs).reshape(ns * nq, -1)      # (nsides*nqs, ND), q fastest
            n = Xq.shape[0]
            ev = lambda u: _as_values(func(Xq, t, u), n, self.nf)
            g0 = ev(np.zeros((n, self.nf)))
            D = np.zeros((n, self.nf, self.nf))
            for cdof in range(self.nf):
                u = np.zeros((n, self.nf)); u[:, cdof] = 1.0
                D[:, :, cdof] = ev(u) - g0
            ut = rng.standard_normal((n, self.nf))
            if not np.allclose(ev(ut), g0 + np.einsum("ndc,nc->nd", D, ut), rtol=1e-10, atol=1e-12 * (1 + np.abs(g0).max())):
                raise ValueError("RobinBC: the flux law is not affine in u; only g0(X,t) + D(X,t) u is supported on the device path")
            c["g0"] = np.asfortranarray(g0.reshape(ns, nq, self.nf).transpose(2, 1, 0))
            c["dvalsdu"] = np.asfortranarray(D.reshape(ns, nq, self.nf, self.nf).transpose(2, 3, 1, 0))


class Source:
    """Source(var_name, func, block_name)  (src/bcs/Sources.jl:17-27): body force density b(X, t) (all NF components)
    on one element block; the assembler adds -int N b dOmega to the residual (:1-5)."""

    def __init__(self, var_name, func, block_name):
        self.var_name, self.func, self.block_name = var_name, func, block_name


class Sources:
    """Sources(mesh, dof, sources) (src/bcs/Sources.jl:80-257): per entry the block index and `vals[NF, NQ, NE]`."""

    def __init__(self, mesh, dof, sources):
        self.nf = dof.nf
        self.funcs, self.blocks, self.vals = [], [], []
        for s in sources:
            dof.dof_index(s.var_name)
            if s.block_name not in mesh.element_block_names:
                raise KeyError(f"Block {s.block_name} not found in mesh")
            self.funcs.append(s.func)
            self.blocks.append(mesh.element_block_names.index(s.block_name))
            self.vals.append(None)

    def __len__(self):
        return len(self.funcs)

    def update_source_values(self, fspace, t):
        """_update_source_values! (:55-66): vals[q, e] = func(X_q, t) at the cell quadrature points"""
        X = np.asarray(fspace.coords)
        for i, (func, b) in enumerate(zip(self.funcs, self.blocks)):
            rf = fspace.ref_fes[b]
            xe = X[:, fspace.elem_conns.block(b) - 1]                    # (ND, NNPE, NE)
            Xq = np.einsum("qa,dae->eqd", rf.N, xe)                      # (NE, NQ, ND)
            n = Xq.shape[0] * Xq.shape[1]
            v = _as_values(func(Xq.reshape(n, -1), t), n, self.nf)
            self.vals[i] = np.asfortranarray(v.reshape(Xq.shape[0], Xq.shape[1], self.nf).transpose(2, 1, 0))


class PeriodicBC:
    """PeriodicBC(var_name, direction, func, side_a_sset, side_b_sset)  (src/bcs/PeriodicBCs.jl:1-15).  `direction` is the
    coordinate ALONG which the two sides run and by which their nodes are matched (:62-81: key = round(X[dir, node] / tol));
    `func(X, t)` is the jump U[b] = U[a] + func(X_b, t) (:253-262)."""

    def __init__(self, var_name, direction, func, side_a_sset, side_b_sset):
        self.var_name, self.direction, self.func = var_name, direction, func
        self.side_a_sset, self.side_b_sset = side_a_sset, side_b_sset


class PeriodicBCs:
    """PeriodicBCs(mesh, dof, periodic_bcs) (src/bcs/PeriodicBCs.jl:17-110): per BC the side-a dofs / nodes (sorted, unique)
    and the side-b dofs / nodes matched to them.  In 3-D the reference's single-coordinate key cannot tell apart the nodes
    of a face (its `_transverse_key` is commented out, :63,:69); here every coordinate except the one the two sides differ
    in is part of the key, which is the same thing in 2-D."""

    def __init__(self, mesh, dof, pbcs, tolerance=1.0e-6):
        X = np.asarray(mesh.nodal_coords)
        nd, nf = X.shape[0], dof.nf
        self.bc_funcs, self.bc_caches = [], []
        for bc in pbcs:
            if bc.direction not in "xyz"[:nd]:
                raise AssertionError(f"direction {bc.direction} on a {nd}-D mesh")
            d = dof.dof_index(bc.var_name)
            a_nodes = np.unique(np.asarray(mesh.sideset_nodes[bc.side_a_sset], dtype=np.int64))   # _unique_sort_perm
            b_nodes = np.asarray(mesh.sideset_nodes[bc.side_b_sset], dtype=np.int64)
            dir_id = "xyz".index(bc.direction)
            if nd == 2:
                axes = [dir_id]
            else:
                normal = int(np.argmax(np.abs(X[:, a_nodes - 1].mean(axis=1) - X[:, b_nodes - 1].mean(axis=1))))
                axes = [i for i in range(nd) if i != normal]
            key = lambda n: tuple(np.rint(X[axes, n - 1] / tolerance).astype(np.int64))
            to_b = {key(n): n for n in b_nodes}
            if len({key(n) for n in a_nodes}) != len(to_b):
                raise AssertionError("Side a and side b have different numbers of nodes")
            matched = np.array([to_b[key(n)] for n in a_nodes], dtype=np.int64)
            self.bc_caches.append(dict(side_a_nodes=a_nodes, side_b_nodes=matched, side_a_dofs=nf * (a_nodes - 1) + d + 1,
                                       side_b_dofs=nf * (matched - 1) + d + 1, vals=np.zeros(len(a_nodes))))
            self.bc_funcs.append(bc.func)

    def __len__(self):
        return len(self.bc_caches)

    def periodic_dofs(self):
        """(side_a_dofs, side_b_dofs) of all BCs concatenated (:240-250): what update_dofs! consumes"""
        if not self.bc_caches:
            return np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64)
        return (np.concatenate([c["side_a_dofs"] for c in self.bc_caches]),
                np.concatenate([c["side_b_dofs"] for c in self.bc_caches]))

    def update_bc_values(self, X, t):
        """update_bc_values!(bcs, X, t) (:252-257, :128-143): vals[n] = func(X[:, side_b_nodes[n]], t)"""
        X = np.asarray(X)
        for func, c in zip(self.bc_funcs, self.bc_caches):
            v = func(X[:, c["side_b_nodes"] - 1].T, t)
            c["vals"] = np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=float), (len(c["side_b_nodes"]),)))

    def values(self):
        return np.concatenate([c["vals"] for c in self.bc_caches]) if self.bc_caches else np.zeros(0)


class InitialCondition:
    """InitialCondition(var_name, func; block_name | nodeset_name | sideset_name)  (src/InitialConditions.jl:10-46).
    `func(X)` is called vectorised: X is (n, ND), returns (n,) or a scalar."""

    def __init__(self, var_name, func, *, block_name=None, nodeset_name=None, sideset_name=None):
        given = [x is not None for x in (block_name, nodeset_name, sideset_name)]
        if sum(given) == 0:
            raise ValueError("block_name, nodeset_name, or sideset_name required as input arguments in DirichletBC")
        if sum(given) != 1:
            raise ValueError("More than one entity type specificed in DirichletBC")
        self.var_name, self.func = var_name, func
        self.block_name, self.nset_name, self.sset_name = block_name, nodeset_name, sideset_name


class InitialConditions:
    """InitialConditions(mesh, dof, ics) (src/InitialConditions.jl:48-224): per IC the dofs / nodes it sets and their
    values; `update_ic_values(X)` evaluates the functions, `update_field_ics(U)` writes U[dofs] = vals (host field --
    e.g. the starting point handed to the integrator through extract_field_unknowns)."""

    def __init__(self, mesh, dof, ics):
        self.ic_funcs, self.ic_caches = [], []
        nf = dof.nf
        for ic in ics:
            d = dof.dof_index(ic.var_name)
            if ic.block_name is not None:
                n = np.unique(mesh.element_conns[ic.block_name])
            elif ic.nset_name is not None:
                n = np.unique(np.asarray(mesh.nodeset_nodes[ic.nset_name], dtype=np.int64))
            else:
                n = np.unique(np.asarray(mesh.sideset_nodes[ic.sset_name], dtype=np.int64))
            self.ic_caches.append(dict(dofs=nf * (n - 1) + d + 1, locations=n, vals=np.zeros(len(n))))
            self.ic_funcs.append(ic.func)

    def __len__(self):
        return len(self.ic_caches)

    def update_ic_values(self, X):
        X = np.asarray(X)
        for func, c in zip(self.ic_funcs, self.ic_caches):
            v = func(X[:, c["locations"] - 1].T)
            c["vals"] = np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=float), (len(c["locations"]),)))

    def update_field_ics(self, U):
        f = U.data_flat if hasattr(U, "data_flat") else np.asarray(U).reshape(-1)
        for c in self.ic_caches:
            f[c["dofs"] - 1] = c["vals"]
