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
