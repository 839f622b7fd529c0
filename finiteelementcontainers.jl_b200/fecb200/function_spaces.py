"""FunctionSpace / functions / DofManager (src/FunctionSpaces.jl, src/Functions.jl, src/DofManagers.jl)."""
from __future__ import annotations

import numpy as np

from .fields import Connectivity, H1Field
from .reference_fe import ReferenceFE


class Lagrange:  # interpolation tag (FunctionSpace(mesh, H1Field, Lagrange))
    pass


class FunctionSpace:
    """FunctionSpace(mesh, H1Field, Lagrange; q_type, q_degree)  (src/FunctionSpaces.jl:158-237).
    Holds coords, the flat Connectivity and one ReferenceFE per block.  The reference's default
    quadrature is `GaussLobattoLegendre` with q_degree 2 (:187); what that means lives in
    ReferenceFiniteElements.jl (not vendored), so tables can also be injected with `ref_fes=`."""

    def __init__(self, mesh, field_type=H1Field, interp=Lagrange, *, q_type="GaussLegendre", q_degree=2, ref_fes=None):
        assert field_type is H1Field, "only H1 spaces are on the hot path"
        self.mesh = mesh
        self.coords = mesh.nodal_coords
        self.block_names = list(mesh.element_block_names)
        self.elem_conns = Connectivity([mesh.element_conns[b] for b in self.block_names])
        if ref_fes is None:
            ref_fes = [ReferenceFE(mesh.element_types[b], q_type, q_degree) for b in self.block_names]
        self.ref_fes = list(ref_fes)

    def num_blocks(self):
        return len(self.block_names)

    def num_nodes(self):
        return self.coords.shape[1]

    def num_dimensions(self):
        return self.coords.shape[0]


class AbstractFunction:
    def names(self):
        return self._names

    def num_fields(self):
        return len(self._names)


class ScalarFunction(AbstractFunction):
    """ScalarFunction(V, name) (src/Functions.jl:36-51)"""

    def __init__(self, fspace, name):
        self.fspace = fspace
        self._names = [str(name)]


class VectorFunction(AbstractFunction):
    """VectorFunction(V, name) (src/Functions.jl:62-87): components name_x, name_y[, name_z]"""

    def __init__(self, fspace, name):
        self.fspace = fspace
        self._names = [f"{name}_{c}" for c in "xyz"[: fspace.num_dimensions()]]


class DofManager:
    """DofManager(var; use_condensed) (src/DofManagers.jl:21-73).  Host copy of the DOF maps
    (1-based Int64 like the reference); after `update_dofs!` they are read back from the library,
    which is the single source of truth for the numbering."""

    def __init__(self, var, use_condensed=False):
        self.var = var
        self.condensed = bool(use_condensed)
        nf, nn = var.num_fields(), var.fspace.num_nodes()
        self.nf, self.nn = nf, nn
        self.dirichlet_dofs = np.zeros(0, dtype=np.int64)
        self.unknown_dofs = np.arange(1, nf * nn + 1, dtype=np.int64)
        self.dof_to_unknown = np.arange(1, nf * nn + 1, dtype=np.int64)
        self.periodic_side_a_dofs = np.zeros(0, dtype=np.int64)
        self.periodic_side_b_dofs = np.zeros(0, dtype=np.int64)

    def __len__(self):
        return self.nf * self.nn

    def size(self):
        return (self.nf, self.nn)

    def dof_index(self, var_name):
        return self.var.names().index(var_name)  # 0-based component
