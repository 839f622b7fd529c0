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

    def __len__(self):
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


class SymmetricTensorFunction(AbstractFunction):
    """SymmetricTensorFunction(V, name; use_spatial_dimension) (src/Functions.jl:120-150): Voigt order xx, yy, zz, yz, xz, xy
    (always 3-D unless use_spatial_dimension: xx, yy, xy in 2-D)"""

    def __init__(self, fspace, name, use_spatial_dimension=False):
        self.fspace = fspace
        nd = fspace.num_dimensions() if use_spatial_dimension else 3
        comps = ["xx", "yy", "zz", "yz", "xz", "xy"] if nd == 3 else ["xx", "yy", "xy"]
        self._names = [f"{name}_{c}" for c in comps]


class TensorFunction(AbstractFunction):
    """TensorFunction(V, name; use_spatial_dimension) (src/Functions.jl:89-118): xx, yy, zz, yz, xz, xy, zy, zx, yx
    (xx, yy, xy, yx in 2-D with use_spatial_dimension)"""

    def __init__(self, fspace, name, use_spatial_dimension=False):
        self.fspace = fspace
        nd = fspace.num_dimensions() if use_spatial_dimension else 3
        comps = ["xx", "yy", "zz", "yz", "xz", "xy", "zy", "zx", "yx"] if nd == 3 else ["xx", "yy", "xy", "yx"]
        self._names = [f"{name}_{c}" for c in comps]


class GeneralFunction(AbstractFunction):
    """GeneralFunction(u, v, ...) (src/Functions.jl:184-207): the fields of several functions on one space side by side"""

    def __init__(self, *funcs):
        assert funcs and all(f.fspace is funcs[0].fspace for f in funcs), "functions must share one FunctionSpace"
        self.fspace = funcs[0].fspace
        self._names = [n for f in funcs for n in f.names()]


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


# ---- host-side field helpers of the DofManager (src/DofManagers.jl:203-213, 349-411, src/bcs/DirichletBCs.jl:411-418) --
# Inside every assemble_* call the library does these three steps on the device (k_update_field); the functions below
# are the same index copies for a caller's own HOST fields (post-processing, initial guesses), like the reference's CPU
# methods.  `U` is an H1Field (or its flat data), indices are the DofManager's 1-based dof ids.

def _flat(U):
    return U.data_flat if hasattr(U, "data_flat") else np.asarray(U).reshape(-1)


def update_field_unknowns(U, dof, Uu):
    """update_field_unknowns!(U, dof, Uu): U[unknown_dofs] = Uu (non-condensed) or Uu[unknown_dofs] (condensed)"""
    f, ud = _flat(U), np.asarray(dof.unknown_dofs) - 1
    Uu = np.asarray(Uu)
    if dof.condensed:
        assert Uu.shape[0] == f.shape[0]
        f[ud] = Uu[ud]
    else:
        assert Uu.shape[0] == len(ud)
        f[ud] = Uu


def extract_field_unknowns(Uu, dof, U):
    """extract_field_unknowns!(Uu, dof, U): Uu[n] = U[unknown_dofs[n]]"""
    Uu[:len(dof.unknown_dofs)] = _flat(U)[np.asarray(dof.unknown_dofs) - 1]


def update_field_dirichlet_bcs(U, bcs):
    """update_field_dirichlet_bcs!(U, bcs): U[dofs[i]] = vals[i], in order (a later BC wins on a shared dof)"""
    f = _flat(U)
    for d, v in zip(np.asarray(bcs.dofs) - 1, np.asarray(bcs.vals)):
        f[d] = v
