"""SparseMatrixAssembler and the assemble_* entry points (src/assemblers/*.jl) on libfecb200.

Python cannot spell `assemble_vector!`; the trailing `!` is dropped.  Everything else keeps the
reference's names, argument order and error behaviour:

    asm = SparseMatrixAssembler(u; sparse_matrix_type=:csr, use_condensed=false, matrix_free=false)
    p   = create_parameters(mesh, asm, physics, props; dirichlet_bcs=dbcs)
    assemble_vector!(asm, residual, Uu, p);            R  = residual(asm)
    assemble_stiffness!(asm, stiffness, Uu, p);        K  = stiffness(asm)
    assemble_matrix_action!(asm, stiffness, Uu, Vu, p) Kv = hvp(asm, Vu)
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import scipy.sparse as sp

from . import _lib
from ._lib import FECError, check, lib
from .bcs import DirichletBCs, NeumannBCs, PeriodicBCs, RobinBCs, Sources, TimeStepper
from .fields import H1Field
from .function_spaces import AbstractFunction, DofManager
from .physics import AbstractPhysics, Poisson, kind_of


class SparseMatrixAssembler:
    """SparseMatrixAssembler(dof_or_var; sparse_matrix_type, use_condensed, use_inplace_methods,
    use_sparse_vector, matrix_free)  (src/assemblers/SparseMatrixAssembler.jl:64-133)."""

    def __init__(self, dof_or_var, *, sparse_matrix_type="csc", use_condensed=False, use_inplace_methods=False,
                 use_sparse_vector=False, matrix_free=False, device=0):
        if isinstance(dof_or_var, AbstractFunction):
            dof_or_var = DofManager(dof_or_var, use_condensed=use_condensed)
        self.dof = dof_or_var
        if sparse_matrix_type not in ("csc", "csr"):
            raise ValueError(f"Unsupported sparse matrix type {sparse_matrix_type}. Only :csc, and :csr are supported.")
        if use_sparse_vector:
            raise NotImplementedError("use_sparse_vector=true is not on the B200 path (SURVEY B6)")
        self.sparse_matrix_type = sparse_matrix_type
        self.use_inplace_methods = bool(use_inplace_methods)
        self.matrix_free = bool(matrix_free)
        self.device = device
        self._h = None
        self._pattern = None
        self._keep = []

    # -- handle lifecycle ----------------------------------------------------------------------
    def _require(self):
        if self._h is None:
            raise FECError("assembler has no device handle yet: call create_parameters(mesh, asm, physics, props; ...)")
        return self._h

    def _create_handle(self, fspace, physics_list, props_list):
        if self._h is not None:
            check(lib.fecb200_destroy(self._h))
            self._h = None
        nb = fspace.num_blocks()
        blocks = (_lib.BlockDesc * nb)()
        keep = []
        for b in range(nb):
            rf = fspace.ref_fes[b]
            ph = physics_list[b]
            conn, cp = _lib.i64(fspace.elem_conns.block(b).reshape(-1, order="F"))
            N, Np = _lib.f64(rf.N)
            dN, dNp = _lib.f64(rf.dN)
            w, wp = _lib.f64(rf.w)
            pr, prp = _lib.f64(props_list[b])
            keep += [conn, N, dN, w, pr]
            d = blocks[b]
            d.elem_type, d.nnpe, d.nelem, d.conn = rf.elem_id, rf.num_cell_dofs, fspace.elem_conns.nelems[b], cp
            d.nq, d.N, d.dN, d.w = rf.num_quadrature_points, Np, dNp, wp
            d.physics_id, d.nprops, d.props, d.nstate = ph.physics_id, len(pr), prp, ph.NS
        X, Xp = _lib.f64(fspace.coords.data_flat)
        mesh = _lib.MeshDesc(fspace.num_nodes(), fspace.num_dimensions(), self.dof.nf, nb, blocks, Xp)
        opts = _lib.Opts(_lib.CSR if self.sparse_matrix_type == "csr" else _lib.CSC, int(self.dof.condensed),
                         int(self.matrix_free), self.device, 0)
        h = _lib.Handle()
        check(lib.fecb200_create(C.byref(mesh), C.byref(opts), C.byref(h)))
        self._h = h
        # block_quadrature_sizes(fspace) (scalar_quadrature_storage, SparseMatrixAssembler.jl:104-108)
        self.block_quadrature_sizes = [(fspace.ref_fes[b].num_quadrature_points, int(fspace.elem_conns.nelems[b])) for b in range(nb)]
        self.mesh_block_names = list(getattr(fspace, "block_names", None) or [f"block_{b + 1}" for b in range(nb)])
        self._pattern = None

    def close(self):
        if self._h is not None:
            lib.fecb200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_matrix_double_buffer(self, enable=True):
        """Opt-in: keep two stiffness value arrays so the zero-fill of assemble_stiffness! (Matrix.jl:39) rides
        inside the element kernel (fecb200_set_matrix_double_buffer)."""
        check(lib.fecb200_set_matrix_double_buffer(self._require(), int(bool(enable))))

    # -- sizes / maps ----------------------------------------------------------------------------
    def sizes(self):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        check(lib.fecb200_sizes(self._require(), C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def _refresh_dof_maps(self):
        ntot, nunk, _ = self.sizes()
        ud = np.empty(nunk, dtype=np.int64)
        d2u = np.empty(ntot, dtype=np.int64)
        check(lib.fecb200_dof_maps_copy(self._h, ud.ctypes.data_as(_lib.c_i64p), d2u.ctypes.data_as(_lib.c_i64p)))
        self.dof.unknown_dofs, self.dof.dof_to_unknown = ud, d2u

    def pattern(self):
        """(n, ptr, idx): rowptr/colval (csr) or colptr/rowval (csc), Int64 1-based."""
        if self._pattern is None:
            n, nnz = C.c_int64(), C.c_int64()
            check(lib.fecb200_pattern_sizes(self._require(), C.byref(n), C.byref(nnz)))
            ptr = np.empty(n.value + 1, dtype=np.int64)
            idx = np.empty(nnz.value, dtype=np.int64)
            check(lib.fecb200_pattern_copy(self._h, ptr.ctypes.data_as(_lib.c_i64p), idx.ctypes.data_as(_lib.c_i64p)))
            self._pattern = (n.value, ptr, idx)
        return self._pattern

    def launch_count(self):
        n = C.c_int64()
        check(lib.fecb200_launch_count(self._require(), C.byref(n)))
        return n.value

    def kernel_form(self, block=0):
        """1 when the block's element kernels take the Walsh-Hadamard form (HEX8, trilinear table on a 2-point rule per
        axis, any numbering), 0 for the plain quadrature loop."""
        f = C.c_int32()
        check(lib.fecb200_block_kernel_form(self._require(), block, C.byref(f)))
        return f.value


class Parameters:
    """The slice of Parameters (src/Parameters.jl:37-73) the hot path touches; device-resident
    fields live inside the handle (`p |> cuda`)."""

    def __init__(self, mesh, asm, physics, props, dirichlet_bcs, times, neumann_bcs=None, sources=None, periodic_bcs=None,
                 robin_bcs=None):
        self.mesh, self.asm = mesh, asm
        self.physics, self.properties = physics, props
        self.dirichlet_bcs = dirichlet_bcs
        self.neumann_bcs, self.sources = neumann_bcs, sources
        self.robin_bcs = robin_bcs
        self.periodic_bcs = periodic_bcs
        self.times = times
        self.coords = mesh.nodal_coords

    @property
    def field(self):
        """p.field (full-length H1Field), copied back from the device"""
        out = np.empty(len(self.asm.dof))
        check(lib.fecb200_field_copy(self.asm._require(), _lib.FIELD_U, _lib.ptr(out)))
        return H1Field(out.reshape(self.asm.dof.nf, -1, order="F"))

    def state(self, block=0, which="new"):
        b = self.asm.dof.var.fspace
        ph = self.physics[block]
        nq, ne = b.ref_fes[block].num_quadrature_points, b.elem_conns.nelems[block]
        out = np.zeros(ph.NS * nq * ne)
        check(lib.fecb200_state_get(self.asm._require(), block, 1 if which == "new" else 0, _lib.ptr(out)))
        return out.reshape(ph.NS, nq, ne, order="F")

    def set_state(self, state, block=0, which="old"):
        s = np.ascontiguousarray(np.asarray(state, dtype=float).reshape(-1, order="F"))
        check(lib.fecb200_state_set(self.asm._require(), block, 1 if which == "new" else 0, _lib.ptr(s)))


def _per_block(x, nb, kind):
    if isinstance(x, dict):
        return list(x.values())
    if isinstance(x, (list, tuple)) and len(x) == nb and (kind is None or all(isinstance(v, kind) for v in x)):
        return list(x)
    return [x] * nb


def create_parameters(mesh, asm, physics, props=None, *, dirichlet_bcs=(), neumann_bcs=(), sources=(), periodic_bcs=(),
                      robin_bcs=(), times=None):
    """create_parameters(mesh, asm, physics, props; dirichlet_bcs, neumann_bcs, sources, times)  (src/Parameters.jl:288-302):
    builds the device handle (the `|> cuda` step), the BC containers, and calls update_dofs!."""
    fspace = asm.dof.var.fspace
    nb = fspace.num_blocks()
    physics_list = _per_block(physics, nb, AbstractPhysics)
    if props is None:
        props_list = [ph.create_properties() for ph in physics_list]
    elif isinstance(props, dict):
        props_list = list(props.values())
    elif (nb > 1 and isinstance(props, (list, tuple)) and len(props) == nb
          and all(isinstance(v, (np.ndarray, list, tuple)) for v in props)):
        props_list = list(props)           # one property vector per block
    else:
        props_list = [props] * nb          # one SVector{NP} shared by every block
    props_list = [np.atleast_1d(np.asarray(p, dtype=float)).reshape(-1) for p in props_list]
    for ph in physics_list:
        if ph.NF != asm.dof.nf:
            raise FECError(f"physics has NF={ph.NF} but the function has {asm.dof.nf} fields")
    asm._create_handle(fspace, physics_list, props_list)
    times = times if times is not None else TimeStepper(0.0, 0.0, 1)
    dbcs = DirichletBCs(mesh, asm.dof, list(dirichlet_bcs))
    nbcs = NeumannBCs(mesh, asm.dof, list(neumann_bcs))
    srcs = Sources(mesh, asm.dof, list(sources))
    pbcs = PeriodicBCs(mesh, asm.dof, list(periodic_bcs))
    rbcs = RobinBCs(mesh, asm.dof, list(robin_bcs))
    p = Parameters(mesh, asm, physics_list, props_list, dbcs, times, nbcs, srcs, pbcs, rbcs)
    update_dofs(asm, dbcs, periodic=pbcs.periodic_dofs())      # update_dofs!(asm, dbcs, pbcs) (Parameters.jl:118-125)
    for i, c in enumerate(nbcs.bc_caches):   # the side sets' geometry is uploaded once (`p |> cuda`)
        sn, snp = _lib.i64(c["side_nodes"].reshape(-1, order="F"))
        Ns, Nsp = _lib.f64(c["Ns"]); dNs, dNsp = _lib.f64(c["dNs"]); ws, wsp = _lib.f64(c["ws"])
        check(lib.fecb200_set_neumann_bc(asm._require(), i, c["side_nodes"].shape[1], c["side_nodes"].shape[0], len(ws),
                                         snp, Nsp, dNsp, wsp))
    for i, c in enumerate(rbcs.bc_caches):
        sn, snp = _lib.i64(c["side_nodes"].reshape(-1, order="F"))
        Ns, Nsp = _lib.f64(c["Ns"]); dNs, dNsp = _lib.f64(c["dNs"]); ws, wsp = _lib.f64(c["ws"])
        check(lib.fecb200_set_robin_bc(asm._require(), i, c["side_nodes"].shape[1], c["side_nodes"].shape[0], len(ws),
                                       snp, Nsp, dNsp, wsp))
    update_bc_values(p)
    _upload_sources(p)
    return p


def _upload_sources(p):
    """evaluate Poisson.func at the quadrature points X_q = sum_a N[q,a] x_a and upload f_q[NQ,NE]"""
    fspace = p.asm.dof.var.fspace
    X = np.asarray(fspace.coords)
    for b, ph in enumerate(p.physics):
        if isinstance(ph, Poisson) and ph.func is not None:
            rf = fspace.ref_fes[b]
            conn = fspace.elem_conns.block(b)                      # (NNPE, NE) 1-based
            xe = X[:, conn - 1]                                    # (ND, NNPE, NE)
            Xq = np.einsum("qa,dae->eqd", rf.N, xe)                # (NE, NQ, ND)
            f = np.asarray(ph.func(Xq.reshape(-1, Xq.shape[2]), p.times.time_current), dtype=float)
            fq = np.ascontiguousarray(np.broadcast_to(f, (Xq.shape[0] * Xq.shape[1],)))
            check(lib.fecb200_set_source_q(p.asm._require(), b, _lib.ptr(fq)))


def update_dofs(asm, dbcs, periodic=((), ())):
    """update_dofs!(asm, dbcs, pbcs)  (SparseMatrixAssembler.jl:228-274)"""
    dd = dbcs.dirichlet_dofs() if isinstance(dbcs, DirichletBCs) else np.asarray(dbcs, dtype=np.int64)
    dd, ddp = _lib.i64(dd)
    pa, pap = _lib.i64(periodic[0])
    pb, pbp = _lib.i64(periodic[1])
    check(lib.fecb200_update_dofs(asm._require(), ddp, len(dd), pap, pbp, len(pa)))
    asm.dof.dirichlet_dofs = dd
    asm.dof.periodic_side_a_dofs, asm.dof.periodic_side_b_dofs = pa, pb
    asm._refresh_dof_maps()
    asm._pattern = None


def update_bc_values(p):
    """update_bc_values!(p, asm) (Parameters.jl:358): evaluate BC closures, push dofs/vals."""
    bcs = p.dirichlet_bcs
    bcs.update_bc_values(p.coords, p.times.time_current)
    d, dp = _lib.i64(bcs.dofs)
    v, vp = _lib.f64(bcs.vals)
    check(lib.fecb200_set_dirichlet_values(p.asm._require(), dp, vp, len(d)))
    if getattr(p, "periodic_bcs", None) is not None and len(p.periodic_bcs):    # jump values U[b] = U[a] + val
        p.periodic_bcs.update_bc_values(p.coords, p.times.time_current)
        # the library keeps one value per (de-duplicated, chain-resolved) pair in registration order and checks the count.
        # Pushed at EVERY update, like the reference rewrites its cache at every update_bc_values!: a time-dependent
        # jump that returns to exactly 0 must not leave a stale value on the device (NULL = all zero).
        pv = p.periodic_bcs.values()
        if np.any(pv != 0.0):
            v, vp = _lib.f64(pv)
            check(lib.fecb200_set_periodic_values(p.asm._require(), vp, len(v)))
        else:
            check(lib.fecb200_set_periodic_values(p.asm._require(), None, 0))
    if p.neumann_bcs is not None and len(p.neumann_bcs):     # update_bc_values!(p.neumann_bcs, asm, X, t)
        p.neumann_bcs.update_bc_values(p.coords, p.times.time_current)
        for i, c in enumerate(p.neumann_bcs.bc_caches):
            v = np.ascontiguousarray(c["vals"].reshape(-1, order="F"))
            check(lib.fecb200_set_neumann_values(p.asm._require(), i, _lib.ptr(v)))
    if getattr(p, "robin_bcs", None) is not None and len(p.robin_bcs):   # update_bc_values!(p.robin_bcs, ...) (Solvers.jl:74)
        p.robin_bcs.update_bc_values(p.coords, p.times.time_current)
        for i, c in enumerate(p.robin_bcs.bc_caches):
            g0 = np.ascontiguousarray(c["g0"].reshape(-1, order="F"))
            D = np.ascontiguousarray(c["dvalsdu"].reshape(-1, order="F"))
            check(lib.fecb200_set_robin_values(p.asm._require(), i, _lib.ptr(g0), _lib.ptr(D)))
    if p.sources is not None and len(p.sources):             # _update_source_values! (Sources.jl:55-66)
        p.sources.update_source_values(p.asm.dof.var.fspace, p.times.time_current)
        for b, vals in zip(p.sources.blocks, p.sources.vals):
            v = np.ascontiguousarray(vals.reshape(-1, order="F"))
            check(lib.fecb200_set_source_values(p.asm._require(), b, _lib.ptr(v)))


def update_time(p):
    """update_time!(p) (Parameters.jl:447)"""
    p.times.time_current += p.times.dt
    check(lib.fecb200_set_time(p.asm._require(), p.times.time_current, p.times.dt))


def create_field(asm_or_dof):
    """create_field(dof) / create_field(asm) (src/DofManagers.jl:148-158)"""
    dof = getattr(asm_or_dof, "dof", asm_or_dof)
    return H1Field.zeros(dof.nf, dof.nn)


def create_unknowns(asm):
    """create_unknowns(asm) (DofManagers.jl:168-175)"""
    return np.zeros(asm.sizes()[2])


# ---- assembly entry points --------------------------------------------------------------------

def assemble_vector(asm, func, Uu, p):
    """assemble_vector!(asm, residual, Uu, p)  (src/assemblers/Vector.jl:4-74)"""
    kind = kind_of(func, (_lib.RESIDUAL,))
    check(lib.fecb200_assemble_vector(asm._require(), kind, _lib.ptr(Uu)))


def assemble_vector_neumann_bc(asm, Uu, p):
    """assemble_vector_neumann_bc!(asm, Uu, p)  (src/assemblers/WeaklyEnforcedBCs.jl:4-15): adds the Neumann loads of
    p.neumann_bcs to the residual storage (no zeroing).  The loads do not depend on Uu; the library integrates them
    once per update_bc_values! and adds the cached vector."""
    check(lib.fecb200_assemble_vector_neumann_bc(asm._require()))


def assemble_vector_robin_bc(asm, Uu, p):
    """assemble_vector_robin_bc!(asm, Uu, p) (src/assemblers/WeaklyEnforcedBCs.jl:20-32): adds the Robin flux at the
    CURRENT p.field to the residual storage (no _update_for_assembly!, like the reference)"""
    check(lib.fecb200_assemble_vector_robin_bc(asm._require()))


def assemble_matrix_robin_bc(asm, Uu, p):
    """assemble_matrix_robin_bc!(asm, Uu, p) (:88-116): adds the Robin tangent to the assembled stiffness values"""
    check(lib.fecb200_assemble_matrix_robin_bc(asm._require()))


def assemble_vector_source(asm, Uu, p):
    """assemble_vector_source!(asm, Uu, p)  (src/assemblers/Source.jl:10-42): adds -int N b of p.sources to the
    residual storage (no zeroing)."""
    check(lib.fecb200_assemble_vector_source(asm._require()))


def assemble_lumped_mass(asm, func, Uu, p):
    """assemble_lumped_mass!(asm, lumped_mass, Uu, p)  (src/assemblers/LumpedMass.jl:32-60): row-sum mass
    rho * JxW * N[a] per dof into the residual storage; read it with lumped_mass(asm)."""
    kind_of(func, (_lib.LUMPED_MASS,))
    check(lib.fecb200_assemble_vector(asm._require(), _lib.LUMPED_MASS, _lib.ptr(Uu)))


def assemble_diagonal(asm, func, Uu, p):
    """assemble_diagonal!(asm, stiffness | mass, Uu, p)  (src/assemblers/Diagonal.jl:16-74): the diagonal of the
    element matrices summed into the residual storage, without assembling the sparse matrix; read it with
    diagonal(asm)."""
    kind = kind_of(func, (_lib.STIFFNESS, _lib.MASS))
    check(lib.fecb200_assemble_vector(asm._require(), _lib.DIAGONAL_STIFFNESS if kind == _lib.STIFFNESS else _lib.DIAGONAL_MASS,
                                      _lib.ptr(Uu)))


def assemble_scalar(asm, func, Uu, p):
    """assemble_scalar!(asm, energy, Uu, p)  (src/assemblers/QuadratureQuantity.jl:4-14): JxW * energy at every
    quadrature point of every block; read it with scalar_values(asm)."""
    kind_of(func, (_lib.ENERGY,))
    check(lib.fecb200_assemble_scalar(asm._require(), _lib.ptr(Uu)))


def scalar_values(asm, block=None):
    """asm.scalar_quadrature_storage: {block name: array (NQ, NE)} (block_view(storage, b)[1, :, :]) or one block."""
    names = list(asm.mesh_block_names)
    out = {}
    for bi, name in enumerate(names):
        if block is not None and block not in (bi, name):
            continue
        nq, ne = asm.block_quadrature_sizes[bi]
        a = np.empty((ne, nq))
        check(lib.fecb200_scalar_values(asm._require(), bi, _lib.ptr(a)))
        out[name] = a.T          # (NQ, NE) view of the column-major [NQ, NE] buffer
    return out if block is None else next(iter(out.values()))


def _check_matrix_assembly_supported(asm, fname):
    if asm.matrix_free:
        raise FECError(f"{fname} called on a matrix-free SparseMatrixAssembler.  Re-create the assembler with "
                       "matrix_free=false to enable matrix assembly.")


def assemble_stiffness(asm, func, Uu, p):
    """assemble_stiffness!(asm, stiffness, Uu, p)  (src/assemblers/Matrix.jl:12-21)"""
    _check_matrix_assembly_supported(asm, "assemble_stiffness!")
    kind_of(func, (_lib.STIFFNESS,))
    check(lib.fecb200_assemble_matrix(asm._require(), _lib.STIFFNESS, _lib.ptr(Uu)))


def assemble_vector_and_stiffness(asm, func_r, func_k, Uu, p):
    """assemble_vector!(asm, residual, Uu, p); assemble_stiffness!(asm, stiffness, Uu, p) in one pass -- the pair
    of calls solve!(::IterativeLinearSolver) makes at every Newton iteration (src/Solvers.jl:133-140)."""
    _check_matrix_assembly_supported(asm, "assemble_stiffness!")
    kind_of(func_r, (_lib.RESIDUAL,))
    kind_of(func_k, (_lib.STIFFNESS,))
    check(lib.fecb200_assemble_vector_and_matrix(asm._require(), _lib.ptr(Uu)))


def assemble_mass(asm, func, Uu, p):
    """assemble_mass!(asm, mass, Uu, p)  (src/assemblers/Matrix.jl:1-10)"""
    _check_matrix_assembly_supported(asm, "assemble_mass!")
    kind_of(func, (_lib.MASS,))
    check(lib.fecb200_assemble_matrix(asm._require(), _lib.MASS, _lib.ptr(Uu)))


def assemble_matrix_action(asm, func, Uu, Vu, p):
    """assemble_matrix_action!(asm, stiffness | mass, Uu, Vu, p)  (MatrixAction.jl:154-238).
    Evaluated matrix-free on the device: identical to K_el * v_el up to rounding."""
    kind = kind_of(func, (_lib.STIFFNESS, _lib.MASS))
    check(lib.fecb200_assemble_action(asm._require(), kind, _lib.ptr(Uu), _lib.ptr(Vu)))


def assemble_matrix_free_action(asm, func_action, Uu, Vu, p):
    """assemble_matrix_free_action!(asm, stiffness_action, Uu, Vu, p)  (MatrixAction.jl:9-77)"""
    kind = kind_of(func_action, (_lib.STIFFNESS, _lib.MASS))
    check(lib.fecb200_assemble_action(asm._require(), kind, _lib.ptr(Uu), _lib.ptr(Vu)))


def assemble_matrix_free_action_full(asm, func_action, U_full, v_full, p):
    """assemble_matrix_free_action_full!(asm, f, U_full, v_full, p)  (MatrixAction.jl:99-149)"""
    kind = kind_of(func_action, (_lib.STIFFNESS, _lib.MASS))
    n = len(asm.dof)
    if np.size(U_full) != n or np.size(v_full) != n:
        raise AssertionError("U_full and v_full must have the full DOF length")
    check(lib.fecb200_assemble_action_full(asm._require(), kind, _lib.ptr(U_full), _lib.ptr(v_full)))


# ---- accessors ---------------------------------------------------------------------------------

def _residual_accessor(asm, out=None):
    """residual(asm)  (src/assemblers/Assemblers.jl:347-371)"""
    out = np.empty(asm.sizes()[2]) if out is None else out
    check(lib.fecb200_residual(asm._require(), _lib.ptr(out)))
    return out


def _vector_values_accessor(asm, out=None):
    """lumped_mass(asm) / diagonal(asm)  (LumpedMass.jl:70-80, Diagonal.jl:76-89): shares the residual storage."""
    out = np.empty(asm.sizes()[2]) if out is None else out
    check(lib.fecb200_vector_values(asm._require(), _lib.ptr(out)))
    return out


diagonal = _vector_values_accessor


def hvp(asm, v, out=None):
    """hvp(asm, v)  (src/assemblers/Assemblers.jl:310-324)"""
    out = np.empty(asm.sizes()[2]) if out is None else out
    check(lib.fecb200_hvp(asm._require(), _lib.ptr(v), _lib.ptr(out)))
    return out


def _sparse(asm, kind):
    n, ptr, idx = asm.pattern()
    nz = np.empty(len(idx))
    check(lib.fecb200_matrix_values(asm._require(), kind, _lib.ptr(nz)))
    cls = sp.csr_matrix if asm.sparse_matrix_type == "csr" else sp.csc_matrix
    ncols = asm.sizes()[2]  # == n except for rank-local (partitioned) assemblers: owned rows x local columns
    return cls((nz, idx - 1, ptr - 1), shape=(n, ncols) if asm.sparse_matrix_type == "csr" else (ncols, n))


def _stiffness_accessor(asm):
    """stiffness(asm)  (Assemblers.jl:376-388): scipy csr/csc matrix with the reference's pattern"""
    if asm.matrix_free:
        n = asm.sizes()[2]
        return sp.csc_matrix((n, n))
    return _sparse(asm, _lib.STIFFNESS)


def _mass_accessor(asm):
    """mass(asm)  (Assemblers.jl:329-341)"""
    if asm.matrix_free:
        n = asm.sizes()[2]
        return sp.csc_matrix((n, n))
    return _sparse(asm, _lib.MASS)


def matrix_multiply(asm, x, out=None, kind=_lib.STIFFNESS):
    """stiffness(asm) * x (or mass(asm) * x with kind=MASS) on the device-resident values: the product the Krylov
    solve forms (src/Solvers.jl:144), without copying the matrix to the host."""
    out = np.empty(asm.sizes()[2]) if out is None else out
    check(lib.fecb200_matrix_multiply(asm._require(), kind, _lib.ptr(x), _lib.ptr(out)))
    return out


def update_field(p, Uu):
    """_update_for_assembly!(p, dof, Uu) (src/Parameters.jl:404-413): p.field <- BC values, unknowns, periodic copies"""
    check(lib.fecb200_update_field(p.asm._require(), _lib.ptr(Uu)))


def full_field(asm, which="residual"):
    """full-length storage of the assembler (residual_storage / stiffness_action_storage)"""
    out = np.empty(len(asm.dof))
    sel = {"u": _lib.FIELD_U, "residual": _lib.FIELD_RESIDUAL, "action": _lib.FIELD_ACTION, "v": _lib.FIELD_V}[which]
    check(lib.fecb200_field_copy(asm._require(), sel, _lib.ptr(out)))
    return out
