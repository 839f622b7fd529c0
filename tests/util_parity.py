"""Shared helpers for the parity tests: build the SAME problem in the product (fecb200, CUDA)
and in the oracle (oracle/fec_oracle.py, numpy)."""
import os

import numpy as np

import fec_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL = 1e-12  # north_star: values within 1e-12 relative (FP64, atomic-order reassociation)


def rel_err(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    scale = max(np.abs(b).max(), 1e-300) if b.size else 1.0
    return (np.abs(a - b).max() / scale) if b.size else 0.0


_ORACLE_PHYS = {
    "poisson": lambda f, nd: O.Poisson(f if f is not None else (lambda X: np.zeros(X.shape[0]))),
    "linear": lambda f, nd: O.LinearElastic(nd),
    "neo": lambda f, nd: O.NeoHookean(nd, "standard"),
    "neo_as_written": lambda f, nd: O.NeoHookean(nd, "as_written"),
    "j2": lambda f, nd: O.J2Plasticity(nd),
    "nonsym": lambda f, nd: O.NonSymmetricTest(nd),
}
_RULES = {"QUAD4": "gauss2", "HEX8": "gauss2", "TRI3": "tri3", "TETRA4": "tet4", "TETRA10": "tet4"}


def product_physics(F, name, nd, func=None):
    form = F.ThreeDimensional() if nd == 3 else F.PlaneStrain()
    return {
        "poisson": lambda: F.Poisson((lambda X, t: func(X)) if func is not None else None),
        "linear": lambda: F.Mechanics(form),
        "neo": lambda: F.NeoHookean(form),
        "neo_as_written": lambda: F.NeoHookean(form, variant="as_written"),
        "j2": lambda: F.J2Plasticity(form),
        "nonsym": lambda: F.NonSymmetricTestPhysics(form),
    }[name]()


def build_pair(F, mesh, phys_name, props, *, condensed, matrix_type, bc_nodes_1based, bc_components=None,
               func=None, matrix_free=False, bc_value=0.0):
    """Returns (asm, p, oasm): product assembler + parameters, oracle assembler.
    `mesh` is a fecb200 mesh; the oracle gets the very same arrays."""
    nd = mesh.num_dimensions()
    nf = 1 if phys_name == "poisson" else nd
    V = F.FunctionSpace(mesh, F.H1Field, F.Lagrange, q_type="GaussLegendre", q_degree=2)
    u = F.ScalarFunction(V, "u") if nf == 1 else F.VectorFunction(V, "displ")
    asm = F.SparseMatrixAssembler(u, sparse_matrix_type=matrix_type, use_condensed=condensed, matrix_free=matrix_free)
    comps = list(range(nf)) if bc_components is None else bc_components
    names = u.names()
    mesh.nodeset_nodes["__parity_bc__"] = np.asarray(bc_nodes_1based, dtype=np.int64)
    dbcs = [F.DirichletBC(names[c], (lambda X, t, v=bc_value: np.full(X.shape[0], v)), nodeset_name="__parity_bc__")
            for c in comps]
    ph = product_physics(F, phys_name, nd, func)
    p = F.create_parameters(mesh, asm, ph, props, dirichlet_bcs=dbcs)
    # ---- oracle twin
    blocks = []
    for b in mesh.element_block_names:
        t = mesh.element_types[b]
        blocks.append(O.Block(mesh.element_conns[b], O.ref_fe_tables(t, _RULES[t]),
                              _ORACLE_PHYS[phys_name](func, nd), props=props if props is not None else ()))
    oasm = O.OracleAssembler(np.asarray(mesh.nodal_coords), blocks, nf, condensed=condensed, matrix_type=matrix_type)
    dd = np.unique(np.concatenate([nf * (np.asarray(bc_nodes_1based) - 1) + c + 1 for c in comps])) \
        if len(bc_nodes_1based) else np.zeros(0, dtype=np.int64)
    oasm.update_dofs(dd)
    oasm.bc_vals[:] = bc_value
    return asm, p, oasm


def perturb(mesh, amp, seed=1234):
    """displace nodes by U(-amp, amp) (SURVEY 8d 'perturbed' variant: non-constant Jacobians)"""
    rng = np.random.default_rng(seed)
    X = np.asarray(mesh.nodal_coords)
    X += rng.uniform(-amp, amp, X.shape)
    return mesh


# ---- the same comparison AT SIZE: the C/OpenMP restatement (oracle/fec_oracle_c.c) does 48^3 neo-Hookean in about a
# second, the numpy oracle does not.  _update_dofs! (SparsityPatterns.jl:160-231) is applied vectorised here.
def c_oracle_reference(mesh, phys_name, props, dirichlet_dofs_1based, bc_vals, Uu, *, source_func=None, nthreads=None):
    """Residual (unknown entries), CSR rowptr / colval / nzval of the NON-condensed assembler, through the C oracle.
    Single-block meshes, no periodic BCs.  Returns dict(R, rowptr, colval, nz, U_full)."""
    import fec_oracle_clib as OC
    bname = mesh.element_block_names[0]
    et = mesh.element_types[bname]
    conn = np.asarray(mesh.element_conns[bname])
    X = np.asarray(mesh.nodal_coords)
    nd, nn = X.shape
    nf = 1 if phys_name == "poisson" else nd
    tabs = O.ref_fe_tables(et, _RULES[et])
    nthreads = nthreads or OC.max_threads()
    dof = O.update_dofs(nf, nn, np.unique(np.asarray(dirichlet_dofs_1based, dtype=np.int64)))
    U = np.zeros(nf * nn)
    U[np.asarray(dirichlet_dofs_1based, dtype=np.int64) - 1] = bc_vals
    U[dof["unknown_dofs"] - 1] = Uu
    fq = None
    if phys_name == "poisson" and source_func is not None:
        x_el = np.transpose(X[:, conn - 1], (2, 1, 0))                      # (NE, NNPE, ND)
        fq = np.stack([source_func(np.einsum("a,eai->ei", tabs[0][q], x_el)) for q in range(len(tabs[2]))], axis=1)
    cp = OC.CProblem(conn, X, tabs, phys_name, nf, props if props is not None else (), source_q=fq)
    R = cp.assemble_vector(U, nthreads=nthreads)
    coo = cp.assemble_matrix_coo(U, 2, nthreads=nthreads)
    Is, Js = cp.pattern()
    d2u = dof["dof_to_unknown"]
    ri, rj = d2u[Is - 1], d2u[Js - 1]
    keep = (ri > 0) & (rj > 0)
    slots = np.nonzero(keep)[0].astype(np.int64) + 1
    ri, rj = np.ascontiguousarray(ri[keep]), np.ascontiguousarray(rj[keep])
    n = len(dof["unknown_dofs"])
    ws = OC.SparseWorkspace(len(ri), n)
    colptr, rowval, nz = ws.sparse_csc(ri, rj, slots, coo)
    rowptr, colval, nzr = ws.csr(colptr, rowval, nz)
    return dict(R=R[dof["unknown_dofs"] - 1], rowptr=rowptr, colval=colval, nz=nzr, U_full=U, dof=dof, cp=cp, n=n)
