"""Parity of every element / physics / quadrature instantiation the library ships, against the oracle:
runtime-NQ kernels (GLL rules), TET4, TET10 matrices, per-block properties on a multi-block mesh."""
import os

import numpy as np
import pytest

import fec_oracle as O
from util_parity import GOLDEN, RTOL, perturb, product_physics, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def F():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import fecb200
    return fecb200


def _tet4_from_tet10(F, n):
    m = F.KuhnTet10Mesh(n)
    conn = m.element_conns["block_1"][:4]
    used = np.unique(conn)
    remap = np.zeros(m.num_nodes() + 1, dtype=np.int64)
    remap[used] = np.arange(1, len(used) + 1)
    t = F.UnstructuredMesh(data=dict(coords=np.asarray(m.nodal_coords)[:, used - 1], block_names=["block_1"], types=["TETRA4"],
                                     conns=[remap[conn]], nodesets={k: remap[v][remap[v] > 0] for k, v in m.nodeset_nodes.items()},
                                     sidesets={}))
    return t


CASES = [
    # (element, physics, q_type, q_degree, oracle rule)
    ("hex", "poisson", "GaussLobattoLegendre", 3, "gll3"),      # 27 points: runtime-NQ vector + matrix kernels
    ("hex", "poisson", "GaussLobattoLegendre", 2, "gll2"),
    ("quad", "poisson", "GaussLobattoLegendre", 3, "gll3"),
    ("quad", "linear", "GaussLegendre", 2, "gauss2"),
    ("quad", "neo", "GaussLegendre", 2, "gauss2"),
    ("tri", "linear", "GaussLegendre", 2, "tri3"),
    ("tri", "poisson", "GaussLegendre", 1, "tri1"),
    ("tet4", "poisson", "GaussLegendre", 2, "tet4"),
    ("tet4", "neo", "GaussLegendre", 2, "tet4"),
    ("tet4", "linear", "GaussLegendre", 1, "tet1"),
    ("tet10", "linear", "GaussLegendre", 2, "tet4"),
    ("tet10", "neo", "GaussLegendre", 2, "tet4"),
    ("tet10", "poisson", "GaussLegendre", 2, "tet4"),
    ("hex", "j2", "GaussLegendre", 2, "gauss2"),                # Walsh form of k_mat2 with the generic per-pair pull-back
    ("hex", "neo", "GaussLobattoLegendre", 2, "gll2"),          # Walsh form with c = 1 (points on the vertices)
    ("hex", "linear", "GaussLobattoLegendre", 2, "gll2"),
]


@pytest.mark.parametrize("el,phys,qt,qd,rule", CASES)
def test_element_family(F, el, phys, qt, qd, rule):
    rng = np.random.default_rng(17)
    if el == "hex":
        mesh = perturb(F.StructuredMesh("hex", (0, 0, 0), (1, 1, 1), (5, 4, 6)), 0.03)
    elif el == "quad":
        mesh = perturb(F.StructuredMesh("quad", (0, 0), (1, 1), (9, 7)), 0.02)
    elif el == "tri":
        mesh = perturb(F.StructuredMesh("tri", (0, 0), (1, 1), (8, 9)), 0.02)
    elif el == "tet4":
        mesh = perturb(_tet4_from_tet10(F, 3), 0.02)
    else:
        mesh = perturb(F.KuhnTet10Mesh(3), 0.015)
    nd = mesh.num_dimensions()
    nf = 1 if phys == "poisson" else nd
    props = {"poisson": None, "j2": np.array([1e3, 10e9, 1e9, 2e8, 1e8])}.get(phys, np.array([1e3, 10e6, 1e6]))
    src = (lambda X: 1.0 + X[:, 0] * X[:, 1]) if phys == "poisson" else None
    V = F.FunctionSpace(mesh, F.H1Field, F.Lagrange, q_type=qt, q_degree=qd)
    u = F.ScalarFunction(V, "u") if nf == 1 else F.VectorFunction(V, "displ")
    asm = F.SparseMatrixAssembler(u, sparse_matrix_type="csr")
    dbcs = [F.DirichletBC(c, lambda X, t: np.full(X.shape[0], 0.01), nodeset_name="bottom") for c in u.names()]
    p = F.create_parameters(mesh, asm, product_physics(F, phys, nd, src), props, dirichlet_bcs=dbcs)
    bname = mesh.element_block_names[0]
    ophys = {"poisson": O.Poisson(src), "linear": O.LinearElastic(nd), "neo": O.NeoHookean(nd), "j2": O.J2Plasticity(nd)}[phys]
    blk = O.Block(mesh.element_conns[bname], O.ref_fe_tables(mesh.element_types[bname], rule), ophys,
                  props=props if props is not None else ())
    oasm = O.OracleAssembler(np.asarray(mesh.nodal_coords), [blk], nf, condensed=False, matrix_type="csr")
    oasm.update_dofs(p.dirichlet_bcs.dirichlet_dofs())
    oasm.bc_vals[:] = 0.01
    N = asm.sizes()[2]
    Uu = (0.15 if phys == "j2" else 0.02) * rng.standard_normal(N)
    Vu = rng.random(N)
    F.assemble_vector(asm, F.residual, Uu, p)
    oasm.assemble_vector(Uu)
    assert rel_err(F.residual(asm), oasm.residual()) < RTOL
    F.assemble_stiffness(asm, F.stiffness, Uu, p)
    oasm.assemble_stiffness(Uu)
    K = F.stiffness(asm)
    n, ptr, idx = asm.pattern()
    optr, oidx, onz = oasm.stiffness()
    assert np.array_equal(ptr, optr) and np.array_equal(idx, oidx)
    assert rel_err(K.data, onz) < RTOL
    F.assemble_matrix_action(asm, F.stiffness, Uu, Vu, p)
    oasm.assemble_matrix_action(Uu, Vu)
    assert rel_err(F.hvp(asm, Vu), oasm.hvp(Vu)) < 1e-11
    F.assemble_mass(asm, F.mass, Uu, p)
    oasm.assemble_stiffness(Uu, kind="mass")
    assert rel_err(F.mass(asm).data, oasm.stiffness()[2]) < RTOL
    asm.close()


def test_per_block_properties(F):
    """two blocks with different property vectors (props per block: Assemblers.jl:193-202)"""
    mesh = F.UnstructuredMesh(os.path.join(GOLDEN, "multi_block_quad4_tri3.npz"))
    V = F.FunctionSpace(mesh, F.H1Field, F.Lagrange)
    u = F.VectorFunction(V, "displ")
    asm = F.SparseMatrixAssembler(u, sparse_matrix_type="csr", use_condensed=True)
    props = [np.array([1e3, 10e9, 1e9]), np.array([2e3, 5e9, 2e9])]
    dbcs = [F.DirichletBC(c, lambda X, t: np.zeros(X.shape[0]), sideset_name="boundary") for c in u.names()]
    p = F.create_parameters(mesh, asm, F.Mechanics(F.PlaneStrain()), props, dirichlet_bcs=dbcs)
    blocks = [O.Block(mesh.element_conns[b], O.ref_fe_tables(mesh.element_types[b], "gauss2" if mesh.element_types[b] == "QUAD4" else "tri3"),
                      O.LinearElastic(2), props=pr) for b, pr in zip(mesh.element_block_names, props)]
    oasm = O.OracleAssembler(np.asarray(mesh.nodal_coords), blocks, 2, condensed=True, matrix_type="csr")
    oasm.update_dofs(p.dirichlet_bcs.dirichlet_dofs())
    Uu = 1e-3 * np.random.default_rng(1).standard_normal(asm.sizes()[2])
    F.assemble_vector(asm, F.residual, Uu, p)
    oasm.assemble_vector(Uu)
    assert rel_err(F.residual(asm), oasm.residual()) < RTOL
    F.assemble_stiffness(asm, F.stiffness, Uu, p)
    oasm.assemble_stiffness(Uu)
    assert rel_err(F.stiffness(asm).data, oasm.stiffness()[2]) < RTOL
    asm.close()


def test_error_behaviour(F):
    """errors follow the reference: matrix assembly on a matrix-free assembler (Matrix.jl:23-28), bad sparse type,
    wrong NF for the physics, closures instead of the shipped element functions."""
    mesh = F.StructuredMesh("hex", (0, 0, 0), (1, 1, 1), (3, 3, 3))
    V = F.FunctionSpace(mesh, F.H1Field, F.Lagrange)
    with pytest.raises(ValueError):
        F.SparseMatrixAssembler(F.ScalarFunction(V, "u"), sparse_matrix_type="coo")
    asm = F.SparseMatrixAssembler(F.ScalarFunction(V, "u"), matrix_free=True)
    with pytest.raises(F.FECError):
        F.create_parameters(mesh, asm, F.NeoHookean(F.ThreeDimensional()), np.array([1e3, 1e7, 1e6]))   # NF mismatch
    p = F.create_parameters(mesh, asm, F.Poisson(None), None)
    Uu = F.create_unknowns(asm)
    with pytest.raises(F.FECError, match="matrix-free"):
        F.assemble_stiffness(asm, F.stiffness, Uu, p)
    with pytest.raises(TypeError):
        F.assemble_vector(asm, lambda *a: 0, Uu, p)
    assert F.stiffness(asm).shape == (27, 27) and F.stiffness(asm).nnz == 0   # _zero_sparse_matrix (Assemblers.jl:393-400)
    asm.close()


@pytest.mark.parametrize("mode", ["classic_env", "permuted_points", "skewed_rule"])
def test_k_mat2_walsh_fallbacks(F, mode, monkeypatch):
    """The Walsh form of the HEX8 kernels is taken only when the block's dN table is the trilinear table on a symmetric
    2-point rule per axis (detect_walsh, common.cuh).  Forced off (FECB200_MAT2_CLASSIC / FECB200_VEC_CLASSIC) and with a
    rule that is not of that kind (one point moved) the plain quadrature loop must run; a rule whose points merely come in
    another order keeps the Walsh form.  Same answer as the oracle on the very same tables in every case."""
    from fecb200.reference_fe import ReferenceFE
    rng = np.random.default_rng(5)
    mesh = perturb(F.StructuredMesh("hex", (0, 0, 0), (1, 1, 1), (6, 5, 4)), 0.03)
    props = np.array([1e3, 10e6, 1e6])
    rfe = ReferenceFE("HEX8", "GaussLegendre", 2)
    tabs = O.ref_fe_tables("HEX8", "gauss2")
    expect = 1
    if mode == "classic_env":
        monkeypatch.setenv("FECB200_MAT2_CLASSIC", "1")
        monkeypatch.setenv("FECB200_VEC_CLASSIC", "1")
    elif mode == "skewed_rule":   # 8 points, but not a tensor rule: point 5 sits elsewhere
        from fecb200.reference_fe import _tensor_shape, _SIGNS
        xi = np.array([[sx, sy, sz] for sz in (-1, 1) for sy in (-1, 1) for sx in (-1, 1)], dtype=float) / np.sqrt(3.0)
        xi[5] = [0.3, -0.45, 0.6]
        rfe.N, rfe.dN = (np.ascontiguousarray(t) for t in _tensor_shape(_SIGNS["HEX8"], xi))
        tabs = (rfe.N, rfe.dN, rfe.w)
        expect = 0
    else:
        perm = np.array([3, 0, 6, 1, 7, 2, 5, 4])
        rfe.N, rfe.dN, rfe.w = (np.ascontiguousarray(rfe.N[perm]), np.ascontiguousarray(rfe.dN[perm]),
                                np.ascontiguousarray(rfe.w[perm]))
        tabs = tuple(np.ascontiguousarray(np.asarray(t)[perm]) for t in tabs)
    V = F.FunctionSpace(mesh, F.H1Field, F.Lagrange, ref_fes=[rfe])
    u = F.VectorFunction(V, "displ")
    asm = F.SparseMatrixAssembler(u, sparse_matrix_type="csr")
    dbcs = [F.DirichletBC(c, lambda X, t: np.full(X.shape[0], 0.01), nodeset_name="bottom") for c in u.names()]
    p = F.create_parameters(mesh, asm, product_physics(F, "neo", 3, None), props, dirichlet_bcs=dbcs)
    bname = mesh.element_block_names[0]
    blk = O.Block(mesh.element_conns[bname], tabs, O.NeoHookean(3), props=props)
    oasm = O.OracleAssembler(np.asarray(mesh.nodal_coords), [blk], 3, condensed=False, matrix_type="csr")
    oasm.update_dofs(p.dirichlet_bcs.dirichlet_dofs())
    oasm.bc_vals[:] = 0.01
    Uu = 0.02 * rng.standard_normal(asm.sizes()[2])
    assert asm.kernel_form(0) == expect   # a permuted rule is still of the Walsh kind; the env switches force the loop at launch
    F.assemble_vector_and_stiffness(asm, F.residual, F.stiffness, Uu, p)   # the fused kernel
    oasm.assemble_vector(Uu)
    oasm.assemble_stiffness(Uu)
    assert rel_err(F.residual(asm), oasm.residual()) < RTOL
    assert rel_err(F.stiffness(asm).data, oasm.stiffness()[2]) < RTOL
    asm.close()


@pytest.mark.parametrize("phys", ["poisson", "neo"])
def test_k_vec_walsh_all_modes(F, phys, monkeypatch):
    """The Walsh form of the vector kernel is the product path only for the mechanics residual (where it is faster);
    FECB200_VEC_WALSH_ALL=1 routes the scalar residual (with its source term) and the matrix-free actions through it
    too, so the whole restatement stays parity-tested."""
    monkeypatch.setenv("FECB200_VEC_WALSH_ALL", "1")
    rng = np.random.default_rng(11)
    mesh = perturb(F.StructuredMesh("hex", (0, 0, 0), (1, 1, 1), (5, 6, 4)), 0.03)
    nf = 1 if phys == "poisson" else 3
    props = None if phys == "poisson" else np.array([1e3, 10e6, 1e6])
    src = (lambda X: 1.0 + X[:, 0] * X[:, 1]) if phys == "poisson" else None
    V = F.FunctionSpace(mesh, F.H1Field, F.Lagrange)
    u = F.ScalarFunction(V, "u") if nf == 1 else F.VectorFunction(V, "displ")
    asm = F.SparseMatrixAssembler(u, sparse_matrix_type="csr")
    dbcs = [F.DirichletBC(c, lambda X, t: np.full(X.shape[0], 0.01), nodeset_name="bottom") for c in u.names()]
    p = F.create_parameters(mesh, asm, product_physics(F, phys, 3, src), props, dirichlet_bcs=dbcs)
    bname = mesh.element_block_names[0]
    ophys = O.Poisson(src) if phys == "poisson" else O.NeoHookean(3)
    blk = O.Block(mesh.element_conns[bname], O.ref_fe_tables("HEX8", "gauss2"), ophys, props=props if props is not None else ())
    oasm = O.OracleAssembler(np.asarray(mesh.nodal_coords), [blk], nf, condensed=False, matrix_type="csr")
    oasm.update_dofs(p.dirichlet_bcs.dirichlet_dofs())
    oasm.bc_vals[:] = 0.01
    N = asm.sizes()[2]
    Uu, Vu = 0.02 * rng.standard_normal(N), rng.random(N)
    F.assemble_vector(asm, F.residual, Uu, p)
    oasm.assemble_vector(Uu)
    assert rel_err(F.residual(asm), oasm.residual()) < RTOL
    F.assemble_matrix_action(asm, F.stiffness, Uu, Vu, p)
    oasm.assemble_matrix_action(Uu, Vu)
    assert rel_err(F.hvp(asm, Vu), oasm.hvp(Vu)) < 1e-11
    asm.close()


@pytest.mark.parametrize("phys", ["neo", "j2"])
def test_walsh_form_in_any_numbering(F, phys):
    """The Walsh form reads the node / point sign triples off the block's own dN table (detect_walsh), so a host that
    numbers the HEX8 nodes and the 8 quadrature points differently (ReferenceFiniteElements.jl is not vendored: its
    ordering is unknown here) still gets the fast kernels -- and the same numbers as the oracle on the same tables."""
    from fecb200.reference_fe import ReferenceFE
    rng = np.random.default_rng(23)
    mesh = perturb(F.StructuredMesh("hex", (0, 0, 0), (1, 1, 1), (6, 4, 5)), 0.03)
    bname = mesh.element_block_names[0]
    pn, pq = rng.permutation(8), rng.permutation(8)
    mesh.element_conns[bname] = np.ascontiguousarray(np.asarray(mesh.element_conns[bname])[pn])   # local node a := old node pn[a]
    rfe = ReferenceFE("HEX8", "GaussLegendre", 2)
    rfe.N = np.ascontiguousarray(rfe.N[pq][:, pn])
    rfe.dN = np.ascontiguousarray(rfe.dN[pq][:, pn, :])
    rfe.w = np.ascontiguousarray(rfe.w[pq])
    N0, dN0, w0 = O.ref_fe_tables("HEX8", "gauss2")
    tabs = (np.ascontiguousarray(np.asarray(N0)[pq][:, pn]), np.ascontiguousarray(np.asarray(dN0)[pq][:, pn, :]),
            np.ascontiguousarray(np.asarray(w0)[pq]))
    props = np.array([1e3, 10e9, 1e9, 2e8, 1e8]) if phys == "j2" else np.array([1e3, 10e6, 1e6])
    V = F.FunctionSpace(mesh, F.H1Field, F.Lagrange, ref_fes=[rfe])
    u = F.VectorFunction(V, "displ")
    asm = F.SparseMatrixAssembler(u, sparse_matrix_type="csr")
    dbcs = [F.DirichletBC(c, lambda X, t: np.full(X.shape[0], 0.01), nodeset_name="bottom") for c in u.names()]
    p = F.create_parameters(mesh, asm, product_physics(F, phys, 3, None), props, dirichlet_bcs=dbcs)
    assert asm.kernel_form(0) == 1, "the permuted table should still be recognised"
    ophys = O.J2Plasticity(3) if phys == "j2" else O.NeoHookean(3)
    blk = O.Block(mesh.element_conns[bname], tabs, ophys, props=props)
    oasm = O.OracleAssembler(np.asarray(mesh.nodal_coords), [blk], 3, condensed=False, matrix_type="csr")
    oasm.update_dofs(p.dirichlet_bcs.dirichlet_dofs())
    oasm.bc_vals[:] = 0.01
    Uu = (0.15 if phys == "j2" else 0.02) * rng.standard_normal(asm.sizes()[2])
    F.assemble_vector(asm, F.residual, Uu, p)                              # Walsh k_vec
    oasm.assemble_vector(Uu)
    assert rel_err(F.residual(asm), oasm.residual()) < RTOL
    F.assemble_vector_and_stiffness(asm, F.residual, F.stiffness, Uu, p)   # Walsh k_mat2, fused
    oasm.assemble_stiffness(Uu)
    n, ptr, idx = asm.pattern()
    optr, oidx, onz = oasm.stiffness()
    assert np.array_equal(ptr, optr) and np.array_equal(idx, oidx)
    assert rel_err(F.residual(asm), oasm.residual()) < RTOL
    assert rel_err(F.stiffness(asm).data, onz) < RTOL
    F.assemble_stiffness(asm, F.stiffness, Uu, p)                          # tangent only
    assert rel_err(F.stiffness(asm).data, onz) < RTOL
    asm.close()


@pytest.mark.parametrize("phys", ["poisson", "neo"])
def test_spmv_versions_agree_with_scipy(F, phys, monkeypatch):
    """y = K x of the device CG (fecb200_matrix_multiply): the packed-adjacency kernel (k_spmv2) and the first version
    (FECB200_SPMV1) against the CSR values times x on the host, with Dirichlet-eliminated rows / columns present."""
    rng = np.random.default_rng(3)
    mesh = perturb(F.StructuredMesh("hex", (0, 0, 0), (1, 1, 1), (6, 5, 7)), 0.03)
    nf = 1 if phys == "poisson" else 3
    props = None if phys == "poisson" else np.array([1e3, 10e6, 1e6])
    src = (lambda X: 1.0 + X[:, 0]) if phys == "poisson" else None
    V = F.FunctionSpace(mesh, F.H1Field, F.Lagrange)
    u = F.ScalarFunction(V, "u") if nf == 1 else F.VectorFunction(V, "displ")
    asm = F.SparseMatrixAssembler(u, sparse_matrix_type="csr")
    names = u.names()
    dbcs = [F.DirichletBC(names[0], lambda X, t: np.zeros(X.shape[0]), nodeset_name="bottom")]   # one component only: partial masks
    if nf == 3:
        dbcs.append(F.DirichletBC(names[2], lambda X, t: np.zeros(X.shape[0]), nodeset_name="left"))
    p = F.create_parameters(mesh, asm, product_physics(F, phys, 3, src), props, dirichlet_bcs=dbcs)
    N = asm.sizes()[2]
    Uu, x = 0.02 * rng.standard_normal(N), rng.standard_normal(N)
    F.assemble_stiffness(asm, F.stiffness, Uu, p)
    K = F.stiffness(asm)
    ref = K @ x
    y2 = np.asarray(F.matrix_multiply(asm, x))
    monkeypatch.setenv("FECB200_SPMV1", "1")
    y1 = np.asarray(F.matrix_multiply(asm, x))
    assert rel_err(y2, ref) < 1e-13 and rel_err(y1, ref) < 1e-13
    asm.close()
