"""GPU parity tests: the CUDA path (through the C ABI, via the fecb200 host mirror) against the
oracle on the same seeded inputs.  Integer outputs (DOF maps, rowptr/colval) must be bit-exact;
FP64 values within 1e-12 relative (north_star; reassociation from atomic ordering and FMA).
Modelled on the reference's test/TestAssemblers.jl:78-430."""
import os

import numpy as np
import pytest

import fec_oracle as O
from util_parity import GOLDEN, RTOL, build_pair, perturb, product_physics, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def F():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import fecb200
    return fecb200


def _boundary_nodes(mesh, names):
    return np.unique(np.concatenate([mesh.nodeset_nodes[n] for n in names]))


def _check_pattern_and_values(F, asm, oasm, K):
    n, ptr, idx = asm.pattern()
    optr, oidx, onz = oasm.stiffness()
    assert n == oasm.n
    assert np.array_equal(ptr, optr), "rowptr/colptr not bit-exact"
    assert np.array_equal(idx, oidx), "colval/rowval not bit-exact"
    assert rel_err(K.data, onz) < RTOL, rel_err(K.data, onz)


def _check_dof_maps(asm, oasm):
    assert np.array_equal(asm.dof.unknown_dofs, oasm.dof["unknown_dofs"])
    assert np.array_equal(asm.dof.dof_to_unknown, oasm.dof["dof_to_unknown"])


SRC3 = lambda X: 3 * np.pi ** 2 * np.sin(np.pi * X[:, 0]) * np.sin(np.pi * X[:, 1]) * np.sin(np.pi * X[:, 2])
SRC2 = lambda X: 2 * np.pi ** 2 * np.sin(np.pi * X[:, 0]) * np.sin(np.pi * X[:, 1])


@pytest.mark.parametrize("matrix_type", ["csr", "csc"])
@pytest.mark.parametrize("condensed", [False, True])
def test_poisson_hex8(F, condensed, matrix_type):
    """BASELINE config 2 at oracle size: residual, CSR/CSC stiffness, matrix action, mass."""
    n = 9
    mesh = perturb(F.StructuredMesh("hex", (0, 0, 0), (1, 1, 1), (n + 1,) * 3), 0.15 / n)
    bc = _boundary_nodes(mesh, ["bottom", "top", "left", "right", "back", "front"])
    asm, p, oasm = build_pair(F, mesh, "poisson", None, condensed=condensed, matrix_type=matrix_type,
                              bc_nodes_1based=bc, func=SRC3)
    _check_dof_maps(asm, oasm)
    rng = np.random.default_rng(42)
    Uu = rng.uniform(-1, 1, asm.sizes()[2])
    Vu = np.random.default_rng(7).uniform(0, 1, asm.sizes()[2])
    F.assemble_vector(asm, F.residual, Uu, p)
    R = F.residual(asm)
    oasm.assemble_vector(Uu)
    assert rel_err(R, oasm.residual()) < RTOL
    F.assemble_stiffness(asm, F.stiffness, Uu, p)
    K = F.stiffness(asm)
    oasm.assemble_stiffness(Uu)
    _check_pattern_and_values(F, asm, oasm, K)
    F.assemble_matrix_action(asm, F.stiffness, Uu, Vu, p)
    Kv = F.hvp(asm, Vu)
    oasm.assemble_matrix_action(Uu, Vu)
    assert rel_err(Kv, oasm.hvp(Vu)) < RTOL
    # matrix-free entry point gives the same numbers (TestAssemblers.jl:279-311)
    F.assemble_matrix_free_action(asm, F.stiffness_action, Uu, Vu, p)
    assert rel_err(F.hvp(asm, Vu), Kv) < 1e-14
    # mass matrix and its action (TestAssemblers.jl:160-178)
    F.assemble_mass(asm, F.mass, Uu, p)
    M = F.mass(asm)
    oasm.assemble_stiffness(Uu, kind="mass")
    _, _, onz = oasm.stiffness()
    assert rel_err(M.data, onz) < RTOL
    F.assemble_matrix_action(asm, F.mass, Uu, Vu, p)
    Mv = F.hvp(asm, Vu)
    oasm.assemble_matrix_action(Uu, Vu, kind="mass")
    assert rel_err(Mv, oasm.hvp(Vu)) < RTOL
    asm.close()


@pytest.mark.parametrize("phys", ["neo", "neo_as_written", "linear"])
@pytest.mark.parametrize("condensed", [False, True])
def test_mechanics_hex8(F, phys, condensed):
    """BASELINE config 3 at oracle size: neo-Hookean residual + tangent + action."""
    n = 6
    mesh = perturb(F.StructuredMesh("hex", (0, 0, 0), (1, 1, 1), (n + 1,) * 3), 0.15 / n)
    props = np.array([1e3, 10e6, 1e6])
    asm, p, oasm = build_pair(F, mesh, phys, props, condensed=condensed, matrix_type="csr",
                              bc_nodes_1based=mesh.nodeset_nodes["bottom"])
    _check_dof_maps(asm, oasm)
    X = np.asarray(mesh.nodal_coords)
    rng = np.random.default_rng(42)
    Ufull = 0.02 * np.stack([np.sin(2 * np.pi * X[1]), np.sin(2 * np.pi * X[2]), np.sin(2 * np.pi * X[0])])
    Ufull += rng.uniform(-1e-3, 1e-3, Ufull.shape) / n
    Uflat = Ufull.reshape(-1, order="F")
    Uu = Uflat.copy() if condensed else Uflat[asm.dof.unknown_dofs - 1]
    Vu = np.random.default_rng(7).uniform(0, 1, len(Uu))
    F.assemble_vector(asm, F.residual, Uu, p)
    R = F.residual(asm)
    oasm.assemble_vector(Uu)
    assert rel_err(R, oasm.residual()) < RTOL
    F.assemble_stiffness(asm, F.stiffness, Uu, p)
    K = F.stiffness(asm)
    oasm.assemble_stiffness(Uu)
    _check_pattern_and_values(F, asm, oasm, K)
    F.assemble_matrix_action(asm, F.stiffness, Uu, Vu, p)
    Kv = F.hvp(asm, Vu)
    oasm.assemble_matrix_action(Uu, Vu)
    assert rel_err(Kv, oasm.hvp(Vu)) < RTOL
    if not condensed:
        assert rel_err(K @ Vu, Kv) < 1e-11
    asm.close()


@pytest.mark.parametrize("phys", ["poisson", "linear"])
@pytest.mark.parametrize("matrix_type", ["csr", "csc"])
@pytest.mark.parametrize("condensed", [False, True])
def test_multi_block_quad4_tri3(F, phys, matrix_type, condensed):
    """The reference's TestAssemblers fixture (test/TestAssemblers.jl:39-76): 280 QUAD4 + 170 TRI3."""
    mesh = F.UnstructuredMesh(os.path.join(GOLDEN, "multi_block_quad4_tri3.npz"))
    props = None if phys == "poisson" else np.array([1e3, 10e9, 1e9])
    asm, p, oasm = build_pair(F, mesh, phys, props, condensed=condensed, matrix_type=matrix_type,
                              bc_nodes_1based=mesh.sideset_nodes["boundary"], func=SRC2 if phys == "poisson" else None)
    _check_dof_maps(asm, oasm)
    rng = np.random.default_rng(3)
    scale = 1.0 if phys == "poisson" else 1e-3
    Uu = scale * rng.uniform(-1, 1, asm.sizes()[2])
    Vu = rng.uniform(0, 1, asm.sizes()[2])
    F.assemble_vector(asm, F.residual, Uu, p)
    oasm.assemble_vector(Uu)
    assert rel_err(F.residual(asm), oasm.residual()) < RTOL
    F.assemble_stiffness(asm, F.stiffness, Uu, p)
    oasm.assemble_stiffness(Uu)
    _check_pattern_and_values(F, asm, oasm, F.stiffness(asm))
    F.assemble_matrix_action(asm, F.stiffness, Uu, Vu, p)
    oasm.assemble_matrix_action(Uu, Vu)
    assert rel_err(F.hvp(asm, Vu), oasm.hvp(Vu)) < RTOL
    F.assemble_mass(asm, F.mass, Uu, p)
    oasm.assemble_stiffness(Uu, kind="mass")
    assert rel_err(F.mass(asm).data, oasm.stiffness()[2]) < RTOL
    asm.close()


def test_poisson_gold_solve(F):
    """BASELINE config 1: test/poisson/poisson.g (16384 QUAD4) Newton + CG on the device against
    test/poisson/poisson.gold (TestPoisson.jl:54-103; exodiff default tolerance).  The same numbers pin the
    reference's 'Laplace with sources' regression (test/laplace_with_source/TestLaplace.jl:24-76): laplace.g / laplace.gold
    are byte-identical to poisson.g / poisson.gold in mesh and nodal values, and its Source("u", f, "block_1") is the
    quadrature-point source array this library receives through fecb200_set_source_q."""
    mesh = F.UnstructuredMesh(os.path.join(GOLDEN, "poisson_g.npz"))
    gold = np.load(os.path.join(GOLDEN, "poisson_g.npz"))["gold_u"]
    for condensed in (False, True):
        V = F.FunctionSpace(mesh, F.H1Field, F.Lagrange)
        u = F.ScalarFunction(V, "u")
        asm = F.SparseMatrixAssembler(u, sparse_matrix_type="csr", use_condensed=condensed)
        dbcs = [F.DirichletBC("u", lambda X, t: np.zeros(X.shape[0]), sideset_name=f"sset_{i}") for i in (1, 2, 3, 4)]
        p = F.create_parameters(mesh, asm, F.Poisson(lambda X, t: SRC2(X)), None, dirichlet_bcs=dbcs)
        solver = F.NewtonSolver(F.IterativeLinearSolver(asm, "cg"))
        integ = F.QuasiStaticIntegrator(solver)
        integ.evolve(p)
        Ufield = p.field.data_flat
        err = np.abs(Ufield - gold).max()
        assert err < 1e-6, err
        assert solver.iterations <= 10
        asm.close()


def test_newton_iteration_count_matches_oracle(F):
    """north_star: 'Newton solves converge in the same number of iterations' (neo-Hookean, small)."""
    n = 4
    mesh = F.StructuredMesh("hex", (0, 0, 0), (1, 1, 1), (n + 1,) * 3)
    props = np.array([1e3, 10e6, 1e6])
    V = F.FunctionSpace(mesh, F.H1Field, F.Lagrange)
    u = F.VectorFunction(V, "displ")
    asm = F.SparseMatrixAssembler(u, sparse_matrix_type="csr", use_condensed=False)
    zero = lambda X, t: np.zeros(X.shape[0])
    pull = lambda X, t: np.full(X.shape[0], 0.1 * t)
    dbcs = [F.DirichletBC(c, zero, nodeset_name="bottom") for c in u.names()] + \
           [F.DirichletBC("displ_x", zero, nodeset_name="top"), F.DirichletBC("displ_z", zero, nodeset_name="top"),
            F.DirichletBC("displ_y", pull, nodeset_name="top")]
    p = F.create_parameters(mesh, asm, F.NeoHookean(F.ThreeDimensional()), props, dirichlet_bcs=dbcs,
                            times=F.TimeStepper(0.0, 1.0, 10))
    solver = F.NewtonSolver(F.IterativeLinearSolver(asm, "cg"))
    integ = F.QuasiStaticIntegrator(solver)
    # oracle twin
    blk = O.Block(mesh.element_conns["block_1"], O.ref_fe_tables("HEX8", "gauss2"), O.NeoHookean(3), props=props)
    oasm = O.OracleAssembler(np.asarray(mesh.nodal_coords), [blk], 3, condensed=False, matrix_type="csr")
    oasm.update_dofs(p.dirichlet_bcs.dirichlet_dofs())
    oUu = oasm.create_unknowns()
    for step in range(2):
        integ.evolve(p)
        bcs = p.dirichlet_bcs
        order = np.argsort(bcs.dofs, kind="stable")
        d_sorted, first = np.unique(bcs.dofs[order], return_index=True)
        # later BCs overwrite earlier ones on shared dofs, as the sequential loop in the reference does
        vals = np.zeros(len(d_sorted))
        for dof_id, val in zip(bcs.dofs, bcs.vals):
            vals[np.searchsorted(d_sorted, dof_id)] = val
        oasm.bc_vals[:] = vals
        oUu, nits, cgits, hist = O.newton_solve(oasm, oUu)
        assert solver.iterations == nits, (solver.iterations, nits)
        assert rel_err(integ.solution, oUu) < 1e-8


def test_j2_tet10_state(F):
    """BASELINE config 4 at oracle size: stateful J2 on tet10: residual + state update + action."""
    mesh = perturb(F.KuhnTet10Mesh(3), 0.02)
    props = np.array([1e3, 10e9, 1e9, 2e8, 1e8])
    asm, p, oasm = build_pair(F, mesh, "j2", props, condensed=False, matrix_type="csr",
                              bc_nodes_1based=mesh.nodeset_nodes["bottom"], matrix_free=True)
    X = np.asarray(mesh.nodal_coords)
    U = 0.15 * np.stack([X[1] ** 2, 0.5 * X[1] * X[0], -0.3 * X[1]])  # yields a good fraction of points
    Uu = U.reshape(-1, order="F")[asm.dof.unknown_dofs - 1]
    rng = np.random.default_rng(5)
    nq, ne = 4, mesh.element_conns["block_1"].shape[1]
    so = np.zeros((7, nq, ne))
    so[:2] = 1e-4 * rng.standard_normal((2, nq, ne)); so[2] = -so[0] - so[1]
    so[3:6] = 1e-4 * rng.standard_normal((3, nq, ne)); so[6] = 1e-4 * rng.random((nq, ne))
    p.set_state(so, which="old")
    oasm.blocks[0].state_old[:] = so
    F.assemble_vector(asm, F.residual, Uu, p)
    R = F.residual(asm)
    oasm.assemble_vector(Uu)
    assert rel_err(R, oasm.residual()) < RTOL
    sn = p.state(which="new")
    osn = oasm.blocks[0].state_new
    assert rel_err(sn, osn) < RTOL
    frac = np.mean(osn[6] > so[6])
    assert 0.05 < frac < 0.999, frac  # some, not all, quadrature points yield
    Vu = rng.random(len(Uu))
    F.assemble_matrix_free_action(asm, F.stiffness_action, Uu, Vu, p)
    Kv = F.hvp(asm, Vu)
    oasm.assemble_matrix_action(Uu, Vu)
    assert rel_err(Kv, oasm.hvp(Vu)) < 1e-11
    with pytest.raises(F.FECError, match="matrix-free"):
        F.assemble_stiffness(asm, F.stiffness, Uu, p)
    asm.close()


def test_full_dof_action_and_invariant(F):
    """assemble_matrix_free_action_full! (TestAssemblers.jl:341-430): (K_full v_full) and the
    'BC slots of V are zero afterwards' invariant (MatrixAction.jl:137-148)."""
    n = 5
    mesh = F.StructuredMesh("hex", (0, 0, 0), (1, 1, 1), (n + 1,) * 3)
    bc = _boundary_nodes(mesh, ["left"])
    asm, p, oasm = build_pair(F, mesh, "poisson", None, condensed=False, matrix_type="csr", bc_nodes_1based=bc)
    rng = np.random.default_rng(11)
    ndof = len(asm.dof)
    U_full, v_full = rng.standard_normal(ndof), rng.standard_normal(ndof)
    F.assemble_matrix_free_action_full(asm, F.stiffness_action, U_full, v_full, p)
    out = F.full_field(asm, "action")
    Xo = np.asarray(mesh.nodal_coords)
    ref = O.assemble_matrix_action(oasm.blocks, Xo, U_full.reshape(1, -1), v_full.reshape(1, -1), 1)
    assert rel_err(out, ref) < RTOL
    V_after = F.full_field(asm, "v")
    assert np.all(V_after[p.dirichlet_bcs.dirichlet_dofs() - 1] == 0.0)
    asm.close()


def test_periodic_vector_path(F):
    """periodic side-b dofs: field update U[b] = U[a] and residual fold-in (Assemblers.jl:357-368)."""
    n = 4
    mesh = F.StructuredMesh("quad", (0, 0), (1, 1), (n + 1, n + 1))
    V = F.FunctionSpace(mesh, F.H1Field, F.Lagrange)
    u = F.ScalarFunction(V, "u")
    asm = F.SparseMatrixAssembler(u, sparse_matrix_type="csr", matrix_free=True)
    p = F.create_parameters(mesh, asm, F.Poisson(lambda X, t: SRC2(X)), None,
                            dirichlet_bcs=[F.DirichletBC("u", lambda X, t: np.zeros(X.shape[0]), nodeset_name="bottom")])
    left, right = mesh.nodeset_nodes["left"][1:], mesh.nodeset_nodes["right"][1:]
    F.update_dofs(asm, p.dirichlet_bcs, periodic=(left, right))
    blk = O.Block(mesh.element_conns["block_1"], O.ref_fe_tables("QUAD4", "gauss2"), O.Poisson(SRC2))
    oasm = O.OracleAssembler(np.asarray(mesh.nodal_coords), [blk], 1, condensed=False)
    oasm.update_dofs(p.dirichlet_bcs.dirichlet_dofs(), left, right)
    assert np.array_equal(asm.dof.unknown_dofs, oasm.dof["unknown_dofs"])
    assert np.array_equal(asm.dof.dof_to_unknown, oasm.dof["dof_to_unknown"])
    Uu = np.random.default_rng(2).standard_normal(asm.sizes()[2])
    F.assemble_vector(asm, F.residual, Uu, p)
    oasm.assemble_vector(Uu)
    assert rel_err(F.residual(asm), oasm.residual()) < RTOL
    asm.close()


def test_device_pointers_zero_copy(F):
    """Uu / outputs may be device buffers (a CuArray on the Julia side): used in place."""
    import torch
    n = 6
    mesh = F.StructuredMesh("hex", (0, 0, 0), (1, 1, 1), (n + 1,) * 3)
    asm, p, oasm = build_pair(F, mesh, "poisson", None, condensed=False, matrix_type="csr",
                              bc_nodes_1based=mesh.nodeset_nodes["left"], func=SRC3)
    Uu = np.random.default_rng(0).standard_normal(asm.sizes()[2])
    dUu = torch.from_numpy(Uu).cuda()
    dR = torch.empty_like(dUu)
    F.assemble_vector(asm, F.residual, dUu, p)
    F.residual(asm, dR)
    torch.cuda.synchronize()
    oasm.assemble_vector(Uu)
    assert rel_err(dR.cpu().numpy(), oasm.residual()) < RTOL
    asm.close()


def test_large_roundtrip_properties(F):
    """Size-independent properties at a size the oracle cannot assemble quickly (64^3 Poisson):
    K*1 = 0 away from constraints (constants are in the kernel of the Laplacian), action linearity,
    K v (assembled, device SpMV through CG machinery) == matrix-free action, symmetry v.Kw = w.Kv."""
    n = 48
    mesh = F.StructuredMesh("hex", (0, 0, 0), (1, 1, 1), (n + 1,) * 3)
    V = F.FunctionSpace(mesh, F.H1Field, F.Lagrange)
    u = F.ScalarFunction(V, "u")
    asm = F.SparseMatrixAssembler(u, sparse_matrix_type="csr", use_condensed=False)
    p = F.create_parameters(mesh, asm, F.Poisson(None), None, dirichlet_bcs=[])
    N = asm.sizes()[2]
    rng = np.random.default_rng(9)
    Uu = rng.standard_normal(N)
    one = np.ones(N)
    F.assemble_matrix_action(asm, F.stiffness, Uu, one, p)
    assert np.abs(F.hvp(asm, one)).max() < 1e-12
    v, w = rng.standard_normal(N), rng.standard_normal(N)
    F.assemble_matrix_action(asm, F.stiffness, Uu, v, p); Kv = F.hvp(asm, v).copy()
    F.assemble_matrix_action(asm, F.stiffness, Uu, w, p); Kw = F.hvp(asm, w).copy()
    F.assemble_matrix_action(asm, F.stiffness, Uu, 2 * v - 3 * w, p)
    assert rel_err(F.hvp(asm, None), 2 * Kv - 3 * Kw) < 1e-12
    assert abs(v @ Kw - w @ Kv) < 1e-10 * abs(v @ Kw)
    F.assemble_stiffness(asm, F.stiffness, Uu, p)
    K = F.stiffness(asm)
    assert rel_err(K @ v, Kv) < 1e-12
    assert abs(K - K.T).max() < 1e-13
    # residual of the homogeneous problem equals K u
    F.assemble_vector(asm, F.residual, Uu, p)
    assert rel_err(F.residual(asm), K @ Uu) < 1e-12
    asm.close()


@pytest.mark.parametrize("phys", ["neo", "linear"])
def test_fused_residual_and_tangent(F, phys):
    """fecb200_assemble_vector_and_matrix == the two separate calls (and the oracle)."""
    n = 6
    mesh = perturb(F.StructuredMesh("hex", (0, 0, 0), (1, 1, 1), (n + 1,) * 3), 0.15 / n)
    props = np.array([1e3, 10e6, 1e6])
    asm, p, oasm = build_pair(F, mesh, phys, props, condensed=False, matrix_type="csr",
                              bc_nodes_1based=mesh.nodeset_nodes["bottom"], bc_value=0.01)
    rng = np.random.default_rng(4)
    Uu = 0.01 * rng.standard_normal(asm.sizes()[2])
    F.assemble_vector(asm, F.residual, Uu, p)
    R1 = F.residual(asm).copy()
    F.assemble_stiffness(asm, F.stiffness, Uu, p)
    K1 = F.stiffness(asm).data.copy()
    F.assemble_vector_and_stiffness(asm, F.residual, F.stiffness, Uu, p)
    R2, K2 = F.residual(asm), F.stiffness(asm).data
    assert rel_err(R2, R1) < RTOL and rel_err(K2, K1) < RTOL
    oasm.assemble_vector(Uu)
    oasm.assemble_stiffness(Uu)
    assert rel_err(R2, oasm.residual()) < RTOL and rel_err(K2, oasm.stiffness()[2]) < RTOL
    asm.close()


def test_async_host_copies(F):
    """fecb200_set_async: pinned host inputs / outputs through the copy streams give the same numbers."""
    import torch
    from fecb200._lib import check, lib
    n = 8
    mesh = F.StructuredMesh("hex", (0, 0, 0), (1, 1, 1), (n + 1,) * 3)
    asm, p, oasm = build_pair(F, mesh, "neo", np.array([1e3, 10e6, 1e6]), condensed=False, matrix_type="csr",
                              bc_nodes_1based=mesh.nodeset_nodes["bottom"])
    N = asm.sizes()[2]
    rng = np.random.default_rng(8)
    U1, U2 = 0.01 * rng.standard_normal(N), 0.01 * rng.standard_normal(N)
    refs = []
    for U in (U1, U2):
        F.assemble_vector_and_stiffness(asm, F.residual, F.stiffness, U, p)
        refs.append(F.residual(asm).copy())
    h = asm._require()
    check(lib.fecb200_set_async(h, 1))
    hU = [torch.from_numpy(U).pin_memory() for U in (U1, U2)]
    hR = [torch.empty(N, dtype=torch.float64).pin_memory() for _ in range(2)]
    for i in range(2):                       # two back-to-back steps, no host synchronisation in between
        F.assemble_vector_and_stiffness(asm, F.residual, F.stiffness, hU[i], p)
        F.residual(asm, hR[i])
    check(lib.fecb200_synchronize(h))
    for i in range(2):
        assert rel_err(hR[i].numpy(), refs[i]) < 1e-14
    check(lib.fecb200_set_async(h, 0))
    asm.close()


@pytest.mark.parametrize("condensed", [False, True])
def test_double_buffered_stiffness(F, condensed):
    """fecb200_set_matrix_double_buffer: the kernel-side zero-fill of the idle value buffer replaces fill!(storage, 0)
    (Matrix.jl:39) without changing any value, over a sequence of assemblies that mixes the fused and the plain
    entry points, the condensed-mode adjustment, the device CG and an update_dofs in between."""
    n = 7
    mesh = perturb(F.StructuredMesh("hex", (0, 0, 0), (1, 1, 1), (n + 1, n + 2, n + 1)), 0.1 / n)
    props = np.array([1e3, 10e6, 1e6])
    asm, p, oasm = build_pair(F, mesh, "neo", props, condensed=condensed, matrix_type="csr",
                              bc_nodes_1based=mesh.nodeset_nodes["bottom"], bc_value=0.01)
    asm.set_matrix_double_buffer(True)
    rng = np.random.default_rng(11)
    N = asm.sizes()[2]
    for it in range(5):
        Uu = 0.01 * rng.standard_normal(N)
        if it % 2 == 0:
            F.assemble_vector_and_stiffness(asm, F.residual, F.stiffness, Uu, p)
            oasm.assemble_vector(Uu)
            assert rel_err(F.residual(asm), oasm.residual()) < RTOL
        else:
            F.assemble_stiffness(asm, F.stiffness, Uu, p)
        K = F.stiffness(asm)
        oasm.assemble_stiffness(Uu)
        assert rel_err(K.data, oasm.stiffness()[2]) < RTOL, it
        if it == 2 and not condensed:   # the solve reads the CURRENT buffer (condensed: penalty rows make CG's true residual stagnate)
            b = rng.random(N)
            x, its, rn = F.IterativeLinearSolver(asm, "cg").solve(b)
            assert np.linalg.norm(K @ x - b) <= 1e-6 * np.linalg.norm(b)
    asm.set_matrix_double_buffer(False)
    Uu = 0.01 * rng.standard_normal(N)
    F.assemble_stiffness(asm, F.stiffness, Uu, p)
    oasm.assemble_stiffness(Uu)
    assert rel_err(F.stiffness(asm).data, oasm.stiffness()[2]) < RTOL
    asm.close()


@pytest.mark.parametrize("case", ["poisson_hex8_csr", "poisson_hex8_csc", "linear_quad_tri", "poisson_quad_tri"])
def test_double_buffered_other_kernels(F, case):
    """The same double buffering through the thread-per-element scalar kernel (Poisson hex8) and the generic
    column-owner kernel on a two-block mesh (each launch clears its share of the idle buffer); a mass assembly in
    between must not disturb the stiffness buffers."""
    if case.startswith("poisson_hex8"):
        n = 8
        mesh = perturb(F.StructuredMesh("hex", (0, 0, 0), (1, 1, 1), (n + 1,) * 3), 0.15 / n)
        bc = _boundary_nodes(mesh, ["bottom", "top"])
        asm, p, oasm = build_pair(F, mesh, "poisson", None, condensed=False, matrix_type=case[-3:], bc_nodes_1based=bc, func=SRC3)
        scale = 1.0
    else:
        mesh = F.UnstructuredMesh(os.path.join(GOLDEN, "multi_block_quad4_tri3.npz"))
        phys = case.split("_")[0]
        asm, p, oasm = build_pair(F, mesh, phys, None if phys == "poisson" else np.array([1e3, 10e9, 1e9]), condensed=False,
                                  matrix_type="csr", bc_nodes_1based=mesh.sideset_nodes["boundary"],
                                  func=SRC2 if phys == "poisson" else None)
        scale = 1.0 if phys == "poisson" else 1e-3
    asm.set_matrix_double_buffer(True)
    rng = np.random.default_rng(5)
    for it in range(4):
        Uu = scale * rng.uniform(-1, 1, asm.sizes()[2])
        F.assemble_stiffness(asm, F.stiffness, Uu, p)
        oasm.assemble_stiffness(Uu)
        assert rel_err(F.stiffness(asm).data, oasm.stiffness()[2]) < RTOL, it
        if it == 1:
            F.assemble_mass(asm, F.mass, Uu, p)
            oasm.assemble_stiffness(Uu, kind="mass")
            assert rel_err(F.mass(asm).data, oasm.stiffness()[2]) < RTOL
    asm.close()


@pytest.mark.parametrize("case", ["neo_hex8", "poisson_hex8", "linear_quad_tri", "poisson_quad_tri"])
@pytest.mark.parametrize("condensed", [False, True])
def test_lumped_mass_and_diagonal(F, case, condensed):
    """assemble_lumped_mass! / assemble_diagonal! (LumpedMass.jl, Diagonal.jl) against the oracle,
    plus the reference's own contract (TestAssemblers.jl:432-518): lumped mass = row sums of the consistent mass."""
    if case.endswith("hex8"):
        n = 6
        mesh = perturb(F.StructuredMesh("hex", (0, 0, 0), (1, 1, 1), (n + 1, n + 2, n + 1)), 0.1 / n)
        phys = case.split("_")[0]
        props = np.array([1e3, 10e6, 1e6]) if phys == "neo" else None
        asm, p, oasm = build_pair(F, mesh, phys, props, condensed=condensed, matrix_type="csr",
                                  bc_nodes_1based=mesh.nodeset_nodes["bottom"], bc_value=0.01, func=SRC3 if phys == "poisson" else None)
        scale = 0.01
    else:
        mesh = F.UnstructuredMesh(os.path.join(GOLDEN, "multi_block_quad4_tri3.npz"))
        phys = case.split("_")[0]
        asm, p, oasm = build_pair(F, mesh, phys, None if phys == "poisson" else np.array([1e3, 10e9, 1e9]), condensed=condensed,
                                  matrix_type="csr", bc_nodes_1based=mesh.sideset_nodes["boundary"],
                                  func=SRC2 if phys == "poisson" else None)
        scale = 1.0 if phys == "poisson" else 1e-3
    rng = np.random.default_rng(12)
    Uu = scale * rng.uniform(-1, 1, asm.sizes()[2])
    F.assemble_lumped_mass(asm, F.lumped_mass, Uu, p)
    ml = F.lumped_mass(asm).copy()
    oasm.assemble_lumped_mass(Uu)
    assert rel_err(ml, oasm.vector_values()) < RTOL
    for func, kind in ((F.stiffness, "stiffness"), (F.mass, "mass")):
        F.assemble_diagonal(asm, func, Uu, p)
        d = F.diagonal(asm).copy()
        oasm.assemble_diagonal(Uu, kind=kind)
        assert rel_err(d, oasm.vector_values()) < RTOL, kind
        if not condensed:   # the true diagonal of the matrix the sparse path assembles
            (F.assemble_stiffness if kind == "stiffness" else F.assemble_mass)(asm, func, Uu, p)
            M = (F.stiffness if kind == "stiffness" else F.mass)(asm)
            assert rel_err(d, M.diagonal()) < RTOL, kind
    # the residual path is untouched by the shared storage
    F.assemble_vector(asm, F.residual, Uu, p)
    oasm.assemble_vector(Uu)
    assert rel_err(F.residual(asm), oasm.residual()) < RTOL
    asm.close()


def test_diagonal_with_state_tet10(F):
    """assemble_diagonal!(stiffness) reads state_old like assemble_stiffness! does (J2 tet10, yielded points)."""
    mesh = perturb(F.KuhnTet10Mesh(2), 0.02)
    props = np.array([1e3, 10e9, 1e9, 2e8, 1e8])
    asm, p, oasm = build_pair(F, mesh, "j2", props, condensed=False, matrix_type="csr",
                              bc_nodes_1based=mesh.nodeset_nodes["bottom"], matrix_free=True)
    X = np.asarray(mesh.nodal_coords)
    U = 0.15 * np.stack([X[1] ** 2, 0.5 * X[1] * X[0], -0.3 * X[1]])
    Uu = U.reshape(-1, order="F")[asm.dof.unknown_dofs - 1]
    rng = np.random.default_rng(6)
    nq, ne = 4, mesh.element_conns["block_1"].shape[1]
    so = np.zeros((7, nq, ne))
    so[:2] = 1e-4 * rng.standard_normal((2, nq, ne)); so[2] = -so[0] - so[1]
    so[3:6] = 1e-4 * rng.standard_normal((3, nq, ne)); so[6] = 1e-4 * rng.random((nq, ne))
    p.set_state(so, which="old")
    oasm.blocks[0].state_old[:] = so
    F.assemble_diagonal(asm, F.stiffness, Uu, p)
    oasm.assemble_diagonal(Uu, kind="stiffness")
    assert rel_err(F.diagonal(asm), oasm.vector_values()) < 1e-11
    F.assemble_lumped_mass(asm, F.lumped_mass, Uu, p)
    oasm.assemble_lumped_mass(Uu)
    assert rel_err(F.lumped_mass(asm), oasm.vector_values()) < RTOL
    asm.close()


def _match_periodic(X, side_a, side_b, axis, tol=1e-9):
    """PeriodicBCContainer (src/bcs/PeriodicBCs.jl:19-110): pair the nodes of two opposite sides by their
    coordinates in the directions other than `axis` (host bookkeeping; the library receives dof pairs)."""
    other = [j for j in range(X.shape[0]) if j != axis]
    key = lambda n: tuple(np.round(X[other, n - 1] / tol).astype(np.int64))
    lookup = {key(n): n for n in side_b}
    a = np.asarray(side_a, dtype=np.int64)
    return a, np.array([lookup[key(n)] for n in a], dtype=np.int64)


@pytest.mark.parametrize("case", ["poisson_quad4", "poisson_quad4_xy", "poisson_quad4_thin", "poisson_hex8", "neo_hex8"])
@pytest.mark.parametrize("matrix_type", ["csr", "csc"])
def test_periodic_matrix_assembly(F, case, matrix_type):
    """Periodic side-b dofs folded into their side-a unknown in the ASSEMBLED matrix (_update_dofs!,
    SparsityPatterns.jl:160-231 with dof_to_unknown_index, DofManagers.jl:188-201): rowptr / colval bit-exact with the
    oracle's sparse! path, values to 1e-12, through the generic, the scalar and the pair-owner kernels.
    '_xy': two periodic directions with corner chains; '_thin': one element between the two sides, so an element
    holds a node and its own periodic image."""
    phys = case.split("_")[0]
    if "quad4" in case:
        nx = 1 if case.endswith("thin") else 5
        mesh = F.StructuredMesh("quad", (0, 0), (1.3, 1), (nx + 1, 6))
        nf, etype, props, func = 1, "QUAD4", None, SRC2
    else:
        mesh = F.StructuredMesh("hex", (0, 0, 0), (1.3, 1, 0.8), (5, 4, 6))
        nf, etype = (1, "HEX8") if phys == "poisson" else (3, "HEX8")
        props, func = (None, SRC3) if phys == "poisson" else (np.array([1e3, 10e6, 1e6]), None)
    X = np.asarray(mesh.nodal_coords)
    left, right = mesh.nodeset_nodes["left"], mesh.nodeset_nodes["right"]
    interior = np.setdiff1d(np.arange(1, X.shape[1] + 1), np.concatenate([mesh.nodeset_nodes[k] for k in mesh.nodeset_nodes]))
    X[:, interior - 1] += np.random.default_rng(3).uniform(-0.02, 0.02, (X.shape[0], len(interior)))  # non-constant Jacobians
    bottom = set(mesh.nodeset_nodes["bottom"].tolist())
    V = F.FunctionSpace(mesh, F.H1Field, F.Lagrange, q_type="GaussLegendre", q_degree=2)
    u = F.ScalarFunction(V, "u") if nf == 1 else F.VectorFunction(V, "displ")
    asm = F.SparseMatrixAssembler(u, sparse_matrix_type=matrix_type, use_condensed=False)
    xy = case.endswith("xy")
    dbcs = [] if xy else [F.DirichletBC(c, lambda X_, t: np.zeros(X_.shape[0]), nodeset_name="bottom") for c in u.names()]
    ph = product_physics(F, phys, X.shape[0], func)
    p = F.create_parameters(mesh, asm, ph, props, dirichlet_bcs=dbcs)
    a_nodes, b_nodes = _match_periodic(X, [n for n in left if xy or n not in bottom], right, axis=0)
    if xy:   # second direction: bottom -> top; the corners form chains that update_dofs! resolves
        a2, b2 = _match_periodic(X, mesh.nodeset_nodes["bottom"], mesh.nodeset_nodes["top"], axis=1)
        a_nodes, b_nodes = np.concatenate([a_nodes, a2]), np.concatenate([b_nodes, b2])
    pa = np.concatenate([nf * (a_nodes - 1) + d + 1 for d in range(nf)])
    pb = np.concatenate([nf * (b_nodes - 1) + d + 1 for d in range(nf)])
    F.update_dofs(asm, p.dirichlet_bcs, periodic=(pa, pb))
    blk = O.Block(mesh.element_conns["block_1"], O.ref_fe_tables(etype, "gauss2"),
                  O.Poisson(func) if phys == "poisson" else O.NeoHookean(3), props=props if props is not None else ())
    oasm = O.OracleAssembler(X, [blk], nf, condensed=False, matrix_type=matrix_type)
    oasm.update_dofs(p.dirichlet_bcs.dirichlet_dofs() if dbcs else [], pa, pb)
    _check_dof_maps(asm, oasm)
    rng = np.random.default_rng(21)
    Uu = (1.0 if phys == "poisson" else 0.01) * rng.uniform(-1, 1, asm.sizes()[2])
    F.assemble_stiffness(asm, F.stiffness, Uu, p)
    oasm.assemble_stiffness(Uu)
    _check_pattern_and_values(F, asm, oasm, F.stiffness(asm))
    F.assemble_vector(asm, F.residual, Uu, p)
    oasm.assemble_vector(Uu)
    assert rel_err(F.residual(asm), oasm.residual()) < RTOL
    if phys == "neo":   # fused entry point: the residual rows still go to the ORIGINAL nodes, folded by the accessor
        F.assemble_vector_and_stiffness(asm, F.residual, F.stiffness, Uu, p)
        oasm.assemble_vector(Uu)    # residual(asm) folds side b into side a IN PLACE (Assemblers.jl:357-361): re-assemble
        assert rel_err(F.residual(asm), oasm.residual()) < RTOL
        assert rel_err(F.stiffness(asm).data, oasm.stiffness()[2]) < RTOL
    F.assemble_mass(asm, F.mass, Uu, p)
    oasm.assemble_stiffness(Uu, kind="mass")
    assert rel_err(F.mass(asm).data, oasm.stiffness()[2]) < RTOL
    # K v through the assembled matrix == the matrix-free action on the same periodic dof maps
    K = F.stiffness(asm)
    Vu = rng.uniform(0, 1, asm.sizes()[2])
    F.assemble_matrix_action(asm, F.stiffness, Uu, Vu, p)
    # (the action path folds nothing in hvp: compare against the oracle's own action instead of K v)
    oasm.assemble_matrix_action(Uu, Vu)
    assert rel_err(F.hvp(asm, Vu), oasm.hvp(Vu)) < RTOL
    # removing the periodic pairs again restores the plain pattern
    F.update_dofs(asm, p.dirichlet_bcs)
    oasm.update_dofs(p.dirichlet_bcs.dirichlet_dofs() if dbcs else [])
    Uu = (1.0 if phys == "poisson" else 0.01) * rng.uniform(-1, 1, asm.sizes()[2])
    F.assemble_stiffness(asm, F.stiffness, Uu, p)
    oasm.assemble_stiffness(Uu)
    _check_pattern_and_values(F, asm, oasm, F.stiffness(asm))
    asm.close()


def test_poisson_periodic_regression(F):
    """The reference's 'test_poisson_periodic' (test/poisson/TestPoissonPBCs.jl:86-127): poisson.g, periodic in x
    (sset_1 <-> sset_3) and y (sset_4 <-> sset_2), no Dirichlet BC, NewtonSolver(IterativeLinearSolver(asm, :cg)) with an
    ASSEMBLED csc matrix; every nodal value within 5e-4 of the analytic solution."""
    mesh = F.UnstructuredMesh(os.path.join(GOLDEN, "poisson_g.npz"))
    X = np.asarray(mesh.nodal_coords)
    f = lambda Xq: ((2 * np.pi) ** 2 * np.cos(2 * np.pi * Xq[:, 0]) + 0.5 * (4 * np.pi) ** 2 * np.cos(4 * np.pi * Xq[:, 1])
                    + 0.25 * ((2 * np.pi) ** 2 + (4 * np.pi) ** 2) * np.sin(2 * np.pi * Xq[:, 0]) * np.sin(4 * np.pi * Xq[:, 1]))
    u_an = np.cos(2 * np.pi * X[0]) + 0.5 * np.cos(4 * np.pi * X[1]) + 0.25 * np.sin(2 * np.pi * X[0]) * np.sin(4 * np.pi * X[1])
    V = F.FunctionSpace(mesh, F.H1Field, F.Lagrange)
    u = F.ScalarFunction(V, "u")
    asm = F.SparseMatrixAssembler(u, sparse_matrix_type="csc", use_condensed=False)
    p = F.create_parameters(mesh, asm, F.Poisson(lambda Xq, t: f(Xq)), None, dirichlet_bcs=[])
    ss = mesh.sideset_nodes
    # which coordinate is constant on a side tells the periodic direction's partner
    def axis_of(nodes):
        return int(np.argmin(np.ptp(X[:, nodes - 1], axis=1)))
    assert axis_of(ss["sset_1"]) == axis_of(ss["sset_3"]) and axis_of(ss["sset_4"]) == axis_of(ss["sset_2"])
    a1, b1 = _match_periodic(X, ss["sset_1"], ss["sset_3"], axis=axis_of(ss["sset_1"]), tol=1e-6)
    a2, b2 = _match_periodic(X, ss["sset_4"], ss["sset_2"], axis=axis_of(ss["sset_4"]), tol=1e-6)
    F.update_dofs(asm, p.dirichlet_bcs, periodic=(np.concatenate([a1, a2]), np.concatenate([b1, b2])))
    solver = F.NewtonSolver(F.IterativeLinearSolver(asm, "cg"))
    integ = F.QuasiStaticIntegrator(solver)
    integ.evolve(p)
    err = np.abs(p.field.data_flat - u_an).max()
    assert err < 5e-4, err
    asm.close()


@pytest.mark.parametrize("case", ["neo_hex8", "linear_hex8", "poisson_hex8", "linear_quad_tri", "poisson_quad_tri"])
def test_assemble_scalar_energy(F, case):
    """assemble_scalar!(asm, energy, Uu, p) (QuadratureQuantity.jl:4-45): JxW * energy at every quadrature point of
    every block, [NQ, NE] in the caller's element order, against the oracle."""
    phys = case.split("_")[0]
    if case.endswith("hex8"):
        n = 5
        mesh = perturb(F.StructuredMesh("hex", (0, 0, 0), (1, 1, 1), (n + 1, n + 2, n + 1)), 0.1 / n)
        props = None if phys == "poisson" else np.array([1e3, 10e6, 1e6])
        asm, p, oasm = build_pair(F, mesh, phys, props, condensed=False, matrix_type="csr", bc_nodes_1based=mesh.nodeset_nodes["bottom"],
                                  bc_value=0.01, func=SRC3 if phys == "poisson" else None, matrix_free=True)
        scale = 1.0 if phys == "poisson" else 0.02
    else:
        mesh = F.UnstructuredMesh(os.path.join(GOLDEN, "multi_block_quad4_tri3.npz"))
        asm, p, oasm = build_pair(F, mesh, phys, None if phys == "poisson" else np.array([1e3, 10e9, 1e9]), condensed=False,
                                  matrix_type="csr", bc_nodes_1based=mesh.sideset_nodes["boundary"],
                                  func=SRC2 if phys == "poisson" else None, matrix_free=True)
        scale = 1.0 if phys == "poisson" else 1e-3
    Uu = scale * np.random.default_rng(17).uniform(-1, 1, asm.sizes()[2])
    F.assemble_scalar(asm, F.energy, Uu, p)
    vals = F.scalar_values(asm)
    oasm.assemble_scalar(Uu)
    assert list(vals) == list(mesh.element_block_names)
    for name, ref in zip(mesh.element_block_names, oasm.scalar_quadrature_storage):
        assert vals[name].shape == ref.shape
        assert rel_err(vals[name], ref) < RTOL, name
    asm.close()


def test_assemble_scalar_without_energy_is_an_error(F):
    mesh = F.KuhnTet10Mesh(2)
    asm, p, oasm = build_pair(F, mesh, "j2", np.array([1e3, 10e9, 1e9, 2e8, 1e8]), condensed=False, matrix_type="csr",
                              bc_nodes_1based=mesh.nodeset_nodes["bottom"], matrix_free=True)
    with pytest.raises(F.FECError, match="no energy"):
        F.assemble_scalar(asm, F.energy, np.zeros(asm.sizes()[2]), p)
    asm.close()


@pytest.mark.parametrize("script,args", [("neohookean_cube.py", ["6"]), ("poisson_periodic.py", ["32"]), ("cantilever_gravity.py", ["3"])])
def test_examples_run(script, args):
    """examples/ (the shape of the reference's examples/mechanics/Cube.jl and examples/poisson/periodic_bc.jl)"""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "examples", script), *args], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "Newton iterations" in r.stdout


# ---- external loads: assemble_vector_source! / assemble_vector_neumann_bc! -------------------

def _oracle_side_nodes(mesh, name):
    return np.asarray(mesh.sideset_side_nodes[name], dtype=np.int64)


_LOAD_CASES = {
    # name: (mesh factory, physics, props, side sets with a load, perturbation)
    "hex8_neo": (lambda F: F.StructuredMesh("hex", (0, 0, 0), (1, 2, 1.5), (5, 4, 6)), "neo", np.array([1e3, 10e6, 1e6]), ["top", "right"], 0.04),
    "quad4_poisson": (lambda F: F.StructuredMesh("quad", (0, 0), (2, 1), (9, 7)), "poisson", None, ["right", "top"], 0.02),
    "tri3_poisson": (lambda F: F.StructuredMesh("tri", (0, 0), (1, 1), (8, 6)), "poisson", None, ["bottom", "left"], 0.02),
    "tet10_linear": (lambda F: F.KuhnTet10Mesh(3), "linear", np.array([1e3, 1e10, 1e9]), ["top", "front"], 0.01),
}


@pytest.mark.parametrize("case", list(_LOAD_CASES))
def test_neumann_and_source_vectors(F, case):
    """assemble_vector! + assemble_vector_source! + assemble_vector_neumann_bc! (the sequence of src/Solvers.jl:133-137)
    against the oracle on perturbed meshes: the loads ADD to the residual storage (Source.jl:1-5), follow the
    time-dependent functions after update_bc_values!, and two calls add twice."""
    make, phys, props, ssets, amp = _LOAD_CASES[case]
    mesh = perturb(make(F), amp)
    nd = mesh.num_dimensions()
    nf = 1 if phys == "poisson" else nd
    flux = lambda X, t: (1.0 + t) * np.stack([np.sin(2.0 * X[:, 0] + d) + X[:, -1] for d in range(nf)], axis=1)
    body = lambda X, t: np.stack([(d + 1.0) * np.cos(X[:, 0]) * (1.0 + X[:, 1]) - 3.0 * t for d in range(nf)], axis=1)
    V = F.FunctionSpace(mesh, F.H1Field, F.Lagrange)
    u = F.ScalarFunction(V, "u") if nf == 1 else F.VectorFunction(V, "displ")
    asm = F.SparseMatrixAssembler(u, sparse_matrix_type="csr", use_condensed=False)
    bc_nodes = mesh.nodeset_nodes["bottom" if "bottom" not in ssets else "right"]
    mesh.nodeset_nodes["__fix__"] = bc_nodes
    dbcs = [F.DirichletBC(c, lambda X, t: np.zeros(X.shape[0]), nodeset_name="__fix__") for c in u.names()]
    nbcs = [F.NeumannBC(u.names()[0], flux, s) for s in ssets]
    srcs = [F.Source(u.names()[0], body, "block_1")]
    p = F.create_parameters(mesh, asm, product_physics(F, phys, nd), props, dirichlet_bcs=dbcs, neumann_bcs=nbcs,
                            sources=srcs, times=F.TimeStepper(0.0, 1.0, 4))
    # oracle twin
    t_el = mesh.element_types["block_1"]
    rule = {"QUAD4": "gauss2", "HEX8": "gauss2", "TRI3": "tri3", "TETRA10": "tet4"}[t_el]
    ophys = {"poisson": O.Poisson(lambda X: np.zeros(len(X))), "neo": O.NeoHookean(3), "linear": O.LinearElastic(3)}[phys]
    X = np.asarray(mesh.nodal_coords)
    blk = O.Block(mesh.element_conns["block_1"], O.ref_fe_tables(t_el, rule), ophys, props=props if props is not None else ())
    oasm = O.OracleAssembler(X, [blk], nf, condensed=False, matrix_type="csr")
    oasm.update_dofs(p.dirichlet_bcs.dirichlet_dofs())
    tabs = O.surface_tables(t_el, "gauss2")
    rng = np.random.default_rng(3)
    Uu = 1e-3 * rng.standard_normal(asm.sizes()[2])
    for step in range(2):
        t = p.times.time_current
        oasm.neumann, oasm.sources = [], []
        for s in ssets:
            sn = _oracle_side_nodes(mesh, s)
            Xq = O.surface_quadrature_points(sn, tabs, X)                      # (nqs, nsides, ND)
            v = flux(Xq.reshape(-1, nd), t).reshape(Xq.shape[0], Xq.shape[1], nf).transpose(2, 0, 1)
            oasm.add_neumann_bc(sn, tabs, v)
        Xq = O.cell_quadrature_points(blk, X)
        oasm.add_source(0, body(Xq.reshape(-1, nd), t).reshape(Xq.shape[0], Xq.shape[1], nf).transpose(2, 0, 1))
        oasm.assemble_vector(Uu)
        R_int = oasm.residual_storage.copy()
        oasm.assemble_vector_source()
        oasm.assemble_vector_neumann_bc()
        F.assemble_vector(asm, F.residual, Uu, p)
        assert rel_err(F.full_field(asm, "residual"), R_int) < RTOL
        F.assemble_vector_source(asm, Uu, p)
        F.assemble_vector_neumann_bc(asm, Uu, p)
        R = F.full_field(asm, "residual")
        assert rel_err(R, oasm.residual_storage) < RTOL, rel_err(R, oasm.residual_storage)
        loads = oasm.residual_storage - R_int
        amp = max(1.0, np.abs(R_int).max() / np.abs(loads).max())               # cancellation of the subtraction
        assert rel_err(R - R_int, loads) < 4 * RTOL * amp                       # the loads alone
        assert rel_err(F.residual(asm), oasm.residual()) < RTOL
        F.assemble_vector_neumann_bc(asm, Uu, p)                                # adds again, never zeroes
        neumann_only = loads - O.assemble_vector_source(np.zeros_like(R_int), blk, oasm.sources[0][1], X, nf)
        assert rel_err(F.full_field(asm, "residual") - R, neumann_only) < 4 * RTOL * max(1.0, np.abs(R).max() / np.abs(neumann_only).max())
        F.update_time(p)
        F.update_bc_values(p)                                                   # re-evaluates flux / body at the new time
    asm.close()


@pytest.mark.parametrize("el", ["quad", "tri", "hex"])
def test_laplace_neumann_known_answer(F, el):
    """The reference's regression (test/laplace_with_source/TestLaplace.jl:429-545, test/poisson/TestPoisson.jl:605-721):
    Laplace, u = 0 on `left`, NeumannBC g = -1 on `right`, Newton + CG through the device solver -> u = x:
    maximum(p.field) ~ 1, minimum ~ 0 (atol 1e-6).  `hex` is the 3-D analogue (not in the reference: its hex8 side
    sets are unfinished)."""
    if el == "hex":
        mesh = F.StructuredMesh("hex", (0., 0., 0.), (1., 1., 1.), (7, 5, 6))
    else:
        mesh = F.StructuredMesh(el, (0., 0.), (1., 1.), (11, 11))
    nd = mesh.num_dimensions()
    for condensed in (False, True):
        V = F.FunctionSpace(mesh, F.H1Field, F.Lagrange)
        u = F.ScalarFunction(V, "u")
        asm = F.SparseMatrixAssembler(u, sparse_matrix_type="csc", use_condensed=condensed)
        dbcs = [F.DirichletBC("u", lambda X, t: np.zeros(X.shape[0]), sideset_name="left")]
        nbcs = [F.NeumannBC("u", lambda X, t: -np.ones((X.shape[0], 1)), "right")]
        p = F.create_parameters(mesh, asm, F.Poisson(None), None, dirichlet_bcs=dbcs, neumann_bcs=nbcs)
        solver = F.NewtonSolver(F.IterativeLinearSolver(asm, "cg"))
        F.QuasiStaticIntegrator(solver).evolve(p)
        U = p.field.data_flat
        assert abs(U.max() - 1.0) < 1e-6 and abs(U.min()) < 1e-6
        assert np.abs(U - np.asarray(mesh.nodal_coords)[0]).max() < 1e-6
        asm.close()


def test_source_reproduces_the_laplace_gold(F):
    """test/laplace_with_source/TestLaplace.jl:24-76 on the device: Laplace physics + Source("u", f, "block_1"),
    Newton + CG, against laplace.gold (== poisson.gold)."""
    mesh = F.UnstructuredMesh(os.path.join(GOLDEN, "poisson_g.npz"))
    gold = np.load(os.path.join(GOLDEN, "poisson_g.npz"))["gold_u"]
    V = F.FunctionSpace(mesh, F.H1Field, F.Lagrange)
    u = F.ScalarFunction(V, "u")
    asm = F.SparseMatrixAssembler(u, sparse_matrix_type="csc", use_condensed=False)
    dbcs = [F.DirichletBC("u", lambda X, t: np.zeros(X.shape[0]), sideset_name=f"sset_{i}") for i in (1, 2, 3, 4)]
    p = F.create_parameters(mesh, asm, F.Poisson(None), None, dirichlet_bcs=dbcs,
                            sources=[F.Source("u", lambda X, t: SRC2(X), "block_1")])
    solver = F.NewtonSolver(F.IterativeLinearSolver(asm, "cg"))
    F.QuasiStaticIntegrator(solver).evolve(p)
    assert np.abs(p.field.data_flat - gold).max() < 1e-6
    asm.close()


def test_newton_with_traction_and_gravity_matches_oracle(F):
    """neo-Hookean block clamped at the bottom, traction on `top`, gravity body force: the device Newton applies the
    external loads after every residual assembly like solve! (src/Solvers.jl:133-137) -- same iteration count and
    solution as the oracle's Newton."""
    mesh = F.StructuredMesh("hex", (0, 0, 0), (1, 1, 1), (5, 5, 5))
    props = np.array([1e3, 10e6, 1e6])
    V = F.FunctionSpace(mesh, F.H1Field, F.Lagrange)
    u = F.VectorFunction(V, "displ")
    asm = F.SparseMatrixAssembler(u, sparse_matrix_type="csr", use_condensed=False)
    zero = lambda X, t: np.zeros(X.shape[0])
    dbcs = [F.DirichletBC(c, zero, nodeset_name="bottom") for c in u.names()]
    trac = lambda X, t: np.tile(np.array([[2.0e4, -5.0e4, 1.0e4]]), (X.shape[0], 1))   # added as +int g N
    grav = lambda X, t: np.tile(np.array([[0.0, -9.81e3, 0.0]]), (X.shape[0], 1))
    p = F.create_parameters(mesh, asm, F.NeoHookean(F.ThreeDimensional()), props, dirichlet_bcs=dbcs,
                            neumann_bcs=[F.NeumannBC("displ_x", trac, "top")], sources=[F.Source("displ_x", grav, "block_1")])
    solver = F.NewtonSolver(F.IterativeLinearSolver(asm, "cg"))
    integ = F.QuasiStaticIntegrator(solver)
    integ.evolve(p)
    X = np.asarray(mesh.nodal_coords)
    blk = O.Block(mesh.element_conns["block_1"], O.ref_fe_tables("HEX8", "gauss2"), O.NeoHookean(3), props=props)
    oasm = O.OracleAssembler(X, [blk], 3, condensed=False, matrix_type="csr")
    oasm.update_dofs(p.dirichlet_bcs.dirichlet_dofs())
    tabs = O.surface_tables("HEX8", "gauss2")
    sn = _oracle_side_nodes(mesh, "top")
    oasm.add_neumann_bc(sn, tabs, np.broadcast_to(np.array([2.0e4, -5.0e4, 1.0e4])[:, None, None], (3, 4, sn.shape[1])))
    oasm.add_source(0, np.broadcast_to(np.array([0.0, -9.81e3, 0.0])[:, None, None], (3, 8, blk.conn.shape[1])))
    oUu, nits, _, _ = O.newton_solve(oasm, oasm.create_unknowns())
    assert np.abs(oUu).max() > 1e-3                         # the loads do deform the block
    assert solver.iterations == nits, (solver.iterations, nits)
    assert rel_err(integ.solution, oUu) < 1e-8
    asm.close()


def test_external_load_error_behaviour(F):
    """error paths of the load entry points: unknown variable (the reference's `_dof_index_from_var_name` throws),
    values never pushed, ids out of order, and no-ops when nothing is registered (Source.jl:21 returns early)."""
    from fecb200 import _lib
    from fecb200._lib import lib
    mesh = F.StructuredMesh("quad", (0., 0.), (1., 1.), (4, 4))
    V = F.FunctionSpace(mesh, F.H1Field, F.Lagrange)
    u = F.ScalarFunction(V, "u")
    asm = F.SparseMatrixAssembler(u, sparse_matrix_type="csr")
    with pytest.raises(ValueError):
        F.create_parameters(mesh, asm, F.Poisson(None), None, neumann_bcs=[F.NeumannBC("nope", lambda X, t: X[:, :1], "top")])
    with pytest.raises(KeyError):
        F.create_parameters(mesh, asm, F.Poisson(None), None, sources=[F.Source("u", lambda X, t: X[:, :1], "block_9")])
    p = F.create_parameters(mesh, asm, F.Poisson(None), None)
    Uu = np.zeros(asm.sizes()[2])
    F.assemble_vector(asm, F.residual, Uu, p)
    F.assemble_vector_neumann_bc(asm, Uu, p)     # nothing registered: no-ops
    F.assemble_vector_source(asm, Uu, p)
    assert np.abs(F.full_field(asm, "residual")).max() == 0.0
    h = asm._require()
    sn, snp = _lib.i64(np.array([1, 2], dtype=np.int64))
    t, tp = _lib.f64(np.array([0.5, 0.5]))
    assert lib.fecb200_set_neumann_bc(h, 3, 1, 2, 1, snp, tp, tp, tp) != 0                 # ids must be 0, 1, 2, ...
    assert lib.fecb200_set_neumann_values(h, 0, None) != 0                                   # unknown id
    assert lib.fecb200_set_neumann_bc(h, 0, 1, 2, 1, snp, tp, tp, tp) == 0
    assert lib.fecb200_assemble_vector_neumann_bc(h) != 0                                    # values never set
    assert b"values were never set" in lib.fecb200_last_error()
    bad, badp = _lib.i64(np.array([1, 99], dtype=np.int64))
    assert lib.fecb200_set_neumann_bc(h, 0, 1, 2, 1, badp, tp, tp, tp) != 0                  # node id out of range
    assert lib.fecb200_clear_neumann_bcs(h) == 0
    assert lib.fecb200_assemble_vector_neumann_bc(h) == 0
    assert lib.fecb200_set_source_values(h, 5, None) != 0                                    # bad block index
    asm.close()


def test_direct_linear_solver_reproduces_the_gold_and_the_neumann_answer(F):
    """NewtonSolver(DirectLinearSolver(asm)) (src/Solvers.jl:37-86, 193-220; the reference's regression tests run both
    linear solvers): device assembly + host `K \\ R`.  poisson.gold to 1e-12 (the direct solve has no Krylov tolerance),
    and the Laplace / Neumann known answer u = x."""
    mesh = F.UnstructuredMesh(os.path.join(GOLDEN, "poisson_g.npz"))
    gold = np.load(os.path.join(GOLDEN, "poisson_g.npz"))["gold_u"]
    V = F.FunctionSpace(mesh, F.H1Field, F.Lagrange)
    u = F.ScalarFunction(V, "u")
    asm = F.SparseMatrixAssembler(u, sparse_matrix_type="csc", use_condensed=False)
    dbcs = [F.DirichletBC("u", lambda X, t: np.zeros(X.shape[0]), sideset_name=f"sset_{i}") for i in (1, 2, 3, 4)]
    p = F.create_parameters(mesh, asm, F.Poisson(lambda X, t: SRC2(X)), None, dirichlet_bcs=dbcs)
    solver = F.NewtonSolver(F.DirectLinearSolver(asm))
    F.QuasiStaticIntegrator(solver).evolve(p)
    assert np.abs(p.field.data_flat - gold).max() < 1e-12
    assert solver.iterations <= 3
    asm.close()
    m2 = F.StructuredMesh("quad", (0., 0.), (1., 1.), (11, 11))
    u2 = F.ScalarFunction(F.FunctionSpace(m2, F.H1Field, F.Lagrange), "u")
    asm2 = F.SparseMatrixAssembler(u2, sparse_matrix_type="csc", use_condensed=True)
    p2 = F.create_parameters(m2, asm2, F.Poisson(None), None,
                             dirichlet_bcs=[F.DirichletBC("u", lambda X, t: np.zeros(X.shape[0]), sideset_name="left")],
                             neumann_bcs=[F.NeumannBC("u", lambda X, t: -np.ones((X.shape[0], 1)), "right")])
    s2 = F.NewtonSolver(F.DirectLinearSolver(asm2))
    F.QuasiStaticIntegrator(s2).evolve(p2)
    assert np.abs(p2.field.data_flat - np.asarray(m2.nodal_coords)[0]).max() < 1e-9
    asm2.close()


# ---- Robin BCs (src/assemblers/WeaklyEnforcedBCs.jl:17-32, 85-180; src/bcs/RobinBCs.jl:72-86) ----
@pytest.mark.parametrize("case", ["poisson_quad4_csr", "poisson_quad4_csc", "linear_hex8_csr", "linear_hex8_csc", "poisson_tri3_csr"])
def test_robin_bc_vector_and_matrix_vs_oracle(F, case):
    """assemble_vector_robin_bc! / assemble_matrix_robin_bc! against the oracle's restatement of the reference loops.
    The hex8 case uses a NON-symmetric dvalsdu, so the transposed COO labelling of the pattern is visible."""
    el, mt = case.split("_")[1], case.split("_")[2]
    rng = np.random.default_rng(5)
    if el == "hex8":
        mesh = perturb(F.StructuredMesh("hex", (0, 0, 0), (1, 1, 1), (5, 4, 6)), 0.03)
        phys, props, nf, sset, fixed = "linear", np.array([1e3, 10e6, 1e6]), 3, "top", "bottom"
        Dm = np.array([[2.0, 0.7, -0.3], [0.1, 1.5, 0.4], [-0.6, 0.2, 3.0]])
        g0 = lambda X: np.stack([np.sin(X[:, 0]), X[:, 1] * X[:, 2], 1.0 + X[:, 0]], axis=1)
    else:
        mesh = perturb(F.StructuredMesh("quad" if el == "quad4" else "tri", (0, 0), (1, 1), (7, 6)), 0.02)
        phys, props, nf, sset, fixed = "poisson", None, 1, "right", "left"
        Dm = np.array([[1.3]])
        g0 = lambda X: (0.5 + np.sin(3 * X[:, 1]))[:, None]
    V = F.FunctionSpace(mesh, F.H1Field, F.Lagrange, q_type="GaussLegendre", q_degree=2)
    u = F.ScalarFunction(V, "u") if nf == 1 else F.VectorFunction(V, "displ")
    asm = F.SparseMatrixAssembler(u, sparse_matrix_type=mt, use_condensed=False)
    dbcs = [F.DirichletBC(c, lambda X, t: np.full(X.shape[0], 0.01), nodeset_name=fixed) for c in u.names()]
    func = lambda X, t, uu: g0(X) + uu @ Dm.T
    rbcs = [F.RobinBC(u.names()[0], func, sset)]
    p = F.create_parameters(mesh, asm, product_physics(F, phys, mesh.num_dimensions()), props, dirichlet_bcs=dbcs, robin_bcs=rbcs)
    N = asm.sizes()[2]
    Uu = 0.05 * rng.standard_normal(N)
    F.assemble_vector(asm, F.residual, Uu, p)
    F.assemble_vector_robin_bc(asm, Uu, p)
    R = F.residual(asm).copy()
    F.assemble_stiffness(asm, F.stiffness, Uu, p)
    F.assemble_matrix_robin_bc(asm, Uu, p)
    K = F.stiffness(asm)
    # ---- oracle twin
    bname = mesh.element_block_names[0]
    et = mesh.element_types[bname]
    from util_parity import _ORACLE_PHYS, _RULES
    blk = O.Block(mesh.element_conns[bname], O.ref_fe_tables(et, _RULES[et]), _ORACLE_PHYS[phys](None, mesh.num_dimensions()),
                  props=props if props is not None else ())
    X = np.asarray(mesh.nodal_coords)
    oasm = O.OracleAssembler(X, [blk], nf, condensed=False, matrix_type=mt)
    oasm.update_dofs(p.dirichlet_bcs.dirichlet_dofs())
    oasm.bc_vals[:] = 0.01
    snodes = np.asarray(mesh.sideset_side_nodes[sset])
    stabs = O.surface_tables(et, "gauss2")
    oasm.assemble_vector(Uu)
    vals, dvals = O.robin_update_bc_values(snodes, stabs, X, oasm._U(), lambda x, t, uu: g0(x[None, :])[0] + Dm @ uu,
                                           lambda x, t, uu: Dm)
    O.assemble_vector_neumann_bc(oasm.residual_storage, snodes, stabs, vals, X, nf)
    oasm.assemble_stiffness(Uu)
    O.assemble_matrix_robin_bc(oasm.stiffness_storage, mesh.element_conns[bname], np.asarray(mesh.sideset_elems[sset]), snodes,
                               stabs, dvals, X, nf)
    assert rel_err(R, oasm.residual()) < RTOL
    _check_pattern_and_values(F, asm, oasm, K)
    # the Robin term really is in there (and, for hex8, non-symmetric)
    F.assemble_stiffness(asm, F.stiffness, Uu, p)
    K0 = F.stiffness(asm)
    assert abs(K - K0).max() > 1e-3
    if el == "hex8":
        assert abs((K - K0) - (K - K0).T).max() > 1e-3
    asm.close()


def test_robin_known_answer(F):
    """-lap u = f on the unit square, u = exp(x) sin(pi y), Robin data du/dn + alpha u = r on all four sides (the set-up of
    the reference's disabled regression, test/poisson/TestPoisson.jl:106-127, with the sign the assembler's convention
    asks for: the flux handed over is g = -du/dn = alpha u - r, cf. TestLaplace.jl:438-440).  Direct solve as in
    solve!(::DirectLinearSolver) (src/Solvers.jl:64-86); second-order convergence to the exact solution."""
    alpha = 1.0
    uex = lambda X: np.exp(X[:, 0]) * np.sin(np.pi * X[:, 1])
    src = lambda X, t: (np.pi ** 2 - 1.0) * uex(X)
    dudn = {"left": lambda X: -uex(X), "right": lambda X: uex(X),
            "bottom": lambda X: -np.pi * np.exp(X[:, 0]) * np.cos(np.pi * X[:, 1]),
            "top": lambda X: np.pi * np.exp(X[:, 0]) * np.cos(np.pi * X[:, 1])}
    errs = []
    for n in (16, 32):
        mesh = F.StructuredMesh("quad", (0, 0), (1, 1), (n + 1, n + 1))
        V = F.FunctionSpace(mesh, F.H1Field, F.Lagrange)
        u = F.ScalarFunction(V, "u")
        asm = F.SparseMatrixAssembler(u, sparse_matrix_type="csc", use_condensed=False)
        rbcs = [F.RobinBC("u", (lambda X, t, uu, s=s: alpha * uu - (dudn[s](X) + alpha * uex(X))[:, None]), s)
                for s in ("left", "right", "bottom", "top")]
        p = F.create_parameters(mesh, asm, F.Poisson(src), None, dirichlet_bcs=[], robin_bcs=rbcs)
        solver = F.NewtonSolver(F.DirectLinearSolver(asm))
        F.QuasiStaticIntegrator(solver).evolve(p)
        U = np.asarray(p.field).reshape(-1)
        errs.append(np.abs(U - uex(np.asarray(mesh.nodal_coords).T)).max())
        asm.close()
    assert errs[1] < 2e-3 and errs[0] / errs[1] > 3.0, errs


@pytest.mark.parametrize("el", ["hex", "quad"])
@pytest.mark.parametrize("matrix_type", ["csr", "csc"])
@pytest.mark.parametrize("condensed", [False, True])
def test_nonsymmetric_tangent_shows_the_transposed_coo_convention(F, el, matrix_type, condensed):
    """SURVEY a-7 / B2: the reference writes K_el column-major into COO slots labelled (i outer, j inner), i.e. the
    assembled matrix holds K_el TRANSPOSED (Assemblers.jl:109-124 vs SparsityPatterns.jl:76-83).  Invisible for the
    shipped (symmetric) laws; the test law has A_ijkl != A_klij.  Values and pattern must match the oracle's literal
    restatement, the matrix must NOT be symmetric, and K^T v must equal the (true) element-wise action."""
    if el == "hex":
        mesh = perturb(F.StructuredMesh("hex", (0, 0, 0), (1, 1, 1), (5, 4, 4)), 0.03)
        fixed = mesh.nodeset_nodes["bottom"]
    else:
        mesh = perturb(F.StructuredMesh("quad", (0, 0), (1, 1), (7, 6)), 0.02)
        fixed = mesh.nodeset_nodes["left"]
    props = np.array([1e3, 10e6, 1e6, 3e6])
    asm, p, oasm = build_pair(F, mesh, "nonsym", props, condensed=condensed, matrix_type=matrix_type,
                              bc_nodes_1based=fixed, bc_value=0.01)
    rng = np.random.default_rng(12)
    N = asm.sizes()[2]
    Uu, Vu = 0.01 * rng.standard_normal(N), rng.random(N)
    F.assemble_vector(asm, F.residual, Uu, p)
    oasm.assemble_vector(Uu)
    assert rel_err(F.residual(asm), oasm.residual()) < RTOL
    F.assemble_stiffness(asm, F.stiffness, Uu, p)
    K = F.stiffness(asm)
    oasm.assemble_stiffness(Uu)
    _check_pattern_and_values(F, asm, oasm, K)
    if not condensed:   # (condensed: the 1e6 tr(K)/n penalty on the constrained diagonal dwarfs every other entry)
        assert abs(K - K.T).max() > 1e-3 * abs(K).max()
    F.assemble_matrix_action(asm, F.stiffness, Uu, Vu, p)
    Kv = F.hvp(asm, Vu).copy()
    oasm.assemble_matrix_action(Uu, Vu)
    assert rel_err(Kv, oasm.hvp(Vu)) < RTOL
    if not condensed:
        assert rel_err(K.T @ Vu, Kv) < 1e-11          # stored matrix = (dR/dU)^T
        assert rel_err(K @ Vu, Kv) > 1e-3
    asm.close()
