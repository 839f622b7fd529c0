"""Small end-to-end case for compute-sanitizer runs (memcheck / racecheck): exercises k_vec (residual, action),
k_mat2 (tangent, fused), k_mat (mass), k_mat_scalar (Poisson), the load kernels (loads.cu), the accessors and the device CG."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "finiteelementcontainers.jl_b200"))
import fecb200 as F  # noqa: E402

n = 5
for phys in ("neo", "poisson", "j2"):
    mesh = F.StructuredMesh("hex", (0, 0, 0), (1, 1, 1), (n + 1, n + 2, n + 1)) if phys != "j2" else F.KuhnTet10Mesh(2)
    V = F.FunctionSpace(mesh, F.H1Field, F.Lagrange)
    u = F.ScalarFunction(V, "u") if phys == "poisson" else F.VectorFunction(V, "displ")
    asm = F.SparseMatrixAssembler(u, sparse_matrix_type="csr", matrix_free=(phys == "j2"))
    dbcs = [F.DirichletBC(c, lambda X, t: np.zeros(X.shape[0]), nodeset_name="bottom") for c in u.names()]
    ph = {"neo": F.NeoHookean(F.ThreeDimensional()), "poisson": F.Poisson(lambda X, t: X[:, 0]), "j2": F.J2Plasticity(F.ThreeDimensional())}[phys]
    props = {"neo": np.array([1e3, 1e7, 1e6]), "poisson": None, "j2": np.array([1e3, 1e10, 1e9, 2e8, 1e8])}[phys]
    nfld = len(u.names())
    nbcs = [F.NeumannBC(u.names()[0], lambda X, t: np.ones((X.shape[0], nfld)), "top")]
    srcs = [F.Source(u.names()[0], lambda X, t: np.tile(np.arange(1.0, nfld + 1.0), (X.shape[0], 1)), "block_1")]
    p = F.create_parameters(mesh, asm, ph, props, dirichlet_bcs=dbcs, neumann_bcs=nbcs, sources=srcs)
    if phys != "j2":
        asm.set_matrix_double_buffer(True)      # TMA zero-fill of the idle value array inside the matrix kernels
    N = asm.sizes()[2]
    rng = np.random.default_rng(0)
    Uu, Vu = 0.01 * rng.standard_normal(N), rng.random(N)
    F.assemble_vector(asm, F.residual, Uu, p)
    F.assemble_vector_source(asm, Uu, p); F.assemble_vector_neumann_bc(asm, Uu, p)   # loads.cu
    F.assemble_vector(asm, F.residual, Uu, p); R = F.residual(asm)
    F.assemble_matrix_action(asm, F.stiffness, Uu, Vu, p); Kv = F.hvp(asm, Vu)
    if phys != "j2":
        F.assemble_stiffness(asm, F.stiffness, Uu, p); K = F.stiffness(asm)
        F.assemble_vector_and_stiffness(asm, F.residual, F.stiffness, Uu, p); R2 = F.residual(asm)
        F.assemble_mass(asm, F.mass, Uu, p); M = F.mass(asm)
        x, its, rn = F.IterativeLinearSolver(asm, "cg").solve(R)
        assert np.allclose(K @ Vu, Kv, rtol=1e-9, atol=1e-9 * np.abs(Kv).max()) and np.allclose(R, R2, rtol=1e-11, atol=1e-11 * np.abs(R).max())
    F.assemble_lumped_mass(asm, F.lumped_mass, Uu, p); ml = F.lumped_mass(asm)
    F.assemble_diagonal(asm, F.stiffness, Uu, p); dk = F.diagonal(asm)
    if phys != "j2":
        F.assemble_scalar(asm, F.energy, Uu, p); en = F.scalar_values(asm)
        assert np.allclose(dk, K.diagonal(), rtol=1e-9)
        # periodic fold of the scatter connectivity (left <-> right), then back
        X = np.asarray(mesh.nodal_coords)
        left, right = mesh.nodeset_nodes["left"], mesh.nodeset_nodes["right"]
        bottom = set(mesh.nodeset_nodes["bottom"].tolist())
        a_n = np.array([n_ for n_ in left if n_ not in bottom])
        key = lambda n_: tuple(np.round(X[1:, n_ - 1] * 1e6).astype(np.int64))
        lk = {key(n_): n_ for n_ in right}
        b_n = np.array([lk[key(n_)] for n_ in a_n])
        nf = asm.dof.nf
        pa = np.concatenate([nf * (a_n - 1) + d + 1 for d in range(nf)]); pb = np.concatenate([nf * (b_n - 1) + d + 1 for d in range(nf)])
        F.update_dofs(asm, p.dirichlet_bcs, periodic=(pa, pb))
        Up = 0.01 * rng.standard_normal(asm.sizes()[2])
        for _ in range(2):
            F.assemble_vector_and_stiffness(asm, F.residual, F.stiffness, Up, p)
        Kp = F.stiffness(asm)
        assert abs(Kp - Kp.T).max() < 1e-8 * abs(Kp).max()
    print(phys, "ok", float(np.abs(R).max()), flush=True)
    asm.close()
