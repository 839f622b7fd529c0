"""A neo-Hookean cantilever under its own weight and an end traction ramped over 5 load steps: body forces
(`Source`) and surface loads (`NeumannBC`) on the B200 path -- the two load terms solve! adds after assemble_vector!
(src/Solvers.jl:133-137), integrated once per load step on the device.

    python examples/cantilever_gravity.py [n]     n = elements through the thickness (default 8; length = 4 n)
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "finiteelementcontainers.jl_b200"))
import fecb200 as F  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
mesh = F.StructuredMesh("hex", (0., 0., 0.), (4., 1., 1.), (4 * n + 1, n + 1, n + 1))
V = F.FunctionSpace(mesh, F.H1Field, F.Lagrange, q_type="GaussLegendre", q_degree=2)
u = F.VectorFunction(V, "displ")
asm = F.SparseMatrixAssembler(u, sparse_matrix_type="csr", use_condensed=False)

rho, K, G = 1.0e2, 10.0e6, 1.0e6
clamp = [F.DirichletBC(c, lambda X, t: np.zeros(X.shape[0]), nodeset_name="left") for c in u.names()]
# Source: b(X, t), the assembler adds -int N b, so b = rho g is the body force itself (Sources.jl:1-5)
weight = F.Source("displ_x", lambda X, t: np.tile([[0.0, -9.81 * rho * t, 0.0]], (X.shape[0], 1)), "block_1")
# NeumannBC: the assembler adds +int N g, so g = -traction (test/laplace_with_source/TestLaplace.jl:438-440)
pull = F.NeumannBC("displ_x", lambda X, t: np.tile([[0.0, 2.0e3 * t, 0.0]], (X.shape[0], 1)), "right")
p = F.create_parameters(mesh, asm, F.NeoHookean(F.ThreeDimensional()), np.array([rho, K, G]), dirichlet_bcs=clamp,
                        neumann_bcs=[pull], sources=[weight], times=F.TimeStepper(0.0, 1.0, 5))
asm.set_matrix_double_buffer(True)
solver = F.NewtonSolver(F.IterativeLinearSolver(asm, "cg"))
integrator = F.QuasiStaticIntegrator(solver)
t0 = time.time()
for step in range(1, 6):
    integrator.evolve(p)
    U = p.field.data_flat.reshape(-1, 3)
    tip = U[mesh.nodeset_nodes["right"] - 1, 1].mean()
    print(f"step {step}  t = {0.2 * step:.1f}  Newton iterations {solver.iterations}  |R| = {solver.residual_norm:.2e}  "
          f"tip deflection = {tip:+.5f}")
print(f"{mesh.element_conns['block_1'].shape[1]} elements, {len(asm.dof)} dofs, 5 load steps in {time.time() - t0:.2f} s")
assert tip < 0.0, "the beam sags under its own weight (the end traction pushes down too)"
asm.close()
