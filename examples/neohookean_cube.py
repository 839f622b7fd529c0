"""Quasi-static pull of a neo-Hookean cube -- the shape of the reference's examples/mechanics/Cube.jl
(NewtonSolver(IterativeLinearSolver(asm, :cg)) inside a QuasiStaticIntegrator, 10 load steps) on the B200 path.

    python examples/neohookean_cube.py [n]        n = elements per edge (default 24)
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "finiteelementcontainers.jl_b200"))
import fecb200 as F  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
mesh = F.StructuredMesh("hex", (0., 0., 0.), (1., 1., 1.), (n + 1, n + 1, n + 1))
V = F.FunctionSpace(mesh, F.H1Field, F.Lagrange, q_type="GaussLegendre", q_degree=2)
u = F.VectorFunction(V, "displ")
asm = F.SparseMatrixAssembler(u, sparse_matrix_type="csr", use_condensed=False)

fixed = lambda X, t: np.zeros(X.shape[0])
displace = lambda X, t: np.full(X.shape[0], 0.1 * t)
dbcs = [F.DirichletBC(c, fixed, nodeset_name="bottom") for c in u.names()]
dbcs += [F.DirichletBC("displ_x", fixed, nodeset_name="top"), F.DirichletBC("displ_z", fixed, nodeset_name="top"),
         F.DirichletBC("displ_y", displace, nodeset_name="top")]
times = F.TimeStepper(0.0, 1.0, 10)
p = F.create_parameters(mesh, asm, F.NeoHookean(F.ThreeDimensional()), np.array([1e3, 10.0e6, 1.0e6]),
                        dirichlet_bcs=dbcs, times=times)
asm.set_matrix_double_buffer(True)

solver = F.NewtonSolver(F.IterativeLinearSolver(asm, "cg"))
integrator = F.QuasiStaticIntegrator(solver)
t0 = time.time()
for step in range(1, 11):
    integrator.evolve(p)
    U = p.field.data_flat.reshape(-1, 3)
    print(f"step {step:2d}  t = {0.1 * step:.1f}  Newton iterations {solver.iterations}  |R| = {solver.residual_norm:.2e}  "
          f"max |u| = {np.abs(U).max():.4f}")
print(f"{mesh.element_conns['block_1'].shape[1]} elements, {len(asm.dof)} dofs, 10 load steps in {time.time() - t0:.2f} s")
# strain energy of the final state (assemble_scalar!)
F.assemble_scalar(asm, F.energy, integrator.solution, p)
print("strain energy:", float(F.scalar_values(asm, "block_1").sum()))
asm.close()
