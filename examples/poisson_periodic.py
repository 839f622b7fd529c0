"""Poisson with periodic boundary conditions in x and y on a structured quad mesh -- the reference's
examples/poisson/periodic_bc.jl / test 'test_poisson_periodic': side-b dofs are folded into their side-a unknown in
the assembled matrix, Newton + CG run on the device.

    python examples/poisson_periodic.py [n]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "finiteelementcontainers.jl_b200"))
import fecb200 as F  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
mesh = F.StructuredMesh("quad", (0., 0.), (1., 1.), (n + 1, n + 1))
X = np.asarray(mesh.nodal_coords)
f = lambda Xq, t: ((2 * np.pi) ** 2 * np.cos(2 * np.pi * Xq[:, 0]) + 0.5 * (4 * np.pi) ** 2 * np.cos(4 * np.pi * Xq[:, 1])
                   + 0.25 * ((2 * np.pi) ** 2 + (4 * np.pi) ** 2) * np.sin(2 * np.pi * Xq[:, 0]) * np.sin(4 * np.pi * Xq[:, 1]))
u_exact = np.cos(2 * np.pi * X[0]) + 0.5 * np.cos(4 * np.pi * X[1]) + 0.25 * np.sin(2 * np.pi * X[0]) * np.sin(4 * np.pi * X[1])

V = F.FunctionSpace(mesh, F.H1Field, F.Lagrange)
u = F.ScalarFunction(V, "u")
asm = F.SparseMatrixAssembler(u, sparse_matrix_type="csc", use_condensed=False)
# PeriodicBC(var, direction, func, side_a, side_b) (src/bcs/PeriodicBCs.jl:1-15): `direction` is the coordinate along which the
# two sides run and by which their nodes are matched; func is the jump U[b] = U[a] + func(X_b, t)
zero = lambda X, t: np.zeros(len(X))
pbcs = [F.PeriodicBC("u", "y", zero, "left", "right"), F.PeriodicBC("u", "x", zero, "bottom", "top")]
p = F.create_parameters(mesh, asm, F.Poisson(f), None, dirichlet_bcs=[], periodic_bcs=pbcs)
solver = F.NewtonSolver(F.IterativeLinearSolver(asm, "cg"))
F.QuasiStaticIntegrator(solver).evolve(p)
err = np.abs(p.field.data_flat - u_exact).max()
print(f"{n}x{n} QUAD4, {asm.sizes()[1]} unknowns, Newton iterations {solver.iterations}, max |u - u_exact| = {err:.3e}")
asm.close()
