"""A/B of the vector kernels (Walsh form vs the quadrature loop, FECB200_VEC_CLASSIC=1): neo-Hookean 192^3 residual and
matrix-free action, Poisson 128^3 residual and action.  Kernel times from the library's events."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "finiteelementcontainers.jl_b200"))
import bench  # noqa: E402
import fecb200 as F  # noqa: E402
from fecb200._lib import check, lib  # noqa: E402


def kms(h, fn, reps=5):
    xs = []
    for _ in range(reps):
        fn()
        f = C.c_float()
        check(lib.fecb200_last_kernel_ms(h, C.byref(f)))
        xs.append(f.value)
    return round(float(np.mean(xs[1:])), 3)


n = int(sys.argv[1]) if len(sys.argv) > 1 else 192
mesh, asm, p, Uu, _ = bench.build_problem(F, n, 0, 1)
h = asm._require()
dUu = torch.from_numpy(Uu).cuda()
Vu = torch.rand(len(Uu), dtype=torch.float64, device="cuda")
check(lib.fecb200_enable_timing(h, 1))
for mode in ("walsh", "classic"):
    if mode == "classic":
        os.environ["FECB200_VEC_CLASSIC"] = "1"
    else:
        os.environ.pop("FECB200_VEC_CLASSIC", None)
    print(mode, "neo", n, "residual", kms(h, lambda: F.assemble_vector(asm, F.residual, dUu, p)),
          "action", kms(h, lambda: F.assemble_matrix_free_action(asm, F.stiffness_action, dUu, Vu, p)), flush=True)
asm.close()

m = 128
mesh = F.StructuredMesh("hex", (0., 0., 0.), (1., 1., 1.), (m + 1,) * 3)
V = F.FunctionSpace(mesh, F.H1Field, F.Lagrange)
u = F.ScalarFunction(V, "u")
asm = F.SparseMatrixAssembler(u, sparse_matrix_type="csr")
src = lambda X, t: 3 * np.pi ** 2 * np.sin(np.pi * X[:, 0]) * np.sin(np.pi * X[:, 1]) * np.sin(np.pi * X[:, 2])
dbcs = [F.DirichletBC("u", lambda X, t: np.zeros(X.shape[0]), nodeset_name=s) for s in ("bottom", "top")]
p = F.create_parameters(mesh, asm, F.Poisson(src), None, dirichlet_bcs=dbcs)
h = asm._require()
N = asm.sizes()[2]
Uu = torch.from_numpy(np.random.default_rng(42).uniform(-1, 1, N)).cuda()
Vu = torch.from_numpy(np.random.default_rng(7).uniform(0, 1, N)).cuda()
check(lib.fecb200_enable_timing(h, 1))
for mode in ("walsh", "classic"):
    if mode == "classic":
        os.environ["FECB200_VEC_CLASSIC"] = "1"
    else:
        os.environ.pop("FECB200_VEC_CLASSIC", None)
    print(mode, "poisson", m, "residual", kms(h, lambda: F.assemble_vector(asm, F.residual, Uu, p)),
          "action", kms(h, lambda: F.assemble_matrix_action(asm, F.stiffness, Uu, Vu, p)), flush=True)
asm.close()
