"""Kernel timings of the other BASELINE.json configurations (parity-test cases, not bench lines):
  config 2: Poisson Q1 hex8, structured 128^3: residual, CSR stiffness, matrix action
  config 4: stateful J2 plasticity, tet10 (64^3 cells x 6 Kuhn tets): residual + matrix-free action
Prints one JSON object; kernel times are CUDA-event timings taken by the library around the element kernel."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "finiteelementcontainers.jl_b200"))
import fecb200 as F  # noqa: E402
from fecb200._lib import check, lib  # noqa: E402

HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def kernel_ms(h, fn, reps=6):
    out = []
    for _ in range(reps):
        fn()
        f = C.c_float()
        check(lib.fecb200_last_kernel_ms(h, C.byref(f)))
        out.append(f.value)
    return float(np.median(out[2:]))


def entry(ne, ms, bytes_per_el):
    return {"kernel_ms": round(ms, 4), "elements_per_s": round(ne / ms * 1e3, 1),
            "GBs_algorithmic": round(bytes_per_el * ne / ms / 1e6, 1), "frac_hbm": round(bytes_per_el * ne / ms / 1e6 / HBM, 4)}


def poisson(n):
    mesh = F.StructuredMesh("hex", (0., 0., 0.), (1., 1., 1.), (n + 1,) * 3)
    V = F.FunctionSpace(mesh, F.H1Field, F.Lagrange)
    u = F.ScalarFunction(V, "u")
    asm = F.SparseMatrixAssembler(u, sparse_matrix_type="csr")
    src = lambda X, t: 3 * np.pi ** 2 * np.sin(np.pi * X[:, 0]) * np.sin(np.pi * X[:, 1]) * np.sin(np.pi * X[:, 2])
    dbcs = [F.DirichletBC("u", lambda X, t: np.zeros(X.shape[0]), nodeset_name=s) for s in ("bottom", "top", "left", "right", "back", "front")]
    t0 = time.time()
    p = F.create_parameters(mesh, asm, F.Poisson(src), None, dirichlet_bcs=dbcs)
    setup = time.time() - t0
    h = asm._require()
    asm.set_matrix_double_buffer(True)   # the zero-fill of the CSR values rides inside the stiffness kernel
    N = asm.sizes()[2]
    Uu = torch.from_numpy(np.random.default_rng(42).uniform(-1, 1, N)).cuda()
    Vu = torch.from_numpy(np.random.default_rng(7).uniform(0, 1, N)).cuda()
    check(lib.fecb200_enable_timing(h, 1))
    ne = mesh.element_conns["block_1"].shape[1]
    nnz = len(asm.pattern()[2])
    out = {"workload": f"poisson_hex8_{n}^3", "elements": ne, "dofs": len(asm.dof), "csr_nnz": nnz, "setup_s": round(setup, 1),
           "residual": entry(ne, kernel_ms(h, lambda: F.assemble_vector(asm, F.residual, Uu, p)), 168.0),
           "stiffness_csr": entry(ne, kernel_ms(h, lambda: F.assemble_stiffness(asm, F.stiffness, Uu, p)), 64 + 24 + 2 * 8.0 * nnz / ne),  # values + in-kernel clear
           "lumped_mass": entry(ne, kernel_ms(h, lambda: F.assemble_lumped_mass(asm, F.lumped_mass, Uu, p)), 64 + 24 + 8 + 8.0),
           "diagonal_stiffness": entry(ne, kernel_ms(h, lambda: F.assemble_diagonal(asm, F.stiffness, Uu, p)), 64 + 24 + 8 + 8.0),
           "matrix_action": entry(ne, kernel_ms(h, lambda: F.assemble_matrix_action(asm, F.stiffness, Uu, Vu, p)), 112.0)}
    # Newton + CG on the device (the reference's solve loop), for the record
    solver = F.NewtonSolver(F.IterativeLinearSolver(asm, "cg"))
    x = np.zeros(N)
    t0 = time.time(); solver.solve(x, p); torch.cuda.synchronize()
    out["newton"] = {"iterations": solver.iterations, "cg_iterations": solver.cg_iterations, "seconds": round(time.time() - t0, 2),
                     "residual_norm": solver.residual_norm}
    asm.close()
    return out


def j2(n):
    mesh = F.KuhnTet10Mesh(n)
    V = F.FunctionSpace(mesh, F.H1Field, F.Lagrange)
    u = F.VectorFunction(V, "displ")
    asm = F.SparseMatrixAssembler(u, sparse_matrix_type="csr", matrix_free=True)
    dbcs = [F.DirichletBC(c, lambda X, t: np.zeros(X.shape[0]), nodeset_name="bottom") for c in u.names()]
    props = np.array([1e3, 10e9, 1e9, 2e8, 1e8])
    t0 = time.time()
    p = F.create_parameters(mesh, asm, F.J2Plasticity(F.ThreeDimensional()), props, dirichlet_bcs=dbcs)
    setup = time.time() - t0
    h = asm._require()
    X = np.asarray(mesh.nodal_coords)
    U = 0.12 * np.stack([X[1] ** 2, 0.5 * X[1] * X[0], -0.3 * X[1]])   # loading that yields roughly half of the points
    Uu = torch.from_numpy(np.ascontiguousarray(U.reshape(-1, order="F")[asm.dof.unknown_dofs - 1])).cuda()
    Vu = torch.rand_like(Uu)
    check(lib.fecb200_enable_timing(h, 1))
    ne = mesh.element_conns["block_1"].shape[1]
    nq = 4
    res_ms = kernel_ms(h, lambda: F.assemble_vector(asm, F.residual, Uu, p))
    sn = p.state(which="new")
    out = {"workload": f"j2_tet10_{n}^3x6", "elements": ne, "nodes": mesh.num_nodes(), "setup_s": round(setup, 1),
           "yield_fraction": round(float(np.mean(sn[6] > 0)), 3),
           "residual": entry(ne, res_ms, 178.0 + 112.0 * nq),
           "matrix_free_action": entry(ne, kernel_ms(h, lambda: F.assemble_matrix_free_action(asm, F.stiffness_action, Uu, Vu, p)),
                                       80 + 1.365 * 96 + 56.0 * nq)}
    asm.close()
    return out


if __name__ == "__main__":
    res = {"hbm_peak_GBs": HBM, "configs": [poisson(int(os.environ.get("POISSON_N", 128))), j2(int(os.environ.get("J2_N", 64)))]}
    print(json.dumps(res), flush=True)
