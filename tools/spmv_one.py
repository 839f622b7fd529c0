"""SpMV (y = K x on the reference-ordered CSR values, fecb200_matrix_multiply) at n^3 neo-Hookean: time and effective
HBM bandwidth (values + node adjacency + column offsets + x gather + y)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "finiteelementcontainers.jl_b200"))
import bench  # noqa: E402
import fecb200 as F  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 192
mesh, asm, p, Uu, _ = bench.build_problem(F, n, 0, 1)
dUu = torch.from_numpy(Uu).cuda()
F.assemble_stiffness(asm, F.stiffness, dUu, p)
x = torch.rand(len(Uu), dtype=torch.float64, device="cuda")
y = torch.empty_like(x)
for _ in range(3):
    F.matrix_multiply(asm, x, y)
import time
from fecb200._lib import check, lib
h = asm._require()
check(lib.fecb200_synchronize(h))          # the library runs on its own stream: time on the host around its synchronize
t0 = time.perf_counter()
for _ in range(20):
    F.matrix_multiply(asm, x, y)
check(lib.fecb200_synchronize(h))
ms = (time.perf_counter() - t0) * 1e3 / 20
nnz = len(asm.pattern()[2]) if len(sys.argv) > 2 else int(81 * len(Uu))
print(f"spmv n {n}: {ms:.3f} ms, values {8 * nnz / 1e9:.2f} GB -> {8 * nnz / ms / 1e6:.0f} GB/s (values only)", flush=True)
asm.close()
