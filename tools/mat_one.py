"""One fused residual + tangent assembly at n^3 (default 192), kernel time from the library's events.
    FECB200_LIB=... FECB200_MAT_KERNEL=mat3 python tools/mat_one.py [n] [reps] [single]"""
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

n = int(sys.argv[1]) if len(sys.argv) > 1 else 192
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
mesh, asm, p, Uu, _ = bench.build_problem(F, n, 0, 1)
if not (len(sys.argv) > 3 and sys.argv[3] == "single"):
    asm.set_matrix_double_buffer(True)
h = asm._require()
dUu = torch.from_numpy(Uu).cuda()
check(lib.fecb200_enable_timing(h, 1))
ms = []
for _ in range(reps):
    F.assemble_vector_and_stiffness(asm, F.residual, F.stiffness, dUu, p)
    f = C.c_float()
    check(lib.fecb200_last_kernel_ms(h, C.byref(f)))
    ms.append(round(f.value, 3))
print("kernel", os.environ.get("FECB200_MAT_KERNEL", "mat2"), "n", n, "ms", ms, flush=True)
asm.close()
