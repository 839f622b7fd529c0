"""Times the vector kernels (residual, matrix-free action) of the neo-Hookean hex8 workload for one
library variant (FECB200_LIB).  Used for tile-size / register-cap sweeps; not part of the bench contract."""
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
mf = "--matrix" not in sys.argv
mesh, asm, p, Uu, _ = bench.build_problem(F, n, 0, 1, matrix_free=mf)
ne = mesh.element_conns["block_1"].shape[1]
h = asm._require()
dUu = torch.from_numpy(Uu).cuda()
dV = torch.rand_like(dUu)
torch.cuda.synchronize()
check(lib.fecb200_enable_timing(h, 1))


def kernel_ms(fn, reps=6):
    out = []
    for _ in range(reps):
        fn()
        f = C.c_float()
        check(lib.fecb200_last_kernel_ms(h, C.byref(f)))
        out.append(f.value)
    return float(np.median(out[2:]))


res = kernel_ms(lambda: F.assemble_vector(asm, F.residual, dUu, p))
act = kernel_ms(lambda: F.assemble_matrix_free_action(asm, F.stiffness_action, dUu, dV, p))
line = f"{os.path.basename(os.environ.get('FECB200_LIB', 'default')):28s} n={n} residual {res:7.3f} ms ({ne/res/1e6:7.1f} Gel/s*1e-3)  action {act:7.3f} ms"
if not mf:
    tan = kernel_ms(lambda: F.assemble_stiffness(asm, F.stiffness, dUu, p))
    line += f"  tangent {tan:7.3f} ms ({ne/tan/1e3:7.1f} Mel/s)"
print(line, flush=True)
asm.close()
