"""Phase-cost experiment for k_mat2 (build with FECB200_DEFINES=-DFEC_MAT2_KO FECB200_VARIANT=ko, run with
FECB200_LIB=.../libfecb200_ko.so): times the fused residual + tangent kernel at 192^3 with one phase knocked out at a
time (results are wrong by design; only the timings mean something).

    python tools/ko_sweep.py [n]
"""
import ctypes as C
import json
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
mesh, asm, p, Uu_h, _ = bench.build_problem(F, n, 0, 1)
h = asm._require()
asm.set_matrix_double_buffer(True)
dUu = torch.from_numpy(Uu_h).cuda()
check(lib.fecb200_enable_timing(h, 1))
MASKS = [(0, "full kernel"), (32, "no zero-fill"), (1, "no RED instructions (loads kept)"), (2, "no scatter read-back, no REDs"),
         (4, "no staging, no scatter"), (8, "no phase K"), (16, "no phase G (gathers, geometry, constitutive)"),
         (4 | 8, "phase G only"), (4 | 16, "phase K only"), (8 | 16, "staging + scatter only"),
         (4 | 8 | 16, "launch + scatter-record fetch only")]
out = []
for mask, name in MASKS:
    os.environ["FECB200_KO"] = str(mask)
    ms = []
    for i in range(6):
        F.assemble_vector_and_stiffness(asm, F.residual, F.stiffness, dUu, p)
        f = C.c_float()
        check(lib.fecb200_last_kernel_ms(h, C.byref(f)))
        ms.append(f.value)
    t = float(np.mean(ms[2:]))
    out.append({"mask": mask, "what": name, "kernel_ms": round(t, 3)})
    print(f"KO {mask:3d}  {t:7.3f} ms   {name}", flush=True)
print(json.dumps(out))
