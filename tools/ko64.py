import ctypes as C, os, sys
import numpy as np, torch
ROOT="/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "finiteelementcontainers.jl_b200"))
import bench, fecb200 as F
from fecb200._lib import check, lib
n=192
mesh, asm, p, Uu_h, _ = bench.build_problem(F, n, 0, 1)
h = asm._require(); asm.set_matrix_double_buffer(True)
dUu = torch.from_numpy(Uu_h).cuda(); check(lib.fecb200_enable_timing(h, 1))
for mask in (0, 64, 1, 2, 4, 24, 24 | 64, 12, 20, 28):
    os.environ["FECB200_KO"]=str(mask); ms=[]
    for i in range(5):
        F.assemble_vector_and_stiffness(asm, F.residual, F.stiffness, dUu, p)
        f=C.c_float(); check(lib.fecb200_last_kernel_ms(h, C.byref(f))); ms.append(f.value)
    print("KO", mask, round(float(np.mean(ms[2:])),3), flush=True)
