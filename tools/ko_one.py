import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "finiteelementcontainers.jl_b200"))
import torch, bench, fecb200 as F
n = 128
mesh, asm, p, Uu_h, _ = bench.build_problem(F, n, 0, 1)
asm.set_matrix_double_buffer(True)
dUu = torch.from_numpy(Uu_h).cuda()
for _ in range(4):
    if os.environ.get("KO_ONE_MODE") == "tangent":
        F.assemble_stiffness(asm, F.stiffness, dUu, p)
    else:
        F.assemble_vector_and_stiffness(asm, F.residual, F.stiffness, dUu, p)
torch.cuda.synchronize()
