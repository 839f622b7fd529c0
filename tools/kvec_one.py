"""Driver for ncu captures of the vector kernels (residual, matrix-free action) at 192^3 neo-Hookean."""
import os
import sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "finiteelementcontainers.jl_b200"))
import torch, bench, fecb200 as F
n = int(sys.argv[1]) if len(sys.argv) > 1 else 192
mesh, asm, p, Uu_h, _ = bench.build_problem(F, n, 0, 1)
dUu = torch.from_numpy(Uu_h).cuda()
Vu = torch.rand(len(Uu_h), dtype=torch.float64, device="cuda")
for _ in range(3):
    F.assemble_vector(asm, F.residual, dUu, p)
    F.assemble_matrix_free_action(asm, F.stiffness_action, dUu, Vu, p)
torch.cuda.synchronize()
