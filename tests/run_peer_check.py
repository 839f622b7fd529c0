"""2-GPU check of the fused peer-memory halo against the NCCL pack/unpack halo (run under torchrun on a
multi-GPU box: `torchrun --nproc-per-node 2 tests/run_peer_check.py`).  Both must give the owner the same
residual (to rounding) on owned nodes."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "finiteelementcontainers.jl_b200"))
import bench  # noqa: E402
import fecb200 as F  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
mesh, asm, p, Uu, part = bench.build_problem(F, n, rank, world)
from fecb200._lib import check, lib  # noqa: E402
stream = torch.cuda.Stream()
check(lib.fecb200_set_stream(asm._require(), stream.cuda_stream))   # library work and NCCL barriers share one stream
with torch.cuda.stream(stream):
    dUu = torch.from_numpy(Uu).cuda()
stream.synchronize()
# reference: NCCL halo
F.assemble_vector_and_stiffness(asm, F.residual, F.stiffness, dUu, p)
part.halo_sum_residual(asm, stream)
R_nccl = F.full_field(asm, "residual").reshape(-1, 3)[:part.n_owned_nodes].copy()
F.assemble_vector(asm, F.residual, dUu, p)
part.halo_sum_residual(asm, stream)
R_nccl2 = F.full_field(asm, "residual").reshape(-1, 3)[:part.n_owned_nodes].copy()
torch.cuda.synchronize(); dist.barrier()
# fused peer path (zero-after-read protocol: the residual accessor re-zeroes the field)
with torch.cuda.stream(stream):
    part.enable_peer_scatter(asm)
    out = torch.empty_like(dUu)
stream.synchronize()
F.residual(asm, out)                     # flush: leaves R zeroed on every rank
errs = []
for fused in (True, False, True):
    part.barrier_on_stream(stream)
    if fused:
        F.assemble_vector_and_stiffness(asm, F.residual, F.stiffness, dUu, p)
    else:
        F.assemble_vector(asm, F.residual, dUu, p)
    part.halo_sum_residual(asm, stream)
    R = F.full_field(asm, "residual").reshape(-1, 3)[:part.n_owned_nodes].copy()
    F.residual(asm, out)
    ref = R_nccl if fused else R_nccl2
    errs.append(float(np.abs(R - ref).max() / np.abs(ref).max()))
t = torch.tensor(errs, device="cuda", dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print("peer-vs-nccl max rel err per pass:", t.tolist(), "OK" if t.max().item() < 1e-12 else "FAIL", flush=True)
dist.barrier()
asm.close()
dist.destroy_process_group()
