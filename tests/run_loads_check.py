"""2-GPU check of the external loads on a partitioned problem (run under torchrun on a multi-GPU box):
residual + body force + traction of the rank-local handles, summed over the halo (NCCL path and fused peer path),
against the serial assembly of the global mesh on the same GPU -- owned nodes only."""
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
from fecb200._lib import check, lib  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dist.all_reduce(torch.zeros(1, device="cuda"))   # creates the communicator before the first point-to-point exchange
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
grid = bench.grid_for(world)
props = bench.NEO_PROPS
trac = lambda X, t: np.stack([2.0e3 * (1.0 + X[:, 0]), -5.0e3 * np.ones(len(X)), 1.0e3 * X[:, 2]], axis=1)
grav = lambda X, t: np.stack([np.zeros(len(X)), -9.81e2 * (1.0 + X[:, 1]), 50.0 * X[:, 0]], axis=1)
zero = lambda X, t: np.zeros(X.shape[0])


def build(mesh, part):
    V = F.FunctionSpace(mesh, F.H1Field, F.Lagrange)
    u = F.VectorFunction(V, "displ")
    asm = F.SparseMatrixAssembler(u, sparse_matrix_type="csr", use_condensed=False, device=local)
    dbcs = [F.DirichletBC(c, zero, nodeset_name="bottom") for c in u.names()]
    p = F.create_parameters(mesh, asm, F.NeoHookean(F.ThreeDimensional()), props, dirichlet_bcs=dbcs,
                            neumann_bcs=[F.NeumannBC("displ_x", trac, "top"), F.NeumannBC("displ_x", trac, "right")],
                            sources=[F.Source("displ_x", grav, "owned" if part is not None else "block_1")])
    if part is not None:
        part.attach(asm)
    return asm, p


# serial reference on the global mesh
gmesh = F.StructuredMesh("hex", (0., 0., 0.), tuple(float(g) for g in grid), tuple(g * n + 1 for g in grid))
gasm, gp = build(gmesh, None)
rng = np.random.default_rng(11)
Ug = 1e-3 * rng.standard_normal(gmesh.num_nodes() * 3)          # a full displacement field (Dirichlet slots overwritten)
Uu_g = Ug[gasm.dof.unknown_dofs - 1]
F.assemble_vector(gasm, F.residual, Uu_g, gp)
F.assemble_vector_source(gasm, Uu_g, gp)
F.assemble_vector_neumann_bc(gasm, Uu_g, gp)
R_serial = F.full_field(gasm, "residual").reshape(-1, 3)
gasm.close()

# partitioned
lmesh, part = F.structured_brick_partition(F, n, grid, rank)
asm, p = build(lmesh, part)
l2g = part.local_to_global - 1
Uu = Ug.reshape(-1, 3)[l2g].reshape(-1)[asm.dof.unknown_dofs - 1]
stream = torch.cuda.Stream()
check(lib.fecb200_set_stream(asm._require(), stream.cuda_stream))
with torch.cuda.stream(stream):
    dUu = torch.from_numpy(np.ascontiguousarray(Uu)).cuda()
stream.synchronize()
ref = R_serial[l2g[:part.n_owned_nodes]]
errs = []
# NCCL halo
F.assemble_vector(asm, F.residual, dUu, p)
F.assemble_vector_source(asm, dUu, p)
F.assemble_vector_neumann_bc(asm, dUu, p)
part.halo_sum_residual(asm, stream)
R = F.full_field(asm, "residual").reshape(-1, 3)[:part.n_owned_nodes]
errs.append(float(np.abs(R - ref).max() / np.abs(ref).max()))
torch.cuda.synchronize(); dist.barrier()
# fused peer halo
with torch.cuda.stream(stream):
    part.enable_peer_scatter(asm)
    out = torch.empty_like(dUu)
stream.synchronize()
F.residual(asm, out)                     # flush: leaves R zeroed on every rank
for _ in range(2):
    part.barrier_on_stream(stream)
    F.assemble_vector(asm, F.residual, dUu, p)
    F.assemble_vector_source(asm, dUu, p)
    F.assemble_vector_neumann_bc(asm, dUu, p)
    part.halo_sum_residual(asm, stream)
    R = F.full_field(asm, "residual").reshape(-1, 3)[:part.n_owned_nodes].copy()
    F.residual(asm, out)
    errs.append(float(np.abs(R - ref).max() / np.abs(ref).max()))
t = torch.tensor(errs, device="cuda", dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print("partitioned loads vs serial, max rel err (nccl, peer, peer):", t.tolist(), "OK" if t.max().item() < 1e-12 else "FAIL", flush=True)
dist.barrier()
asm.close()
dist.destroy_process_group()
