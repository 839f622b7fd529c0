"""N-GPU check of the collective plane INSIDE libfecb200 (fecb200_comm_init / fecb200_halo_sum / fecb200_comm_peer_enable,
distributed CG and Newton), driven through ctypes only: the ranks are plain spawned processes, the 128-byte ncclUniqueId
travels through a multiprocessing queue, torch.distributed is never initialised.  Everything is compared against a
SERIAL assembly / solve of the same global mesh (each rank builds it on its own GPU), not against another product path.

    python tests/run_comm_check.py [nranks=2] [n_per_rank=16] [metis|brick]

Reference model: ext/PartitionedArraysExt.jl:449-481 (assembly of PVector / PSparseMatrix), :522-540 (distributed
solve), src/Solvers.jl:128-220."""
import ctypes as C
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "finiteelementcontainers.jl_b200"))

PROPS = np.array([1e3, 10.0e6, 1.0e6])
GRID = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}


def _problem(F, mesh, device, part=None):
    V = F.FunctionSpace(mesh, F.H1Field, F.Lagrange, q_type="GaussLegendre", q_degree=2)
    u = F.VectorFunction(V, "displ")
    asm = F.SparseMatrixAssembler(u, sparse_matrix_type="csr", use_condensed=False, device=device)
    zero = lambda X, t: np.zeros(X.shape[0])
    pull = lambda X, t: np.full(X.shape[0], 0.02 * t)
    dbcs = [F.DirichletBC(c, zero, nodeset_name="bottom") for c in u.names()]
    dbcs += [F.DirichletBC("displ_x", zero, nodeset_name="top"), F.DirichletBC("displ_z", zero, nodeset_name="top"),
             F.DirichletBC("displ_y", pull, nodeset_name="top")]
    p = F.create_parameters(mesh, asm, F.NeoHookean(F.ThreeDimensional()), PROPS, dirichlet_bcs=dbcs,
                            times=F.TimeStepper(0.0, 1.0, 4))
    if part is not None:
        part.attach(asm)
    X = np.asarray(mesh.nodal_coords)
    U = 0.02 * np.stack([np.sin(2 * np.pi * X[1]), np.sin(2 * np.pi * X[2]), np.sin(2 * np.pi * X[0])])
    Uu = np.ascontiguousarray(U.reshape(-1, order="F")[asm.dof.unknown_dofs - 1])
    return asm, p, Uu


def worker(rank, world, n, how, q_id, q_out):
    try:
        _worker(rank, world, n, how, q_id, q_out)
    except BaseException:   # a rank that dies silently leaves the others blocked in NCCL: report and let the parent kill them
        import traceback
        q_out.put((rank, {"error": traceback.format_exc()}))


def _say(rank, msg):
    print(f"[rank {rank}] {msg}", file=sys.stderr, flush=True)


def _worker(rank, world, n, how, q_id, q_out):
    import torch
    torch.cuda.set_device(rank)
    import fecb200 as F
    from fecb200 import _lib
    from fecb200._lib import check, lib
    from fecb200.partition import metis_cell_partition, structured_brick_partition, structured_cell_partition
    res = {}
    g = GRID[world]
    E = tuple(gi * n for gi in g)
    if how == "brick":
        ix, iy, iz = np.meshgrid(np.arange(g[0]), np.arange(g[1]), np.arange(g[2]), indexing="ij")
        lm, part = structured_cell_partition(F, E, ix + g[0] * (iy + g[1] * iz), n, rank, h=1.0 / n)
    else:
        c = 4
        cp = metis_cell_partition(tuple(e // c for e in E), world)     # deterministic: every rank computes the same map
        lm, part = structured_cell_partition(F, E, cp, c, rank, h=1.0 / n)
    asm, p, Uu = _problem(F, lm, rank, part)
    h = asm._require()
    # ---- the only out-of-band exchange: rank 0's ncclUniqueId
    if rank == 0:
        buf = C.create_string_buffer(128)
        check(lib.fecb200_comm_unique_id(C.cast(buf, C.c_void_p)))
        for _ in range(world - 1):
            q_id.put(bytes(buf.raw))
        uid = bytes(buf.raw)
    else:
        uid = q_id.get(timeout=120)
    part.comm_init(asm, unique_id=uid)
    _say(rank, "communicator up")
    # ---- serial twin on this rank's GPU
    gmesh = F.StructuredMesh("hex", (0., 0., 0.), tuple(e / n for e in E), tuple(e + 1 for e in E))
    gasm, gp, gUu = _problem(F, gmesh, rank)
    F.assemble_vector_and_stiffness(gasm, F.residual, F.stiffness, gUu, gp)
    Rg = F.full_field(gasm, "residual").reshape(-1, 3).copy()
    l2g = part.local_to_global - 1
    own = l2g[:part.n_owned_nodes]
    nown = C.c_int64()
    check(lib.fecb200_owned_length(h, C.byref(nown)))
    nown = nown.value
    ud = asm.dof.unknown_dofs - 1
    ug_all = gasm.dof.dof_to_unknown[3 * l2g[ud // 3] + ud % 3] - 1
    assert (ug_all >= 0).all()
    ug = ug_all[:nown]
    rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))

    _say(rank, "serial twin assembled")
    # 1. NCCL halo (pack / grouped send-recv / add inside the library)
    F.assemble_vector_and_stiffness(asm, F.residual, F.stiffness, Uu, p)
    check(lib.fecb200_halo_sum(h, _lib.FIELD_RESIDUAL))
    R = F.full_field(asm, "residual").reshape(-1, 3)[:part.n_owned_nodes]
    res["R_nccl_halo_vs_serial"] = rel(R, Rg[own])
    res["residual_accessor_vs_serial"] = rel(F.residual(asm)[:nown], F.residual(gasm)[ug])
    _say(rank, "1 done")
    # 2. owned Jacobian rows: K v and K 1 through the distributed SpMV (ghost refresh inside)
    vg = np.random.default_rng(7).uniform(0, 1, gasm.sizes()[2])
    for name, xg in (("Kv", vg), ("rowsum", np.ones_like(vg))):
        xl = np.ascontiguousarray(xg[ug_all])
        xl[nown:] = -123.0        # ghost entries deliberately wrong: the library must refresh them from their owners
        yl = F.matrix_multiply(asm, xl)[:nown]
        res[name + "_vs_serial"] = rel(yl, F.matrix_multiply(gasm, xg)[ug])
    _say(rank, "2 done")
    # 3. owner -> ghost update of a nodal field
    F.update_field(p, Uu)
    check(lib.fecb200_halo_update(h, _lib.FIELD_U))
    F.update_field(gp, gUu)
    res["halo_update_U"] = rel(F.full_field(asm, "u").reshape(-1, 3), F.full_field(gasm, "u").reshape(-1, 3)[l2g])
    _say(rank, "3 done")
    # 4. fused peer-memory halo (IPC handles + ghost ids exchanged over NCCL inside the library)
    part.enable_peer_scatter(asm)
    F.residual(asm)                                   # flush: R is zero on every rank ...
    check(lib.fecb200_comm_barrier(h))                # ... before anyone scatters
    for it in range(2):
        if it == 0:
            F.assemble_vector_and_stiffness(asm, F.residual, F.stiffness, Uu, p)
        else:
            F.assemble_vector(asm, F.residual, Uu, p)
        check(lib.fecb200_halo_sum(h, _lib.FIELD_RESIDUAL))
        R = F.full_field(asm, "residual").reshape(-1, 3)[:part.n_owned_nodes].copy()
        res[f"R_peer_halo_vs_serial_{it}"] = rel(R, Rg[own])
        F.residual(asm)
        check(lib.fecb200_comm_barrier(h))
    _say(rank, "4 done")
    # 5. distributed CG on the assembled tangent: same iterates as the serial solve
    bg = np.random.default_rng(3).uniform(-1, 1, gasm.sizes()[2])
    xs, its_s, _ = F.IterativeLinearSolver(gasm, "cg").solve(bg)
    bl = np.ascontiguousarray(bg[ug_all])
    xl, its_l, _ = F.IterativeLinearSolver(asm, "cg").solve(bl)
    res["cg_iterations"] = (int(its_l), int(its_s))
    res["cg_solution_vs_serial"] = rel(xl, xs[ug_all])       # ghost entries included: refreshed at the end of the solve
    _say(rank, f"5 done {res['cg_iterations']}")
    # 6. distributed Newton load step (peer halo on): iteration counts equal the serial solve's
    for pp in (p, gp):
        F.update_time(pp); F.update_bc_values(pp)
    sol_l, sol_g = np.zeros(asm.sizes()[2]), np.zeros(gasm.sizes()[2])
    sl = F.NewtonSolver(F.IterativeLinearSolver(asm, "cg")); sl.solve(sol_l, p)
    sg = F.NewtonSolver(F.IterativeLinearSolver(gasm, "cg")); sg.solve(sol_g, gp)
    res["newton_iterations"] = (sl.iterations, sg.iterations)
    res["newton_cg_iterations"] = (sl.cg_iterations, sg.cg_iterations)
    res["newton_solution_vs_serial"] = rel(sol_l, sol_g[ug_all])
    res["stats"] = dict(owned_elements=part.n_owned_elements, neighbours=part.neighbors, ghosts=len(l2g) - part.n_owned_nodes)
    check(lib.fecb200_comm_barrier(h))
    gasm.close()
    asm.close()
    q_out.put((rank, res))


def main():
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    how = sys.argv[3] if len(sys.argv) > 3 else "metis"
    ctx = mp.get_context("spawn")
    q_id, q_out = ctx.Queue(), ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, world, n, how, q_id, q_out)) for r in range(world)]
    for pr in procs:
        pr.start()
    results, failed = {}, False
    try:
        for _ in range(world):
            r, res = q_out.get(timeout=int(os.environ.get("COMM_CHECK_TIMEOUT", 240)))
            results[r] = res
            if "error" in res:
                failed = True
                break
    except Exception as e:   # queue.Empty: a rank is stuck
        print("comm check: no result within the time limit:", repr(e), flush=True)
        failed = True
    finally:
        for pr in procs:
            pr.join(timeout=1 if failed else 60)
            if pr.is_alive():
                pr.kill()
    if failed:
        for r, res in results.items():
            if "error" in res:
                print(f"rank {r} raised:\n{res['error']}", flush=True)
        print(f"comm check ({world} ranks, {how}): FAIL", flush=True)
        sys.exit(1)
    ok = len(results) == world
    for r in sorted(results):
        res = results[r]
        print(f"rank {r}: {res}", flush=True)
        for k, v in res.items():
            if k.endswith("_vs_serial") or k.startswith("R_") or k == "halo_update_U":
                tol = 1e-8 if k.startswith(("cg_", "newton_")) else 1e-12
                ok &= bool(v < tol)
        ok &= res["newton_iterations"][0] == res["newton_iterations"][1]
        ok &= abs(res["cg_iterations"][0] - res["cg_iterations"][1]) <= 1
    print(f"comm check ({world} ranks, {how}):", "OK" if ok else "FAIL", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
