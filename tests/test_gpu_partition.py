"""GPU test of the rank-local (partitioned) assembly: all ranks are emulated one after another on a single
GPU, the halo exchange is routed by hand between their device buffers.  Checks, against ONE serial handle on
the global mesh: owned residual entries after the ghost->owner sum, and every owned Jacobian row (assembled
locally with the halo-element block, no exchange) -- values to 1e-12, structure exactly."""
import numpy as np
import pytest

from util_parity import rel_err

pytestmark = pytest.mark.gpu
PROPS = np.array([1e3, 10e6, 1e6])


@pytest.fixture(scope="module")
def F():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import fecb200
    return fecb200


def _problem(F, mesh, matrix_free=False):
    V = F.FunctionSpace(mesh, F.H1Field, F.Lagrange)
    u = F.VectorFunction(V, "displ")
    asm = F.SparseMatrixAssembler(u, sparse_matrix_type="csr", matrix_free=matrix_free)
    zero = lambda X, t: np.zeros(X.shape[0])
    dbcs = [F.DirichletBC(c, zero, nodeset_name="bottom") for c in u.names()] + \
           [F.DirichletBC("displ_y", lambda X, t: np.full(X.shape[0], 0.05), nodeset_name="top")]
    p = F.create_parameters(mesh, asm, F.NeoHookean(F.ThreeDimensional()), PROPS, dirichlet_bcs=dbcs)
    return asm, p


def _field(X):
    return 0.02 * np.stack([np.sin(2 * np.pi * X[1]), np.sin(2 * np.pi * X[2]), np.sin(2 * np.pi * X[0])])


@pytest.mark.parametrize("nparts,kind", [(2, "metis"), (4, "metis"), (4, "bricks")])
def test_partitioned_assembly_matches_serial(F, nparts, kind):
    import torch
    from fecb200 import _lib
    from fecb200._lib import check, lib
    n = 4
    if kind == "bricks":
        grid = (2, 2, 1)
        gmesh = F.StructuredMesh("hex", (0, 0, 0), (2, 2, 1), (2 * n + 1, 2 * n + 1, n + 1))
        locals_ = [F.structured_brick_partition(F, n, grid, r) for r in range(nparts)]
    else:
        gmesh = F.StructuredMesh("hex", (0, 0, 0), (1, 1, 1), (2 * n + 1, n + 3, n + 1))
        ep = F.metis_partition_elements(gmesh, nparts)
        locals_ = [F.partition_mesh(gmesh, ep, nparts, r) for r in range(nparts)]
    # ---- serial reference on the global mesh
    gasm, gp = _problem(F, gmesh)
    Xg = np.asarray(gmesh.nodal_coords)
    Ug = _field(Xg).reshape(-1, order="F")
    Uug = Ug[gasm.dof.unknown_dofs - 1]
    F.assemble_vector(gasm, F.residual, Uug, gp)
    Rg = F.full_field(gasm, "residual").reshape(-1, 3)
    F.assemble_stiffness(gasm, F.stiffness, Uug, gp)
    Kg = F.stiffness(gasm).tocsr()
    # ---- ranks
    ranks = []
    for r, (lm, part) in enumerate(locals_):
        asm, p = _problem(F, lm)
        part.attach(asm)
        l2g_dof = (3 * (part.local_to_global[:, None] - 1) + np.arange(3)[None, :]).reshape(-1)   # 0-based global dof per local dof
        Uu = Ug[l2g_dof][asm.dof.unknown_dofs - 1]
        F.assemble_vector(asm, F.residual, Uu, p)
        send = torch.zeros(max(1, sum(part._send_counts)), dtype=torch.float64, device="cuda")
        check(lib.fecb200_halo_pack(asm._require(), _lib.FIELD_RESIDUAL, _lib.ptr(send)))
        torch.cuda.synchronize()
        ranks.append(dict(asm=asm, p=p, part=part, Uu=Uu, send=send, l2g_dof=l2g_dof))
    # route: segment of rank a addressed to b  ->  segment of rank b coming from a
    for b, rb in enumerate(ranks):
        pb = rb["part"]
        recv = torch.zeros(max(1, sum(pb._recv_counts)), dtype=torch.float64, device="cuda")
        off = 0
        for nb, cnt in zip(pb.neighbors, pb._recv_counts):
            if cnt:
                pa = ranks[nb]["part"]
                so = sum(c for r_, c in zip(pa.neighbors, pa._send_counts) if r_ < b)
                assert pa._send_counts[pa.neighbors.index(b)] == cnt
                recv[off:off + cnt] = ranks[nb]["send"][so:so + cnt]
            off += cnt
        check(lib.fecb200_halo_unpack_add(rb["asm"]._require(), _lib.FIELD_RESIDUAL, _lib.ptr(recv)))
        torch.cuda.synchronize()
    tot_rows = 0
    for r, rk in enumerate(ranks):
        asm, part = rk["asm"], rk["part"]
        R = F.full_field(asm, "residual").reshape(-1, 3)
        own = slice(0, part.n_owned_nodes)
        assert rel_err(R[own], Rg[part.local_to_global[own] - 1]) < 1e-12
        # owned Jacobian rows, assembled locally
        F.assemble_stiffness(asm, F.stiffness, rk["Uu"], rk["p"])
        K = F.stiffness(asm).tocsr()
        owned_unknown_rows = [d for d in asm.dof.unknown_dofs if (d - 1) // 3 < part.n_owned_nodes]
        assert K.shape == (len(owned_unknown_rows), len(asm.dof.unknown_dofs))
        tot_rows += K.shape[0]
        # local unknown column -> global unknown column
        col_l2g = gasm.dof.dof_to_unknown[rk["l2g_dof"][asm.dof.unknown_dofs - 1]] - 1
        assert np.all(col_l2g >= 0)
        for i, d in enumerate(owned_unknown_rows):
            grow = gasm.dof.dof_to_unknown[rk["l2g_dof"][d - 1]] - 1
            a, b_ = K.indptr[i], K.indptr[i + 1]
            ga, gb = Kg.indptr[grow], Kg.indptr[grow + 1]
            gcols = col_l2g[K.indices[a:b_]]
            order = np.argsort(gcols)
            assert np.array_equal(gcols[order], Kg.indices[ga:gb]), (r, i)
            assert rel_err(K.data[a:b_][order], Kg.data[ga:gb]) < 1e-12
    assert tot_rows == Kg.shape[0]   # every global row is owned by exactly one rank
    for rk in ranks:
        rk["asm"].close()
    gasm.close()
