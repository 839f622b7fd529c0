"""Host-side multi-GPU logic on CPU: METIS partitioning, owned/ghost numbering, halo lists, and the
exchange itself over torch.distributed (gloo, world_size 2).  The per-rank element work is done by the
numpy oracle here; on the GPU box the same Partition drives libfecb200 over NCCL (tests/test_gpu_partition.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import fec_oracle as O
import fecb200 as F
from fecb200.partition import (exchange, partition_mesh, structured_brick_partition, metis_partition_elements, metis_partition_graph,
                               metis_cell_partition, structured_cell_partition)

PROPS = np.array([1e3, 10e6, 1e6])


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _global_field(X):
    return 0.02 * np.stack([np.sin(2 * np.pi * X[1]), np.sin(2 * np.pi * X[2]), np.sin(2 * np.pi * X[0])])


def _serial_residual(mesh):
    X = np.asarray(mesh.nodal_coords)
    blk = O.Block(mesh.element_conns[mesh.element_block_names[0]], O.ref_fe_tables("HEX8", "gauss2"), O.NeoHookean(3), props=PROPS)
    return O.assemble_vector([blk], X, _global_field(X), 3)


def _rank_residual_with_exchange(lm, part):
    """owned-element residual on the local mesh, then ghost -> owner accumulation through `exchange`"""
    X = np.asarray(lm.nodal_coords)
    blk = O.Block(lm.element_conns["owned"], O.ref_fe_tables("HEX8", "gauss2"), O.NeoHookean(3), props=PROPS)
    R = O.assemble_vector([blk], X, _global_field(X), 3).reshape(-1, 3)
    send = torch.from_numpy(np.concatenate([R[part.send[r] - 1].reshape(-1) for r in part.neighbors if r in part.send]
                                           or [np.zeros(0)]))
    sc = [3 * len(part.send.get(r, ())) for r in part.neighbors]
    rc = [3 * len(part.recv.get(r, ())) for r in part.neighbors]
    recv = torch.zeros(max(1, sum(rc)), dtype=torch.float64)
    exchange(part.neighbors, send, sc, recv, rc)
    off = 0
    for r, n in zip(part.neighbors, rc):
        if n:
            np.add.at(R, part.recv[r] - 1, recv[off:off + n].numpy().reshape(-1, 3))
        off += n
    return R


def _worker(rank, world, port, mode, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 4
        if mode == "metis":
            gmesh = F.StructuredMesh("hex", (0, 0, 0), (1, 1, 1), (n + 1, n + 1, 2 * n + 1))
            epart = metis_partition_elements(gmesh, world)
            lm, part = partition_mesh(gmesh, epart, world, rank)
        else:
            gmesh = F.StructuredMesh("hex", (0, 0, 0), (2, 1, 1), (2 * n + 1, n + 1, n + 1))
            lm, part = structured_brick_partition(F, n, (2, 1, 1), rank)
            # same global numbering and coordinates as the global StructuredMesh
            assert np.allclose(np.asarray(lm.nodal_coords), np.asarray(gmesh.nodal_coords)[:, part.local_to_global - 1])
        R = _rank_residual_with_exchange(lm, part)
        Rg = _serial_residual(gmesh).reshape(-1, 3)
        own = slice(0, part.n_owned_nodes)
        err = np.abs(R[own] - Rg[part.local_to_global[own] - 1]).max() / np.abs(Rg).max()
        # ownership covers every node exactly once across ranks
        cnt = torch.zeros(gmesh.num_nodes(), dtype=torch.float64)
        cnt[torch.from_numpy(part.local_to_global[own] - 1)] += 1
        dist.all_reduce(cnt)
        q.put((rank, float(err), bool((cnt == 1).all()), part.n_owned_elements, part.has_halo_block))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["metis", "bricks"])
def test_halo_exchange_two_ranks_gloo(mode):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sum(r[3] for r in res) == (4 * 4 * 8)
    for rank, err, cover, _, halo in res:
        assert err < 1e-13, (rank, err)
        assert cover
    # the lower rank owns the interface nodes, so it (and only it) needs a halo element layer
    assert {r[0]: r[4] for r in res} == {0: True, 1: False}


def test_metis_partition_balance_and_graph():
    mesh = F.StructuredMesh("hex", (0, 0, 0), (1, 1, 1), (9, 9, 9))
    ep = metis_partition_elements(mesh, 8)
    counts = np.bincount(ep, minlength=8)
    assert counts.sum() == 512 and counts.min() > 40 and counts.max() < 90
    # ext/MetisExt.jl:6-14: graph built from the pattern's (I, J): here the node adjacency of a 1-D chain
    nv = 64
    xadj = np.zeros(nv + 1, dtype=np.int64); adj = []
    for v in range(nv):
        nb = [u for u in (v - 1, v + 1) if 0 <= u < nv]
        adj += nb; xadj[v + 1] = xadj[v] + len(nb)
    part = metis_partition_graph(xadj, np.array(adj), 4)
    assert np.bincount(part, minlength=4).min() >= 12


def test_partition_invariants_four_parts():
    """every element owned once; ghost rows of owned nodes are complete with the halo block;
    send/recv lists of neighbouring ranks are mirror images in GLOBAL ids."""
    mesh = F.StructuredMesh("hex", (0, 0, 0), (1, 1, 1), (7, 7, 7))
    ep = metis_partition_elements(mesh, 4)
    parts = [partition_mesh(mesh, ep, 4, r) for r in range(4)]
    conn = mesh.element_conns["block_1"]
    assert sum(p.n_owned_elements for _, p in parts) == conn.shape[1]
    for r, (lm, p) in enumerate(parts):
        assert np.all(p.local_to_owner[:p.n_owned_nodes] == r) and np.all(p.local_to_owner[p.n_owned_nodes:] != r)
        # every global element touching an owned node is present locally (owned or halo)
        owned_g = set(p.local_to_global[:p.n_owned_nodes])
        need = {e for e in range(conn.shape[1]) if owned_g & set(conn[:, e])}
        have = set(p.owned_elements) | set(p.halo_elements)
        assert need == have
        for s, (_, ps) in enumerate(parts):
            if s == r:
                continue
            a = p.local_to_global[p.send[s] - 1] if s in p.send else np.zeros(0, dtype=np.int64)
            b = ps.local_to_global[ps.recv[r] - 1] if r in ps.recv else np.zeros(0, dtype=np.int64)
            assert np.array_equal(a, b)
            # consistent!: r's ghosts owned by s == s's owned nodes that r holds, same order (global id)
            a = p.local_to_global[p.ghost_by_owner[s] - 1] if s in p.ghost_by_owner else np.zeros(0, dtype=np.int64)
            b = ps.local_to_global[ps.own_ghosted[r] - 1] if r in ps.own_ghosted else np.zeros(0, dtype=np.int64)
            assert np.array_equal(a, b)


def test_brick_partition_matches_general_builder():
    n, grid = 3, (2, 2, 1)
    gmesh = F.StructuredMesh("hex", (0, 0, 0), (2, 2, 1), (2 * n + 1, 2 * n + 1, n + 1))
    # element owner from brick coordinates, StructuredMesh element order (ex outer, ez inner)
    ex, ey, ez = np.meshgrid(np.arange(2 * n), np.arange(2 * n), np.arange(n), indexing="ij")
    epart = ((ex // n) + 2 * (ey // n)).reshape(-1)
    for rank in range(4):
        lm_b, pb = structured_brick_partition(F, n, grid, rank)
        lm_g, pg = partition_mesh(gmesh, epart, 4, rank)
        assert pb.n_owned_nodes == pg.n_owned_nodes and pb.n_owned_elements == pg.n_owned_elements
        assert np.array_equal(np.sort(pb.local_to_global), np.sort(pg.local_to_global))
        assert np.array_equal(pb.local_to_global[:pb.n_owned_nodes], pg.local_to_global[:pg.n_owned_nodes])
        assert pb.neighbors == pg.neighbors
        for r in pb.neighbors:
            for a, b in ((pb.send, pg.send), (pb.recv, pg.recv)):
                ga = pb.local_to_global[a[r] - 1] if r in a else np.zeros(0, dtype=np.int64)
                gb = pg.local_to_global[b[r] - 1] if r in b else np.zeros(0, dtype=np.int64)
                assert np.array_equal(ga, gb)
        # same set of (global) owned and halo elements
        def gl(lm, p, blk):
            return {tuple(sorted(p.local_to_global[c - 1])) for c in lm.element_conns[blk].T} if blk in lm.element_conns else set()
        assert gl(lm_b, pb, "owned") == gl(lm_g, pg, "owned") and gl(lm_b, pb, "halo") == gl(lm_g, pg, "halo")
        assert np.allclose(np.asarray(lm_b.nodal_coords)[:, :pb.n_owned_nodes], np.asarray(lm_g.nodal_coords)[:, :pg.n_owned_nodes])


@pytest.mark.parametrize("how", ["bricks", "metis"])
def test_rank_local_sidesets_tile_the_global_ones(how):
    """surface loads on a partition: the side sets of the rank-local meshes are faces of OWNED elements only and,
    mapped back to global node ids, tile the global side set exactly once (no face lost, none counted twice), so
    summing the ranks' Neumann vectors over the halo gives the serial vector."""
    n, grid, P = 4, (2, 1, 1), 2
    gmesh = F.StructuredMesh("hex", (0., 0., 0.), (2., 1., 1.), (2 * n + 1, n + 1, n + 1))
    locals_ = []
    if how == "bricks":
        for r in range(P):
            locals_.append(structured_brick_partition(F, n, grid, r))
    else:
        epart = metis_partition_elements(gmesh, P)
        for r in range(P):
            locals_.append(partition_mesh(gmesh, epart, P, r))
    tabs = O.surface_tables("HEX8", "gauss2")
    Xg = np.asarray(gmesh.nodal_coords)
    for name in ("top", "right", "front", "left"):
        key = lambda cols: sorted(tuple(sorted(c)) for c in cols.T.tolist())
        want = key(np.asarray(gmesh.sideset_side_nodes[name]))
        got = []
        Rsum = np.zeros(3 * gmesh.num_nodes())
        for lm, part in locals_:
            sn = np.asarray(lm.sideset_side_nodes[name])
            n_owned_el = lm.element_conns["owned"].shape[1]
            assert np.all(np.asarray(lm.sideset_elems[name]) <= n_owned_el)          # faces of owned elements only
            if sn.size:
                gsn = part.local_to_global[sn - 1]
                got += key(gsn)
                # the rank's Neumann vector, scattered to global numbering (what the halo sum + owner add up to)
                Xl = np.asarray(lm.nodal_coords)
                Rl = O.assemble_vector_neumann_bc(np.zeros(3 * Xl.shape[1]), sn, tabs, np.ones((3, 4, sn.shape[1])), Xl, 3)
                np.add.at(Rsum.reshape(-1, 3), part.local_to_global - 1, Rl.reshape(-1, 3))
        assert sorted(got) == want
        Rg = O.assemble_vector_neumann_bc(np.zeros(3 * gmesh.num_nodes()), np.asarray(gmesh.sideset_side_nodes[name]), tabs,
                                          np.ones((3, 4, len(want))), Xg, 3)
        assert np.allclose(Rsum, Rg, rtol=1e-13, atol=1e-15)


def test_metis_partition_of_the_sparsity_pattern():
    """Metis.partition(pattern, nparts) (ext/MetisExt.jl:6-14) on the DOF graph of the CSR pattern: balanced parts, every
    dof assigned, edge cut well below a random assignment's."""
    m = O.structured_mesh("quad", (0., 0.), (1., 1.), (17, 17))
    dof = O.update_dofs(1, m["coords"].shape[1], [])
    pat = O.matrix_pattern([m["conn"]], 1, dof, condensed=True)
    n = m["coords"].shape[1]
    colptr, rowval, _ = O.sparse_csc(pat["Is"], pat["Js"], np.ones(len(pat["Is"])), n)
    part = F.metis_partition_pattern((n, colptr, rowval), 4)
    assert part.shape == (n,) and set(np.unique(part)) == {0, 1, 2, 3}
    counts = np.bincount(part, minlength=4)
    assert counts.max() <= 1.1 * n / 4 + 2
    rows = np.repeat(np.arange(n), np.diff(colptr))
    cut = np.count_nonzero(part[rows] != part[rowval - 1])
    rng = np.random.default_rng(0)
    rnd = rng.integers(0, 4, n)
    assert cut < 0.25 * np.count_nonzero(rnd[rows] != rnd[rowval - 1])


@pytest.mark.parametrize("nel,c,P", [((8, 8, 4), 2, 4), ((12, 6, 6), 3, 3), ((8, 8, 8), 2, 8)])
def test_metis_cell_partition_matches_general_builder(nel, c, P):
    """BASELINE config 5 builder: METIS on the coarse-cell graph + rank-local materialisation == partition_mesh on
    the global mesh with the same element owners (numbering, ownership, halo block, send / recv lists)."""
    Ex, Ey, Ez = nel
    cells = (Ex // c, Ey // c, Ez // c)
    cp = metis_cell_partition(cells, P)
    counts = np.bincount(cp.reshape(-1), minlength=P)
    assert counts.min() > 0 and counts.max() <= 1.35 * counts.mean() + 1
    h = 1.0 / min(nel)
    gmesh = F.StructuredMesh("hex", (0, 0, 0), (Ex * h, Ey * h, Ez * h), (Ex + 1, Ey + 1, Ez + 1))
    ex, ey, ez = np.meshgrid(np.arange(Ex), np.arange(Ey), np.arange(Ez), indexing="ij")   # StructuredMesh element order
    epart = cp[ex // c, ey // c, ez // c].reshape(-1)
    for rank in range(P):
        lm_b, pb = structured_cell_partition(F, nel, cp, c, rank)
        lm_g, pg = partition_mesh(gmesh, epart, P, rank)
        assert pb.n_owned_nodes == pg.n_owned_nodes and pb.n_owned_elements == pg.n_owned_elements
        assert np.array_equal(pb.local_to_global, pg.local_to_global)
        assert np.array_equal(pb.local_to_owner, pg.local_to_owner)
        assert pb.neighbors == pg.neighbors
        for r in pb.neighbors:
            for a, b in ((pb.send, pg.send), (pb.recv, pg.recv)):
                ga = pb.local_to_global[a[r] - 1] if r in a else np.zeros(0, dtype=np.int64)
                gb = pg.local_to_global[b[r] - 1] if r in b else np.zeros(0, dtype=np.int64)
                assert np.array_equal(ga, gb)

        def gl(lm, p, blk):
            return {tuple(sorted(p.local_to_global[cc - 1])) for cc in lm.element_conns[blk].T} if blk in lm.element_conns else set()
        assert gl(lm_b, pb, "owned") == gl(lm_g, pg, "owned") and gl(lm_b, pb, "halo") == gl(lm_g, pg, "halo")
        assert np.allclose(np.asarray(lm_b.nodal_coords), np.asarray(lm_g.nodal_coords))
        # owner -> ghost update lists (every ghost, incl. the far nodes of halo elements): same in both builders
        assert set(pb.own_ghosted) == set(pg.own_ghosted) and set(pb.ghost_by_owner) == set(pg.ghost_by_owner)
        for r in pb.own_ghosted:
            assert np.array_equal(pb.own_ghosted[r], pg.own_ghosted[r])
        assert sum(len(v) for v in pb.ghost_by_owner.values()) == len(pb.local_to_global) - pb.n_owned_nodes
        for name in ("bottom", "top", "left", "right", "back", "front"):
            assert np.array_equal(np.sort(pb.local_to_global[lm_b.nodeset_nodes[name] - 1]),
                                  np.sort(pg.local_to_global[lm_g.nodeset_nodes[name] - 1]))
