"""Domain decomposition for multi-GPU assembly (ext/PartitionedArraysExt.jl, ext/MetisExt.jl).

One process per GPU.  The mesh is split by ELEMENTS (METIS k-way on the element dual graph, or bricks
for structured meshes); a node is owned by the lowest rank among the owners of the elements that touch
it (test/ext/script.jl:29-32 uses `minimum` the same way).  Every rank holds a rank-local mesh:

    local nodes    = owned nodes first, then ghosts            (PartitionedArrays OwnAndGhostIndices)
    block "owned"  = the elements this rank owns               -> residual / action assembly
    block "halo"   = neighbour-owned elements touching an owned node -> only the Jacobian uses them,
                     so each rank assembles ALL of its owned rows locally, with no exchange.

Residual: contributions to ghost nodes are packed on the device, exchanged with NCCL send/recv and
added into the owner's entries (the reverse halo / "assemble" step of PVector, :469-481).
`local_to_global` / `local_to_owner` are exactly what `LocalIndices(n_global, rank, l2g, l2o)` takes
(ext/PartitionedArraysExt.jl:228-232).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, lib
from .fields import H1Field
from .meshes import AbstractMesh


def metis_partition_elements(mesh, nparts, ncommon=None):
    """epart (0-based part id per element, blocks concatenated) from METIS_PartMeshDual."""
    conns = [mesh.element_conns[b] - 1 for b in mesh.element_block_names]
    ne = sum(c.shape[1] for c in conns)
    nn = mesh.num_nodes()
    eptr = np.zeros(ne + 1, dtype=np.int64)
    eptr[1:] = np.cumsum(np.concatenate([np.full(c.shape[1], c.shape[0]) for c in conns]))
    eind = np.concatenate([np.ascontiguousarray(c.T).reshape(-1) for c in conns]).astype(np.int64)
    if ncommon is None:
        ncommon = 2 if mesh.num_dimensions() == 2 else 3
    epart = np.zeros(ne, dtype=np.int64)
    npart = np.zeros(nn, dtype=np.int64)
    check(lib.fecb200_metis_part_mesh_dual(ne, nn, _lib.i64(eptr)[1], _lib.i64(eind)[1], ncommon, nparts,
                                           epart.ctypes.data_as(_lib.c_i64p), npart.ctypes.data_as(_lib.c_i64p)))
    return epart


def metis_partition_graph(xadj, adjncy, nparts):
    """Metis.partition(graph, nparts) (ext/MetisExt.jl:6-14): 0-based CSR adjacency -> part ids."""
    xadj, xp = _lib.i64(xadj)
    adjncy, ap = _lib.i64(adjncy)
    part = np.zeros(len(xadj) - 1, dtype=np.int64)
    check(lib.fecb200_metis_part_graph(len(part), xp, ap, nparts, part.ctypes.data_as(_lib.c_i64p)))
    return part


def metis_partition_pattern(pattern, nparts):
    """Metis.partition(pattern::SparseMatrixPattern, nparts) (ext/MetisExt.jl:6-14): partition of the DOF graph of the
    sparsity pattern.  `pattern` is an assembler (its `pattern()` is used) or a `(n, ptr, idx)` triple, Int64 1-based as
    the library returns it; the diagonal is dropped (METIS graphs have no self loops).  Returns 0-based part ids."""
    n, ptr, idx = pattern.pattern() if hasattr(pattern, "pattern") else pattern
    ptr, idx = np.asarray(ptr, dtype=np.int64) - 1, np.asarray(idx, dtype=np.int64) - 1
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(ptr))
    keep = rows != idx
    xadj = np.zeros(n + 1, dtype=np.int64)
    np.add.at(xadj, rows[keep] + 1, 1)
    return metis_partition_graph(np.cumsum(xadj), idx[keep], nparts)


class _LocalMesh(AbstractMesh):
    """Rank-local mesh: block `owned` (+ `halo`: neighbour-owned elements, matrix assembly only).  Side sets are the
    faces of OWNED elements lying in the node sets, built on first use (surface loads of halo elements belong to their
    owner rank)."""

    def _lazy_sidesets(self):
        if "_ss" not in self.__dict__:
            self.__dict__["_ss"] = True
            self._set_sidesets(self._sidesets_from_nodesets(self.sideset_nodes, blocks=["owned"]))

    def __getattr__(self, name):
        if name in ("sideset_elems", "sideset_sides", "sideset_side_nodes"):
            self._lazy_sidesets()
            return self.__dict__[name]
        raise AttributeError(name)


class Partition:
    """Rank-local decomposition data + halo exchange."""

    def __init__(self, rank, nparts, local_to_global, local_to_owner, n_owned_nodes, n_owned_elements,
                 n_global_elements, n_global_nodes, send, recv, has_halo_block, own_ghosted=None):
        self.rank, self.nparts = rank, nparts
        self.local_to_global = np.asarray(local_to_global, dtype=np.int64)   # 1-based global node ids
        self.local_to_owner = np.asarray(local_to_owner, dtype=np.int64)     # 0-based owner rank per local node
        self.n_owned_nodes = int(n_owned_nodes)
        self.n_owned_elements = int(n_owned_elements)
        self.n_global_elements = int(n_global_elements)
        self.n_global_nodes = int(n_global_nodes)
        # {neighbour rank: 1-based local node ids}, both sorted by GLOBAL node id so the two sides agree
        self.send = {int(r): np.asarray(v, dtype=np.int64) for r, v in sorted(send.items()) if len(v)}
        self.recv = {int(r): np.asarray(v, dtype=np.int64) for r, v in sorted(recv.items()) if len(v)}
        self.neighbors = sorted(set(self.send) | set(self.recv))
        self.has_halo_block = has_halo_block
        # owner -> ghost update (consistent!): EVERY ghost node by owner, and (own_ghosted) my owned nodes every other
        # rank holds as ghosts -- a superset of the residual lists: the far nodes of halo elements are Jacobian columns
        # but never receive residual contributions.  Both sorted by global id.
        l2g, l2o = self.local_to_global, self.local_to_owner
        gh = np.arange(self.n_owned_nodes, len(l2g))
        self.ghost_by_owner = {int(r): gh[l2o[gh] == r][np.argsort(l2g[gh[l2o[gh] == r]], kind="stable")] + 1
                               for r in np.unique(l2o[gh])} if len(gh) else {}
        self.own_ghosted = None if own_ghosted is None else \
            {int(r): np.asarray(v, dtype=np.int64) for r, v in sorted(own_ghosted.items()) if len(v)}
        self._sendbuf = self._recvbuf = None
        self._nf = None

    # ---- library wiring ---------------------------------------------------------------------------
    def attach(self, asm):
        """partition_setup (ghost rows dropped, halo block skipped by vector assembly) + halo lists"""
        h = asm._require()
        nb = asm.dof.var.fspace.num_blocks()
        flags = (C.c_int32 * nb)(*([0] + [1] * (nb - 1) if self.has_halo_block else [0] * nb))
        check(lib.fecb200_partition_setup(h, self.n_owned_nodes, flags))
        asm._pattern = None
        nbrs = np.array(self.neighbors, dtype=np.int32)
        sp = np.zeros(len(nbrs) + 1, dtype=np.int64)
        rp = np.zeros(len(nbrs) + 1, dtype=np.int64)
        sn, rn = [], []
        for i, r in enumerate(nbrs):
            s, v = self.send.get(int(r), np.zeros(0, dtype=np.int64)), self.recv.get(int(r), np.zeros(0, dtype=np.int64))
            sn.append(s); rn.append(v)
            sp[i + 1], rp[i + 1] = sp[i] + len(s), rp[i] + len(v)
        sn = np.concatenate(sn) if sn else np.zeros(0, dtype=np.int64)
        rn = np.concatenate(rn) if rn else np.zeros(0, dtype=np.int64)
        check(lib.fecb200_halo_setup(h, len(nbrs), nbrs.ctypes.data_as(_lib.c_i32p), _lib.i64(sp)[1], _lib.i64(sn)[1],
                                     _lib.i64(rp)[1], _lib.i64(rn)[1]))
        if self.own_ghosted is not None:
            gr = sorted(set(self.own_ghosted) | set(self.ghost_by_owner))
            op = np.zeros(len(gr) + 1, dtype=np.int64)
            gp = np.zeros(len(gr) + 1, dtype=np.int64)
            on, gn = [], []
            for i, r in enumerate(gr):
                a, b = self.own_ghosted.get(r, np.zeros(0, dtype=np.int64)), self.ghost_by_owner.get(r, np.zeros(0, dtype=np.int64))
                on.append(a); gn.append(b)
                op[i + 1], gp[i + 1] = op[i] + len(a), gp[i] + len(b)
            on = np.concatenate(on) if on else np.zeros(0, dtype=np.int64)
            gn = np.concatenate(gn) if gn else np.zeros(0, dtype=np.int64)
            gra = np.array(gr, dtype=np.int32)
            check(lib.fecb200_ghost_setup(h, len(gr), gra.ctypes.data_as(_lib.c_i32p), _lib.i64(op)[1], _lib.i64(on)[1],
                                          _lib.i64(gp)[1], _lib.i64(gn)[1]))
        self._nf = asm.dof.nf
        self._asm_handle = h
        self._send_counts = [int(sp[i + 1] - sp[i]) * self._nf for i in range(len(nbrs))]
        self._recv_counts = [int(rp[i + 1] - rp[i]) * self._nf for i in range(len(nbrs))]

    def comm_init(self, asm, unique_id=None):
        """fecb200_comm_init: the library's own NCCL communicator for this handle.  The 128-byte ncclUniqueId of rank 0
        is the one thing that travels out of band: `unique_id` (bytes, e.g. from MPI_Bcast / a file), or, when None,
        a torch.distributed object broadcast.  After this, halo sums, barriers, the peer-memory set-up and the
        distributed CG / Newton run inside the library -- torch.distributed is not in the data path."""
        if unique_id is None:
            import torch.distributed as dist
            box = [None]
            if self.rank == 0:
                buf = C.create_string_buffer(128)
                check(lib.fecb200_comm_unique_id(C.cast(buf, C.c_void_p)))
                box[0] = bytes(buf.raw)
            dist.broadcast_object_list(box, src=0)
            unique_id = box[0]
        assert len(unique_id) == 128
        buf = C.create_string_buffer(unique_id, 128)
        check(lib.fecb200_comm_init(asm._require(), self.rank, self.nparts, C.cast(buf, C.c_void_p)))
        self._comm = True

    def enable_peer_scatter(self, asm):
        """Switch the residual halo to the fused NVLink path: kernels add ghost contributions directly into the
        owner's residual through peer-mapped memory (CUDA IPC).  With the library communicator (comm_init) the whole
        exchange of IPC handles and owner-local ghost ids happens inside libfecb200; the torch.distributed variant
        below is kept for hosts that drive the halo themselves."""
        if getattr(self, "_comm", False):
            check(lib.fecb200_comm_peer_enable(asm._require(), _lib.FIELD_RESIDUAL))
            self._peer = True
            return
        import torch
        import torch.distributed as dist
        h = asm._require()
        dev = torch.device("cuda", torch.cuda.current_device())
        sc = [len(self.send.get(r, ())) for r in self.neighbors]
        rc = [len(self.recv.get(r, ())) for r in self.neighbors]
        # owner-local ids of my ghosts: every neighbour tells me ITS local ids of the nodes it receives from me
        mine = np.concatenate([self.recv[r] for r in self.neighbors if r in self.recv] or [np.zeros(0, dtype=np.int64)])
        mine_t = torch.from_numpy(np.ascontiguousarray(mine)).to(dev)
        theirs_t = torch.zeros(max(1, sum(sc)), dtype=torch.int64, device=dev)
        exchange(self.neighbors, mine_t, rc, theirs_t, sc)
        theirs = theirs_t.cpu().numpy()
        # IPC handles of every rank's residual field
        buf = (C.c_ubyte * 64)()
        check(lib.fecb200_ipc_export(h, _lib.FIELD_RESIDUAL, C.cast(buf, C.c_void_p)))
        mine_h = torch.tensor(list(buf), dtype=torch.uint8, device=dev)
        allh = [torch.zeros(64, dtype=torch.uint8, device=dev) for _ in range(self.nparts)]
        dist.all_gather(allh, mine_h)
        peers = [r for r in self.neighbors if r in self.send]
        assert len(peers) <= 8
        handles = b"".join(bytes(allh[r].cpu().numpy().tolist()) for r in peers)
        n_ghost = len(self.local_to_global) - self.n_owned_nodes
        gpeer = np.full(n_ghost, -1, dtype=np.int32)
        gnode = np.zeros(n_ghost, dtype=np.int64)
        off = 0
        for r, ns in zip(self.neighbors, sc):
            if ns:
                g = self.send[r] - 1 - self.n_owned_nodes
                gpeer[g] = peers.index(r)
                gnode[g] = theirs[off:off + ns] - 1
            off += ns
        hb = C.create_string_buffer(handles, max(64, len(handles)))
        # node count of every peer's local mesh (range check of the ghost ids inside the library)
        nn_t = torch.tensor([len(self.local_to_global)], dtype=torch.int64, device=dev)
        all_nn = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(self.nparts)]
        dist.all_gather(all_nn, nn_t)
        peer_nn = np.array([int(all_nn[r].item()) for r in peers] or [0], dtype=np.int64)
        check(lib.fecb200_peer_attach(h, _lib.FIELD_RESIDUAL, len(peers), C.cast(hb, C.c_void_p),
                                      peer_nn.ctypes.data_as(_lib.c_i64p), gpeer.ctypes.data_as(_lib.c_i32p),
                                      gnode.ctypes.data_as(_lib.c_i64p), n_ghost))
        self._peer = True
        self._bar = torch.zeros(1, dtype=torch.float32, device=dev)

    def barrier_on_stream(self, stream=None):
        """stream-ordered cross-rank barrier (a 4-byte NCCL all-reduce): orders the peer scatter phases"""
        if getattr(self, "_comm", False):
            check(lib.fecb200_comm_barrier(self._asm_handle))   # on the handle's stream
            return
        import torch
        import torch.distributed as dist
        ctx = torch.cuda.stream(stream) if stream is not None else _nullctx()
        with ctx:
            dist.all_reduce(self._bar)

    def halo_sum_residual(self, asm, stream=None, field=_lib.FIELD_RESIDUAL):
        """ghost -> owner accumulation of the assembled residual (device pack, NCCL send/recv, device add)"""
        if getattr(self, "_comm", False):
            # library communicator: pack / grouped ncclSend+ncclRecv / add, or the closing barrier of the peer halo
            check(lib.fecb200_halo_sum(asm._require(), field))
            return
        import torch
        if getattr(self, "_peer", False):
            # fused path: the kernels already added the ghost rows into their owners; just wait for everyone
            self.barrier_on_stream(stream)
            return
        h = asm._require()
        if self._sendbuf is None:
            dev = torch.device("cuda", torch.cuda.current_device())
            self._sendbuf = torch.empty(max(1, sum(self._send_counts)), dtype=torch.float64, device=dev)
            self._recvbuf = torch.empty(max(1, sum(self._recv_counts)), dtype=torch.float64, device=dev)
        ctx = torch.cuda.stream(stream) if stream is not None else _nullctx()
        with ctx:
            check(lib.fecb200_halo_pack(h, field, _lib.ptr(self._sendbuf)))
            exchange(self.neighbors, self._sendbuf, self._send_counts, self._recvbuf, self._recv_counts)
            check(lib.fecb200_halo_unpack_add(h, field, _lib.ptr(self._recvbuf)))


class _nullctx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def exchange(neighbors, sendbuf, send_counts, recvbuf, recv_counts):
    """One batched point-to-point exchange with every neighbour (ncclGroupStart/Send/Recv/End under
    torch.distributed; works with gloo on CPU tensors for the host-logic tests)."""
    import torch.distributed as dist
    ops, so, ro = [], 0, 0
    for r, ns, nr in zip(neighbors, send_counts, recv_counts):
        if ns:
            ops.append(dist.P2POp(dist.isend, sendbuf[so:so + ns], r))
        if nr:
            ops.append(dist.P2POp(dist.irecv, recvbuf[ro:ro + nr], r))
        so += ns
        ro += nr
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


# --------------------------------------------------------------------------------------------------
# general builder: any single-block mesh + an element partition (e.g. from METIS)
# --------------------------------------------------------------------------------------------------
def partition_mesh(mesh, epart, nparts, rank):
    """Rank-local mesh + Partition from a global mesh and `epart` (0-based owner of every element).
    Every rank can call this with the same global data (no communication needed)."""
    assert len(mesh.element_block_names) == 1, "partition_mesh handles single-block meshes"
    bname = mesh.element_block_names[0]
    conn = mesh.element_conns[bname] - 1                     # (NNPE, NE) 0-based
    nn = mesh.num_nodes()
    epart = np.asarray(epart, dtype=np.int64)
    owner = np.full(nn, nparts, dtype=np.int64)
    np.minimum.at(owner, conn.reshape(-1), np.broadcast_to(epart, conn.shape).reshape(-1))
    mine_e = np.nonzero(epart == rank)[0]
    owned_nodes = np.nonzero(owner == rank)[0]
    own_mask = np.zeros(nn, dtype=bool)
    own_mask[owned_nodes] = True
    halo_e = np.nonzero(own_mask[conn].any(axis=0) & (epart != rank))[0]
    touched_own = np.unique(conn[:, mine_e])
    all_local = np.unique(np.concatenate([touched_own, np.unique(conn[:, halo_e]) if len(halo_e) else touched_own,
                                          owned_nodes]))
    ghosts = all_local[~own_mask[all_local]]
    l2g = np.concatenate([owned_nodes, ghosts])
    g2l = np.full(nn, -1, dtype=np.int64)
    g2l[l2g] = np.arange(len(l2g))
    # halo lists (both sides sorted by global id)
    res_ghosts = touched_own[~own_mask[touched_own]]
    send = {int(r): g2l[res_ghosts[owner[res_ghosts] == r]] + 1 for r in np.unique(owner[res_ghosts])}
    recv = {}
    for r in range(nparts):
        if r == rank:
            continue
        er = np.nonzero(epart == r)[0]
        if not len(er):
            continue
        tr = np.unique(conn[:, er])
        hit = tr[own_mask[tr]]
        if len(hit):
            recv[r] = g2l[hit] + 1
    lm = _LocalMesh()
    lm.nodal_coords = H1Field(np.asarray(mesh.nodal_coords)[:, l2g])
    et = mesh.element_types[bname]
    lm.element_block_names = ["owned"] + (["halo"] if len(halo_e) else [])
    lm.element_types = {b: et for b in lm.element_block_names}
    lm.element_conns = {"owned": g2l[conn[:, mine_e]] + 1}
    if len(halo_e):
        lm.element_conns["halo"] = g2l[conn[:, halo_e]] + 1

    def restrict(sets):
        out = {}
        for k, v in sets.items():
            loc = g2l[np.asarray(v, dtype=np.int64) - 1]
            out[k] = loc[loc >= 0] + 1
        return out

    lm.nodeset_nodes = restrict(mesh.nodeset_nodes)
    lm.sideset_nodes = restrict(mesh.sideset_nodes)
    own_ghosted = {}
    for r in range(nparts):       # rank r's local node set = nodes of its owned + halo elements; my owned nodes in it
        if r == rank:
            continue
        er = epart == r
        if not er.any():
            continue
        own_r = owner == r
        loc_r = np.zeros(nn, dtype=bool)
        loc_r[conn[:, er].reshape(-1)] = True
        halo_r = own_r[conn].any(axis=0) & ~er
        loc_r[conn[:, halo_r].reshape(-1)] = True
        hit = np.nonzero(loc_r & own_mask)[0]          # ascending global id
        if len(hit):
            own_ghosted[r] = g2l[hit] + 1
    part = Partition(rank, nparts, l2g + 1, owner[l2g], len(owned_nodes), len(mine_e), conn.shape[1], nn, send, recv,
                     bool(len(halo_e)), own_ghosted)
    part.owned_elements, part.halo_elements = mine_e, halo_e
    return lm, part


# --------------------------------------------------------------------------------------------------
# structured bricks for the weak-scaling benchmark: every rank builds ONLY its own piece
# --------------------------------------------------------------------------------------------------
def structured_brick_partition(F, n, grid, rank):
    """Rank-local hex8 mesh of the global (gx*n) x (gy*n) x (gz*n) element grid split into gx*gy*gz bricks
    of n^3 elements (what METIS returns on this grid, without building the 57M-node global mesh on every
    rank).  Global numbering = StructuredMesh's (nodes x fastest; elements ex outer, ez inner)."""
    gx, gy, gz = grid
    P = gx * gy * gz
    px, py, pz = rank % gx, (rank // gx) % gy, rank // (gx * gy)
    Ng = np.array([gx * n + 1, gy * n + 1, gz * n + 1], dtype=np.int64)
    pc, g = np.array([px, py, pz]), np.array(grid)
    lo = pc * n                                       # first global node index of the brick
    hal = (pc < g - 1).astype(np.int64)               # one halo element layer on the high sides
    nloc = n + 1 + hal                                # local node box
    ii, jj, kk = np.meshgrid(np.arange(nloc[0]), np.arange(nloc[1]), np.arange(nloc[2]), indexing="ij")
    ig, jg, kg = ii + lo[0], jj + lo[1], kk + lo[2]
    gid = (ig + Ng[0] * (jg + Ng[1] * kg)).reshape(-1)   # 0-based global node id

    def owner_of(i, j, k):
        ox = np.minimum(np.maximum(0, (i - 1) // n), gx - 1)
        oy = np.minimum(np.maximum(0, (j - 1) // n), gy - 1)
        oz = np.minimum(np.maximum(0, (k - 1) // n), gz - 1)
        return ox + gx * (oy + gy * oz)

    own = owner_of(ig, jg, kg).reshape(-1)
    is_own = own == rank
    order = np.concatenate([np.nonzero(is_own)[0][np.argsort(gid[is_own], kind="stable")],
                            np.nonzero(~is_own)[0][np.argsort(gid[~is_own], kind="stable")]])
    box2loc = np.empty(len(gid), dtype=np.int64)
    box2loc[order] = np.arange(len(gid))
    l2g, l2o = gid[order], own[order]
    n_owned = int(is_own.sum())
    box = lambda i, j, k: (i * nloc[1] + j) * nloc[2] + k   # index into the meshgrid('ij') raveling

    def conn_of(ex, ey, ez):  # local box element coordinates -> (8, NE) local node ids (1-based)
        c = [box(ex, ey, ez), box(ex + 1, ey, ez), box(ex + 1, ey + 1, ez), box(ex, ey + 1, ez),
             box(ex, ey, ez + 1), box(ex + 1, ey, ez + 1), box(ex + 1, ey + 1, ez + 1), box(ex, ey + 1, ez + 1)]
        return box2loc[np.stack(c)] + 1

    nel = nloc - 1
    ex, ey, ez = np.meshgrid(np.arange(nel[0]), np.arange(nel[1]), np.arange(nel[2]), indexing="ij")
    ex, ey, ez = ex.reshape(-1), ey.reshape(-1), ez.reshape(-1)
    inside = (ex < n) & (ey < n) & (ez < n)
    lm = _LocalMesh()
    h = 1.0 / n
    X = np.stack([ig.reshape(-1) * h, jg.reshape(-1) * h, kg.reshape(-1) * h])[:, order]
    lm.nodal_coords = H1Field(X)
    lm.element_block_names = ["owned"] + (["halo"] if (~inside).any() else [])
    lm.element_types = {b: "HEX8" for b in lm.element_block_names}
    lm.element_conns = {"owned": conn_of(ex[inside], ey[inside], ez[inside])}
    if (~inside).any():
        lm.element_conns["halo"] = conn_of(ex[~inside], ey[~inside], ez[~inside])
    jgo, igo, kgo = jg.reshape(-1)[order], ig.reshape(-1)[order], kg.reshape(-1)[order]
    loc = np.arange(1, len(l2g) + 1)
    lm.nodeset_nodes = {"bottom": loc[jgo == 0], "top": loc[jgo == Ng[1] - 1], "left": loc[igo == 0],
                        "right": loc[igo == Ng[0] - 1], "back": loc[kgo == 0], "front": loc[kgo == Ng[2] - 1]}
    lm.sideset_nodes = dict(lm.nodeset_nodes)
    # residual halo: ghosts touched by OWNED elements = nodes of the brick box not owned by this rank
    inbrick = ((ii <= n) & (jj <= n) & (kk <= n)).reshape(-1)[order]
    send = {}
    cand = np.nonzero(inbrick & (l2o != rank))[0]
    for r in np.unique(l2o[cand]):
        sel = cand[l2o[cand] == r]
        send[int(r)] = sel[np.argsort(l2g[sel], kind="stable")] + 1
    # owned nodes that a neighbour's brick touches (that neighbour sends them to us)
    recv = {}
    for dz in (0, 1):
        for dy in (0, 1):
            for dx in (0, 1):
                if dx == dy == dz == 0:
                    continue
                q = pc + np.array([dx, dy, dz])
                if np.any(q >= g):
                    continue
                r = int(q[0] + gx * (q[1] + gy * q[2]))
                qlo = q * n
                sel = np.nonzero((l2o == rank) & (igo >= qlo[0]) & (igo <= qlo[0] + n) & (jgo >= qlo[1]) &
                                 (jgo <= qlo[1] + n) & (kgo >= qlo[2]) & (kgo <= qlo[2] + n))[0]
                if len(sel):
                    recv[r] = sel[np.argsort(l2g[sel], kind="stable")] + 1
    part = Partition(rank, P, l2g + 1, l2o, n_owned, int(inside.sum()), int(np.prod(g * n)), int(np.prod(Ng)), send, recv,
                     bool((~inside).any()))
    return lm, part


# --------------------------------------------------------------------------------------------------
# METIS partition of the structured weak-scaling grid (BASELINE config 5) without the global mesh:
# METIS k-way runs on the dual graph of COARSE cells (c^3 elements each; 384^3 elements -> 48^3 cells),
# every rank then materialises only its own ragged piece.
# --------------------------------------------------------------------------------------------------
def metis_cell_partition(cells, nparts):
    """METIS_PartGraphKway (ext/MetisExt.jl:6-14 partitions the same way, on the DOF graph) on the face-adjacency
    graph of a cx x cy x cz grid of cells.  Returns part ids shaped (cx, cy, cz), 0-based."""
    cx, cy, cz = (int(v) for v in cells)
    nv = cx * cy * cz
    ids = np.arange(nv, dtype=np.int64).reshape(cx, cy, cz)
    src, dst = [], []
    for ax in range(3):
        lo = [slice(None)] * 3
        hi = [slice(None)] * 3
        lo[ax], hi[ax] = slice(0, -1), slice(1, None)
        a, b = ids[tuple(lo)].reshape(-1), ids[tuple(hi)].reshape(-1)
        src += [a, b]
        dst += [b, a]
    src, dst = np.concatenate(src), np.concatenate(dst)
    order = np.lexsort((dst, src))
    src, dst = src[order], dst[order]
    xadj = np.zeros(nv + 1, dtype=np.int64)
    np.add.at(xadj, src + 1, 1)
    part = metis_partition_graph(np.cumsum(xadj), dst, nparts)
    return part.reshape(cx, cy, cz)


def _dilate_to_nodes(emask):
    """element mask (ex, ey, ez) -> mask of the nodes those elements touch (ex+1, ey+1, ez+1)"""
    out = np.zeros(tuple(s + 1 for s in emask.shape), dtype=bool)
    for dx in (0, 1):
        for dy in (0, 1):
            for dz in (0, 1):
                out[dx:dx + emask.shape[0], dy:dy + emask.shape[1], dz:dz + emask.shape[2]] |= emask
    return out


def structured_cell_partition(F, nel, cell_part, c, rank, h=None):
    """Rank-local hex8 mesh + Partition of the structured grid of nel = (Ex, Ey, Ez) elements whose c^3-element
    cells are assigned to ranks by `cell_part` (shape nel // c).  Same numbering and ownership rules as
    partition_mesh (global node id x fastest; node owner = lowest rank among the touching elements' owners), but only
    this rank's bounding box is ever materialised."""
    Ex, Ey, Ez = (int(v) for v in nel)
    cell_part = np.asarray(cell_part)
    assert cell_part.shape == (Ex // c, Ey // c, Ez // c) and Ex % c == Ey % c == Ez % c == 0
    P = int(cell_part.max()) + 1
    h = 1.0 / min(Ex, Ey, Ez) if h is None else h
    mine = np.argwhere(cell_part == rank)
    assert len(mine), f"rank {rank} owns no cell"
    # element box = my cells + TWO elements of margin on every side (clipped to the domain): halo elements sit in the
    # first layer, and the owner of THEIR nodes depends on the second
    elo = np.maximum(mine.min(axis=0) * c - 2, 0)
    ehi = np.minimum((mine.max(axis=0) + 1) * c + 2, [Ex, Ey, Ez])           # exclusive
    bx, by, bz = (ehi - elo).tolist()
    BIG = np.int16(P)
    # owner of every element of the box (int16), from the cell map
    ix, iy, iz = (np.arange(elo[a], ehi[a]) // c for a in range(3))
    eown = cell_part[np.ix_(ix, iy, iz)].astype(np.int16)
    # node owner = min over the (up to 8) touching elements; elements outside the DOMAIN do not exist, elements outside
    # the BOX only matter for nodes on the box boundary, which are never local (the margin guarantees it)
    nown = np.full((bx + 1, by + 1, bz + 1), BIG, dtype=np.int16)
    for dx in (0, 1):
        for dy in (0, 1):
            for dz in (0, 1):
                v = nown[dx:dx + bx, dy:dy + by, dz:dz + bz]
                np.minimum(v, eown, out=v)
    e_mine = eown == rank
    n_touched = _dilate_to_nodes(e_mine)                 # nodes of my elements (owned or ghost)
    n_owned = n_touched & (nown == rank)
    # halo elements: not mine, touching one of my owned nodes
    touch = np.zeros((bx, by, bz), dtype=bool)
    for dx in (0, 1):
        for dy in (0, 1):
            for dz in (0, 1):
                touch |= n_owned[dx:dx + bx, dy:dy + by, dz:dz + bz]
    e_halo = touch & ~e_mine
    n_local = n_touched | _dilate_to_nodes(e_halo)
    n_ghost = n_local & ~n_owned
    # box-boundary sanity: a local node on the low/high face of the box must be on the domain boundary there
    for a, (lo_, hi_, E_) in enumerate(zip(elo, ehi, (Ex, Ey, Ez))):
        sl_lo = [slice(None)] * 3; sl_lo[a] = 0
        sl_hi = [slice(None)] * 3; sl_hi[a] = -1
        assert lo_ == 0 or not n_owned[tuple(sl_lo)].any()
        assert hi_ == E_ or not n_owned[tuple(sl_hi)].any()
    # local numbering: owned nodes by ascending global id, then ghosts by ascending global id.  Arrays are indexed
    # [x][y][z]; global id = x + Nx (y + Ny z), so sort keys come from a transposed (z, y, x) walk.
    Nx, Ny, Nz = Ex + 1, Ey + 1, Ez + 1

    def ordered(mask):
        kz, ky, kx = np.nonzero(mask.transpose(2, 1, 0))   # ascending (z, y, x) == ascending global id
        return kx, ky, kz

    ox, oy, oz = ordered(n_owned)
    gx, gy, gz = ordered(n_ghost)
    lx, ly, lz = np.concatenate([ox, gx]), np.concatenate([oy, gy]), np.concatenate([oz, gz])
    n_own = len(ox)
    box2loc = np.full((bx + 1, by + 1, bz + 1), -1, dtype=np.int32)
    box2loc[lx, ly, lz] = np.arange(len(lx), dtype=np.int32)
    GX, GY, GZ = lx + elo[0], ly + elo[1], lz + elo[2]
    l2g = (GX + Nx * (GY + Ny * GZ)).astype(np.int64)
    l2o = nown[lx, ly, lz].astype(np.int64)

    def conn_of(emask):
        ex, ey, ez = np.nonzero(emask)
        c8 = [box2loc[ex, ey, ez], box2loc[ex + 1, ey, ez], box2loc[ex + 1, ey + 1, ez], box2loc[ex, ey + 1, ez],
              box2loc[ex, ey, ez + 1], box2loc[ex + 1, ey, ez + 1], box2loc[ex + 1, ey + 1, ez + 1], box2loc[ex, ey + 1, ez + 1]]
        out = np.stack(c8).astype(np.int64) + 1
        assert out.min() >= 1
        return out

    lm = _LocalMesh()
    lm.nodal_coords = H1Field(np.stack([GX * h, GY * h, GZ * h]).astype(float))
    has_halo = bool(e_halo.any())
    lm.element_block_names = ["owned"] + (["halo"] if has_halo else [])
    lm.element_types = {b: "HEX8" for b in lm.element_block_names}
    lm.element_conns = {"owned": conn_of(e_mine)}
    if has_halo:
        lm.element_conns["halo"] = conn_of(e_halo)
    loc = np.arange(1, len(l2g) + 1)
    lm.nodeset_nodes = {"bottom": loc[GY == 0], "top": loc[GY == Ny - 1], "left": loc[GX == 0], "right": loc[GX == Nx - 1],
                        "back": loc[GZ == 0], "front": loc[GZ == Nz - 1]}
    lm.sideset_nodes = dict(lm.nodeset_nodes)
    # halo lists, both sides sorted by global id (= local order inside the owned and the ghost range)
    send, recv = {}, {}
    res_ghost = n_touched & ~n_owned                       # ghosts my OWNED elements add to
    gl = box2loc[ordered(res_ghost)].astype(np.int64)
    go = l2o[gl]
    for r in np.unique(go):
        send[int(r)] = gl[go == r] + 1
    for r in np.unique(eown):
        r = int(r)
        if r == rank:
            continue
        hit = _dilate_to_nodes(eown == r) & n_owned        # my owned nodes that rank r's elements touch
        if hit.any():
            recv[r] = box2loc[ordered(hit)].astype(np.int64) + 1
    # owner -> ghost update lists: my owned nodes inside rank r's local node set (its elements + ITS halo elements)
    own_ghosted = {}
    for r in np.unique(eown):
        r = int(r)
        if r == rank:
            continue
        e_r = eown == r
        touch_r = np.zeros((bx, by, bz), dtype=bool)
        n_r = nown == r
        for dx in (0, 1):
            for dy in (0, 1):
                for dz in (0, 1):
                    touch_r |= n_r[dx:dx + bx, dy:dy + by, dz:dz + bz]
        loc_r = _dilate_to_nodes(e_r | (touch_r & ~e_r))
        hit = loc_r & n_owned
        if hit.any():
            own_ghosted[r] = box2loc[ordered(hit)].astype(np.int64) + 1
    part = Partition(rank, P, l2g + 1, l2o, n_own, int(e_mine.sum()), Ex * Ey * Ez, Nx * Ny * Nz, send, recv, has_halo,
                     own_ghosted)
    part.n_halo_elements = int(e_halo.sum())
    return lm, part
