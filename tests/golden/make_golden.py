"""Generate the committed golden fixtures from the reference's own test data.

Run in the BUILD container only (needs /root/reference):  python tests/golden/make_golden.py
The Exodus files are NetCDF-classic and are read with scipy; nothing here executes the
reference (it is Julia and cannot run in this container).  Outputs are small .npz files:

  poisson_g.npz   test/poisson/poisson.g  (coords, conn, node sets, side-set nodes) +
                  test/poisson/poisson.gold nodal variable `u`   (TestPoisson.jl:54-103)
  multi_block_quad4_tri3.npz   test/poisson/multi_block_mesh_quad4_tri3.g  (TestAssemblers.jl:44)
  cube_g.npz      examples/mechanics/cube.g  (8 HEX8)
  mechanics_coarse_g.npz  test/mechanics/mechanics_coarse.g
"""
import os
import numpy as np
from scipy.io import netcdf_file

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))

# Exodus side -> local node table (1-based sides, 0-based local nodes)
SIDE_NODES = {
    "QUAD4": [(0, 1), (1, 2), (2, 3), (3, 0)],
    "QUAD": [(0, 1), (1, 2), (2, 3), (3, 0)],
    "TRI3": [(0, 1), (1, 2), (2, 0)],
    "HEX8": [(0, 1, 5, 4), (1, 2, 6, 5), (2, 3, 7, 6), (0, 4, 7, 3), (0, 3, 2, 1), (4, 5, 6, 7)],
}


def names(var):
    out = []
    for row in var.data:
        s = b"".join(row).split(b"\x00")[0].decode().strip()
        out.append(s)
    return out


def read_exo(path):
    nc = netcdf_file(path, "r", mmap=False)
    d = {}
    nd = nc.dimensions["num_dim"]
    coords = np.stack([np.array(nc.variables["coord" + "xyz"[i]].data, dtype=float) for i in range(nd)])
    d["coords"] = coords
    nblk = nc.dimensions["num_el_blk"]
    conns, types = [], []
    for b in range(1, nblk + 1):
        v = nc.variables[f"connect{b}"]
        conns.append(np.array(v.data, dtype=np.int64).T.copy())  # (NNPE, NE) 1-based
        types.append(v.elem_type.decode().strip().upper())
    d["n_blocks"] = nblk
    bnames = names(nc.variables["eb_names"]) if "eb_names" in nc.variables else [""] * nblk
    ids = np.array(nc.variables["eb_prop1"].data)
    for b in range(nblk):
        d[f"conn_{b}"] = conns[b]
        d[f"type_{b}"] = types[b]
        d[f"block_name_{b}"] = bnames[b] if bnames[b] else f"block_{ids[b]}"
    # node sets
    nns = nc.dimensions.get("num_node_sets", 0) or 0
    nsn = names(nc.variables["ns_names"]) if nns else []
    nsid = np.array(nc.variables["ns_prop1"].data) if nns else []
    d["nodeset_names"] = np.array([nsn[i] if nsn[i] else f"nset_{nsid[i]}" for i in range(nns)])
    for i in range(nns):
        d[f"nodeset_{i}"] = np.array(nc.variables[f"node_ns{i+1}"].data, dtype=np.int64)
    # side sets -> unique node lists in order of first appearance (Julia `unique`)
    nss = nc.dimensions.get("num_side_sets", 0) or 0
    ssn = names(nc.variables["ss_names"]) if nss else []
    ssid = np.array(nc.variables["ss_prop1"].data) if nss else []
    d["sideset_names"] = np.array([ssn[i] if ssn[i] else f"sset_{ssid[i]}" for i in range(nss)])
    offs = np.cumsum([0] + [c.shape[1] for c in conns])
    for i in range(nss):
        el = np.array(nc.variables[f"elem_ss{i+1}"].data, dtype=np.int64)
        sd = np.array(nc.variables[f"side_ss{i+1}"].data, dtype=np.int64)
        nodes = []
        for e, s in zip(el, sd):
            b = int(np.searchsorted(offs, e - 1, side="right") - 1)
            loc = SIDE_NODES[types[b]][s - 1]
            nodes.extend(conns[b][list(loc), e - 1 - offs[b]].tolist())
        _, first = np.unique(nodes, return_index=True)
        d[f"sideset_nodes_{i}"] = np.array(nodes, dtype=np.int64)[np.sort(first)]
    return d, nc


if __name__ == "__main__":
    d, _ = read_exo(f"{REF}/test/poisson/poisson.g")
    g = netcdf_file(f"{REF}/test/poisson/poisson.gold", "r", mmap=False)
    d["gold_u"] = np.array(g.variables["vals_nod_var1"].data[0], dtype=float)
    np.savez_compressed(f"{OUT}/poisson_g.npz", **d)
    for src, dst in [("test/poisson/multi_block_mesh_quad4_tri3.g", "multi_block_quad4_tri3.npz"),
                     ("examples/mechanics/cube.g", "cube_g.npz"),
                     ("test/mechanics/mechanics_coarse.g", "mechanics_coarse_g.npz")]:
        d, _ = read_exo(f"{REF}/{src}")
        np.savez_compressed(f"{OUT}/{dst}", **d)
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(f"{OUT}/{f}"))
