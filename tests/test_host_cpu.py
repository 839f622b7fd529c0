"""CPU-only tests: host-side mirror logic, and that libfecb200.so loads and exports every symbol
include/fecb200.h declares (no compute calls without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

import fec_oracle as O
import fecb200 as F
from fecb200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def test_header_symbols_exported():
    hdr = open(os.path.join(ROOT, "include", "fecb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(fecb200_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 30
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in fecb200.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.fecb200_version() == 100


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = F.StructuredMesh("hex", (0, 0, 0), (1, 1, 1), (3, 3, 3))
    V = F.FunctionSpace(m, F.H1Field, F.Lagrange)
    asm = F.SparseMatrixAssembler(F.ScalarFunction(V, "u"))
    with pytest.raises(F.FECError, match="no CPU fallback"):
        F.create_parameters(m, asm, F.Poisson(None))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "finiteelementcontainers.jl_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".jl")):
                src = open(os.path.join(dp, f)).read()
                assert "fec_oracle" not in src and "oracle/" not in src, f"{f} references the oracle"


@pytest.mark.parametrize("el,counts", [("hex", (4, 3, 5)), ("quad", (4, 6)), ("tri", (3, 5))])
def test_structured_mesh_matches_oracle(el, counts):
    nd = len(counts)
    m = F.StructuredMesh(el, (0.0,) * nd, (1.0,) * nd, counts)
    o = O.structured_mesh(el, (0.0,) * nd, (1.0,) * nd, counts)
    if el != "hex" or len(set(counts)) == 1:
        assert np.array_equal(np.asarray(m.nodal_coords), o["coords"])
    assert np.array_equal(m.element_conns["block_1"], o["conn"])
    for k, v in o["nodesets"].items():
        assert np.array_equal(m.nodeset_nodes[k], v), k


def test_structured_mesh_known_answers():
    # test/TestMesh.jl:96-128
    m = F.StructuredMesh("quad", (0., 0.), (1., 1.), (3, 3))
    assert np.array_equal(m.element_conns["block_1"], [[1, 4, 2, 5], [2, 5, 3, 6], [5, 8, 6, 9], [4, 7, 5, 8]])
    m = F.StructuredMesh("tri", (0., 0.), (1., 1.), (3, 3))
    assert np.array_equal(m.element_conns["block_1"],
                          [[1, 1, 4, 4, 2, 2, 5, 5], [2, 5, 5, 8, 3, 6, 6, 9], [5, 4, 8, 7, 6, 5, 9, 8]])
    with pytest.raises(ValueError):
        F.StructuredMesh("bad element", (0., 0.), (1., 1.), (3, 3))
    with pytest.raises(IndexError):
        F.StructuredMesh("tri3", (0., 0.), (0., 1.), (3, 3))
    with pytest.raises(AssertionError):
        F.StructuredMesh("tet", (0., 0., 0.), (1., 1., 1.), (3, 3, 3))


def test_hex_coords_cube():
    m = F.StructuredMesh("hex", (0, 0, 0), (1, 1, 1), (3, 3, 3))
    o = O.structured_mesh("hex", (0, 0, 0), (1, 1, 1), (3, 3, 3))
    assert np.array_equal(np.asarray(m.nodal_coords), o["coords"])


def test_tet10_mesh_matches_oracle():
    m = F.KuhnTet10Mesh(2)
    o = O.kuhn_tet10_mesh(2)
    assert np.array_equal(np.asarray(m.nodal_coords), o["coords"])
    assert np.array_equal(m.element_conns["block_1"], o["conn"])


@pytest.mark.parametrize("el,rule,qt,qd", [("HEX8", "gauss2", "GaussLegendre", 2), ("HEX8", "gll2", "GaussLobattoLegendre", 2),
                                            ("HEX8", "gll3", "GaussLobattoLegendre", 3), ("QUAD4", "gauss2", "GaussLegendre", 2),
                                            ("TRI3", "tri3", "GaussLegendre", 2), ("TETRA4", "tet4", "GaussLegendre", 2),
                                            ("TETRA10", "tet4", "GaussLegendre", 2), ("TETRA10", "tet1", "GaussLegendre", 1)])
def test_reference_tables_match_oracle(el, rule, qt, qd):
    rf = F.ReferenceFE(el, qt, qd)
    N, dN, w = O.ref_fe_tables(el, rule)
    assert np.allclose(rf.N, N, atol=1e-15) and np.allclose(rf.dN, dN, atol=1e-15) and np.allclose(rf.w, w, atol=1e-15)


def test_h1field_layout():
    # src/Fields.jl:36-40: data[(n-1)*NF + d]
    f = F.H1Field(np.arange(6.0).reshape(2, 3))
    assert f[1, 2] == 5.0
    assert list(f.data_flat) == [0, 3, 1, 4, 2, 5]
    z = F.H1Field.zeros(3, 4)
    z.data_flat[3 * 2 + 1] = 7.0
    assert z[1, 2] == 7.0


def test_connectivity_offsets():
    # test/TestFields.jl:1-37
    a = np.arange(1, 13).reshape(4, 3, order="F")
    b = np.arange(1, 7).reshape(3, 2, order="F")
    c = F.Connectivity([a, b])
    assert c.offsets == [1, 13] and c.nepes == [4, 3] and c.nelems == [3, 2]
    assert np.array_equal(c.block(1), b) and np.array_equal(c.block(0), a)


def test_unstructured_fixture_and_bcs():
    m = F.UnstructuredMesh(os.path.join(GOLDEN, "multi_block_quad4_tri3.npz"))
    assert m.num_nodes() == 406 and [m.element_conns[b].shape for b in m.element_block_names] == [(4, 280), (3, 170)]
    assert set(m.element_types.values()) == {"QUAD4", "TRI3"}
    V = F.FunctionSpace(m, F.H1Field, F.Lagrange)
    u = F.VectorFunction(V, "displ")
    assert u.names() == ["displ_x", "displ_y"]
    dof = F.DofManager(u)
    bcs = F.DirichletBCs(m, dof, [F.DirichletBC("displ_x", lambda X, t: 0.0, sideset_name="boundary"),
                                  F.DirichletBC("displ_y", lambda X, t: 0.1 * t, sideset_name="boundary")])
    dd = bcs.dirichlet_dofs()
    assert len(dd) == 160 and np.all(np.diff(dd) > 0)
    bcs.update_bc_values(m.nodal_coords, 2.0)
    assert np.allclose(bcs.vals[:80], 0.0) and np.allclose(bcs.vals[80:], 0.2)
    with pytest.raises(ValueError):
        F.DirichletBC("u", lambda X, t: 0.0)
    with pytest.raises(ValueError):
        F.DirichletBC("u", lambda X, t: 0.0, nodeset_name="a", sideset_name="b")


def test_element_function_tokens():
    from fecb200.physics import kind_of
    assert kind_of(F.residual, (_lib.RESIDUAL,)) == _lib.RESIDUAL
    with pytest.raises(TypeError, match="no CPU fallback"):
        kind_of(lambda *a: 0, (_lib.RESIDUAL,))
    with pytest.raises(ValueError):
        kind_of(F.stiffness, (_lib.RESIDUAL,))


# ---- side sets, surface tables and the load containers of the host mirror (no GPU needed) ----------------------

@pytest.mark.parametrize("el,counts", [("quad", (6, 4)), ("tri", (5, 7)), ("hex", (4, 3, 5))])
def test_structured_sidesets_match_the_oracle_and_the_nodesets(el, counts):
    """QUAD4 / TRI3: the reference's own loops (src/meshes/StructuredMesh.jl:257-330, 355-431); HEX8: the faces that lie on
    the named boundary.  Every side's nodes are in the node set of the same name, and the generic finder
    (_sidesets_from_nodesets, used for meshes without side-set records) recovers the same sides."""
    nd = len(counts)
    m = F.StructuredMesh(el, (0.,) * nd, (1.,) * nd, counts)
    osets = O.structured_sidesets(el, counts)
    conn = m.element_conns["block_1"]
    found = m._sidesets_from_nodesets(m.nodeset_nodes)
    for name, (oe, os_) in osets.items():
        assert np.array_equal(m.sideset_elems[name], oe) and np.array_equal(m.sideset_sides[name], os_)
        sn = O.side_nodes(m.element_types["block_1"], conn, oe, os_)
        assert np.array_equal(m.sideset_side_nodes[name], sn)
        assert set(np.unique(sn)) == set(m.nodeset_nodes[name].tolist())
        fe, fs = found[name]
        assert sorted(zip(fe.tolist(), fs.tolist())) == sorted(zip(oe.tolist(), os_.tolist()))


@pytest.mark.parametrize("et", ["QUAD4", "TRI3", "HEX8", "TETRA4", "TETRA10"])
def test_surface_tables_match_the_oracle_and_integrate_exactly(et):
    """the host's surface tables (what a Julia host reads from ReferenceFiniteElements) against the oracle's: partition of
    unity, zero-sum gradients, weights = measure of the reference side (2 for [-1,1], 4 for [-1,1]^2, 1/2 for the triangle)"""
    Ns, dNs, ws = F.ReferenceFE(et).surface_tables()
    oN, odN, ow = O.surface_tables(et, "gauss2")
    assert np.allclose(Ns, oN, atol=1e-15) and np.allclose(dNs, odN, atol=1e-15) and np.allclose(ws, ow, atol=1e-15)
    assert np.allclose(Ns.sum(axis=1), 1.0) and np.allclose(dNs.sum(axis=1), 0.0, atol=1e-14)
    assert np.isclose(ws.sum(), {"QUAD4": 2.0, "TRI3": 2.0, "HEX8": 4.0, "TETRA4": 0.5, "TETRA10": 0.5}[et])


def test_neumann_and_source_containers_evaluate_at_the_oracles_points():
    """NeumannBCs / Sources (src/bcs/NeumannBCs.jl:60-71,157-171, Sources.jl:55-66): vals[d, q, e] = func(X_q, t)[d] at the
    same quadrature points, in the same [NF, NQ, n] layout, as the oracle's restatement."""
    m = F.StructuredMesh("hex", (0., 0., 0.), (1., 2., 1.5), (4, 3, 5))
    rng = np.random.default_rng(2)
    X = np.asarray(m.nodal_coords)
    X += rng.uniform(-0.02, 0.02, X.shape)
    V = F.FunctionSpace(m, F.H1Field, F.Lagrange)
    dof = F.DofManager(F.VectorFunction(V, "displ"))
    f = lambda X, t: np.stack([X[:, 0] + t, X[:, 1] * X[:, 2], np.sin(X[:, 0]) - t], axis=1)
    nb = F.NeumannBCs(m, dof, [F.NeumannBC("displ_y", f, "top"), F.NeumannBC("displ_x", f, "left")])
    nb.update_bc_values(X, 0.3)
    tabs = O.surface_tables("HEX8", "gauss2")
    for c, name in zip(nb.bc_caches, ("top", "left")):
        sn = np.asarray(m.sideset_side_nodes[name])
        Xq = O.surface_quadrature_points(sn, tabs, X)                          # (nqs, nsides, ND)
        want = f(Xq.reshape(-1, 3), 0.3).reshape(Xq.shape[0], Xq.shape[1], 3).transpose(2, 0, 1)
        assert c["vals"].shape == want.shape and np.allclose(c["vals"], want, rtol=1e-14, atol=1e-14)
        assert c["vals"].flags.f_contiguous                                     # [NF, nqs, nsides] as the ABI expects
    src = F.Sources(m, dof, [F.Source("displ_x", f, "block_1")])
    src.update_source_values(V, 0.3)
    blk = O.Block(m.element_conns["block_1"], O.ref_fe_tables("HEX8", "gauss2"), O.LinearElastic(3))
    Xq = O.cell_quadrature_points(blk, X)
    want = f(Xq.reshape(-1, 3), 0.3).reshape(Xq.shape[0], Xq.shape[1], 3).transpose(2, 0, 1)
    assert np.allclose(src.vals[0], want, rtol=1e-14, atol=1e-14)
    with pytest.raises(ValueError):
        F.NeumannBCs(m, dof, [F.NeumannBC("nope", f, "top")])
    with pytest.raises(KeyError):
        F.Sources(m, dof, [F.Source("displ_x", f, "block_7")])
    # scalar-valued functions broadcast for NF = 1
    dof1 = F.DofManager(F.ScalarFunction(V, "u"))
    nb1 = F.NeumannBCs(m, dof1, [F.NeumannBC("u", lambda X, t: -1.0, "right")])
    nb1.update_bc_values(X, 0.0)
    assert nb1.bc_caches[0]["vals"].shape[0] == 1 and np.all(nb1.bc_caches[0]["vals"] == -1.0)


def test_field_helpers_of_the_dofmanager():
    """update_field_unknowns! / extract_field_unknowns! / update_field_dirichlet_bcs! on host fields
    (src/DofManagers.jl:203-213, 349-411; src/bcs/DirichletBCs.jl:411-418) against the oracle's _update_field."""
    m = F.StructuredMesh("quad", (0., 0.), (1., 1.), (5, 4))
    V = F.FunctionSpace(m, F.H1Field, F.Lagrange)
    u = F.VectorFunction(V, "displ")
    rng = np.random.default_rng(4)
    for condensed in (False, True):
        dof = F.DofManager(u, use_condensed=condensed)
        dbcs = F.DirichletBCs(m, dof, [F.DirichletBC("displ_x", lambda X, t: X[:, 1] + t, nodeset_name="left"),
                                       F.DirichletBC("displ_y", lambda X, t: np.full(len(X), 2.0), nodeset_name="left"),
                                       F.DirichletBC("displ_x", lambda X, t: np.full(len(X), 7.0), nodeset_name="bottom")])
        dbcs.update_bc_values(m.nodal_coords, 0.5)
        dd = dbcs.dirichlet_dofs()
        od = O.update_dofs(2, m.num_nodes(), dd)
        dof.unknown_dofs, dof.dof_to_unknown, dof.dirichlet_dofs = od["unknown_dofs"], od["dof_to_unknown"], dd
        n = len(dof) if condensed else len(dof.unknown_dofs)
        Uu = rng.standard_normal(n)
        U = F.H1Field.zeros(2, m.num_nodes())
        F.update_field_dirichlet_bcs(U, dbcs)
        F.update_field_unknowns(U, dof, Uu)
        blk = O.Block(m.element_conns["block_1"], O.ref_fe_tables("QUAD4", "gauss2"), O.LinearElastic(2))
        oasm = O.OracleAssembler(np.asarray(m.nodal_coords), [blk], 2, condensed=condensed)
        oasm.update_dofs(dd)
        vals = np.zeros(len(dd))
        for d, v in zip(dbcs.dofs, dbcs.vals):                 # later BCs win on shared dofs (bottom-left corner)
            vals[np.searchsorted(dd, d)] = v
        oasm.bc_vals[:] = vals
        oasm._update_field(oasm.field, Uu)
        assert np.array_equal(U.data_flat, oasm.field)
        back = np.zeros(n)
        F.extract_field_unknowns(back, dof, U)
        ud = dof.unknown_dofs - 1
        assert np.array_equal(back[:len(ud)], U.data_flat[ud])
        assert U[0, m.nodeset_nodes["bottom"][0] - 1] == 7.0      # corner node: the bottom BC was applied last


def test_postprocessor_writes_a_readable_exodus_file(tmp_path):
    """PostProcessor / write_times / write_field / close (src/PostProcessors.jl:27-49, src/meshes/Exodus.jl:148-282), the
    output side of the reference's regression tests (TestPoisson.jl:88-97): the file is Exodus II over NetCDF classic and
    reads back -- mesh, sets and the nodal field -- through this package's own Exodus reader."""
    from scipy.io import netcdf_file
    g = np.load(os.path.join(GOLDEN, "poisson_g.npz"))
    mesh = F.UnstructuredMesh(os.path.join(GOLDEN, "poisson_g.npz"))
    V = F.FunctionSpace(mesh, F.H1Field, F.Lagrange)
    u = F.ScalarFunction(V, "u")
    out = str(tmp_path / "poisson_out.e")
    pp = F.PostProcessor(mesh, out, u)
    F.write_times(pp, 1, 0.0)
    F.write_field(pp, 1, ("u",), F.H1Field(g["gold_u"]))
    F.write_times(pp, 2, 0.5)
    F.write_field(pp, 2, ("u",), F.H1Field(2.0 * g["gold_u"]))
    with pytest.raises(KeyError):
        F.write_field(pp, 1, ("v",), F.H1Field(g["gold_u"]))
    F.close(pp)
    assert open(out, "rb").read(4) == b"CDF\x02"
    back = F.UnstructuredMesh(out)
    assert np.array_equal(np.asarray(back.nodal_coords), np.asarray(mesh.nodal_coords))
    assert back.element_block_names == mesh.element_block_names and back.element_types == mesh.element_types
    assert np.array_equal(back.element_conns["block_1"], mesh.element_conns["block_1"])
    for k in mesh.sideset_elems:
        assert np.array_equal(back.sideset_elems[k], mesh.sideset_elems[k]) and np.array_equal(back.sideset_sides[k], mesh.sideset_sides[k])
    nc = netcdf_file(out, "r", mmap=False)
    assert b"".join(nc.variables["name_nod_var"].data[0]).split(b"\x00")[0] == b"u"
    assert np.allclose(nc.variables["time_whole"].data, [0.0, 0.5])
    vals = np.array(nc.variables["vals_nod_var1"].data)
    assert vals.shape == (2, mesh.num_nodes()) and np.array_equal(vals[0], g["gold_u"]) and np.array_equal(vals[1], 2.0 * g["gold_u"])
    assert nc.variables["connect1"].elem_type.decode().strip() == "QUAD4"
    nc.close()
    # vector functions: one variable per component, 3-D structured mesh with generated side sets
    m3 = F.StructuredMesh("hex", (0., 0., 0.), (1., 1., 1.), (3, 4, 3))
    d3 = F.VectorFunction(F.FunctionSpace(m3, F.H1Field, F.Lagrange), "displ")
    out3 = str(tmp_path / "cube.exo")
    pp = F.PostProcessor(m3, out3, d3, extra_nodal_names=["pressure"])
    assert pp.nodal_names == ["displ_x", "displ_y", "displ_z", "pressure"]
    U = F.H1Field(np.arange(3.0 * m3.num_nodes()).reshape(3, -1, order="F"))
    F.write_field(pp, 1, d3.names(), U)
    pp.close()
    b3 = F.UnstructuredMesh(out3)
    assert np.array_equal(b3.element_conns["block_1"], m3.element_conns["block_1"])
    assert set(b3.sideset_elems) == set(m3.sideset_elems)
    nc = netcdf_file(out3, "r", mmap=False)
    assert np.array_equal(np.array(nc.variables["vals_nod_var2"].data)[0], np.asarray(U)[1])
    nc.close()
    with pytest.raises(RuntimeError):
        F.PostProcessor(m3, str(tmp_path / "cube.vtk"), d3)


def test_periodic_bc_container_matches_nodes_like_the_reference():
    """PeriodicBCs (src/bcs/PeriodicBCs.jl:17-110): side-a nodes sorted and unique, side-b nodes matched through the
    coordinate along the sides; the pairs feed update_dofs! (same maps as the oracle's), jumps come from func(X_b, t)."""
    m = F.StructuredMesh("quad", (0., 0.), (1., 1.), (6, 5))
    V = F.FunctionSpace(m, F.H1Field, F.Lagrange)
    dof = F.DofManager(F.ScalarFunction(V, "u"))
    X = np.asarray(m.nodal_coords)
    pb = F.PeriodicBCs(m, dof, [F.PeriodicBC("u", "y", lambda X, t: 0.25 * X[:, 1] + t, "left", "right"),
                                F.PeriodicBC("u", "x", lambda X, t: np.zeros(len(X)), "bottom", "top")])
    c = pb.bc_caches[0]
    assert np.array_equal(c["side_a_nodes"], np.sort(m.nodeset_nodes["left"]))
    assert np.allclose(X[1, c["side_a_nodes"] - 1], X[1, c["side_b_nodes"] - 1])          # matched by y along the side
    assert np.allclose(X[0, c["side_b_nodes"] - 1], 1.0) and np.allclose(X[0, c["side_a_nodes"] - 1], 0.0)
    c2 = pb.bc_caches[1]
    assert np.allclose(X[0, c2["side_a_nodes"] - 1], X[0, c2["side_b_nodes"] - 1]) and np.allclose(X[1, c2["side_b_nodes"] - 1], 1.0)
    pa, pbd = pb.periodic_dofs()
    assert len(pa) == len(pbd) == 6 + 5 and np.array_equal(pa[:5], c["side_a_dofs"])
    od = O.update_dofs(1, m.num_nodes(), [], pa, pbd)                                       # the chains resolve (corner nodes)
    assert len(od["unknown_dofs"]) == m.num_nodes() - len(np.unique(pbd))
    pb.update_bc_values(X, 0.5)
    assert np.allclose(pb.bc_caches[0]["vals"], 0.25 * X[1, c["side_b_nodes"] - 1] + 0.5)
    with pytest.raises(AssertionError):
        F.PeriodicBCs(m, dof, [F.PeriodicBC("u", "z", lambda X, t: 0.0, "left", "right")])
    # 3-D: both transverse coordinates take part in the match
    m3 = F.StructuredMesh("hex", (0., 0., 0.), (1., 1., 1.), (4, 3, 5))
    d3 = F.DofManager(F.VectorFunction(F.FunctionSpace(m3, F.H1Field, F.Lagrange), "displ"))
    p3 = F.PeriodicBCs(m3, d3, [F.PeriodicBC("displ_y", "y", lambda X, t: 0.0, "left", "right")])
    X3 = np.asarray(m3.nodal_coords)
    a, b = p3.bc_caches[0]["side_a_nodes"], p3.bc_caches[0]["side_b_nodes"]
    assert np.allclose(X3[1:, a - 1], X3[1:, b - 1]) and len(np.unique(b)) == len(b) == 15
    assert np.array_equal(p3.bc_caches[0]["side_b_dofs"], 3 * (b - 1) + 2)


def test_initial_conditions_and_accessors():
    """InitialCondition(s) (src/InitialConditions.jl) and the small accessor functions the reference exports."""
    m = F.StructuredMesh("hex", (0., 0., 0.), (1., 1., 1.), (3, 4, 3))
    V = F.FunctionSpace(m, F.H1Field, F.Lagrange)
    u = F.VectorFunction(V, "displ")
    dof = F.DofManager(u)
    ics = F.InitialConditions(m, dof, [F.InitialCondition("displ_y", lambda X: 0.1 * X[:, 1], block_name="block_1"),
                                      F.InitialCondition("displ_x", lambda X: 2.0, nodeset_name="top")])
    F.update_ic_values(ics, m.nodal_coords)
    U = F.create_field_like(dof) if hasattr(F, "create_field_like") else F.H1Field.zeros(3, m.num_nodes())
    F.update_field_ics(U, ics)
    X = np.asarray(m.nodal_coords)
    assert np.allclose(U[1], 0.1 * X[1]) and np.all(U[2] == 0.0)
    top = m.nodeset_nodes["top"] - 1
    assert np.all(U[0, top] == 2.0) and np.count_nonzero(U[0]) == len(top)
    with pytest.raises(ValueError):
        F.InitialCondition("displ_x", lambda X: 0.0)
    assert F.num_dimensions(m) == 3 and F.num_nodes(m) == 36 and F.element_blocks(m) == ["block_1"]
    assert F.num_fields(u) == 3 and F.num_fields(U) == 3 and F.num_entities(U) == 36 and F.num_fields(dof) == 3
    assert F.connectivity(V.elem_conns, 1).shape == (8, 12) and F.num_elements(V) == 12 and F.num_elements(V, 1) == 12
    assert F.current_time(F.TimeStepper(0.5, 1.0, 5)) == 0.5
    assert np.array_equal(F.nodal_coordinates(m), m.nodal_coords) and "top" in F.nodesets(m) and "top" in F.sidesets(m)


def test_reference_bc_tests_restated():
    """test/TestBCs.jl:21-103 on the same mesh (poisson.g): DirichletBC input handling, container construction on a block /
    node set / side set, the unknown-variable error, update_bc_values! with bc_func(_, t) = 2 t^2 at t = 3 -> 18, and
    update_field_dirichlet_bcs!."""
    mesh = F.UnstructuredMesh(os.path.join(GOLDEN, "poisson_g.npz"))
    fspace = F.FunctionSpace(mesh, F.H1Field, F.Lagrange)
    f1 = lambda X, t: 5.0 * t
    with pytest.raises(ValueError):
        F.DirichletBC("u", f1)
    with pytest.raises(ValueError):
        F.DirichletBC("u", f1, block_name="some_block", nodeset_name="some_nodeset")
    bc = F.DirichletBC("my_var", f1, block_name="my_block")
    assert (bc.block_name, bc.nset_name, bc.sset_name, bc.var_name, bc.func) == ("my_block", None, None, "my_var", f1)
    bc = F.DirichletBC("my_var", f1, nodeset_name="my_nset")
    assert (bc.block_name, bc.nset_name, bc.sset_name) == (None, "my_nset", None)
    bc = F.DirichletBC("my_var", f1, sideset_name="my_sset")
    assert (bc.block_name, bc.nset_name, bc.sset_name) == (None, None, "my_sset")
    dof = F.DofManager(F.VectorFunction(fspace, "displ"))
    F.DirichletBCs(mesh, dof, [F.DirichletBC("displ_x", f1, block_name="block_1")])
    F.DirichletBCs(mesh, dof, [F.DirichletBC("displ_x", f1, sideset_name="sset_1")])
    with pytest.raises(ValueError):
        F.DirichletBCs(mesh, dof, [F.DirichletBC("bad_var_name", f1, sideset_name="sset_1")])
    bcs = F.DirichletBCs(mesh, dof, [F.DirichletBC("displ_x", lambda X, t: 2.0 * t ** 2, sideset_name="sset_1")])
    bcs.update_bc_values(mesh.nodal_coords, 3.0)
    assert np.allclose(bcs.vals, 18.0)
    U = F.create_field(dof)
    F.update_field_dirichlet_bcs(U, bcs)
    dd = F.dirichlet_dofs(bcs)
    assert np.allclose(U.data_flat[dd - 1], 18.0) and np.count_nonzero(U.data_flat) == len(dd)
    assert np.all((dd - 1) % 2 == 0)                      # displ_x dofs only: NF*(n-1) + 1


def test_reference_function_dof_and_ic_tests_restated():
    """test/TestFunctions.jl:1-45 (component names of every function type), test/TestDofManagers.jl:17-27 (sizes) and
    test/TestICs.jl:9-60 (InitialCondition input, container init + update on block / node set / side set), on poisson.g."""
    mesh = F.UnstructuredMesh(os.path.join(GOLDEN, "poisson_g.npz"))
    fspace = F.FunctionSpace(mesh, F.H1Field, F.Lagrange)
    u = F.ScalarFunction(fspace, "u")
    assert len(u) == 1 and u.names() == ["u"]
    u = F.VectorFunction(fspace, "u")
    assert len(u) == 2 and u.names() == ["u_x", "u_y"]
    u = F.TensorFunction(fspace, "u")
    assert len(u) == 9 and u.names() == ["u_xx", "u_yy", "u_zz", "u_yz", "u_xz", "u_xy", "u_zy", "u_zx", "u_yx"]
    u = F.TensorFunction(fspace, "u", use_spatial_dimension=True)
    assert len(u) == 4 and u.names() == ["u_xx", "u_yy", "u_xy", "u_yx"]
    u = F.SymmetricTensorFunction(fspace, "u")
    assert len(u) == 6 and u.names() == ["u_xx", "u_yy", "u_zz", "u_yz", "u_xz", "u_xy"]
    u = F.SymmetricTensorFunction(fspace, "u", use_spatial_dimension=True)
    assert len(u) == 3 and u.names() == ["u_xx", "u_yy", "u_xy"]
    g = F.GeneralFunction(F.VectorFunction(fspace, "u"), F.ScalarFunction(fspace, "t"))
    assert len(g) == 3 and g.names() == ["u_x", "u_y", "t"]
    dof1 = F.DofManager(F.VectorFunction(fspace, "u"))
    assert dof1.size() == (2, 16641) and len(dof1) == 2 * 16641 and F.create_field(dof1).shape == (2, 16641)
    # ICs
    dof = F.DofManager(F.VectorFunction(fspace, "displ"))
    f3 = lambda X: 3.0
    ic = F.InitialCondition("my_var", f3, block_name="my_block")
    assert (ic.block_name, ic.nset_name, ic.sset_name, ic.func, ic.var_name) == ("my_block", None, None, f3, "my_var")
    for kw, nodes in (({"block_name": "block_1"}, np.arange(1, 16642)),
                      ({"nodeset_name": "nset_1"}, mesh.nodeset_nodes.get("nset_1")),
                      ({"sideset_name": "sset_1"}, mesh.sideset_nodes["sset_1"])):
        if nodes is None:
            continue
        ics = F.InitialConditions(mesh, dof, [F.InitialCondition("displ_x", f3, **kw)])
        U = F.create_field(dof)
        F.update_ic_values(ics, mesh.nodal_coords)
        assert np.all(ics.ic_caches[0]["vals"] == 3.0)
        F.update_field_ics(U, ics)
        assert np.all(U[0, np.asarray(nodes) - 1] == 3.0) and np.all(U[1] == 0.0)


def test_reference_field_tests_restated():
    """test/TestFields.jl:1-98: Connectivity block views / per-element connectivity with the reference's 1-based offsets,
    and H1Field indexing (flat column-major `field[n]`, dual `field[d, n]`), fill and similar."""
    conns_in = [np.array([[1, 5, 9], [2, 6, 10], [3, 7, 11], [4, 8, 12]]),
                np.array([[13, 16, 19, 22, 25], [14, 17, 20, 23, 26], [15, 18, 21, 24, 27]])]
    conn = F.Connectivity(conns_in)
    b1, b2 = F.connectivity(conn, 1), F.connectivity(conn, 2)
    assert b1.shape == (4, 3) and b2.shape == (3, 5)
    elem = lambda nnpe, e, off: conn.data[off - 1 + (e - 1) * nnpe: off - 1 + e * nnpe].tolist()   # Fields.jl:183-188
    assert conn.offsets == [1, 13]
    assert [elem(4, e, 1) for e in (1, 2, 3)] == [[1, 2, 3, 4], [5, 6, 7, 8], [9, 10, 11, 12]]
    assert [elem(3, e, 13) for e in (1, 2, 3, 4, 5)] == [[13, 14, 15], [16, 17, 18], [19, 20, 21], [22, 23, 24], [25, 26, 27]]
    rng = np.random.default_rng(1)
    data = rng.random((2, 20))
    field = F.H1Field(data)
    assert field.dtype == data.dtype and field.ndim == 2 and field.shape == data.shape
    assert F.num_fields(field) == 2 and F.num_entities(field) == 20
    assert np.array_equal(field.data_flat, data.reshape(-1, order="F"))            # field[n] == data[n] (column-major)
    assert all(field[d, n] == data[d, n] for n in range(20) for d in range(2))
    data2 = rng.random((2, 20))
    for n in range(20):
        for d in range(2):
            field[d, n] = data2[d, n]
    assert np.array_equal(np.asarray(field), data2) and np.array_equal(field.data_flat, data2.reshape(-1, order="F"))
    field.fill(3.9)
    assert np.all(field == 3.9)
    new = F.H1Field(np.empty_like(field))
    new[...] = field
    assert type(new) is type(field) and np.all(new == field)


def test_reference_physics_test_restated():
    """test/TestPhysics.jl:1-8: num_fields / num_properties / num_states of an AbstractPhysics{1, 2, 3}; and the shipped
    physics' parameter counts (TestPoissonCommon.jl:4, TestMechanicsCommon.jl:3, TestMechanicsWithState.jl:15)."""
    class MyPhysics(F.AbstractPhysics):
        NF, NP, NS = 1, 2, 3
    ph = MyPhysics()
    assert (F.num_fields(ph), F.num_properties(ph), F.num_states(ph)) == (1, 2, 3)
    assert len(F.create_initial_state(ph)) == 3
    assert (F.num_fields(F.Poisson(None)), F.num_properties(F.Poisson(None)), F.num_states(F.Poisson(None))) == (1, 0, 0)
    mech = F.Mechanics(F.ThreeDimensional())
    assert F.num_fields(mech) == 3 and F.num_states(mech) == 0 and len(F.create_properties(mech)) == F.num_properties(mech) == 3
    j2 = F.J2Plasticity(F.ThreeDimensional())
    assert F.num_states(j2) == 7 and len(F.create_initial_state(j2)) == 7
