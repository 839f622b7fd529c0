"""Pins the C restatement (oracle/fec_oracle_c.c, the CPU-baseline port) to the numpy oracle."""
import numpy as np
import pytest

import fec_oracle as O
import fec_oracle_clib as OC


def _rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("phys,el,nf", [("poisson", "hex", 1), ("neo", "hex", 3), ("neo_as_written", "hex", 3),
                                        ("linear", "hex", 3), ("j2", "tet10", 3), ("poisson", "quad", 1),
                                        ("linear", "quad", 2)])
def test_c_port_matches_numpy_oracle(phys, el, nf):
    rng = np.random.default_rng(0)
    if el == "hex":
        m, tabs = O.structured_mesh("hex", (0, 0, 0), (1, 1, 1), (5, 5, 5)), O.ref_fe_tables("HEX8", "gauss2")
    elif el == "quad":
        m, tabs = O.structured_mesh("quad", (0, 0), (1, 1), (7, 6)), O.ref_fe_tables("QUAD4", "gauss2")
    else:
        m, tabs = O.kuhn_tet10_mesh(2), O.ref_fe_tables("TETRA10", "tet4")
    X = m["coords"] + 0.02 * rng.standard_normal(m["coords"].shape)
    nd, nn = X.shape
    props = {"poisson": (), "j2": (1e3, 10e9, 1e9, 2e8, 1e8)}.get(phys, (1e3, 10e6, 1e6))
    src = lambda Xq: np.sin(Xq[:, 0]) + Xq[:, 1]
    ophys = {"poisson": O.Poisson(src), "neo": O.NeoHookean(nd), "neo_as_written": O.NeoHookean(nd, "as_written"),
             "linear": O.LinearElastic(nd), "j2": O.J2Plasticity(nd)}[phys]
    blk = O.Block(m["conn"], tabs, ophys, props=props)
    amp = 0.15 if phys == "j2" else 0.02
    U = amp * rng.standard_normal((nf, nn))
    V = rng.random((nf, nn))
    so = None
    if phys == "j2":
        so = 1e-4 * rng.standard_normal(blk.state_old.shape)
        so[2] = -so[0] - so[1]; so[6] = np.abs(so[6])
        blk.state_old[:] = so
    fq = None
    if phys == "poisson":
        x_el = np.transpose(X[:, m["conn"] - 1], (2, 1, 0))
        fq = np.stack([src(np.einsum("a,eai->ei", tabs[0][q], x_el)) for q in range(len(tabs[2]))], axis=1)  # (NE,NQ)
    cp = OC.CProblem(m["conn"], X, tabs, phys, nf, props, source_q=fq, state_old=so)
    Uf, Vf = U.reshape(-1, order="F"), V.reshape(-1, order="F")
    assert _rel(cp.assemble_vector(Uf, nthreads=2), O.assemble_vector([blk], X, U, nf)) < 1e-12
    if phys == "j2":
        assert _rel(cp.state_new_ref(), blk.state_new) < 1e-12
    coo = cp.assemble_matrix_coo(Uf, 2, nthreads=2)
    coo_ref = O.assemble_matrix_coo([blk], X, U, nf)
    assert _rel(coo, coo_ref) < 1e-12
    assert _rel(cp.assemble_matrix_coo(Uf, 3), O.assemble_matrix_coo([blk], X, U, nf, "mass")) < 1e-12
    assert _rel(cp.assemble_action(Uf, Vf, 2, nthreads=2), O.assemble_matrix_action([blk], X, U, V, nf)) < 1e-12
    # pattern + sparse! + CSR
    Is, Js = cp.pattern()
    pat = O.matrix_pattern([m["conn"]], nf)
    assert np.array_equal(Is, pat["Is"]) and np.array_equal(Js, pat["Js"])
    n = nf * nn
    ws = OC.SparseWorkspace(len(Is), n)
    slots = np.arange(1, len(Is) + 1, dtype=np.int64)
    colptr, rowval, nz = ws.sparse_csc(Is, Js, slots, coo)
    rcolptr, rrowval, rnz = O.sparse_csc(pat["Is"], pat["Js"], coo_ref, n)
    assert np.array_equal(colptr, rcolptr) and np.array_equal(rowval, rrowval)
    assert _rel(nz, rnz) < 1e-13
    rowptr, colval, nzr = ws.csr(colptr, rowval, nz)
    rrowptr, rcolval, rnzr = O.csc_to_csr(rcolptr, rrowval, rnz, n)
    assert np.array_equal(rowptr, rrowptr) and np.array_equal(colval, rcolval) and _rel(nzr, rnzr) < 1e-13


def test_at_size_helper_matches_numpy_assembler():
    """util_parity.c_oracle_reference (C oracle + vectorised _update_dofs!, used by tests/test_gpu_at_size.py at 48^3)
    against the numpy OracleAssembler on a small mesh: unknown residual, rowptr / colval bit-exact, values."""
    import fecb200 as F
    from util_parity import c_oracle_reference, perturb
    n = 5
    mesh = perturb(F.StructuredMesh("hex", (0, 0, 0), (1, 1, 1), (n + 1, n + 2, n + 1)), 0.1 / n)
    props = np.array([1e3, 10e6, 1e6])
    X = np.asarray(mesh.nodal_coords)
    nn = X.shape[1]
    bot, top = mesh.nodeset_nodes["bottom"], mesh.nodeset_nodes["top"]
    dd = np.concatenate([3 * (bot - 1) + 1, 3 * (bot - 1) + 2, 3 * (bot - 1) + 3, 3 * (top - 1) + 2])
    vals = np.concatenate([np.zeros(3 * len(bot)), np.full(len(top), 0.01)])
    blk = O.Block(mesh.element_conns["block_1"], O.ref_fe_tables("HEX8", "gauss2"), O.NeoHookean(3), props=props)
    oasm = O.OracleAssembler(X, [blk], 3, condensed=False, matrix_type="csr")
    oasm.update_dofs(dd)
    order = np.argsort(dd, kind="stable")
    oasm.bc_vals[:] = vals[order]
    Uu = 1e-3 * np.random.default_rng(0).standard_normal(oasm.n)
    oasm.assemble_vector(Uu)
    oasm.assemble_stiffness(Uu)
    rowptr, colval, nz = oasm.stiffness()
    ref = c_oracle_reference(mesh, "neo", props, dd, vals, Uu, nthreads=2)
    assert np.array_equal(ref["rowptr"], rowptr) and np.array_equal(ref["colval"], colval)
    assert _rel(ref["nz"], nz) < 1e-13 and _rel(ref["R"], oasm.residual()) < 1e-13
