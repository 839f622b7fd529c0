"""PostProcessor (src/PostProcessors.jl:15-82, src/meshes/Exodus.jl:148-282): nodal-field output to an Exodus II file.

    pp = PostProcessor(mesh, "out.e", u)          # copy_mesh + nodal variable names of the functions
    write_times(pp, 1, 0.0)
    write_field(pp, 1, ("u",), p.field)           # H1Field (NF, NN): one nodal variable per component
    close(pp)                                      # pp.close()

The reference writes through Exodus.jl (libexodus); the image has neither, so the file is produced directly as NetCDF
classic (64-bit offset) with scipy -- the same container format the reference's meshes use (`CDF\\x02`), laid out per
the Exodus II schema (dimensions num_dim / num_nodes / num_elem / num_el_blk / time_step, variables coordx.., connectN,
eb_prop1, time_whole, name_nod_var, vals_nod_varN, node / side sets), so exodiff-style tools and this package's own
reader (meshes.UnstructuredMesh) can read it back.  Host-side IO only; nothing here is on the assembly path."""
from __future__ import annotations

import numpy as np

from .fields import H1Field

_LEN_STRING, _LEN_LINE = 33, 81
_EXO_TYPE = {"QUAD4": "QUAD4", "TRI3": "TRI3", "HEX8": "HEX8", "TETRA4": "TETRA4", "TETRA10": "TETRA10"}


def _chars(strings, width):
    out = np.zeros((len(strings), width), dtype="S1")
    for i, s in enumerate(strings):
        b = s.encode()[: width - 1]
        out[i, : len(b)] = np.frombuffer(b, dtype="S1")
    return out


class PostProcessor:
    """PostProcessor(mesh, output_file, vars...; extra_nodal_names)"""

    def __init__(self, mesh, output_file, *vars, extra_nodal_names=()):
        from scipy.io import netcdf_file
        if not (".e" in output_file or ".exo" in output_file):
            raise RuntimeError(f"Unsupported file type with extension {output_file.rsplit('.', 1)[-1]}")
        self.output_file_name = output_file
        self.nodal_names = [n for v in vars for n in v.names()] + list(extra_nodal_names)
        self.element_names, self.global_names = [], []
        nc = netcdf_file(output_file, "w", version=2)
        self.field_output_db = nc
        X = np.asarray(mesh.nodal_coords)
        nd, nn = X.shape
        blocks = list(mesh.element_block_names)
        ne = sum(mesh.element_conns[b].shape[1] for b in blocks)
        nc.title = b"fecb200 PostProcessor"
        nc.api_version = np.float32(8.03)
        nc.version = np.float32(8.03)
        nc.floating_point_word_size = np.int32(8)
        nc.file_size = np.int32(1)
        nc.int64_status = np.int32(0)
        nc.createDimension("time_step", None)          # the record dimension has to come first (scipy's classic writer)
        nc.createDimension("len_string", _LEN_STRING)
        nc.createDimension("len_line", _LEN_LINE)
        nc.createDimension("four", 4)
        nc.createDimension("num_dim", nd)
        nc.createDimension("num_nodes", nn)
        nc.createDimension("num_elem", ne)
        nc.createDimension("num_el_blk", len(blocks))
        # ---- copy_mesh: coordinates, blocks, node sets, side sets
        for i in range(nd):
            v = nc.createVariable("coord" + "xyz"[i], "d", ("num_nodes",))
            v[:] = X[i]
        v = nc.createVariable("coor_names", "c", ("num_dim", "len_string"))
        v[:] = _chars(list("xyz"[:nd]), _LEN_STRING)
        v = nc.createVariable("eb_names", "c", ("num_el_blk", "len_string"))
        v[:] = _chars(blocks, _LEN_STRING)
        v = nc.createVariable("eb_status", "i", ("num_el_blk",))
        v[:] = np.ones(len(blocks), dtype=np.int32)
        v = nc.createVariable("eb_prop1", "i", ("num_el_blk",))
        v[:] = np.arange(1, len(blocks) + 1, dtype=np.int32)
        v.name = b"ID"
        for b, name in enumerate(blocks):
            c = mesh.element_conns[name]
            nc.createDimension(f"num_el_in_blk{b + 1}", c.shape[1])
            nc.createDimension(f"num_nod_per_el{b + 1}", c.shape[0])
            v = nc.createVariable(f"connect{b + 1}", "i", (f"num_el_in_blk{b + 1}", f"num_nod_per_el{b + 1}"))
            v[:] = np.ascontiguousarray(c.T, dtype=np.int32)
            v.elem_type = _EXO_TYPE[mesh.element_types[name]].encode()
        nsets = {k: np.asarray(v_, dtype=np.int64) for k, v_ in mesh.nodeset_nodes.items() if not k.startswith("__")}
        if nsets:
            nc.createDimension("num_node_sets", len(nsets))
            v = nc.createVariable("ns_names", "c", ("num_node_sets", "len_string"))
            v[:] = _chars(list(nsets), _LEN_STRING)
            v = nc.createVariable("ns_status", "i", ("num_node_sets",))
            v[:] = np.ones(len(nsets), dtype=np.int32)
            v = nc.createVariable("ns_prop1", "i", ("num_node_sets",))
            v[:] = np.arange(1, len(nsets) + 1, dtype=np.int32)
            v.name = b"ID"
            for i, nodes in enumerate(nsets.values()):
                if len(nodes):
                    nc.createDimension(f"num_nod_ns{i + 1}", len(nodes))
                    v = nc.createVariable(f"node_ns{i + 1}", "i", (f"num_nod_ns{i + 1}",))
                    v[:] = nodes.astype(np.int32)
        ssets = {k: (np.asarray(mesh.sideset_elems[k]), np.asarray(mesh.sideset_sides[k]))
                 for k in getattr(mesh, "sideset_elems", {}) if not k.startswith("__")}
        ssets = {k: v_ for k, v_ in ssets.items() if len(v_[0])}
        if ssets:
            nc.createDimension("num_side_sets", len(ssets))
            v = nc.createVariable("ss_names", "c", ("num_side_sets", "len_string"))
            v[:] = _chars(list(ssets), _LEN_STRING)
            v = nc.createVariable("ss_status", "i", ("num_side_sets",))
            v[:] = np.ones(len(ssets), dtype=np.int32)
            v = nc.createVariable("ss_prop1", "i", ("num_side_sets",))
            v[:] = np.arange(1, len(ssets) + 1, dtype=np.int32)
            v.name = b"ID"
            for i, (el, sd) in enumerate(ssets.values()):
                nc.createDimension(f"num_side_ss{i + 1}", len(el))
                v = nc.createVariable(f"elem_ss{i + 1}", "i", (f"num_side_ss{i + 1}",))
                v[:] = el.astype(np.int32)
                v = nc.createVariable(f"side_ss{i + 1}", "i", (f"num_side_ss{i + 1}",))
                v[:] = sd.astype(np.int32)
        # ---- result variables
        self._time = nc.createVariable("time_whole", "d", ("time_step",))
        self._vals = []
        if self.nodal_names:
            nc.createDimension("num_nod_var", len(self.nodal_names))
            v = nc.createVariable("name_nod_var", "c", ("num_nod_var", "len_string"))
            v[:] = _chars(self.nodal_names, _LEN_STRING)
            for i in range(len(self.nodal_names)):
                self._vals.append(nc.createVariable(f"vals_nod_var{i + 1}", "d", ("time_step", "num_nodes")))
        self._nn = nn
        self._time[0] = 0.0            # Exodus.write_time(exo, 1, 0.0) in the reference's constructor

    def close(self):
        if self.field_output_db is not None:
            # every record variable needs the same number of records: variables never written at the last time step
            # (e.g. an extra nodal name the caller did not fill) are padded with zeros
            nrec = self._time.shape[0]
            for v in self._vals:
                if v.shape[0] < nrec:
                    v[nrec - 1, :] = np.zeros(self._nn)
            self.field_output_db.close()
            self.field_output_db = None


def write_times(pp, time_index, time_val):
    """write_times(pp, time_index, time_val): 1-based time index"""
    pp._time[time_index - 1] = float(time_val)


def write_field(pp, time_index, field_names, field):
    """write_field(pp, time_index, field_names, field::H1Field): one nodal variable per component"""
    f = np.asarray(field if isinstance(field, H1Field) else H1Field(field))
    names = [field_names] if isinstance(field_names, str) else list(field_names)
    if len(names) != f.shape[0]:
        raise AssertionError("length(field_names) == num_fields(field)")
    for d, name in enumerate(names):
        if name not in pp.nodal_names:
            raise KeyError(f"nodal variable {name} was not registered with the PostProcessor")
        pp._vals[pp.nodal_names.index(name)][time_index - 1, :] = f[d]


def close(pp):
    pp.close()
