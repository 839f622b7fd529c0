"""ctypes binding of libfecb200.so (include/fecb200.h).

This is the same binding a Julia host makes with `ccall` (see INTEGRATION.md and
../julia/FECB200.jl); Python is used here because the build container has no Julia.
There is NO fallback: if the shared library is missing the import fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FECB200_LIB", os.path.join(os.path.dirname(_HERE), "lib", "libfecb200.so"))


class FECError(RuntimeError):
    """Non-zero status from libfecb200 (the Julia shim raises `error(...)` the same way)."""


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"libfecb200.so not found at {LIB_PATH}. Build it with "
        "`python finiteelementcontainers.jl_b200/build.py` (nvcc, sm_100a). "
        "fecb200 has no CPU fallback.")

lib = C.CDLL(LIB_PATH)

# element / physics / kind enums (keep in sync with include/fecb200.h)
QUAD4, TRI3, HEX8, TET4, TET10 = 1, 2, 3, 4, 5
PHYS_POISSON, PHYS_LINEAR_ELASTIC, PHYS_NEOHOOKEAN, PHYS_NEOHOOKEAN_AS_WRITTEN, PHYS_J2_PLASTICITY = 1, 2, 3, 4, 5
PHYS_TEST_NONSYMMETRIC = 6
RESIDUAL, STIFFNESS, MASS = 1, 2, 3
LUMPED_MASS, DIAGONAL_STIFFNESS, DIAGONAL_MASS = 4, 5, 6
ENERGY = 7   # host-side token only (fecb200_assemble_scalar has no kind argument)
CSC, CSR = 1, 2
FIELD_U, FIELD_RESIDUAL, FIELD_ACTION, FIELD_V = 1, 2, 3, 4

c_i64p = C.POINTER(C.c_int64)
c_i32p = C.POINTER(C.c_int32)
c_f64p = C.POINTER(C.c_double)


class BlockDesc(C.Structure):
    _fields_ = [("elem_type", C.c_int32), ("nnpe", C.c_int32), ("nelem", C.c_int64), ("conn", c_i64p),
                ("nq", C.c_int32), ("N", c_f64p), ("dN", c_f64p), ("w", c_f64p),
                ("physics_id", C.c_int32), ("nprops", C.c_int32), ("props", c_f64p), ("nstate", C.c_int32)]


class MeshDesc(C.Structure):
    _fields_ = [("nnodes", C.c_int64), ("ndim", C.c_int32), ("nf", C.c_int32), ("nblocks", C.c_int32),
                ("blocks", C.POINTER(BlockDesc)), ("coords", c_f64p)]


class Opts(C.Structure):
    _fields_ = [("matrix_type", C.c_int32), ("condensed", C.c_int32), ("matrix_free", C.c_int32),
                ("device", C.c_int32), ("tile_elems", C.c_int32), ("reserved", C.c_int32 * 3)]


Handle = C.c_void_p
VP = C.c_void_p  # double* that may be host or device

# every symbol the header declares, with its signature
SIGNATURES = {
    "fecb200_last_error": (C.c_char_p, []),
    "fecb200_version": (C.c_int, []),
    "fecb200_create": (C.c_int, [C.POINTER(MeshDesc), C.POINTER(Opts), C.POINTER(Handle)]),
    "fecb200_destroy": (C.c_int, [Handle]),
    "fecb200_set_stream": (C.c_int, [Handle, C.c_void_p]),
    "fecb200_synchronize": (C.c_int, [Handle]),
    "fecb200_set_async": (C.c_int, [Handle, C.c_int32]),
    "fecb200_update_dofs": (C.c_int, [Handle, c_i64p, C.c_int64, c_i64p, c_i64p, C.c_int64]),
    "fecb200_sizes": (C.c_int, [Handle, c_i64p, c_i64p, c_i64p]),
    "fecb200_dof_maps_copy": (C.c_int, [Handle, c_i64p, c_i64p]),
    "fecb200_pattern_sizes": (C.c_int, [Handle, c_i64p, c_i64p]),
    "fecb200_pattern_copy": (C.c_int, [Handle, c_i64p, c_i64p]),
    "fecb200_set_dirichlet_values": (C.c_int, [Handle, c_i64p, c_f64p, C.c_int64]),
    "fecb200_set_periodic_values": (C.c_int, [Handle, c_f64p, C.c_int64]),
    "fecb200_set_time": (C.c_int, [Handle, C.c_double, C.c_double]),
    "fecb200_set_source_q": (C.c_int, [Handle, C.c_int32, VP]),
    "fecb200_state_set": (C.c_int, [Handle, C.c_int32, C.c_int32, VP]),
    "fecb200_state_get": (C.c_int, [Handle, C.c_int32, C.c_int32, VP]),
    "fecb200_state_swap": (C.c_int, [Handle]),
    "fecb200_assemble_vector": (C.c_int, [Handle, C.c_int32, VP]),
    "fecb200_residual": (C.c_int, [Handle, VP]),
    "fecb200_vector_values": (C.c_int, [Handle, VP]),
    "fecb200_assemble_scalar": (C.c_int, [Handle, VP]),
    "fecb200_scalar_values": (C.c_int, [Handle, C.c_int32, VP]),
    "fecb200_assemble_matrix": (C.c_int, [Handle, C.c_int32, VP]),
    "fecb200_assemble_vector_and_matrix": (C.c_int, [Handle, VP]),
    "fecb200_set_matrix_double_buffer": (C.c_int, [Handle, C.c_int32]),
    "fecb200_matrix_values": (C.c_int, [Handle, C.c_int32, VP]),
    "fecb200_matrix_values_device": (C.c_int, [Handle, C.c_int32, C.POINTER(C.c_void_p)]),
    "fecb200_assemble_action": (C.c_int, [Handle, C.c_int32, VP, VP]),
    "fecb200_assemble_action_full": (C.c_int, [Handle, C.c_int32, VP, VP]),
    "fecb200_hvp": (C.c_int, [Handle, VP, VP]),
    "fecb200_field_copy": (C.c_int, [Handle, C.c_int32, VP]),
    "fecb200_update_field": (C.c_int, [Handle, VP]),
    "fecb200_matrix_multiply": (C.c_int, [Handle, C.c_int32, VP, VP]),
    "fecb200_set_neumann_bc": (C.c_int, [Handle, C.c_int32, C.c_int64, C.c_int32, C.c_int32, c_i64p, c_f64p, c_f64p, c_f64p]),
    "fecb200_set_neumann_values": (C.c_int, [Handle, C.c_int32, VP]),
    "fecb200_clear_neumann_bcs": (C.c_int, [Handle]),
    "fecb200_assemble_vector_neumann_bc": (C.c_int, [Handle]),
    "fecb200_set_robin_bc": (C.c_int, [Handle, C.c_int32, C.c_int64, C.c_int32, C.c_int32, c_i64p, c_f64p, c_f64p, c_f64p]),
    "fecb200_set_robin_values": (C.c_int, [Handle, C.c_int32, VP, VP]),
    "fecb200_clear_robin_bcs": (C.c_int, [Handle]),
    "fecb200_assemble_vector_robin_bc": (C.c_int, [Handle]),
    "fecb200_assemble_matrix_robin_bc": (C.c_int, [Handle]),
    "fecb200_set_source_values": (C.c_int, [Handle, C.c_int32, VP]),
    "fecb200_assemble_vector_source": (C.c_int, [Handle]),
    "fecb200_cg_solve": (C.c_int, [Handle, VP, VP, C.c_double, C.c_double, C.c_int64, C.c_int32,
                                   c_i64p, c_f64p]),
    "fecb200_newton_solve": (C.c_int, [Handle, VP, C.c_int32, C.c_double, C.c_int32, c_i32p, c_i64p, c_f64p]),
    "fecb200_halo_setup": (C.c_int, [Handle, C.c_int32, c_i32p, c_i64p, c_i64p, c_i64p, c_i64p]),
    "fecb200_ghost_setup": (C.c_int, [Handle, C.c_int32, c_i32p, c_i64p, c_i64p, c_i64p, c_i64p]),
    "fecb200_metis_part_mesh_dual": (C.c_int, [C.c_int64, C.c_int64, c_i64p, c_i64p, C.c_int64, C.c_int64, c_i64p, c_i64p]),
    "fecb200_metis_part_graph": (C.c_int, [C.c_int64, c_i64p, c_i64p, C.c_int64, c_i64p]),
    "fecb200_partition_setup": (C.c_int, [Handle, C.c_int64, c_i32p]),
    "fecb200_halo_pack": (C.c_int, [Handle, C.c_int32, VP]),
    "fecb200_halo_send_size": (C.c_int, [Handle, c_i64p]),
    "fecb200_halo_unpack_add": (C.c_int, [Handle, C.c_int32, VP]),
    "fecb200_halo_recv_size": (C.c_int, [Handle, c_i64p]),
    "fecb200_ipc_export": (C.c_int, [Handle, C.c_int32, C.c_void_p]),
    "fecb200_peer_attach": (C.c_int, [Handle, C.c_int32, C.c_int32, C.c_void_p, c_i64p, c_i32p, c_i64p, C.c_int64]),
    "fecb200_comm_unique_id": (C.c_int, [C.c_void_p]),
    "fecb200_comm_init": (C.c_int, [Handle, C.c_int32, C.c_int32, C.c_void_p]),
    "fecb200_comm_destroy": (C.c_int, [Handle]),
    "fecb200_halo_sum": (C.c_int, [Handle, C.c_int32]),
    "fecb200_halo_update": (C.c_int, [Handle, C.c_int32]),
    "fecb200_halo_update_unknowns": (C.c_int, [Handle, VP]),
    "fecb200_owned_length": (C.c_int, [Handle, c_i64p]),
    "fecb200_comm_barrier": (C.c_int, [Handle]),
    "fecb200_comm_allreduce_sum": (C.c_int, [Handle, c_f64p, C.c_int32]),
    "fecb200_comm_peer_enable": (C.c_int, [Handle, C.c_int32]),
    "fecb200_peer_detach": (C.c_int, [Handle]),
    "fecb200_launch_count": (C.c_int, [Handle, c_i64p]),
    "fecb200_block_kernel_form": (C.c_int, [Handle, C.c_int32, C.POINTER(C.c_int32)]),
    "fecb200_enable_timing": (C.c_int, [Handle, C.c_int32]),
    "fecb200_last_kernel_ms": (C.c_int, [Handle, C.POINTER(C.c_float)]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _f = getattr(lib, _name)  # AttributeError here = header/library mismatch: fail loudly
    _f.restype = _res
    _f.argtypes = _args


def check(status: int):
    if status != 0:
        raise FECError(lib.fecb200_last_error().decode("utf-8", "replace"))


def i64(a):
    a = np.ascontiguousarray(a, dtype=np.int64)
    return a, a.ctypes.data_as(c_i64p)


def f64(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(c_f64p)


def ptr(x):
    """Raw address of a numpy array (host) or of anything exposing data_ptr() (torch tensor,
    host or CUDA) -- device buffers are used in place by the library (zero copy)."""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        assert x.dtype == np.float64 and x.flags.c_contiguous
        return C.c_void_p(x.ctypes.data)
    if hasattr(x, "data_ptr"):
        return C.c_void_p(x.data_ptr())
    if isinstance(x, int):
        return C.c_void_p(x)
    raise TypeError(f"cannot take the address of {type(x)}")
