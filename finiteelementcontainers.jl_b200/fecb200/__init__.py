"""fecb200 -- Python host mirror of the FiniteElementContainers.jl assembly interface on top of
libfecb200.so (hand-written CUDA, sm_100a, FP64).  Same names / argument meaning as the reference
for the hot path: DofManager, FunctionSpace/H1Field, SparseMatrixAssembler, assemble_vector!,
assemble_stiffness!, assemble_matrix_action!, residual/stiffness/hvp accessors, Physics tags.
"""
from . import _lib
from ._lib import FECError
from .fields import H1Field, Connectivity
from .meshes import StructuredMesh, UnstructuredMesh, KuhnTet10Mesh
from .reference_fe import ReferenceFE
from .function_spaces import (FunctionSpace, Lagrange, ScalarFunction, VectorFunction, TensorFunction,
                              SymmetricTensorFunction, GeneralFunction, DofManager, update_field_unknowns,
                              extract_field_unknowns, update_field_dirichlet_bcs)
from .bcs import (DirichletBC, DirichletBCs, InitialCondition, InitialConditions, NeumannBC, NeumannBCs, PeriodicBC,
                  PeriodicBCs, RobinBC, RobinBCs, Source, Sources, TimeStepper)
from .physics import (AbstractPhysics, Poisson, Mechanics, NeoHookean, J2Plasticity, NonSymmetricTestPhysics, ThreeDimensional, PlaneStrain,
                      residual, residual_b, stiffness, stiffness_b, mass, mass_b, stiffness_action,
                      stiffness_action_b, mass_action, mass_action_b, lumped_mass, energy)
from .assemblers import (SparseMatrixAssembler, Parameters, create_parameters, update_dofs, update_bc_values,
                         update_time, create_field, create_unknowns, assemble_vector, assemble_stiffness,
                         assemble_mass, assemble_vector_and_stiffness, assemble_matrix_action, assemble_matrix_free_action,
                         assemble_matrix_free_action_full, hvp, full_field, assemble_lumped_mass, assemble_diagonal, diagonal, assemble_scalar, scalar_values,
                         assemble_vector_neumann_bc, assemble_vector_source, assemble_vector_robin_bc, assemble_matrix_robin_bc,
                         matrix_multiply, update_field)
from .solvers import DirectLinearSolver, IterativeLinearSolver, NewtonSolver, QuasiStaticIntegrator
from .postprocessors import PostProcessor, write_times, write_field, close
from .partition import (Partition, partition_mesh, structured_brick_partition, metis_partition_elements,
                        metis_partition_graph, metis_partition_pattern)



# small accessors the reference exports as functions (src/FiniteElementContainers.jl:35-171)
def num_dimensions(mesh):
    return mesh.num_dimensions()


def num_nodes(mesh):
    return mesh.num_nodes()


def nodal_coordinates(mesh):
    return mesh.nodal_coords


def element_blocks(mesh):
    return list(mesh.element_block_names)


def nodesets(mesh):
    return mesh.nodeset_nodes


def sidesets(mesh):
    return mesh.sideset_nodes


def num_fields(x):
    """num_fields(field | function | dof | physics)"""
    if isinstance(x, AbstractPhysics):
        return x.NF
    return x.num_fields() if hasattr(x, "num_fields") else (x.nf if hasattr(x, "nf") else x.shape[0])


def num_properties(physics):
    """num_properties(::AbstractPhysics{NF, NP, NS}) (src/Physics.jl:20-30)"""
    return physics.NP


def num_states(physics):
    return physics.NS


def create_properties(physics):
    return physics.create_properties()


def create_initial_state(physics):
    return physics.create_initial_state()


def num_entities(field):
    return field.shape[1]


def connectivity(conn, b):
    """connectivity(conn, b) with the reference's 1-based block index (src/Fields.jl:163-168)"""
    return conn.block(b - 1)


def num_elements(fspace, b=None):
    return sum(fspace.elem_conns.nelems) if b is None else fspace.elem_conns.nelems[b - 1]


def current_time(times):
    return times.time_current


def dirichlet_dofs(bcs):
    return bcs.dirichlet_dofs()


def update_ic_values(ics, X):
    ics.update_ic_values(X)


def update_field_ics(U, ics):
    ics.update_field_ics(U)


__all__ = [n for n in dir() if not n.startswith("_")]
