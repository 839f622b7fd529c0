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
from .function_spaces import (FunctionSpace, Lagrange, ScalarFunction, VectorFunction, DofManager, update_field_unknowns,
                              extract_field_unknowns, update_field_dirichlet_bcs)
from .bcs import (DirichletBC, DirichletBCs, NeumannBC, NeumannBCs, PeriodicBC, PeriodicBCs, Source, Sources,
                  TimeStepper)
from .physics import (AbstractPhysics, Poisson, Mechanics, NeoHookean, J2Plasticity, ThreeDimensional, PlaneStrain,
                      residual, residual_b, stiffness, stiffness_b, mass, mass_b, stiffness_action,
                      stiffness_action_b, mass_action, mass_action_b, lumped_mass, energy)
from .assemblers import (SparseMatrixAssembler, Parameters, create_parameters, update_dofs, update_bc_values,
                         update_time, create_field, create_unknowns, assemble_vector, assemble_stiffness,
                         assemble_mass, assemble_vector_and_stiffness, assemble_matrix_action, assemble_matrix_free_action,
                         assemble_matrix_free_action_full, hvp, full_field, assemble_lumped_mass, assemble_diagonal, diagonal, assemble_scalar, scalar_values,
                         assemble_vector_neumann_bc, assemble_vector_source)
from .solvers import DirectLinearSolver, IterativeLinearSolver, NewtonSolver, QuasiStaticIntegrator
from .postprocessors import PostProcessor, write_times, write_field, close
from .partition import (Partition, partition_mesh, structured_brick_partition, metis_partition_elements,
                        metis_partition_graph)

__all__ = [n for n in dir() if not n.startswith("_")]
