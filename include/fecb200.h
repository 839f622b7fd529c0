/*
 * fecb200.h -- C ABI of libfecb200.so: the B200-native FE assembly hot path of
 * FiniteElementContainers.jl (element-wise residual / stiffness / matrix-action assembly
 * over H1 spaces), hand-written CUDA for sm_100a, FP64.
 *
 * This is the drop-in boundary.  The reference has no FFI layer: its
 * seam is Julia multiple dispatch on the assembler type (ext/CUDAExt.jl:8-31 is the only
 * backend hook).  A Julia `B200Assembler <: AbstractAssembler` therefore binds the entry
 * points below with `ccall` (see INTEGRATION.md and finiteelementcontainers.jl_b200/julia/).
 * Every function cites the reference interface it replaces (path:line under the
 * reference tree).
 *
 * Conventions
 *   - all functions return 0 on success, non-zero on error; fecb200_last_error() gives text.
 *     No exceptions cross the boundary (the reference throws Julia exceptions; the shim
 *     converts a non-zero status into `error(...)`).
 *   - every index array crossing the boundary is Int64 and 1-based, as in the reference
 *     (Connectivity: src/Fields.jl:129-160; patterns: src/assemblers/SparsityPatterns.jl:30-51).
 *   - dof id of (field d, node n) = NF*(n-1) + d  (src/DofManagers.jl:41-58).
 *   - pointers marked [host|device] may be host or device memory (detected with
 *     cudaPointerGetAttributes); device pointers (e.g. a CuArray) are used in place.
 *   - the handle owns all device memory it allocates; the caller owns every buffer it passes.
 *   - one handle <-> one CUDA device + one stream; calls on one handle are not re-entrant.
 *   - there is NO CPU fallback: creation fails if no CUDA device is present.
 */
#ifndef FECB200_H
#define FECB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fecb200_handle fecb200_handle;

/* element topologies (Exodus node ordering; src/FunctionSpaces.jl:1-29 names) */
enum { FECB200_QUAD4 = 1, FECB200_TRI3 = 2, FECB200_HEX8 = 3, FECB200_TET4 = 4, FECB200_TET10 = 5 };

/* CUDA-side physics (replaces the Julia closures of src/Physics.jl:1-18, which cannot cross a C ABI) */
enum {
  FECB200_PHYS_POISSON = 1,            /* test/poisson/TestPoissonCommon.jl:4-139   NF=1          */
  FECB200_PHYS_LINEAR_ELASTIC = 2,     /* test/mechanics/TestMechanicsCommon.jl:3-236  props (rho,K,G) */
  FECB200_PHYS_NEOHOOKEAN = 3,         /* TestMechanicsLargeDeformation.jl:17-27, stress-free U(J)  */
  FECB200_PHYS_NEOHOOKEAN_AS_WRITTEN = 4, /* the script's U(J) verbatim               */
  FECB200_PHYS_J2_PLASTICITY = 5,      /* NS=7 state, props (rho,K,G,sigma_y,H); hooks only in the reference */
  FECB200_PHYS_TEST_NONSYMMETRIC = 6   /* TEST law, props (rho,K,G,beta): linear elasticity + beta d_ij T_kl, a tangent WITHOUT
                                          major symmetry, so that the transposed COO labelling of the pattern
                                          (Assemblers.jl:109-124 vs SparsityPatterns.jl:76-83) shows in a parity test */
};

/* which element-level function is assembled: the `f` argument of assemble_vector! etc.
 * (mapped by identity on the Julia side: residual / stiffness / mass / stiffness_action) */
enum { FECB200_RESIDUAL = 1, FECB200_STIFFNESS = 2, FECB200_MASS = 3,
       /* vector flavours of fecb200_assemble_vector */
       FECB200_LUMPED_MASS = 4, FECB200_DIAGONAL_STIFFNESS = 5, FECB200_DIAGONAL_MASS = 6 };

/* sparse_matrix_type of SparseMatrixAssembler (src/assemblers/SparseMatrixAssembler.jl:64-76) */
enum { FECB200_CSC = 1, FECB200_CSR = 2 };

/* field selectors for fecb200_field_copy */
enum { FECB200_FIELD_U = 1, FECB200_FIELD_RESIDUAL = 2, FECB200_FIELD_ACTION = 3, FECB200_FIELD_V = 4 };

/* One element block: FunctionSpace block + its ReferenceFE tables + Physics + props
 * (src/FunctionSpaces.jl:158-237, src/assemblers/Assemblers.jl:129-131).  The reference
 * element tables are NOT hard-coded in the library: ReferenceFiniteElements.jl is not
 * vendored by the reference, so the host passes `ref_fe.cell_interps[q]` as arrays. */
typedef struct {
  int32_t elem_type;      /* FECB200_HEX8 ...                                            */
  int32_t nnpe;           /* nodes per element                                            */
  int64_t nelem;
  const int64_t* conn;    /* [nnpe*nelem], 1-based node ids, element-major (Connectivity.data) */
  int32_t nq;             /* quadrature points                                            */
  const double* N;        /* [nq*nnpe]        N[q*nnpe + a]                                */
  const double* dN;       /* [nq*nnpe*ndim]   dN[(q*nnpe + a)*ndim + j] = dN_a/dxi_j       */
  const double* w;        /* [nq]                                                         */
  int32_t physics_id;     /* FECB200_PHYS_*                                               */
  int32_t nprops;
  const double* props;    /* [nprops]  SVector{NP} props of the block (Assemblers.jl:193-202) */
  int32_t nstate;         /* NS of AbstractPhysics{NF,NP,NS}; state arrays are [NS,NQ,NE] */
} fecb200_block_desc;

typedef struct {
  int64_t nnodes;
  int32_t ndim;
  int32_t nf;             /* dofs per node (NF)                                           */
  int32_t nblocks;
  const fecb200_block_desc* blocks;
  const double* coords;   /* [ndim*nnodes]  H1Field data: coords[(n-1)*ndim + j]  (src/Fields.jl:36-40) */
} fecb200_mesh_desc;

typedef struct {
  int32_t matrix_type;    /* FECB200_CSC | FECB200_CSR                                    */
  int32_t condensed;      /* DofManager{Condensed}  (src/DofManagers.jl:21-58)            */
  int32_t matrix_free;    /* SparseMatrixAssembler(...; matrix_free=true) (:82-88)        */
  int32_t device;         /* CUDA device ordinal                                          */
  int32_t tile_elems;     /* elements per CTA tile for the vector kernels; 0 = default    */
  int32_t reserved[3];
} fecb200_opts;

const char* fecb200_last_error(void);
int fecb200_version(void);

/* SparseMatrixAssembler(dof; ...) + Parameters device upload: builds DOF maps, element tiles,
 * node adjacency / CSR pattern and the element->CSR slot maps on the device.
 * replaces: SparseMatrixAssembler ctor (SparseMatrixAssembler.jl:64-124), SparseMatrixPattern
 * (SparsityPatterns.jl:53-117), `asm |> cuda` / `p |> cuda` (ext/CUDAExt.jl:8-14). */
int fecb200_create(const fecb200_mesh_desc* mesh, const fecb200_opts* opts, fecb200_handle** out);
int fecb200_destroy(fecb200_handle* h);

/* use an existing CUDA stream (cudaStream_t) for all work of this handle; NULL = own stream */
int fecb200_set_stream(fecb200_handle* h, void* cuda_stream);
int fecb200_synchronize(fecb200_handle* h);
/* Opt-in asynchronous HOST copies.  Default (off): every call is complete on return.  On: host inputs are uploaded
 * and host outputs downloaded on dedicated copy streams, ordered against the compute stream with events, so the
 * H2D of Uu overlaps the zero-fill of the CSR values and the D2H of a result overlaps the next assembly.  Host
 * buffers should be page-locked; host OUTPUTS are valid only after fecb200_synchronize(); host INPUTS must stay
 * untouched until then. */
int fecb200_set_async(fecb200_handle* h, int32_t on);

/* update_dofs!(asm, dbcs, pbcs)  (SparseMatrixAssembler.jl:228-274, DofManagers.jl:227-298,
 * SparsityPatterns.jl:160-231).  dirichlet_dofs need not be sorted/unique. */
int fecb200_update_dofs(fecb200_handle* h, const int64_t* dirichlet_dofs, int64_t n_dirichlet,
                        const int64_t* periodic_side_a, const int64_t* periodic_side_b, int64_t n_periodic);

/* sizes: n_total = NF*NN, n_unknowns = length(dof.unknown_dofs); length(Uu) = n_unknowns
 * (non-condensed) or n_total (condensed)  (DofManagers.jl:168-175) */
int fecb200_sizes(fecb200_handle* h, int64_t* n_total_dofs, int64_t* n_unknowns, int64_t* len_Uu);
/* DofManager arrays, 1-based Int64: unknown_dofs[n_unknowns], dof_to_unknown[n_total] (-1 Dirichlet,
 * -2 periodic side b).  "DOF numbering bit-exact" is checked on these. Either may be NULL. */
int fecb200_dof_maps_copy(fecb200_handle* h, int64_t* unknown_dofs, int64_t* dof_to_unknown);

/* sparsity pattern of stiffness(asm)/mass(asm): n x n with nnz stored entries, explicit zeros kept.
 * ptr[n+1], idx[nnz] Int64 1-based: (rowptr, colval) for CSR, (colptr, rowval) for CSC, exactly what
 * SparseArrays.sparse! / SparseMatrixCSR(csc) produce (SparsityPatterns.jl:301-329). [host] */
int fecb200_pattern_sizes(fecb200_handle* h, int64_t* n, int64_t* nnz);
int fecb200_pattern_copy(fecb200_handle* h, int64_t* ptr, int64_t* idx);

/* Dirichlet values: U[dofs[i]] = vals[i] before every assemble (update_field_dirichlet_bcs!,
 * src/bcs/DirichletBCs.jl:411-418; values come from update_bc_values!, Parameters.jl:358). */
int fecb200_set_dirichlet_values(fecb200_handle* h, const int64_t* dofs, const double* vals, int64_t n);
/* periodic offsets: U[b] = U[a] + val (src/bcs/PeriodicBCs.jl:253-262), one per resolved pair; NULL = all zero (n ignored) */
int fecb200_set_periodic_values(fecb200_handle* h, const double* vals, int64_t n);
int fecb200_set_time(fecb200_handle* h, double t, double dt);
/* Poisson source f(X_q, t) pre-evaluated at quadrature points on the host (the closure
 * physics.func of TestPoissonCommon.jl:4-6 cannot cross the ABI; same pattern as
 * src/bcs/Sources.jl:55-66).  fq[q + nq*e], element order of the block's conn. [host|device] */
int fecb200_set_source_q(fecb200_handle* h, int32_t block, const double* fq);

/* state variables [NS,NQ,NE] per block (Parameters.jl:1-23, Assemblers.jl:248-251).
 * which = 0 state_old, 1 state_new. */
int fecb200_state_set(fecb200_handle* h, int32_t block, int32_t which, const double* state);
int fecb200_state_get(fecb200_handle* h, int32_t block, int32_t which, double* state);
int fecb200_state_swap(fecb200_handle* h); /* state_old <- state_new at the end of a load step */

/* assemble_vector!(asm, residual, Uu, p)  (src/assemblers/Vector.jl:4-74): zero storage,
 * _update_for_assembly! (Parameters.jl:404-413), element kernel, nodal scatter.
 * Uu [host|device], length len_Uu. */
int fecb200_assemble_vector(fecb200_handle* h, int32_t kind, const double* Uu);
/* The other vector flavours go through the same call and, like the reference, into the residual storage:
 *   kind = FECB200_LUMPED_MASS         assemble_lumped_mass!(asm, lumped_mass, Uu, p)  (src/assemblers/LumpedMass.jl:32-60):
 *                                      row-sum mass rho*JxW*N[a] per dof (partition of unity)
 *   kind = FECB200_DIAGONAL_STIFFNESS  assemble_diagonal!(asm, stiffness, Uu, p)       (src/assemblers/Diagonal.jl:16-74,
 *          FECB200_DIAGONAL_MASS       assemble_diagonal!(asm, mass, Uu, p)             Assemblers.jl:42-45): diag of K_el / M_el
 * fecb200_vector_values = lumped_mass(asm) / diagonal(asm) (LumpedMass.jl:70-80, Diagonal.jl:76-89): the full
 * storage (condensed) or its unknown-dof subset -- no constraint scaling, no periodic fold.  out [host|device],
 * length len_Uu. */
int fecb200_vector_values(fecb200_handle* h, double* out);
/* assemble_scalar!(asm, energy, Uu, p) (src/assemblers/QuadratureQuantity.jl:4-45): the quadrature-point values
 * storage[1, q, e] = JxW * energy_q of every block (Assemblers.jl:47-51), no scatter.  Energies: Poisson
 * 1/2 |grad u|^2 - u f (TestPoissonCommon.jl:8-16), linear elastic psi (TestMechanicsCommon.jl:14-50), neo-Hookean psi
 * (TestMechanicsLargeDeformation.jl:17-27); the J2 law has none (error).
 * fecb200_scalar_values copies block `block` (0-based) as [NQ, NE] column-major (q fastest) in the caller's element
 * order = block_view(asm.scalar_quadrature_storage, b).  out [host|device], length NQ*NE of that block. */
int fecb200_assemble_scalar(fecb200_handle* h, const double* Uu);
int fecb200_scalar_values(fecb200_handle* h, int32_t block, double* out);
/* residual(asm) (src/assemblers/Assemblers.jl:347-371): condensed -> R*(1-c); else periodic
 * fold + gather of unknowns.  out [host|device], length len_Uu. */
int fecb200_residual(fecb200_handle* h, double* out);

/* assemble_stiffness!/assemble_mass! (src/assemblers/Matrix.jl:1-75) fused with the sparse!
 * realisation: values are accumulated straight into the CSR/CSC nzval (no COO is materialised). */
int fecb200_assemble_matrix(fecb200_handle* h, int32_t kind, const double* Uu);
/* Fused assemble_vector!(asm, residual, Uu, p) + assemble_stiffness!(asm, stiffness, Uu, p): what one Newton
 * iteration of solve!(::IterativeLinearSolver) asks for back to back at the same Uu (src/Solvers.jl:133-140).
 * One field update; blocks whose tangent kernel already holds dN_X and the constitutive state at every quadrature
 * point produce the residual in the same pass.  Results are identical (to rounding) to the two separate calls;
 * fecb200_residual / fecb200_matrix_values read them as usual. */
int fecb200_assemble_vector_and_matrix(fecb200_handle* h, const double* Uu);
/* stiffness(asm)/mass(asm) (Assemblers.jl:329-388): applies the condensed-mode constraint
 * adjustment (assemblers/Utils.jl:53-148: row/col scaling by (1-c), penalty 1e6*tr(K)/n on
 * constrained diagonals) and copies nzval [nnz] out. [host|device] */
int fecb200_matrix_values(fecb200_handle* h, int32_t kind, double* nzval_out);
/* device pointer to the handle's nzval storage (valid until destroy / update_dofs; with double buffering
 * enabled, until the next assembly of that matrix) */
int fecb200_matrix_values_device(fecb200_handle* h, int32_t kind, double** nzval_dev);
/* Opt-in double buffering of the stiffness nzval storage: replaces the stand-alone fill!(storage, 0) pass of
 * assemble_stiffness! (src/assemblers/Matrix.jl:39) in steady state.  Each assembly accumulates into the buffer
 * the previous assembly's kernels cleared and clears the other one with TMA bulk stores that drain under the
 * element kernel.  Costs a second nzval array; values equal those of the single-buffer path (up to the
 * summation order of the atomics, as always). */
int fecb200_set_matrix_double_buffer(fecb200_handle* h, int32_t enable);

/* assemble_matrix_action!(asm, f, Uu, Vu, p) and assemble_matrix_free_action! (MatrixAction.jl:9-77,
 * 154-238): always evaluated matrix-free.  kind = FECB200_STIFFNESS or FECB200_MASS. */
int fecb200_assemble_action(fecb200_handle* h, int32_t kind, const double* Uu, const double* Vu);
/* assemble_matrix_free_action_full! (MatrixAction.jl:99-149): caller passes full-length U, v */
int fecb200_assemble_action_full(fecb200_handle* h, int32_t kind, const double* U_full, const double* v_full);
/* hvp(asm, v) (Assemblers.jl:310-324): condensed -> (1-c)Av + c v ; else gather unknowns */
int fecb200_hvp(fecb200_handle* h, const double* v, double* out);

/* copy of a full-length nodal field (p.field, residual_storage, stiffness_action_storage) */
int fecb200_field_copy(fecb200_handle* h, int32_t which, double* out);
/* _update_for_assembly!(p, dof, Uu) on its own (src/Parameters.jl:404-413): Dirichlet values, unknowns and periodic
 * copies into p.field, one launch.  evolve! calls it after the solve so that the stored field is the converged,
 * BC-enforced one (src/integrators/QuasiStaticIntegrator.jl:21-24).  Uu [host|device], length len_Uu. */
int fecb200_update_field(fecb200_handle* h, const double* Uu);
/* y = stiffness(asm) * x / mass(asm) * x on the assembled values (the product Krylov forms, src/Solvers.jl:144):
 * device SpMV over the reference-ordered CSR, constraint adjustment applied first in condensed mode.  x, y
 * [host|device], length len_Uu.  CSC handles hold K^T row-wise; the product is K x for the (symmetric) operators
 * assembled here.  On a partitioned handle x's ghost entries are refreshed from their owners first and only owned
 * rows of y are produced (ghost entries of y are 0). */
int fecb200_matrix_multiply(fecb200_handle* h, int32_t kind, const double* x, double* y);

/* ---- external loads of the residual: the two calls solve! makes right after
 * assemble_vector! (src/Solvers.jl:66-69, 133-137).  Neither zeroes the residual storage, both add to it
 * (Source.jl:1-5); fecb200_newton_solve applies them after every residual assembly, as solve! does.
 *
 * Neumann BCs = NeumannBCContainer (src/bcs/NeumannBCs.jl:29-48, _setup_sideset BoundaryConditions.jl:337-411):
 *   id          0, 1, 2, ... in registration order (re-registering an id replaces it)
 *   side_nodes  [nnps, nsides] column-major, Int64 1-based: surface_connectivity of every side
 *   Ns, dNs, ws surface tables of the block's ReferenceFE (ReferenceFiniteElements.jl is un-vendored, so they
 *               travel like the cell tables): Ns[q*nnps + a], dNs[(q*nnps + a)*(ND-1) + k], ws[q]
 *   vals        [NF, nqs, nsides] = Matrix{SVector{NF,Float64}}(nqs, nsides): func(X_q, t) evaluated by the host
 *               (update_bc_values!, NeumannBCs.jl:60-71, 157-171)  [host|device]
 * assemble_vector_neumann_bc! (src/assemblers/WeaklyEnforcedBCs.jl:4-15, 61-83):
 *   R[(n,d)] += sum_q JxW_s(q) Ns[q][n] vals[d,q,e],  JxW_s = |sum_a x_a dNs_a| ws (edges) or |t_0 x t_1| ws (faces). */
int fecb200_set_neumann_bc(fecb200_handle* h, int32_t id, int64_t nsides, int32_t nnps, int32_t nqs,
                           const int64_t* side_nodes, const double* Ns, const double* dNs, const double* ws);
int fecb200_set_neumann_values(fecb200_handle* h, int32_t id, const double* vals);
int fecb200_clear_neumann_bcs(fecb200_handle* h);
int fecb200_assemble_vector_neumann_bc(fecb200_handle* h);
/* Robin BCs = RobinBCContainer (src/bcs/RobinBCs.jl:29-86): same side-set geometry as a Neumann BC; the flux law
 * func(X_q, t, u_q) of the reference is a closure differentiated with ForwardDiff (:72-75) and cannot cross the ABI, so
 * the host hands over its AFFINE form at the surface quadrature points (update_bc_values!, :77-86):
 *   g0       [NF, nqs, nsides]      vals(q, e)    = g0 + dvalsdu * u_q,   u_q = sum_a Ns[q][a] U_a
 *   dvalsdu  [NF, NF, nqs, nsides]  dvalsdu(q, e) = d func / d u          (column-major SMatrix{NF,NF} per point)
 * (the reference's own Robin regression, test/poisson/TestPoisson.jl:106-127, is of this form: a(x) - alpha u).
 * fecb200_assemble_vector_robin_bc = assemble_vector_robin_bc! (src/assemblers/WeaklyEnforcedBCs.jl:20-32, 61-83):
 *   R[(n,d)] += sum_q JxW_s Ns[q][n] vals_d(q, e)  at the CURRENT p.field (no _update_for_assembly!, like the reference).
 * fecb200_assemble_matrix_robin_bc = assemble_matrix_robin_bc! (:88-180): K_el = sum_q JxW_s Ns_i Ns_j dvalsdu[di,dj]
 *   added to the assembled stiffness values with the pattern's (transposed) COO labelling; call it after
 *   fecb200_assemble_matrix(STIFFNESS) and before fecb200_matrix_values, as solve!(::DirectLinearSolver) does
 *   (src/Solvers.jl:73-79).  Not available on partitioned handles. */
int fecb200_set_robin_bc(fecb200_handle* h, int32_t id, int64_t nsides, int32_t nnps, int32_t nqs,
                         const int64_t* side_nodes, const double* Ns, const double* dNs, const double* ws);
int fecb200_set_robin_values(fecb200_handle* h, int32_t id, const double* g0, const double* dvalsdu);
int fecb200_clear_robin_bcs(fecb200_handle* h);
int fecb200_assemble_vector_robin_bc(fecb200_handle* h);
int fecb200_assemble_matrix_robin_bc(fecb200_handle* h);
/* Body forces = SourceContainer.vals (src/bcs/Sources.jl:38-66): vals [NF, NQ, NE] of one block, element order of
 * the block's conn; NULL removes the block's source  [host|device].
 * assemble_vector_source! (src/assemblers/Source.jl:10-64): R[(n,d)] += - sum_q JxW(q) N[q][n] vals[d,q,e]. */
int fecb200_set_source_values(fecb200_handle* h, int32_t block, const double* vals);
int fecb200_assemble_vector_source(fecb200_handle* h);

/* ---- callers of the path: device-resident Krylov / Newton ---------------
 * krylov_solve!(ws, stiffness(asm), residual(asm)) with CG (src/Solvers.jl:128-153); Krylov.jl
 * defaults atol = rtol = sqrt(eps), itmax = 2n.  Solves K x = b on the device using the handle's
 * assembled matrix (matrix_free = 0) or the matrix-free operator at the current U (matrix_free = 1). */
int fecb200_cg_solve(fecb200_handle* h, const double* b, double* x, double atol, double rtol,
                     int64_t itmax, int32_t matrix_free, int64_t* iters_out, double* rnorm_out);
/* solve!(NewtonSolver, Uu, p) (src/Solvers.jl:193-220): <= max_iters, |dU|,|R|,|R|/|R0| < tol.
 * Uu is updated in place [host|device]. */
int fecb200_newton_solve(fecb200_handle* h, double* Uu, int32_t max_iters, double tol,
                         int32_t matrix_free, int32_t* newton_iters_out, int64_t* cg_iters_out,
                         double* rnorm_out);

/* ---- multi-GPU (ext/PartitionedArraysExt.jl:223-233, 449-481): one handle per rank over the
 * rank-local mesh (owned nodes first, then ghosts).  The halo lists say which local nodes are
 * sent to / received from each neighbour rank.  Buffers are packed/unpacked on the device; the
 * exchange itself is driven by the host (NCCL send/recv or peer memory). ---------------------- */
/* METIS k-way partition of the element dual graph (elements sharing >= ncommon nodes are adjacent).
 * eptr[ne+1], eind: 0-based CSR element -> node lists; epart[ne], npart[nn] receive 0-based part ids.
 * (The reference partitions with the SEACAS `decomp` tool, ext/PartitionedArraysExt.jl:41-49, or METIS on
 * the DOF graph, ext/MetisExt.jl:6-14.)  Host-only; needs no handle. */
int fecb200_metis_part_mesh_dual(int64_t ne, int64_t nn, const int64_t* eptr, const int64_t* eind, int64_t ncommon,
                                 int64_t nparts, int64_t* epart, int64_t* npart);
/* Metis.partition(graph, nparts) of ext/MetisExt.jl:6-14: xadj/adjncy 0-based CSR adjacency (no self loops) */
int fecb200_metis_part_graph(int64_t nv, const int64_t* xadj, const int64_t* adjncy, int64_t nparts, int64_t* part);

/* Rank-local view: local nodes [1, n_owned_nodes] are owned, the rest are ghosts (PartitionedArrays
 * OwnAndGhostIndices order).  Blocks flagged in block_is_halo hold neighbour-owned elements that touch owned
 * nodes: they are skipped by vector assembly (their residual arrives through the halo exchange) and included in
 * matrix assembly, so every OWNED row of the Jacobian is complete without communication; ghost rows are not
 * stored, ghost columns are.  Rebuilds the CSR structure. */
int fecb200_partition_setup(fecb200_handle* h, int64_t n_owned_nodes, const int32_t* block_is_halo);
int fecb200_halo_setup(fecb200_handle* h, int32_t n_neighbors, const int32_t* neighbor_ranks,
                       const int64_t* send_ptr, const int64_t* send_nodes,  /* ghosts I hold -> owner   */
                       const int64_t* recv_ptr, const int64_t* recv_nodes); /* my owned, ghosted by nbr */
/* Lists of the owner -> ghost update (consistent!, ext/PartitionedArraysExt.jl:449-459), per neighbour rank: my OWNED
 * nodes that the neighbour holds as ghosts, and my ghosts it owns -- EVERY ghost, including the far nodes of halo
 * elements (columns of owned Jacobian rows that no owned element touches, hence absent from the residual lists
 * above).  Both sides list a pair's nodes in the same (global-id) order.  Without this call halo_update falls back
 * to the residual lists reversed. */
int fecb200_ghost_setup(fecb200_handle* h, int32_t n_neighbors, const int32_t* neighbor_ranks,
                        const int64_t* own_ptr, const int64_t* own_nodes, const int64_t* ghost_ptr, const int64_t* ghost_nodes);
/* pack field values (NF per node) of the send nodes into the caller's device buffer (halo_send_size doubles) */
int fecb200_halo_pack(fecb200_handle* h, int32_t which_field, double* sendbuf_dev);
int fecb200_halo_send_size(fecb200_handle* h, int64_t* n_doubles);
/* owner side: add received ghost contributions into the owned entries */
int fecb200_halo_unpack_add(fecb200_handle* h, int32_t which_field, const double* recvbuf_dev);
int fecb200_halo_recv_size(fecb200_handle* h, int64_t* n_doubles);

/* Peer-memory halo (one process per GPU, NVLink / NVSwitch): instead of pack -> NCCL send/recv -> unpack, the
 * assembly kernels add the contributions of ghost nodes straight into the OWNER's field with system-scope REDs
 * over peer-mapped memory.
 *   fecb200_ipc_export      64-byte cudaIpcMemHandle_t of this handle's field (host exchanges it, e.g. all_gather)
 *   fecb200_peer_attach     open the neighbours' handles; ghost_peer[g] / ghost_node[g] give, for ghost g =
 *                           local node n_owned + g, the index into `handles` (or -1) and the 0-based node id on
 *                           the owner.  Needs fecb200_partition_setup first.
 * The host must order steps across ranks: every rank's field is zeroed before any rank scatters, and all ranks
 * finished scattering before an owner reads (two stream-ordered barriers per assembly). */
int fecb200_ipc_export(fecb200_handle* h, int32_t which_field, void* handle64);
/* peer_n_nodes[i] = node count of peer i's local mesh: every ghost_node is range-checked against its owner before
 * any handle is opened (a stale exchange list must not turn into a stray write into another process). */
int fecb200_peer_attach(fecb200_handle* h, int32_t which_field, int32_t n_peers, const void* handles64,
                        const int64_t* peer_n_nodes, const int32_t* ghost_peer, const int64_t* ghost_node, int64_t n_ghosts);
int fecb200_peer_detach(fecb200_handle* h);

/* ---- collective plane: include/fecb200.h alone drives N GPUs, one process per GPU, NCCL over
 * NVLink.  The host moves ONE thing out of band: rank 0's 128-byte ncclUniqueId (MPI_Bcast in the reference's
 * setting, ext/PartitionedArraysExt.jl:15-37).  Everything below is stream-ordered on the handle's stream.
 *   fecb200_comm_unique_id       ncclGetUniqueId (rank 0)
 *   fecb200_comm_init            ncclCommInitRank for this handle's device; needs fecb200_halo_setup for the halo calls
 *   fecb200_halo_sum             ghost -> owner sum of a nodal field = assembly of a PVector (:469-481): device pack,
 *                                one grouped ncclSend/ncclRecv per neighbour, device add; with the peer-memory halo
 *                                enabled for that field it is the closing barrier only
 *   fecb200_halo_update          owner -> ghost copy of a nodal field = consistent! (:449-459)
 *   fecb200_halo_update_unknowns the same for a device vector in the Uu layout (ghost unknowns follow the owned ones)
 *   fecb200_owned_length         leading entries of a Uu-shaped vector that belong to owned nodes (own_values)
 *   fecb200_comm_barrier         stream-ordered barrier (4-byte all-reduce): orders the two phases of the peer halo
 *   fecb200_comm_allreduce_sum   sum of 1..4 host scalars over the ranks (distributed dots / norms, :522-540)
 *   fecb200_comm_peer_enable     exchanges the IPC handles and the owner-local ghost ids over NCCL inside the library
 *                                and attaches the peer-memory halo for that field (fecb200_peer_attach)
 * With a communicator attached, fecb200_cg_solve / fecb200_newton_solve / fecb200_matrix_multiply run distributed:
 * dots over owned entries + ncclAllReduce, ghost refresh before every operator application, ghost -> owner sum of the
 * residual; iteration counts equal the serial solve's. */
int fecb200_comm_unique_id(void* id128);
int fecb200_comm_init(fecb200_handle* h, int32_t rank, int32_t nranks, const void* id128);
int fecb200_comm_destroy(fecb200_handle* h);
int fecb200_halo_sum(fecb200_handle* h, int32_t which_field);
int fecb200_halo_update(fecb200_handle* h, int32_t which_field);
int fecb200_halo_update_unknowns(fecb200_handle* h, double* Uu_dev);
int fecb200_owned_length(fecb200_handle* h, int64_t* n);
int fecb200_comm_barrier(fecb200_handle* h);
int fecb200_comm_allreduce_sum(fecb200_handle* h, double* vals, int32_t n);
int fecb200_comm_peer_enable(fecb200_handle* h, int32_t which_field);

/* ---- instrumentation: kernels launched by this handle since creation (bench `gpu_launches`) */
int fecb200_launch_count(fecb200_handle* h, int64_t* n);
/* which form of the element kernels block `block` (0-based) takes: 1 = the Walsh-Hadamard form (HEX8 blocks whose
 * ReferenceFE table -- `ref_fe.cell_interps` of the block's FunctionSpace entry, src/FunctionSpaces.jl -- is the trilinear
 * one on a symmetric 2-point rule per axis, in any node / point numbering), 0 = the plain quadrature loop. */
int fecb200_block_kernel_form(fecb200_handle* h, int32_t block, int32_t* form);
/* last kernel timing (ms) measured with CUDA events on the handle's stream around the dominant
 * element kernel of the last assemble_* call; enabled by fecb200_enable_timing(h, 1) */
int fecb200_enable_timing(fecb200_handle* h, int32_t on);
int fecb200_last_kernel_ms(fecb200_handle* h, float* ms);

#ifdef __cplusplus
}
#endif
#endif /* FECB200_H */
