# FECB200.jl -- Julia host shim: drives libfecb200.so (include/fecb200.h) through `ccall` behind
# FiniteElementContainers.jl's own entry points.  The reference's only backend seam is multiple dispatch on
# the assembler type (ext/CUDAExt.jl:8-31), so the shim is a new `AbstractAssembler` subtype plus methods of
# the reference's generic functions specialised on it.
#
# NOTE: the build container has no Julia, so this file is the binding a maintainer adds (INTEGRATION.md);
# it is kept in sync with fecb200/_lib.py, the ctypes twin that IS exercised by the test-suite.
module FECB200

using FiniteElementContainers
import FiniteElementContainers: AbstractAssembler, DofManager, assemble_vector!, assemble_stiffness!, assemble_mass!,
                                assemble_lumped_mass!, assemble_diagonal!, lumped_mass, diagonal, assemble_scalar!,
                                assemble_matrix_action!, assemble_matrix_free_action!, assemble_matrix_free_action_full!,
                                residual, stiffness, mass, hvp, update_dofs!, update_bc_values!, update_time!,
                                create_unknowns, function_space, assemble_vector_neumann_bc!, assemble_vector_source!,
                                assemble_vector_robin_bc!, assemble_matrix_robin_bc!, surface_connectivity,
                                _update_for_assembly!
using SparseArrays, SparseMatricesCSR
import ReferenceFiniteElements

const LIB = get(ENV, "FECB200_LIB", joinpath(@__DIR__, "..", "lib", "libfecb200.so"))

# enums of include/fecb200.h
const QUAD4, TRI3, HEX8, TET4, TET10 = Int32(1), Int32(2), Int32(3), Int32(4), Int32(5)
const RESIDUAL, STIFFNESS, MASS = Int32(1), Int32(2), Int32(3)
const LUMPED_MASS, DIAGONAL_STIFFNESS, DIAGONAL_MASS = Int32(4), Int32(5), Int32(6)
const CSC, CSR = Int32(1), Int32(2)
const FIELD_U, FIELD_RESIDUAL, FIELD_ACTION, FIELD_V = Int32(1), Int32(2), Int32(3), Int32(4)

struct BlockDesc
  elem_type::Int32; nnpe::Int32; nelem::Int64; conn::Ptr{Int64}
  nq::Int32; N::Ptr{Float64}; dN::Ptr{Float64}; w::Ptr{Float64}
  physics_id::Int32; nprops::Int32; props::Ptr{Float64}; nstate::Int32
end
struct MeshDesc
  nnodes::Int64; ndim::Int32; nf::Int32; nblocks::Int32; blocks::Ptr{BlockDesc}; coords::Ptr{Float64}
end
struct Opts
  matrix_type::Int32; condensed::Int32; matrix_free::Int32; device::Int32; tile_elems::Int32
  reserved::NTuple{3, Int32}
end

check(status::Cint) = status == 0 || error(unsafe_string(ccall((:fecb200_last_error, LIB), Cstring, ())))

"""
Assembler whose storage lives in a libfecb200 handle.  Arbitrary user closures cannot run on the device:
`f` is mapped BY IDENTITY (`f === residual`, `stiffness`, `mass`, `stiffness_action`, ...) and the physics by
type (`physics_id(::Poisson) = 1`, ...); anything else raises an error -- there is no CPU fallback.
"""
mutable struct B200Assembler{Dof <: DofManager} <: AbstractAssembler{Dof}
  dof::Dof
  handle::Ptr{Cvoid}
  sparse_matrix_type::Symbol
  matrix_free::Bool
  keep::Vector{Any}             # host arrays referenced by the descriptors during create
end

# users register their physics types:  FECB200.physics_id(::Poisson) = Int32(1)   (include/fecb200.h enums)
physics_id(physics) = error("no CUDA implementation registered for $(typeof(physics)) (fecb200 has no CPU fallback)")
elem_id(name::String) = Dict("QUAD4" => QUAD4, "TRI3" => TRI3, "HEX8" => HEX8, "TETRA4" => TET4, "TETRA10" => TET10)[name]

kind(f) = f === FiniteElementContainers.residual ? RESIDUAL :
          f === FiniteElementContainers.stiffness || f === FiniteElementContainers.stiffness_action ? STIFFNESS :
          f === FiniteElementContainers.mass || f === FiniteElementContainers.mass_action ? MASS :
          error("fecb200 assembles only the shipped element functions; got $f")

function B200Assembler(dof::DofManager, p; sparse_matrix_type = :csr, matrix_free = false, device = 0)
  fspace = function_space(dof)
  nb = FiniteElementContainers.num_blocks(fspace)
  keep = Any[]
  blocks = Vector{BlockDesc}(undef, nb)
  for b in 1:nb
    ref_fe = values(fspace.ref_fes)[b]
    conn = collect(vec(FiniteElementContainers.connectivity(fspace.elem_conns, b)))     # Int64, 1-based
    nq = FiniteElementContainers.num_cell_quadrature_points(ref_fe)
    # ref_fe.cell_interps[q] = (N, grad_N_xi, w), flattened as N[q*nnpe + a], dN[(q*nnpe + a)*nd + j]
    nnpe = size(ref_fe.cell_interps[1].N, 1)
    nd = size(ref_fe.cell_interps[1].∇N_ξ, 2)
    N  = Float64[ref_fe.cell_interps[q].N[a] for q in 1:nq for a in 1:nnpe]
    dN = Float64[ref_fe.cell_interps[q].∇N_ξ[a, j] for q in 1:nq for a in 1:nnpe for j in 1:nd]
    w  = Float64[ref_fe.cell_interps[q].JxW for q in 1:nq]
    props = collect(Float64, values(p.properties)[b])
    physics = values(p.physics)[b]
    push!(keep, conn, N, dN, w, props)
    blocks[b] = BlockDesc(elem_id(fspace.elem_types[b]), Int32(nnpe), fspace.elem_conns.nelems[b], pointer(conn),
                          Int32(nq), pointer(N), pointer(dN), pointer(w), physics_id(physics), Int32(length(props)),
                          pointer(props), Int32(FiniteElementContainers.num_states(physics)))
  end
  X = fspace.coords.data
  mesh = MeshDesc(size(fspace.coords, 2), Int32(size(fspace.coords, 1)), Int32(size(dof, 1)), Int32(nb),
                  pointer(blocks), pointer(X))
  opts = Opts(sparse_matrix_type == :csr ? CSR : CSC, Int32(FiniteElementContainers._is_condensed(dof)),
              Int32(matrix_free), Int32(device), Int32(0), (Int32(0), Int32(0), Int32(0)))
  h = Ref{Ptr{Cvoid}}(C_NULL)
  GC.@preserve keep blocks X check(ccall((:fecb200_create, LIB), Cint, (Ref{MeshDesc}, Ref{Opts}, Ref{Ptr{Cvoid}}), mesh, opts, h))
  asm = B200Assembler{typeof(dof)}(dof, h[], sparse_matrix_type, matrix_free, keep)
  finalizer(a -> ccall((:fecb200_destroy, LIB), Cint, (Ptr{Cvoid},), a.handle), asm)
  return asm
end

function update_dofs!(asm::B200Assembler, dbcs, pbcs)
  ddofs = length(dbcs) > 0 ? FiniteElementContainers.dirichlet_dofs(dbcs) : Int[]
  a, b = FiniteElementContainers.periodic_dofs(pbcs)
  check(ccall((:fecb200_update_dofs, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Int64, Ptr{Int64}, Ptr{Int64}, Int64),
              asm.handle, ddofs, length(ddofs), a, b, length(a)))
  push_dirichlet_values!(asm, dbcs)
end

# ---- values that change in time: update_time!(p) / update_bc_values!(p, asm) (src/Parameters.jl:358, 447) -------------
# The host containers stay the reference's own; after the reference's update has filled them, the device copies are
# refreshed.  Time-dependent Dirichlet values therefore reach the device at EVERY update_bc_values!, not only at
# update_dofs!.
function push_dirichlet_values!(asm::B200Assembler, dbcs)
  cache = dbcs.bc_cache
  check(ccall((:fecb200_set_dirichlet_values, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Float64}, Int64),
              asm.handle, cache.dofs, cache.vals, length(cache.dofs)))
end
function push_periodic_values!(asm::B200Assembler, pbcs)
  vals = length(pbcs) > 0 ? FiniteElementContainers.periodic_values(pbcs) : Float64[]
  if any(!iszero, vals)
    check(ccall((:fecb200_set_periodic_values, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64), asm.handle, vals, length(vals)))
  else   # NULL = every jump zero (a jump that returned to 0 must not leave a stale value on the device)
    check(ccall((:fecb200_set_periodic_values, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64), asm.handle, C_NULL, 0))
  end
end
function update_time!(asm::B200Assembler, p)
  FiniteElementContainers.update_time!(p)
  check(ccall((:fecb200_set_time, LIB), Cint, (Ptr{Cvoid}, Float64, Float64), asm.handle, p.times.time_current, p.times.Δt))
end
function update_bc_values!(p, asm::B200Assembler)
  invoke(update_bc_values!, Tuple{typeof(p), AbstractAssembler}, p, asm)     # the reference fills its host caches
  push_dirichlet_values!(asm, p.dirichlet_bcs)
  push_periodic_values!(asm, asm.periodic_bcs)
  push_neumann_values!(asm, p)
  push_source_values!(asm, p)
  push_robin_values!(asm, p)
  push_poisson_sources!(asm, p)
end
# p.field <- BC values, unknowns, periodic copies in one launch (src/Parameters.jl:404-413); evolve! calls it after the solve
_update_for_assembly!(p, asm::B200Assembler, Uu) =
  GC.@preserve Uu check(ccall((:fecb200_update_field, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), asm.handle, _ptr(Uu)))
function field!(out::Vector{Float64}, asm::B200Assembler, which::Int32 = FIELD_U)
  check(ccall((:fecb200_field_copy, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}), asm.handle, which, out))
  return out
end

# Poisson.func (test/poisson/TestPoissonCommon.jl:4-6, used at :75-83) is a closure: evaluated HERE at the quadrature
# points X_q = sum_a N_a x_a of every element, like _update_source_values! does for Sources (src/bcs/Sources.jl:55-66),
# and uploaded as f_q[NQ, NE].  Register the source of a physics type with  FECB200.source_function(ph::Poisson) = ph.func
source_function(physics) = nothing
function push_poisson_sources!(asm::B200Assembler, p)
  fspace = function_space(asm.dof)
  X = fspace.coords
  t = p.times.time_current
  for b in 1:FiniteElementContainers.num_blocks(fspace)
    f = source_function(values(p.physics)[b])
    f === nothing && continue
    ref_fe = values(fspace.ref_fes)[b]
    conns = FiniteElementContainers.connectivity(fspace.elem_conns, b)
    nq, ne = FiniteElementContainers.num_cell_quadrature_points(ref_fe), size(conns, 2)
    fq = Matrix{Float64}(undef, nq, ne)
    for e in 1:ne, q in 1:nq
      N = ref_fe.cell_interps[q].N
      Xq = sum(N[a] * X[:, conns[a, e]] for a in axes(conns, 1))
      fq[q, e] = f(Xq, t)
    end
    check(ccall((:fecb200_set_source_q, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}), asm.handle, Int32(b - 1), fq))
  end
end

# state variables [NS, NQ, NE] per block (src/Parameters.jl:1-23): which = 0 state_old, 1 state_new
function set_state!(asm::B200Assembler, b::Integer, state::Array{Float64, 3}; which = 0)
  check(ccall((:fecb200_state_set, LIB), Cint, (Ptr{Cvoid}, Int32, Int32, Ptr{Float64}), asm.handle, Int32(b - 1), Int32(which), state))
end
function get_state!(state::Array{Float64, 3}, asm::B200Assembler, b::Integer; which = 1)
  check(ccall((:fecb200_state_get, LIB), Cint, (Ptr{Cvoid}, Int32, Int32, Ptr{Float64}), asm.handle, Int32(b - 1), Int32(which), state))
  return state
end
swap_state!(asm::B200Assembler) = check(ccall((:fecb200_state_swap, LIB), Cint, (Ptr{Cvoid},), asm.handle))

# Uu / Vu / outputs may be Vector{Float64} (host) or CuArray{Float64} (device, used in place)
_ptr(x::Vector{Float64}) = pointer(x)
_ptr(x) = reinterpret(Ptr{Float64}, pointer(x))          # CuArray: device pointer

function assemble_vector!(asm::B200Assembler, f::F, Uu, p) where F <: Function
  GC.@preserve Uu check(ccall((:fecb200_assemble_vector, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}), asm.handle, kind(f), _ptr(Uu)))
end
# assemble_lumped_mass! (src/assemblers/LumpedMass.jl:32-60) / assemble_diagonal! (src/assemblers/Diagonal.jl:16-74)
function assemble_lumped_mass!(asm::B200Assembler, f::F, Uu, p) where F <: Function
  f === FiniteElementContainers.lumped_mass || error("fecb200 assembles only the shipped element functions; got $f")
  GC.@preserve Uu check(ccall((:fecb200_assemble_vector, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}), asm.handle, LUMPED_MASS, _ptr(Uu)))
end
function assemble_diagonal!(asm::B200Assembler, f::F, Uu, p) where F <: Function
  k = kind(f) == STIFFNESS ? DIAGONAL_STIFFNESS : kind(f) == MASS ? DIAGONAL_MASS : error("assemble_diagonal!: stiffness or mass")
  GC.@preserve Uu check(ccall((:fecb200_assemble_vector, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}), asm.handle, k, _ptr(Uu)))
end
function _vector_values(asm::B200Assembler)
  out = zeros(_sizes(asm)[3])
  check(ccall((:fecb200_vector_values, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), asm.handle, out))
  return out
end
lumped_mass(asm::B200Assembler) = _vector_values(asm)
diagonal(asm::B200Assembler) = _vector_values(asm)

# assemble_scalar!(asm, energy, Uu, p) (src/assemblers/QuadratureQuantity.jl:4-14); values per block as [NQ, NE]
function assemble_scalar!(asm::B200Assembler, f::F, Uu, p) where F <: Function
  f === FiniteElementContainers.energy || error("fecb200 assembles only the shipped element functions; got $f")
  GC.@preserve Uu check(ccall((:fecb200_assemble_scalar, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), asm.handle, _ptr(Uu)))
end
function scalar_values(asm::B200Assembler, b::Integer, nq::Integer, ne::Integer)
  out = Matrix{Float64}(undef, nq, ne)
  check(ccall((:fecb200_scalar_values, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}), asm.handle, Int32(b - 1), out))
  return out
end

# External loads of the residual (src/Solvers.jl:133-137).  The containers stay the reference's own
# (p.neumann_bcs::NeumannBCs, p.sources::Sources, filled by update_bc_values! on the host); their geometry is pushed
# once, their values whenever they change, and each assemble call adds the device-cached load vector.
# surface_tables(ref_fe): (Ns[nnps, nqs], dNs[ND-1, nnps, nqs], ws[nqs]) read from ref_fe.surface_interps of
# ReferenceFiniteElements (column-major Julia arrays are exactly the row-major [q][a][k] tables of the header)
function surface_tables(ref_fe)
  si = ref_fe.surface_interps
  nqs = ReferenceFiniteElements.num_surface_quadrature_points(ref_fe)
  Ns = reduce(hcat, [collect(si[q, 1].N_reduced) for q in 1:nqs])
  dNs = cat([permutedims(collect(si[q, 1].∇N_ξ_reduced)) for q in 1:nqs]...; dims = 3)
  ws = [si[q, 1].w for q in 1:nqs]
  return Ns, dNs, ws
end
function push_neumann_bcs!(asm::B200Assembler, p)
  fspace = function_space(asm.dof)
  for (i, cache) in enumerate(p.neumann_bcs.bc_caches)
    ref_fe = fspace.ref_fes[p.neumann_bcs.block_ids[i]]
    nqs = size(cache.vals, 1); nsides = length(cache.sides)
    snodes = reduce(hcat, [collect(surface_connectivity(ref_fe, cache.element_conns.data, cache.sides[e], e, 1)) for e in 1:nsides])
    Ns, dNs, ws = surface_tables(ref_fe)   # N_reduced, its parametric gradient and the weights of the surface rule
    check(ccall((:fecb200_set_neumann_bc, LIB), Cint,
                (Ptr{Cvoid}, Int32, Int64, Int32, Int32, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                asm.handle, Int32(i - 1), nsides, Int32(size(snodes, 1)), Int32(nqs), Int64.(snodes), Ns, dNs, ws))
  end
  push_neumann_values!(asm, p)
end
function push_neumann_values!(asm::B200Assembler, p)
  for (i, cache) in enumerate(p.neumann_bcs.bc_caches)
    check(ccall((:fecb200_set_neumann_values, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}), asm.handle, Int32(i - 1),
                reinterpret(Float64, vec(cache.vals))))
  end
end

# Robin BCs (src/bcs/RobinBCs.jl): geometry like a Neumann BC; the flux law in affine form g0 + D u_q.  The reference
# holds the closure and its ForwardDiff Jacobian (RobinBCFunction, :66-75): g0 = func(X_q, t, 0), D = dfuncdu(X_q, t, 0);
# a law that is not affine in u cannot run on the device.
function push_robin_bcs!(asm::B200Assembler, p)
  fspace = function_space(asm.dof)
  for (i, cache) in enumerate(p.robin_bcs.bc_caches)
    ref_fe = fspace.ref_fes[p.robin_bcs.block_ids[i]]
    nqs = size(cache.vals, 1); nsides = length(cache.sides)
    snodes = reduce(hcat, [collect(surface_connectivity(ref_fe, cache.element_conns.data, cache.sides[e], e, 1)) for e in 1:nsides])
    Ns, dNs, ws = surface_tables(ref_fe)
    check(ccall((:fecb200_set_robin_bc, LIB), Cint,
                (Ptr{Cvoid}, Int32, Int64, Int32, Int32, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                asm.handle, Int32(i - 1), nsides, Int32(size(snodes, 1)), Int32(nqs), Int64.(snodes), Ns, dNs, ws))
  end
  push_robin_values!(asm, p)
end
function push_robin_values!(asm::B200Assembler, p)
  fspace = function_space(asm.dof)
  NF = size(asm.dof, 1)
  t = p.times.time_current
  for (i, cache) in enumerate(p.robin_bcs.bc_caches)
    ref_fe = fspace.ref_fes[p.robin_bcs.block_ids[i]]
    func = p.robin_bcs.bc_funcs[i]
    nqs, nsides = size(cache.vals)
    g0 = Array{Float64, 3}(undef, NF, nqs, nsides)
    D = Array{Float64, 4}(undef, NF, NF, nqs, nsides)
    u0 = zeros(SVector{NF, Float64})
    for e in 1:nsides
      conn = FiniteElementContainers.connectivity(ref_fe, cache.element_conns.data, e, 1)
      X_el = FiniteElementContainers._element_level_fields(fspace.coords, ref_fe, conn)
      for q in 1:nqs
        interps = FiniteElementContainers.MappedH1OrL2SurfaceInterpolants(ref_fe, X_el, q, cache.sides[e])
        g0[:, q, e] .= func.func(interps.X_q, t, u0)
        D[:, :, q, e] .= func.dfuncdu(interps.X_q, t, u0)
      end
    end
    check(ccall((:fecb200_set_robin_values, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}), asm.handle, Int32(i - 1), g0, D))
  end
end
assemble_vector_robin_bc!(asm::B200Assembler, Uu, p) =
  check(ccall((:fecb200_assemble_vector_robin_bc, LIB), Cint, (Ptr{Cvoid},), asm.handle))
assemble_matrix_robin_bc!(asm::B200Assembler, Uu, p) =
  check(ccall((:fecb200_assemble_matrix_robin_bc, LIB), Cint, (Ptr{Cvoid},), asm.handle))
function push_source_values!(asm::B200Assembler, p)
  for (b, block_id) in enumerate(p.sources.block_id_to_source)
    block_id == -1 && continue
    check(ccall((:fecb200_set_source_values, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}), asm.handle, Int32(b - 1),
                reinterpret(Float64, vec(p.sources.source_caches[block_id].vals))))
  end
end
assemble_vector_neumann_bc!(asm::B200Assembler, Uu, p) =
  check(ccall((:fecb200_assemble_vector_neumann_bc, LIB), Cint, (Ptr{Cvoid},), asm.handle))
assemble_vector_source!(asm::B200Assembler, Uu, p) =
  check(ccall((:fecb200_assemble_vector_source, LIB), Cint, (Ptr{Cvoid},), asm.handle))

function assemble_stiffness!(asm::B200Assembler, f::F, Uu, p) where F <: Function
  GC.@preserve Uu check(ccall((:fecb200_assemble_matrix, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}), asm.handle, STIFFNESS, _ptr(Uu)))
end
function assemble_mass!(asm::B200Assembler, f::F, Uu, p) where F <: Function
  GC.@preserve Uu check(ccall((:fecb200_assemble_matrix, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}), asm.handle, MASS, _ptr(Uu)))
end
function assemble_matrix_action!(asm::B200Assembler, f::F, Uu, Vu, p) where F <: Function
  GC.@preserve Uu Vu check(ccall((:fecb200_assemble_action, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}),
                                 asm.handle, kind(f), _ptr(Uu), _ptr(Vu)))
end
assemble_matrix_free_action!(asm::B200Assembler, f::F, Uu, Vu, p) where F <: Function = assemble_matrix_action!(asm, f, Uu, Vu, p)
# assemble_matrix_free_action_full! (src/assemblers/MatrixAction.jl:99-149): caller passes full-length U, v
function assemble_matrix_free_action_full!(asm::B200Assembler, f::F, U_full, v_full, p) where F <: Function
  GC.@preserve U_full v_full check(ccall((:fecb200_assemble_action_full, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}),
                                         asm.handle, kind(f), _ptr(U_full), _ptr(v_full)))
end
# one pass for what solve! asks back to back at the same Uu (src/Solvers.jl:133-140)
function assemble_vector_and_stiffness!(asm::B200Assembler, Uu, p)
  GC.@preserve Uu check(ccall((:fecb200_assemble_vector_and_matrix, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), asm.handle, _ptr(Uu)))
end
# y = stiffness(asm) * x on the device-resident values (the product Krylov forms, src/Solvers.jl:144)
function matrix_multiply!(y, asm::B200Assembler, x; kind = STIFFNESS)
  GC.@preserve x y check(ccall((:fecb200_matrix_multiply, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}), asm.handle, kind, _ptr(x), _ptr(y)))
  return y
end

function _sizes(asm::B200Assembler)
  a, b, c = Ref{Int64}(0), Ref{Int64}(0), Ref{Int64}(0)
  check(ccall((:fecb200_sizes, LIB), Cint, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}, Ref{Int64}), asm.handle, a, b, c))
  return a[], b[], c[]
end
create_unknowns(asm::B200Assembler) = zeros(_sizes(asm)[3])

function residual(asm::B200Assembler)
  out = zeros(_sizes(asm)[3])
  check(ccall((:fecb200_residual, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), asm.handle, out))
  return out
end
function hvp(asm::B200Assembler, v)
  out = zeros(_sizes(asm)[3])
  GC.@preserve v check(ccall((:fecb200_hvp, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), asm.handle, _ptr(v), out))
  return out
end

function _sparse(asm::B200Assembler, k::Int32)
  n, nnz = Ref{Int64}(0), Ref{Int64}(0)
  check(ccall((:fecb200_pattern_sizes, LIB), Cint, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}), asm.handle, n, nnz))
  ptr, idx, nz = Vector{Int64}(undef, n[] + 1), Vector{Int64}(undef, nnz[]), Vector{Float64}(undef, nnz[])
  check(ccall((:fecb200_pattern_copy, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), asm.handle, ptr, idx))
  check(ccall((:fecb200_matrix_values, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}), asm.handle, k, nz))
  return asm.sparse_matrix_type == :csr ? SparseMatrixCSR{1}(n[], n[], ptr, idx, nz) : SparseMatrixCSC(n[], n[], ptr, idx, nz)
end
stiffness(asm::B200Assembler) = asm.matrix_free ? spzeros(_sizes(asm)[3], _sizes(asm)[3]) : _sparse(asm, STIFFNESS)
mass(asm::B200Assembler) = asm.matrix_free ? spzeros(_sizes(asm)[3], _sizes(asm)[3]) : _sparse(asm, MASS)


# ---- multi-GPU (ext/PartitionedArraysExt.jl:223-233, 449-481, 522-540): one Julia process (MPI rank) per GPU -----------
# `l2o` / own-ghost layout are PartitionedArrays' LocalIndices; the lists below are its assembly / consistency neighbours.
#   partition!(asm, n_owned_nodes, block_is_halo)      rank-local view: owned rows only, halo-element block in the Jacobian
#   halo_setup!(asm, ranks, send, recv)                ghosts my owned elements add to  /  my owned nodes neighbours add to
#   ghost_setup!(asm, ranks, own, ghost)               consistent!: my owned nodes each neighbour ghosts / my ghosts by owner
#   comm_init!(asm, comm)                              ncclUniqueId of rank 0, MPI.Bcast, ncclCommInitRank inside the library
#   halo_sum!(asm, FIELD_RESIDUAL)                     assembly of a PVector (ghost -> owner), NCCL or fused peer memory
# After comm_init!, fecb200_cg_solve / fecb200_newton_solve / matrix_multiply! run distributed (dots over owned entries +
# ncclAllReduce); iteration counts equal the serial solve's (tests/run_comm_check.py).
function _csr_lists(lists::Vector{Vector{Int64}})
  ptr = Int64[0; cumsum(length.(lists))]
  return ptr, reduce(vcat, lists; init = Int64[])
end
partition!(asm::B200Assembler, n_owned_nodes::Integer, block_is_halo::Vector{Int32}) =
  check(ccall((:fecb200_partition_setup, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Int32}), asm.handle, n_owned_nodes, block_is_halo))
function halo_setup!(asm::B200Assembler, ranks::Vector{Int32}, send::Vector{Vector{Int64}}, recv::Vector{Vector{Int64}})
  sp, sn = _csr_lists(send); rp, rn = _csr_lists(recv)
  check(ccall((:fecb200_halo_setup, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}),
              asm.handle, Int32(length(ranks)), ranks, sp, sn, rp, rn))
end
function ghost_setup!(asm::B200Assembler, ranks::Vector{Int32}, own::Vector{Vector{Int64}}, ghost::Vector{Vector{Int64}})
  op, on = _csr_lists(own); gp, gn = _csr_lists(ghost)
  check(ccall((:fecb200_ghost_setup, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}),
              asm.handle, Int32(length(ranks)), ranks, op, on, gp, gn))
end
function comm_init!(asm::B200Assembler, rank::Integer, nranks::Integer, bcast!)   # bcast!(id::Vector{UInt8}) = MPI.Bcast!(id, 0, comm)
  id = zeros(UInt8, 128)
  rank == 0 && check(ccall((:fecb200_comm_unique_id, LIB), Cint, (Ptr{UInt8},), id))
  bcast!(id)
  check(ccall((:fecb200_comm_init, LIB), Cint, (Ptr{Cvoid}, Int32, Int32, Ptr{UInt8}), asm.handle, Int32(rank), Int32(nranks), id))
end
halo_sum!(asm::B200Assembler, which::Int32 = FIELD_RESIDUAL) = check(ccall((:fecb200_halo_sum, LIB), Cint, (Ptr{Cvoid}, Int32), asm.handle, which))
halo_update!(asm::B200Assembler, which::Int32 = FIELD_U) = check(ccall((:fecb200_halo_update, LIB), Cint, (Ptr{Cvoid}, Int32), asm.handle, which))
comm_barrier!(asm::B200Assembler) = check(ccall((:fecb200_comm_barrier, LIB), Cint, (Ptr{Cvoid},), asm.handle))
enable_peer_halo!(asm::B200Assembler, which::Int32 = FIELD_RESIDUAL) = check(ccall((:fecb200_comm_peer_enable, LIB), Cint, (Ptr{Cvoid}, Int32), asm.handle, which))
function owned_length(asm::B200Assembler)
  n = Ref{Int64}(0)
  check(ccall((:fecb200_owned_length, LIB), Cint, (Ptr{Cvoid}, Ref{Int64}), asm.handle, n))
  return n[]
end

# ---- device-resident solvers (src/Solvers.jl:128-220): Uu is updated in place ---------------------------------------------
function newton_solve!(asm::B200Assembler, Uu; max_iters = 10, tol = 1e-12, matrix_free = asm.matrix_free)
  nit, cgit, rn = Ref{Int32}(0), Ref{Int64}(0), Ref{Float64}(0.0)
  GC.@preserve Uu check(ccall((:fecb200_newton_solve, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int32, Float64, Int32, Ref{Int32}, Ref{Int64}, Ref{Float64}),
                              asm.handle, _ptr(Uu), Int32(max_iters), tol, Int32(matrix_free), nit, cgit, rn))
  return nit[], cgit[], rn[]
end
function cg_solve!(x, asm::B200Assembler, b; atol = -1.0, rtol = -1.0, itmax = 0, matrix_free = asm.matrix_free)
  its, rn = Ref{Int64}(0), Ref{Float64}(0.0)
  GC.@preserve x b check(ccall((:fecb200_cg_solve, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Float64, Float64, Int64, Int32, Ref{Int64}, Ref{Float64}),
                               asm.handle, _ptr(b), _ptr(x), atol, rtol, itmax, Int32(matrix_free), its, rn))
  return its[], rn[]
end

# 1: the block's element kernels take the Walsh-Hadamard form (HEX8, trilinear table on a 2-point rule per axis, any
# node / point numbering of the ReferenceFE tables handed over at construction), 0: the plain quadrature loop
function block_kernel_form(asm::B200Assembler, block::Integer)
  f = Ref{Int32}(0)
  check(ccall((:fecb200_block_kernel_form, LIB), Cint, (Ptr{Cvoid}, Int32, Ref{Int32}), asm.handle, Int32(block - 1), f))
  return Int(f[])
end

end # module
