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
                                assemble_matrix_action!, assemble_matrix_free_action!, residual, stiffness, mass, hvp,
                                update_dofs!, create_unknowns, function_space, assemble_vector_neumann_bc!,
                                assemble_vector_source!, surface_connectivity
using SparseArrays, SparseMatricesCSR
import ReferenceFiniteElements

const LIB = get(ENV, "FECB200_LIB", joinpath(@__DIR__, "..", "lib", "libfecb200.so"))

# enums of include/fecb200.h
const QUAD4, TRI3, HEX8, TET4, TET10 = Int32(1), Int32(2), Int32(3), Int32(4), Int32(5)
const RESIDUAL, STIFFNESS, MASS = Int32(1), Int32(2), Int32(3)
const LUMPED_MASS, DIAGONAL_STIFFNESS, DIAGONAL_MASS = Int32(4), Int32(5), Int32(6)
const CSC, CSR = Int32(1), Int32(2)

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
  # Dirichlet values come from update_bc_values!(p, asm): push cache.dofs / cache.vals
  cache = dbcs.bc_cache
  check(ccall((:fecb200_set_dirichlet_values, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Float64}, Int64),
              asm.handle, cache.dofs, cache.vals, length(cache.dofs)))
end

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
    check(ccall((:fecb200_set_neumann_values, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}), asm.handle, Int32(i - 1),
                reinterpret(Float64, vec(cache.vals))))
  end
end
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

end # module
