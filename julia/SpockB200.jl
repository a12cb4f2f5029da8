# SpockB200.jl -- Julia binding of libspock_b200.so for kul-optec/spock.jl.
#
# Drop-in for the hot path only: `build_model` keeps building the reference's Julia problem structs
# (scen_tree, cost, dynamics, rms, constraints) and this file replaces what happens after them:
#   build_cpock / build_spock  (src/models/cpock.jl:1, src/models/spock.jl:1)   -> spk_create
#   solve_model!               (src/models/cpock.jl:169, src/models/spock.jl:231) -> spk_solve
# `model.solver_state.z`, `.v` stay host `Vector{Float64}` in the reference layout and are synchronised with the
# device around each solve, so examples/*.jl and test/*.jl read their results exactly as before.
#
# NOTE: Julia is not available in the build image of this repository; this file is mechanical (one ccall per
# entry point of include/spock_b200.h) and has not been executed there.  The same C ABI is exercised by the Python
# ctypes binding spock_b200/_lib.py, which is what the test-suite runs.
module SpockB200

import LinearAlgebra as LA

const LIB = get(ENV, "SPOCK_B200_LIB", joinpath(@__DIR__, "..", "spock_b200", "csrc", "libspock_b200.so"))

# spk_problem, field for field (include/spock_b200.h)
struct SpkProblem
  N::Int32; d::Int32; nx::Int32; nu::Int32
  A::Ptr{Float64}; B::Ptr{Float64}
  cost_dense::Int32; n_sqrtQ::Int32; sqrtQ::Ptr{Float64}; sqrtR::Ptr{Float64}
  n_sqrtQN::Int32; sqrtQN::Ptr{Float64}
  ny::Int32; E::Ptr{Float64}; F::Ptr{Float64}; b::Ptr{Float64}
  n_cones::Int32; cone_type::Ptr{Int32}; cone_dim::Ptr{Int32}
  x_min::Float64; x_max::Float64; u_min::Float64; u_max::Float64
  algorithm::Int32; L_norm::Float64; factor_mode::Int32; device::Int32
  rank::Int32; world_size::Int32; comm_id::Ptr{UInt8}
end

struct SpkResult
  passes::Int32; printed_iter::Int32; converged::Int32
  xi1::Float64; xi2::Float64; xi::Float64
  k1_count::Int32; k2_count::Int32; fallback_count::Int32
  solve_ms::Float64
end

check(rc) = rc == 0 ? nothing : error(unsafe_string(ccall((:spk_last_error, LIB), Cstring, ())))

# MOI cone types -> spk_cone (risk_measures.jl:5)
cone_code(c) = c isa Main.MOI.Zeros ? 0 : c isa Main.MOI.Nonnegatives ? 1 : c isa Main.MOI.Nonpositives ? 2 :
               c isa Main.MOI.Reals ? 3 : c isa Main.MOI.SecondOrderCone ? 4 : error("unsupported cone")

flat(ms::Vector{<:LA.Diagonal}) = reduce(vcat, [Vector{Float64}(m.diag) for m in ms])
flat(ms::Vector{<:AbstractMatrix}) = reduce(vcat, [vec(Matrix{Float64}(m)) for m in ms])   # column-major already

mutable struct Model
  handle::Ptr{Cvoid}
  nz::Int; nv::Int
  z::Vector{Float64}; v::Vector{Float64}       # == model.solver_state.z / .v of the reference
  x_inds::UnitRange{Int}; u_inds::UnitRange{Int}; s_inds::UnitRange{Int}; tau_inds::UnitRange{Int}; y_inds::UnitRange{Int}
  keep::Vector{Any}
end

"""
    build_model(scen_tree, cost, dynamics, rms, constraints, algorithm = 1; L_norm = 0.0, factor_mode = 0)

Same arguments as `spock.build_model` (src/model.jl:273); `algorithm` 1 = CP, 2 = SP.
`map(sqrt, cost.Q)` stays here like in cpock.jl:119-121.
"""
function build_model(scen_tree, cost, dynamics, rms, constraints, algorithm::Integer = 1; L_norm = 0.0, factor_mode = 0)
  nx, nu = size(dynamics.B[1])
  d = length(dynamics.A)
  uniform(ms) = all(m -> m == ms[1], ms) ? ms[1:1] : ms
  sQ, sR, sQN = map(sqrt, uniform(cost.Q)), map(sqrt, uniform(cost.R)), map(sqrt, uniform(cost.QN))
  dense = !(eltype(sQ) <: LA.Diagonal)
  A = reduce(vcat, [vec(a) for a in dynamics.A]); B = reduce(vcat, [vec(b) for b in dynamics.B])
  rm = rms[1]
  ct = Int32[cone_code(c) for c in rm.K.subcones]; cd = Int32[Main.MOI.dimension(c) for c in rm.K.subcones]
  E, F, b = vec(copy(rm.E)), vec(copy(rm.F)), copy(rm.b)
  fQ, fR, fQN = flat(sQ), flat(sR), flat(sQN)
  keep = Any[A, B, fQ, fR, fQN, E, F, b, ct, cd]
  p = SpkProblem(scen_tree.N, d, nx, nu, pointer(A), pointer(B), dense, length(sQ), pointer(fQ), pointer(fR),
                 length(sQN), pointer(fQN), length(b), pointer(E), pointer(F), pointer(b), length(ct), pointer(ct),
                 pointer(cd), constraints.x_min, constraints.x_max, constraints.u_min, constraints.u_max,
                 algorithm, L_norm, factor_mode, -1, 0, 1, C_NULL)
  h = Ref{Ptr{Cvoid}}(C_NULL)
  GC.@preserve keep check(ccall((:spk_create, LIB), Cint, (Ref{SpkProblem}, Ref{Ptr{Cvoid}}), p, h))
  sz = zeros(Int64, 5); check(ccall((:spk_sizes, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}), h[], sz))
  zo = zeros(Int64, 5); vo = zeros(Int64, 11)
  check(ccall((:spk_offsets, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), h[], zo, vo))
  nz, nv = sz[4], sz[5]
  ends = [zo[2:5]; nz]
  m = Model(h[], nz, nv, zeros(nz), zeros(nv), (zo[1]+1:ends[1]), (zo[2]+1:ends[2]), (zo[3]+1:ends[3]),
            (zo[4]+1:ends[4]), (zo[5]+1:ends[5]), keep)
  finalizer(x -> ccall((:spk_destroy, LIB), Cint, (Ptr{Cvoid},), x.handle), m)
  return m
end

"""
    solve_model!(model, x0; tol = 1e-3, verbose = false, z0 = nothing, v0 = nothing, gamma = nothing, sigma = nothing)

src/models/cpock.jl:169-189 / spock.jl:231-251: z, v persist on the device between solves (warm start).
"""
function solve_model!(m::Model, x0::AbstractVector{Float64}; tol = 1e-3, verbose::Bool = false, z0 = nothing,
                      v0 = nothing, gamma = nothing, sigma = nothing, max_iter = 0)
  (z0 === nothing) == (v0 === nothing) || error("z0 and v0 must be given together")
  res = Ref{SpkResult}()
  cap = max_iter > 0 ? max_iter : 5000
  xi = verbose ? zeros(cap) : Float64[]; xi1 = copy(xi); xi2 = copy(xi); bc = verbose ? zeros(Int32, cap) : Int32[]
  nul(a) = a === nothing || isempty(a) ? C_NULL : pointer(a)
  GC.@preserve z0 v0 xi xi1 xi2 bc check(ccall((:spk_solve, LIB), Cint,
    (Ptr{Cvoid}, Ptr{Float64}, Float64, Int32, Float64, Float64, Int32, Ptr{Float64}, Ptr{Float64}, Ref{SpkResult},
     Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Int32),
    m.handle, x0, tol, max_iter, something(gamma, 0.0), something(sigma, 0.0), verbose, nul(z0), nul(v0), res,
    nul(xi), nul(xi1), nul(xi2), nul(bc), verbose ? cap : 0))
  check(ccall((:spk_get_iterate, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), m.handle, m.z, m.v))
  println("Finished in $(res[].printed_iter) iterations.")        # cp.jl:230 / sp.jl:468
  if verbose   # verbose=LOG side effect of the reference (cp.jl:88-96, sp.jl:306-317)
    k = res[].passes
    open(io -> foreach(x -> println(io, x), xi[1:k]), "examples/output/xi_b200.dat", "a")
  end
  return nothing
end

"""
    get_state(model) / set_state!(model, state)

Everything besides (z, v) that `solve_model!` carries over from the previous solve (model.jl:80-226): checkpoint /
resume of a warm-started MPC loop.  Layout: include/spock_b200.h.
"""
function get_state(m::Model)
  n = Ref{Int64}(0); check(ccall((:spk_state_size, LIB), Cint, (Ptr{Cvoid}, Ref{Int64}), m.handle, n))
  st = zeros(n[]); check(ccall((:spk_get_state, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), m.handle, st))
  return st
end
set_state!(m::Model, st::Vector{Float64}) = check(ccall((:spk_set_state, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), m.handle, st))

"""
    estimate_L_norm!(model; iters = 50)

Replaces the hard-coded `L_norm` of cpock.jl:62-63 / spock.jl:63-64 by a power-iteration estimate of ||L||^2.
"""
function estimate_L_norm!(m::Model; iters = 50)
  out = Ref{Float64}(0.0)
  check(ccall((:spk_estimate_L_norm, LIB), Cint, (Ptr{Cvoid}, Int32, Ref{Float64}), m.handle, iters, out))
  check(ccall((:spk_set_L_norm, LIB), Cint, (Ptr{Cvoid}, Float64), m.handle, out[]))
  return out[]
end

"""    gather_iterate!(model): multi-GPU only -- completes model.z / model.v on every rank (collective)."""
function gather_iterate!(m::Model)
  check(ccall((:spk_gather_iterate, LIB), Cint, (Ptr{Cvoid},), m.handle))
  check(ccall((:spk_get_iterate, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), m.handle, m.z, m.v))
end

end # module
