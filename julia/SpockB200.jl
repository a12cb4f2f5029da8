# SpockB200.jl -- Julia binding of libspock_b200.so for kul-optec/spock.jl: a drop-in for the hot path.
#
# HOW IT IS USED.  This file is `include`d INSIDE `module spock` (src/spock.jl) IN PLACE OF the eight includes that make
# up the reference's CPU hot path:
#
#     include("model.jl")                                    # unchanged: SolverOptions, CP/SP, SILENT/LOG, build_model, ...
#   - include("model_algorithms/cp.jl")                      # \
#   - include("model_algorithms/sp.jl")                      #  |
#   - include("model_algorithms/qnewton_directions/restarted_broyden.jl")
#   - include("model_algorithms/qnewton_directions/anderson.jl")      #  > replaced by
#   - include("model_dynamics/implicit_l.jl")                #  |
#   - include("models/cpock.jl")                             #  |
#   - include("models/spock.jl")                             # /
#   + include("../julia/SpockB200.jl")
#   (the last two lines of src/spock.jl, `precompile(solve_model!, (CPOCK, ...))`, become `(CPOCK_B200{Int64,Float64}, ...)`)
#
# `build_model` (src/model.jl:273-297) is untouched and keeps dispatching on `SolverOptions(CP | SP, qnewton)` to
# `build_cpock` / `build_spock`, which this file defines on top of `spk_create`; `solve_model!` keeps the reference's
# signature (src/models/cpock.jl:169-189, spock.jl:231-251) on top of `spk_solve`.  The models carry the field names every
# caller in examples/ and test/ reads -- `model.solver_state.{z, v, nz, nv, L_norm, res_0}`,
# `model.solver_state_internal.{x,u,s,tau,y}_inds` and the `v*_offset`s, `model.problem_definition.{x0, nx, nu, ...}` -- so
# examples/server_heat/{residuals,scaling,mpc_simulation}.jl and test/runtests.jl need no edit.  `z`, `v` stay host
# `Vector{Float64}` in the reference layout and are synchronised with the device around each solve.  `verbose = LOG`
# appends to the reference's files (cp.jl:88-96, sp.jl:306-317).  The inner operator interface that cp.jl / sp.jl
# dispatched on (`L!`, `L_transpose!`, `spock_mul!`, `spock_dot`, `spock_norm`, `prox_f!`, `prox_h_conj!`) is kept for
# the reference's unit tests (test/model_algorithms/prox.jl:53-81), each one `spk_apply` / `spk_dot` call.
#
# B200-specific knobs come from the environment so that no call site changes:
#   SPOCK_B200_LIB      path of libspock_b200.so
#   SPOCK_B200_DEVICES  "0,1,2,3": partition the tree over these GPUs inside this one Julia process (spk_problem.devices)
#   SPOCK_B200_FACTORS  "per_node" (default, like model.jl:99-102) | "per_stage"
#
# NOTE: Julia is not available in the build image of this repository; this file is mechanical (one ccall per entry point
# of include/spock_b200.h) and has not been executed there.  The identical C ABI is exercised by the Python ctypes binding
# spock_b200/_lib.py (struct layout checked against the header by tests/test_abi.py), which is what the test-suite runs.

const SPK_LIB = get(ENV, "SPOCK_B200_LIB", joinpath(@__DIR__, "..", "spock_b200", "csrc", "libspock_b200.so"))

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
  n_devices::Int32; devices::Ptr{Int32}
  n_rms::Int32
  mode_of_node::Ptr{Int32}
end

struct SpkResult
  passes::Int32; printed_iter::Int32; converged::Int32
  xi1::Float64; xi2::Float64; xi::Float64
  k1_count::Int32; k2_count::Int32; fallback_count::Int32
  solve_ms::Float64
end

spk_check(rc) = rc == 0 ? nothing : error(unsafe_string(ccall((:spk_last_error, SPK_LIB), Cstring, ())))

# MOI cone types -> spk_cone (risk_measures.jl:5)
spk_cone_code(c) = c isa MOI.Zeros ? 0 : c isa MOI.Nonnegatives ? 1 : c isa MOI.Nonpositives ? 2 :
                   c isa MOI.Reals ? 3 : c isa MOI.SecondOrderCone ? 4 : error("unsupported cone")

spk_flat(ms::Vector{<:LA.Diagonal}) = reduce(vcat, [Vector{Float64}(m.diag) for m in ms])
spk_flat(ms::Vector{<:AbstractMatrix}) = reduce(vcat, [vec(Matrix{Float64}(m)) for m in ms])   # column-major already

# ---- state types: the field names of GENERIC_SOLVER_STATE (model.jl:23-44) and CP_/SP_IMPLICITL_STATE_INTERNAL
#      (model.jl:80-125, 137-226) that callers read; the work vectors of the CPU path live on the device instead.
struct B200_SOLVER_STATE <: SOLVER_STATE
  nz :: Int64
  nv :: Int64
  L_norm :: Float64
  z :: Vector{Float64}
  v :: Vector{Float64}
  res_0 :: Vector{Float64}
end

mutable struct B200_STATE_INTERNAL <: SOLVER_STATE_INTERNAL
  handle :: Ptr{Cvoid}
  x_inds :: Vector{Int64}
  u_inds :: Vector{Int64}
  s_inds :: Vector{Int64}
  tau_inds :: Vector{Int64}
  y_inds :: Vector{Int64}
  v2_offset :: Int64
  v3_offset :: Int64
  v4_offset :: Int64
  v5_offset :: Int64
  v6_offset :: Int64
  v7_offset :: Int64
  v11_offset :: Int64
  v12_offset :: Int64
  v13_offset :: Int64
  v14_offset :: Int64
  last_result :: Union{SpkResult, Nothing}
  keep :: Vector{Any}
end

# model.jl:127-131 / 228-232 keep defining CPOCK / SPOCK (their internal-state types hold the CPU work vectors); the
# B200-backed models are separate types with the same three fields, returned by build_cpock / build_spock below
struct CPOCK_B200{TI, TF} <: CUSTOM_SOLVER_MODEL
  solver_state :: B200_SOLVER_STATE
  solver_state_internal :: B200_STATE_INTERNAL
  problem_definition :: GENERIC_PROBLEM_DEFINITION{TF, TI}
end
struct SPOCK_B200{TI, TF} <: CUSTOM_SOLVER_MODEL
  solver_state :: B200_SOLVER_STATE
  solver_state_internal :: B200_STATE_INTERNAL
  problem_definition :: GENERIC_PROBLEM_DEFINITION{TF, TI}
end
const MODEL_B200 = Union{CPOCK_B200, SPOCK_B200}    # plays the role of MODEL_IMPLICITL (model.jl:261)

function spk_build(scen_tree, cost, dynamics, rms, constraints, algorithm :: Int)
  nx, nu = size(dynamics.B[1])
  d = length(dynamics.A)
  # sqrt of the cost matrices stays on this side (cpock.jl:119-121).  Q and R share ONE count in the C ABI (n_sqrtQ):
  # collapse to a single entry only when BOTH are uniform over the nodes, otherwise pass both per node.
  uniform(ms) = all(m -> m == ms[1], ms)
  uq = uniform(cost.Q) && uniform(cost.R)
  Qs, Rs = uq ? cost.Q[1:1] : cost.Q, uq ? cost.R[1:1] : cost.R
  QNs = uniform(cost.QN) ? cost.QN[1:1] : cost.QN
  sQ, sR, sQN = map(sqrt, Qs), map(sqrt, Rs), map(sqrt, QNs)
  @assert length(sQ) == length(sR)
  dense = !(eltype(sQ) <: LA.Diagonal)
  A = reduce(vcat, [vec(Matrix{Float64}(a)) for a in dynamics.A]); B = reduce(vcat, [vec(Matrix{Float64}(b)) for b in dynamics.B])
  rm = rms[1]                      # ny and the sub-cone list are uniform (implicit_l.jl:14, 19); E, F, b may differ per node
  ct = Int32[spk_cone_code(c) for c in rm.K.subcones]; cd = Int32[MOI.dimension(c) for c in rm.K.subcones]
  uniform_rms = all(r -> r.E == rm.E && r.F == rm.F && r.b == rm.b, rms)
  rl = uniform_rms ? rms[1:1] : rms          # get_nonuniform_rms_avar_v2 (risk_measures.jl:127-136): one entry per non-leaf node
  E = reduce(vcat, [vec(Matrix{Float64}(r.E)) for r in rl]); F = reduce(vcat, [vec(Matrix{Float64}(r.F)) for r in rl])
  b = reduce(vcat, [Vector{Float64}(r.b) for r in rl])
  fQ, fR, fQN = spk_flat(sQ), spk_flat(sR), spk_flat(sQN)
  devs = haskey(ENV, "SPOCK_B200_DEVICES") ? Int32[parse(Int32, s) for s in split(ENV["SPOCK_B200_DEVICES"], ",")] : Int32[]
  fmode = get(ENV, "SPOCK_B200_FACTORS", "per_node") == "per_stage" ? 1 : 0
  # node_info[j].w as the caller left it (scenario_tree.jl:77); NULL when it is the generator's (j % d) + 1
  wmap = Int32[j == 1 ? 0 : scen_tree.node_info[j].w for j in 1:scen_tree.n]
  default_w = all(j -> wmap[j] == (j % d) + 1, 2:scen_tree.n)
  keep = Any[A, B, fQ, fR, fQN, E, F, b, ct, cd, devs, wmap]
  p = SpkProblem(scen_tree.N, d, nx, nu, pointer(A), pointer(B), dense, length(sQ), pointer(fQ), pointer(fR),
                 length(sQN), pointer(fQN), length(rm.b), pointer(E), pointer(F), pointer(b), length(ct), pointer(ct),
                 pointer(cd), constraints.x_min, constraints.x_max, constraints.u_min, constraints.u_max,
                 algorithm, 0.0, fmode, -1, 0, 1, C_NULL, length(devs), isempty(devs) ? C_NULL : pointer(devs), length(rl),
                 default_w ? Ptr{Int32}(C_NULL) : pointer(wmap))
  h = Ref{Ptr{Cvoid}}(C_NULL)
  GC.@preserve keep spk_check(ccall((:spk_create, SPK_LIB), Cint, (Ref{SpkProblem}, Ref{Ptr{Cvoid}}), p, h))
  sz = zeros(Int64, 5); spk_check(ccall((:spk_sizes, SPK_LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}), h[], sz))
  zo = zeros(Int64, 5); vo = zeros(Int64, 11)
  spk_check(ccall((:spk_offsets, SPK_LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), h[], zo, vo))
  nz, nv = sz[4], sz[5]
  ends = [zo[2:5]; nz]
  inds(k) = collect(zo[k] + 1 : ends[k])                                  # z_to_x ... z_to_y, implicit_l.jl:106-158
  L_norm = algorithm == 2 ? 3.28 : 10.0                                   # spock.jl:64 / cpock.jl:63
  state = B200_SOLVER_STATE(nz, nv, L_norm, zeros(nz), zeros(nv), [-Inf, -Inf])
  internal = B200_STATE_INTERNAL(h[], inds(1), inds(2), inds(3), inds(4), inds(5),
                                 vo[2], vo[3], vo[4], vo[5], vo[6], vo[7], vo[8], vo[9], vo[10], vo[11], nothing, keep)
  finalizer(x -> ccall((:spk_destroy, SPK_LIB), Cint, (Ptr{Cvoid},), x.handle), internal)
  pd = GENERIC_PROBLEM_DEFINITION(zeros(nx), nx, nu, scen_tree, rms, cost, dynamics, constraints)   # model.jl:46-55
  return state, internal, pd
end

# build_cpock (src/models/cpock.jl:1) / build_spock (src/models/spock.jl:1): what build_model (model.jl:273-297) calls
function build_cpock(scen_tree, cost, dynamics, rms, constraints)
  s, i, pd = spk_build(scen_tree, cost, dynamics, rms, constraints, 1)
  return CPOCK_B200{Int64, Float64}(s, i, pd)
end
function build_spock(scen_tree, cost, dynamics, rms, constraints, qnewton = nothing)   # qnewton is ignored by the reference too (spock.jl:212-224)
  s, i, pd = spk_build(scen_tree, cost, dynamics, rms, constraints, 2)
  return SPOCK_B200{Int64, Float64}(s, i, pd)
end

spk_nul(a) = (a === nothing || isempty(a)) ? C_NULL : pointer(a)

"""
    solve_model!(model, x0; tol = 1e-3, verbose = SILENT, z0 = nothing, v0 = nothing, gamma = nothing, sigma = nothing)

src/models/cpock.jl:169-189 / src/models/spock.jl:231-251.  z, v (and SuperMann's Anderson memory) persist on the device
between solves -- the reference's implicit warm start.  Returns `nothing`; the result is `model.solver_state.z` / `.v`.
"""
function solve_model!(model :: MODEL_B200, x0 :: AbstractArray{TF, 1}; tol :: TF = 1e-3, verbose :: VERBOSE_LEVEL = SILENT,
                      z0 :: Union{Vector{Float64}, Nothing} = nothing, v0 :: Union{Vector{Float64}, Nothing} = nothing,
                      gamma :: Union{Float64, Nothing} = nothing, sigma :: Union{Float64, Nothing} = nothing) where {TF <: Real}
  h = model.solver_state_internal.handle
  is_sp = model isa SPOCK_B200
  (z0 === nothing) == (v0 === nothing) || error("z0 and v0 must be given together")   # cpock.jl:180-183 would throw in copyto!
  x0v = Vector{Float64}(x0)
  log = verbose === LOG
  cap = is_sp ? 1000 : 5000                                                               # MAX_ITER sp.jl:360 / cp.jl:190
  xi = log ? zeros(cap) : Float64[]; xi1 = copy(xi); xi2 = copy(xi); bc = log ? zeros(Int32, cap) : Int32[]
  res = Ref{SpkResult}()
  GC.@preserve x0v z0 v0 xi xi1 xi2 bc spk_check(ccall((:spk_solve, SPK_LIB), Cint,
    (Ptr{Cvoid}, Ptr{Float64}, Float64, Int32, Float64, Float64, Int32, Ptr{Float64}, Ptr{Float64}, Ref{SpkResult},
     Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Int32),
    h, x0v, Float64(tol), 0, something(gamma, 0.0), something(sigma, 0.0), log, spk_nul(z0), spk_nul(v0), res,
    spk_nul(xi), spk_nul(xi1), spk_nul(xi2), spk_nul(bc), log ? cap : 0))
  copyto!(model.problem_definition.x0, x0v)                                               # cpock.jl:185
  spk_check(ccall((:spk_get_iterate, SPK_LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), h, model.solver_state.z, model.solver_state.v))
  model.solver_state_internal.last_result = res[]
  if log                             # the reference appends one value per iteration to these files (cp.jl:88-96, sp.jl:306-317)
    k = res[].passes
    sfx = is_sp ? "sp" : "cp"
    open(io -> writedlm(io, xi[1:k]), "examples/output/xi_$(sfx).dat", "a")
    open(io -> writedlm(io, xi1[1:k]), "examples/output/xi1_$(sfx).dat", "a")
    open(io -> writedlm(io, xi2[1:k]), "examples/output/xi2_$(sfx).dat", "a")
    is_sp && open(io -> writedlm(io, bc[1:k]), "examples/output/xi_backtrack_count.dat", "a")
  end
  println("Finished in $(res[].printed_iter) iterations.")                                # cp.jl:230 / sp.jl:468
  return nothing
end

# ---- inner operator interface (src/model_dynamics/implicit_l.jl), in place on caller-owned vectors like the reference
function spk_apply!(model, op, inp :: Vector{Float64}, out :: Vector{Float64}, scalar = 0.0)
  spk_check(ccall((:spk_apply, SPK_LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}, Float64),
                  model.solver_state_internal.handle, op, inp, out, scalar))
end
function L!(model :: MODEL_B200, z :: AbstractArray{TF, 1}, v :: AbstractArray{TF, 1}) where {TF <: Real}        # :177
  (length(z) == model.solver_state.nz && length(v) == model.solver_state.nv) || error("z or v has wrong length")    # :184-186
  out = zeros(model.solver_state.nv); spk_apply!(model, 0, Vector{Float64}(z), out); copyto!(v, out)
end
function L_transpose!(model :: MODEL_B200, z :: AbstractArray{TF, 1}, v :: AbstractArray{TF, 1}) where {TF <: Real}   # :320
  (length(z) == model.solver_state.nz && length(v) == model.solver_state.nv) || error("z or v has wrong length")    # :327-329
  out = zeros(model.solver_state.nz); spk_apply!(model, 1, Vector{Float64}(v), out); copyto!(z, out)
end
function spock_mul!(model :: MODEL_B200, C :: AbstractArray{TF, 1}, trans :: Bool, B :: AbstractArray{TF, 1}, alpha :: TF, beta :: TF) where {TF <: Real}   # :451-481
  Cv = Vector{Float64}(C)
  spk_check(ccall((:spk_mul, SPK_LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}, Float64, Float64, Ptr{Float64}),
                  model.solver_state_internal.handle, trans ? 1 : 0, Vector{Float64}(B), alpha, beta, Cv))
  copyto!(C, Cv)
end
function spock_dot(model :: MODEL_B200, a_z, a_v, b_z, b_v, gamma, sigma)                                         # :485-516
  out = Ref{Float64}(0.0)
  spk_check(ccall((:spk_dot, SPK_LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Float64, Float64, Ref{Float64}),
                  model.solver_state_internal.handle, Vector{Float64}(a_z), Vector{Float64}(a_v), Vector{Float64}(b_z), Vector{Float64}(b_v), gamma, sigma, out))
  return out[]
end
spock_norm(model :: MODEL_B200, a_z, a_v, gamma, sigma) = sqrt(spock_dot(model, a_z, a_v, a_z, a_v, gamma, sigma))   # :538-557
function prox_f!(model :: MODEL_B200, arg :: AbstractArray{TF, 1}, gamma :: TF) where {TF <: Real}                 # :729-748
  spk_check(ccall((:spk_set_x0, SPK_LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), model.solver_state_internal.handle, model.problem_definition.x0))
  out = zeros(length(arg)); spk_apply!(model, 2, Vector{Float64}(arg), out, gamma); copyto!(arg, out)
end
function prox_h_conj!(model :: MODEL_B200, arg :: AbstractArray{TF, 1}, sigma :: TF) where {TF <: Real}            # :911-951
  out = zeros(length(arg)); spk_apply!(model, 3, Vector{Float64}(arg), out, sigma); copyto!(arg, out)
end

# ---- extensions without a reference counterpart
"""Everything besides (z, v) that `solve_model!` carries over from the previous solve (model.jl:80-226): checkpoint / resume."""
function get_state(model :: MODEL_B200)
  h = model.solver_state_internal.handle
  n = Ref{Int64}(0); spk_check(ccall((:spk_state_size, SPK_LIB), Cint, (Ptr{Cvoid}, Ref{Int64}), h, n))
  st = zeros(n[]); spk_check(ccall((:spk_get_state, SPK_LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), h, st))
  return st
end
set_state!(model :: MODEL_B200, st :: Vector{Float64}) =
  spk_check(ccall((:spk_set_state, SPK_LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), model.solver_state_internal.handle, st))

"""Replaces the hard-coded `L_norm` of cpock.jl:62-63 / spock.jl:63-64 by a power-iteration estimate of ||L||^2."""
function estimate_L_norm!(model :: MODEL_B200; iters = 50)
  h = model.solver_state_internal.handle
  out = Ref{Float64}(0.0)
  spk_check(ccall((:spk_estimate_L_norm, SPK_LIB), Cint, (Ptr{Cvoid}, Int32, Ref{Float64}), h, iters, out))
  spk_check(ccall((:spk_set_L_norm, SPK_LIB), Cint, (Ptr{Cvoid}, Float64), h, out[]))
  return out[]
end
