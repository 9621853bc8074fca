# SparlectraB200.jl - Julia host glue for libsblx.so (the B200 batched rectangular NR engine).
#
# UNTESTED AT RUNTIME: the build image has no Julia (SURVEY.md section 0).  It is written against
# include/sblx.h and the reference interfaces it plugs into, and mirrors, call for call, the Python
# host layer (sparlectra.jl_b200/sblx/api.py) that IS exercised by the GPU parity tests.
#
# Plug-in points in Sparlectra (paths relative to the Sparlectra.jl tree):
#   * `B200BatchSolver <: Sparlectra.AbstractExternalSolver` + `Sparlectra.solvePf` method
#       src/solver_interface.jl:134-145; template ext/SparlectraAnalyticLoadFlowExt.jl:44-136
#   * `runContingenciesBatched!` - same signature / return type as `runContingencies!`
#       src/contingency.jl:257-319; per-case seam `_run_one_contingency` src/contingency.jl:183-223
module SparlectraB200

using Sparlectra
using SparseArrays

const LIBSBLX = get(ENV, "SBLX_LIB", joinpath(@__DIR__, "..", "lib", "libsblx.so"))

# struct sblx_options (include/sblx.h) - field order and types must match exactly
Base.@kwdef mutable struct SblxOptions
  struct_size::Int32 = 0
  index_base::Int32 = 1
  max_iter::Int32 = 30
  qlimits_enabled::Int32 = 1
  qlimit_start_iter::Int32 = 2
  cooldown_iters::Int32 = 0
  guard_max_switches::Int32 = 10
  freeze_after_repeated_switching::Int32 = 1
  device::Int32 = -1
  max_slots::Int32 = 0
  ordering::Int32 = 0
  relax_small::Int32 = 18
  factor_mode::Int32 = 0
  reserved1::Int32 = 0
  tol::Float64 = 1e-8
  damp::Float64 = 1.0
  q_hyst_pu::Float64 = 0.0
  vm_min_pu::Float64 = 0.9
  vm_max_pu::Float64 = 1.1
  base_mva::Float64 = 100.0
  mem_fraction::Float64 = 0.6
  relax_zero_frac::Float64 = 0.08
  panel_smem_kb::Float64 = 200.0
end

# struct sblx_results: 11 pointers
struct SblxResults
  converged::Ptr{UInt8}
  status::Ptr{Int32}
  iterations::Ptr{Int32}
  n_switch_events::Ptr{Int32}
  island_count::Ptr{Int32}
  final_mismatch::Ptr{Float64}
  vmin_pu::Ptr{Float64}
  vmax_pu::Ptr{Float64}
  max_loading_pct::Ptr{Float64}
  V::Ptr{Float64}
  bus_type_final::Ptr{UInt8}
end

struct SblxError <: Exception
  code::Int
  msg::String
end

mutable struct Engine
  h::Ptr{Cvoid}
  n::Int
  function Engine(h::Ptr{Cvoid}, n::Int)
    e = new(h, n)
    finalizer(x -> (x.h != C_NULL && ccall((:sblx_destroy, LIBSBLX), Cvoid, (Ptr{Cvoid},), x.h); x.h = C_NULL), e)
    return e
  end
end

_check(rc::Integer, h::Ptr{Cvoid} = C_NULL) = rc == 0 ? nothing :
  throw(SblxError(rc, unsafe_string(ccall((:sblx_last_error, LIBSBLX), Cstring, (Ptr{Cvoid},), h))))

const _BUSTYPE = Dict(:PQ => UInt8(1), :PV => UInt8(2), :Slack => UInt8(3))

"Branch two-port table exactly as createYBUS stamps it (src/equicircuit.jl:583-645, src/branch.jl:297-310)."
function _branch_table(net::Sparlectra.Net)
  nb = length(net.branchVec)
  from = Vector{Int32}(undef, nb); to = similar(from)
  y11 = Vector{ComplexF64}(undef, nb); y12 = similar(y11); y21 = similar(y11); y22 = similar(y11)
  rating = fill(NaN, nb); status = zeros(UInt8, nb)
  for (k, br) in enumerate(net.branchVec)
    from[k] = br.fromBus; to[k] = br.toBus
    y11[k], y12[k], y21[k], y22[k] = Sparlectra.calcAdmittance(br, br.comp.cVN, net.baseMVA)
    rating[k] = (br.sn_MVA === nothing) ? NaN : Float64(br.sn_MVA)
    status[k] = (br.status == 1 && Sparlectra._branch_terminal_state(br) === :closed) ? 0x01 : 0x00
  end
  return from, to, y11, y12, y21, y22, rating, status
end

"""
    Engine(net; tol, maxIte, kwargs...) -> Engine

Builds the engine from a net whose node voltages are the warm start (the solved template of
`runContingencies!`, src/contingency.jl:275-287).  Requires a net without isolated buses (the full
n x n Ybus of `runpf_rectangular!`, rectangular_network_solver.jl:470-475).
"""
function Engine(net::Sparlectra.Net; tol::Float64 = 1e-8, maxIte::Int = 30, vm_min_pu = 0.9, vm_max_pu = 1.1,
                lock_pv_to_pq_buses::AbstractVector{Int} = Int[], kwargs...)
  # remaining kwargs are sblx_options fields under the runpf! keyword names they mirror
  # (damp, qlimits_enabled, qlimit_start_iter, guard_max_switches, ...); anything else is an error
  Sparlectra.refreshBusTypesFromProsumers!(net)
  isempty(net.isoNodes) || error("SparlectraB200.Engine: base net must not contain isolated buses")
  Y = Sparlectra.createYBUS(net = net, sparse = true)
  n = size(Y, 1)
  V0, slack = Sparlectra.initialVrect(net; flatstart = net.flatstart)
  S = Sparlectra.buildComplexSVec(net)
  Vset = Sparlectra._bus_voltage_setpoints_from_prosumers(net)
  V0[slack] = Sparlectra._apply_voltage_magnitude_preserving_angle(V0[slack], Vset[slack])
  bt = UInt8[_BUSTYPE[Symbol(Sparlectra.getNodeType(nd))] for nd in net.nodeVec]
  qmin, qmax = Sparlectra.getQLimits_pu(net)
  qload = Sparlectra.build_qload_pu(net)
  from, to, y11, y12, y21, y22, rating, status = _branch_table(net)
  opt = SblxOptions(; max_iter = maxIte, tol = tol, vm_min_pu = vm_min_pu, vm_max_pu = vm_max_pu,
                    cooldown_iters = net.cooldown_iters, q_hyst_pu = net.q_hyst_pu, base_mva = net.baseMVA, kwargs...)
  opt.struct_size = sizeof(SblxOptions)
  colptr = Vector{Int64}(Y.colptr); rowval = Vector{Int64}(Y.rowval); nz = Vector{ComplexF64}(Y.nzval)
  href = Ref{Ptr{Cvoid}}(C_NULL)
  GC.@preserve colptr rowval nz bt Vset S V0 qmin qmax qload from to y11 y12 y21 y22 rating status opt begin
    rc = ccall((:sblx_create, LIBSBLX), Cint,
      (Ref{Ptr{Cvoid}}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Int64, Ptr{UInt8}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
       Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
       Ptr{Float64}, Ptr{UInt8}, Ref{SblxOptions}),
      href, n, colptr, rowval, reinterpret(Float64, nz), slack, bt, Vset, reinterpret(Float64, S), reinterpret(Float64, V0),
      Vector{Float64}(qmin), Vector{Float64}(qmax), qload, length(from), from, to, reinterpret(Float64, y11), reinterpret(Float64, y12),
      reinterpret(Float64, y21), reinterpret(Float64, y22), rating, status, opt)
    _check(rc)
  end
  e = Engine(href[], n)
  if !isempty(lock_pv_to_pq_buses)      # runpf!(...; lock_pv_to_pq_buses = [...]) (rectangular_network_solver.jl:1811)
    lk = Vector{Int32}(lock_pv_to_pq_buses)
    GC.@preserve lk _check(ccall((:sblx_set_locked_pv, LIBSBLX), Cint, (Ptr{Cvoid}, Int64, Ptr{Int32}), e.h, length(lk), lk), e.h)
  end
  return e
end

struct BatchResults
  converged::Vector{UInt8}; status::Vector{Int32}; iterations::Vector{Int32}; n_switch_events::Vector{Int32}
  island_count::Vector{Int32}; final_mismatch::Vector{Float64}; vmin_pu::Vector{Float64}; vmax_pu::Vector{Float64}
  max_loading_pct::Vector{Float64}; V::Matrix{ComplexF64}
end

"Solve a batch of N-k outage scenarios: `cases[s]` = branch indices (1-based) taken out in scenario s."
function solve_outages(e::Engine, cases::Vector{Vector{Int}})
  B = length(cases)
  optr = Int32[0]; obr = Int32[]
  for c in cases
    append!(obr, Int32.(c)); push!(optr, Int32(length(obr)))
  end
  isempty(obr) && push!(obr, Int32(0))
  r = BatchResults(zeros(UInt8, B), zeros(Int32, B), zeros(Int32, B), zeros(Int32, B), zeros(Int32, B), zeros(B), zeros(B), zeros(B),
                   zeros(B), zeros(ComplexF64, e.n, B))
  GC.@preserve optr obr r begin
    res = SblxResults(pointer(r.converged), pointer(r.status), pointer(r.iterations), pointer(r.n_switch_events), pointer(r.island_count),
                      pointer(r.final_mismatch), pointer(r.vmin_pu), pointer(r.vmax_pu), pointer(r.max_loading_pct),
                      Ptr{Float64}(pointer(r.V)), Ptr{UInt8}(C_NULL))
    _check(ccall((:sblx_solve_outages, LIBSBLX), Cint, (Ptr{Cvoid}, Int64, Ptr{Int32}, Ptr{Int32}, Ref{SblxResults}), e.h, B, optr, obr, res), e.h)
  end
  return r
end

# ---------------------------------------------------------------------------------------------
# AbstractExternalSolver backend: a batch of one through sblx_create + sblx_solve_one on a PFModel
# ---------------------------------------------------------------------------------------------
Base.@kwdef struct B200BatchSolver <: Sparlectra.AbstractExternalSolver
  maxIte::Int = 30
  qlimits::Bool = false
end

function Sparlectra.solvePf(solver::B200BatchSolver, model::Sparlectra.PFModel; tol::Float64 = 1e-8, kwargs...)
  n = length(model.V0)
  Y = model.Ybus isa SparseMatrixCSC ? model.Ybus : sparse(model.Ybus)
  bt = UInt8[_BUSTYPE[t] for t in model.busType]
  opt = SblxOptions(; max_iter = solver.maxIte, tol = tol, qlimits_enabled = solver.qlimits ? 1 : 0, base_mva = model.baseMVA)
  opt.struct_size = sizeof(SblxOptions)
  colptr = Vector{Int64}(Y.colptr); rowval = Vector{Int64}(Y.rowval); nz = Vector{ComplexF64}(Y.nzval)
  qmin = isempty(model.qmin_pu) ? fill(-Inf, n) : Vector{Float64}(model.qmin_pu)
  qmax = isempty(model.qmax_pu) ? fill(Inf, n) : Vector{Float64}(model.qmax_pu)
  href = Ref{Ptr{Cvoid}}(C_NULL)
  Vout = zeros(ComplexF64, n); conv = Ref{UInt8}(0); it = Ref{Int32}(0); res = Ref{Float64}(0.0); st = Ref{Int32}(0)
  Sspec = Vector{ComplexF64}(model.Sspec); V0 = Vector{ComplexF64}(model.V0); Vset = Vector{Float64}(model.Vset)
  GC.@preserve colptr rowval nz bt Vset Sspec V0 qmin qmax opt Vout begin
    _check(ccall((:sblx_create, LIBSBLX), Cint,
      (Ref{Ptr{Cvoid}}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Int64, Ptr{UInt8}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
       Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
       Ptr{Float64}, Ptr{UInt8}, Ref{SblxOptions}),
      href, n, colptr, rowval, reinterpret(Float64, nz), model.slack_idx, bt, Vset, reinterpret(Float64, Sspec), reinterpret(Float64, V0),
      qmin, qmax, C_NULL, 0, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, opt))
    h = href[]
    try
      _check(ccall((:sblx_solve_one, LIBSBLX), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ref{UInt8}, Ref{Int32}, Ref{Float64}, Ref{Int32}),
                   h, C_NULL, Ptr{Float64}(pointer(Vout)), conv, it, res, st), h)
    finally
      ccall((:sblx_destroy, LIBSBLX), Cvoid, (Ptr{Cvoid},), h)
    end
  end
  return Sparlectra.PFSolution(V = Vout, converged = conv[] != 0, iters = Int(it[]), residual_inf = res[], meta = (status = Int(st[]),))
end

# ---------------------------------------------------------------------------------------------
# runContingencies! with the batched backend (src/contingency.jl:257-319)
# ---------------------------------------------------------------------------------------------
const _STATUS_TEXT = Dict(7 => "islanded without reference", 8 => "unknown branch")

function runContingenciesBatched!(net::Sparlectra.Net, cases::Vector{Sparlectra.ContingencyCase};
                                  vm_min_pu::Float64 = 0.9, vm_max_pu::Float64 = 1.1, maxIte::Int = 30, tol::Float64 = 1e-8,
                                  retry_flat_start::Bool = false, kwargs...)
  vm_min_pu < vm_max_pu || throw(ArgumentError("runContingencies!: vm_min_pu must be below vm_max_pu."))
  isempty(cases) && return Sparlectra.ContingencyResult[]
  template = deepcopy(net)                                   # the caller's net is never mutated
  base_ok = try
    _, erg = Sparlectra.runpf!(template, maxIte, tol, 0; islands_enabled = true)
    erg == 0
  catch
    false
  end
  if !base_ok
    template = deepcopy(net); template.flatstart = true
  end
  eng = Engine(template; tol = tol, maxIte = maxIte, vm_min_pu = vm_min_pu, vm_max_pu = vm_max_pu, kwargs...)
  idx = [Sparlectra._resolve_contingency_branch(template, c.element) for c in cases]
  raw = solve_outages(eng, [i === nothing ? [0] : [i] for i in idx])
  if retry_flat_start
    # one second batch from a flat start for the cases that did not converge, iterations summed (contingency.jl:195-207)
    again = [k for k in eachindex(cases) if idx[k] !== nothing && raw.converged[k] == 0 && 1 <= raw.status[k] <= 6]
    if !isempty(again)
      flat = deepcopy(template); flat.flatstart = true
      eng2 = Engine(flat; tol = tol, maxIte = maxIte, vm_min_pu = vm_min_pu, vm_max_pu = vm_max_pu, kwargs...)
      raw2 = solve_outages(eng2, [[idx[k]] for k in again])
      for (j, k) in enumerate(again)
        it1 = raw.iterations[k]
        for f in (:converged, :status, :iterations, :n_switch_events, :island_count, :final_mismatch, :vmin_pu, :vmax_pu, :max_loading_pct)
          getfield(raw, f)[k] = getfield(raw2, f)[j]
        end
        raw.V[:, k] .= view(raw2.V, :, j)
        raw.iterations[k] += it1
      end
    end
  end
  busnames = Dict{Int,String}(i => name for (name, i) in template.busDict)
  out = Vector{Sparlectra.ContingencyResult}(undef, length(cases))
  for (k, c) in enumerate(cases)
    if idx[k] === nothing
      out[k] = Sparlectra.ContingencyResult(c.name, false, 0, NaN, NaN, NaN, String[], String[], 0, "unknown branch $(c.element) (not found in net.branchVec)")
    elseif raw.converged[k] == 0
      msg = get(_STATUS_TEXT, Int(raw.status[k]), "power flow did not converge (status 1)")
      out[k] = Sparlectra.ContingencyResult(c.name, false, Int(raw.iterations[k]), NaN, NaN, NaN, String[], String[], Int(raw.island_count[k]), msg)
    else
      vm = abs.(view(raw.V, :, k))
      viol = String[get(busnames, b, string(b)) for b in eachindex(vm) if vm[b] < vm_min_pu || vm[b] > vm_max_pu]
      out[k] = Sparlectra.ContingencyResult(c.name, true, Int(raw.iterations[k]), raw.vmax_pu[k], raw.vmin_pu[k], raw.max_loading_pct[k],
                                            String[], viol, Int(raw.island_count[k]), nothing)
    end
  end
  return out
end

export B200BatchSolver, runContingenciesBatched!, Engine, solve_outages

end # module
