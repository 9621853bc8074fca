# file: ext/SparlectraB200Ext.jl
# Package extension: bridges Sparlectra's external-solver interface (`PFModel` / `PFSolution` /
# `AbstractExternalSolver` / `solvePf`, src/solver_interface.jl:134-145) and its contingency API
# (`runContingencies!`, src/contingency.jl:257-319) to libsblx.so, the B200 batched rectangular NR engine.
#
# Loads when the session has `using LibSblx` in addition to `using Sparlectra` (weak dependency via Julia package
# extensions - the mechanism and layout of ext/SparlectraAnalyticLoadFlowExt.jl, Project.toml:26-30).  The upstream
# side (Project.toml entries, the `b200_solver` / `runContingenciesBatched!` reachability stubs, the
# `power_flow.solver = :b200` dispatch) is `upstream.patch` next to this directory.
#
# UNTESTED AT RUNTIME: the build image has no Julia.  The same call sequence is exercised through ctypes by
# sparlectra.jl_b200/sblx/api.py (`runContingenciesBatched`, `solvePf`, `MultiEngine`) in the GPU parity tests.
module SparlectraB200Ext

using Sparlectra
using LibSblx
using SparseArrays

const _BUSTYPE = Dict(:PQ => UInt8(1), :PV => UInt8(2), :Slack => UInt8(3))

"""
    B200BatchSolver <: AbstractExternalSolver

Runs a `PFModel` as a batch of one on the GPU (`sblx_solve_one`).  Use `Sparlectra.b200_solver(...)` to construct one
without depending on this extension module directly.
"""
Base.@kwdef struct B200BatchSolver <: Sparlectra.AbstractExternalSolver
  maxIte::Int = 30
  qlimits::Bool = false
  device::Int = -1
end

function Sparlectra.solvePf(solver::B200BatchSolver, model::Sparlectra.PFModel; tol::Float64 = 1e-8, kwargs...)
  n = length(model.V0)
  Y = model.Ybus isa SparseMatrixCSC ? model.Ybus : sparse(model.Ybus)
  bt = UInt8[_BUSTYPE[t] for t in model.busType]
  opt = LibSblx.Options(; max_iter = solver.maxIte, tol = tol, qlimits_enabled = solver.qlimits ? 1 : 0, base_mva = model.baseMVA,
                        device = solver.device)
  qmin = isempty(model.qmin_pu) ? fill(-Inf, n) : Vector{Float64}(model.qmin_pu)
  qmax = isempty(model.qmax_pu) ? fill(Inf, n) : Vector{Float64}(model.qmax_pu)
  h = LibSblx.create(n, Vector{Int64}(Y.colptr), Vector{Int64}(Y.rowval), Vector{ComplexF64}(Y.nzval), Int(model.slack_idx), bt,
                     Vector{Float64}(model.Vset), Vector{ComplexF64}(model.Sspec), Vector{ComplexF64}(model.V0), qmin, qmax, nothing,
                     Int32[], Int32[], nothing, nothing, nothing, nothing, nothing, nothing, opt)
  try
    V, conv, it, res, st = LibSblx.solve_one(h)
    return Sparlectra.PFSolution(V = V, converged = conv, iters = it, residual_inf = res, meta = (solver = :b200, status = st))
  finally
    LibSblx.destroy!(h)
  end
end

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

# runpf! keywords (rectangular_network_solver.jl:1783-1883) the engine mirrors -> sblx_options fields
const _PF_KW = (:damp => :damp, :qlimits_enabled => :qlimits_enabled, :qlimit_start_iter => :qlimit_start_iter,
                :qlimit_guard_max_switches => :guard_max_switches, :freeze_after_repeated_switching => :freeze_after_repeated_switching)

"""
    engine_from_net(net; tol, maxIte, vm_min_pu, vm_max_pu, n_devices, lock_pv_to_pq_buses, pf_kwargs...) -> LibSblx.Handle

Builds the engine from a net whose node voltages are the warm start (the solved template of `runContingencies!`,
src/contingency.jl:275-287).  Requires a net without isolated buses (the full n x n Ybus of `runpf_rectangular!`,
rectangular_network_solver.jl:470-475).  `pf_kwargs` are the `runpf!` keywords the contingency path forwards; the ones
the engine mirrors become `sblx_options` fields, any other keyword is an error (never silently dropped).
"""
function engine_from_net(net::Sparlectra.Net; tol::Float64 = 1e-8, maxIte::Int = 30, vm_min_pu = 0.9, vm_max_pu = 1.1,
                         n_devices::Int = 1, lock_pv_to_pq_buses::AbstractVector{Int} = Int[], pf_kwargs...)
  Sparlectra.refreshBusTypesFromProsumers!(net)
  isempty(net.isoNodes) || error("SparlectraB200Ext: the base net must not contain isolated buses")
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
  opt = LibSblx.Options(; max_iter = maxIte, tol = tol, vm_min_pu = vm_min_pu, vm_max_pu = vm_max_pu, cooldown_iters = net.cooldown_iters,
                        q_hyst_pu = net.q_hyst_pu, base_mva = net.baseMVA)
  known = Dict(_PF_KW)
  for (k, v) in pf_kwargs
    haskey(known, k) || throw(ArgumentError("runContingenciesBatched!: runpf! keyword $(k) is not supported by the :b200 backend"))
    setfield!(opt, known[k], convert(fieldtype(LibSblx.Options, known[k]), v isa Bool ? Int32(v) : v))
  end
  h = LibSblx.create(n, Vector{Int64}(Y.colptr), Vector{Int64}(Y.rowval), Vector{ComplexF64}(Y.nzval), slack, bt, Vector{Float64}(Vset),
                     Vector{ComplexF64}(S), Vector{ComplexF64}(V0), Vector{Float64}(qmin), Vector{Float64}(qmax), Vector{Float64}(qload),
                     from, to, y11, y12, y21, y22, rating, status, opt; n_devices = n_devices)
  isempty(lock_pv_to_pq_buses) || LibSblx.set_locked_pv!(h, Vector{Int32}(lock_pv_to_pq_buses))
  return h
end

const _STATUS_TEXT = Dict(7 => "islanded without reference", 8 => "unknown branch")

"""
    run_contingencies_batched!(net, cases; vm_min_pu, vm_max_pu, maxIte, tol, retry_flat_start, n_devices, pf_kwargs...)

`runContingencies!` (src/contingency.jl:257-319) with the batched backend - same arguments, same
`Vector{ContingencyResult}` in input order, the caller's `net` is never mutated.  The base case is solved by the
reference `runpf!` WITH the caller's `pf_kwargs` (contingency.jl:275-287 passes them), the template goes to the GPU(s)
once, ALL cases are solved in one library call; `overloaded_branches` and `voltage_violations` come from the device-side
lists (isolated buses excluded, contingency.jl:150-156).
"""
function run_contingencies_batched!(net::Sparlectra.Net, cases::Vector{Sparlectra.ContingencyCase};
                                    vm_min_pu::Float64 = 0.9, vm_max_pu::Float64 = 1.1, maxIte::Int = 30, tol::Float64 = 1e-8,
                                    retry_flat_start::Bool = false, n_devices::Int = 1, pf_kwargs...)
  vm_min_pu < vm_max_pu || throw(ArgumentError("runContingencies!: vm_min_pu must be below vm_max_pu."))
  isempty(cases) && return Sparlectra.ContingencyResult[]
  template = deepcopy(net)
  base_ok = try
    _, erg = Sparlectra.runpf!(template, maxIte, tol, 0; islands_enabled = true, pf_kwargs...)
    erg == 0
  catch
    false
  end
  if !base_ok
    template = deepcopy(net); template.flatstart = true
  end
  mk(t) = engine_from_net(t; tol = tol, maxIte = maxIte, vm_min_pu = vm_min_pu, vm_max_pu = vm_max_pu, n_devices = n_devices, pf_kwargs...)
  eng = mk(template)
  idx = [Sparlectra._resolve_contingency_branch(template, c.element) for c in cases]
  raw = LibSblx.solve_outages(eng, [i === nothing ? [0] : [i] for i in idx])
  LibSblx.destroy!(eng)
  if retry_flat_start
    # one second batch from a flat start for the cases that did not converge, iterations summed (contingency.jl:195-207)
    again = [k for k in eachindex(cases) if idx[k] !== nothing && raw.converged[k] == 0 && 1 <= raw.status[k] <= 6]
    if !isempty(again)
      flat = deepcopy(template); flat.flatstart = true
      eng2 = mk(flat)
      raw2 = LibSblx.solve_outages(eng2, [[idx[k]] for k in again])
      LibSblx.destroy!(eng2)
      for (j, k) in enumerate(again)
        it1 = raw.iterations[k]
        for f in (:converged, :status, :iterations, :n_switch_events, :island_count, :final_mismatch, :vmin_pu, :vmax_pu, :max_loading_pct,
                  :n_voltage_violations, :n_overloaded, :violations, :overloads)
          getfield(raw, f)[k] = getfield(raw2, f)[j]
        end
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
      msg = get(_STATUS_TEXT, Int(raw.status[k]), "power flow did not converge (status 1)")   # runpf! returns erg = 1 for every rejection reason
      out[k] = Sparlectra.ContingencyResult(c.name, false, Int(raw.iterations[k]), NaN, NaN, NaN, String[], String[], Int(raw.island_count[k]), msg)
    else
      viol = String[get(busnames, b, Sparlectra.getCompName(template.nodeVec[b].comp)) for b in raw.violations[k]]
      over = String[Sparlectra.getCompName(template.branchVec[b].comp) for b in raw.overloads[k]]
      out[k] = Sparlectra.ContingencyResult(c.name, true, Int(raw.iterations[k]), raw.vmax_pu[k], raw.vmin_pu[k], raw.max_loading_pct[k],
                                            over, viol, Int(raw.island_count[k]), nothing)
    end
  end
  return out
end

"""
    run_injection_sweep(net, n_draws; sigma = 0.1, lo = 0.5, hi = 1.5, seed = 20260724, maxIte = 30, tol = 1e-8, pf_kwargs...)

The serial Monte-Carlo loop of examples/powerflow/mc_probabilistic_powerflow.jl:73-99 (draw load factors, `runpf!`, collect
the voltage band and loading) as ONE device-side sweep: the base case is solved by the reference `runpf!`, every draw
starts from it, the factors are generated on the GPU.  Returns the per-draw summary (`LibSblx.BatchResults`).
"""
function run_injection_sweep(net::Sparlectra.Net, n_draws::Int; sigma::Float64 = 0.1, lo::Float64 = 0.5, hi::Float64 = 1.5,
                             seed::Integer = 20260724, maxIte::Int = 30, tol::Float64 = 1e-8, pf_kwargs...)
  template = deepcopy(net)
  _, erg = Sparlectra.runpf!(template, maxIte, tol, 0; islands_enabled = true, pf_kwargs...)
  erg == 0 || error("run_injection_sweep: the base case did not converge")
  eng = engine_from_net(template; tol = tol, maxIte = maxIte, pf_kwargs...)
  try
    return LibSblx.solve_injection_sweep(eng, n_draws; seed = seed, sigma = sigma, lo = lo, hi = hi)
  finally
    LibSblx.destroy!(eng)
  end
end

end # module SparlectraB200Ext
