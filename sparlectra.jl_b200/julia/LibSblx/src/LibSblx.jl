# LibSblx.jl - raw binding of libsblx.so (include/sblx.h, ABI 0.2.0).  No dependency on Sparlectra: this is the weak
# dependency whose presence activates `ext/SparlectraB200Ext.jl` inside Sparlectra (Project.toml [weakdeps] /
# [extensions], the mechanism of ext/SparlectraAnalyticLoadFlowExt.jl, Project.toml:26-30).
#
# UNTESTED AT RUNTIME: the build image has no Julia.  Struct layouts follow include/sblx.h field by field; the same
# layouts are exercised through ctypes by sparlectra.jl_b200/sblx/lib.py (tests/test_host_side.py checks their sizes).
module LibSblx

const libsblx = get(ENV, "SBLX_LIB", joinpath(@__DIR__, "..", "..", "..", "lib", "libsblx.so"))

# struct sblx_options - field order and types must match include/sblx.h exactly
Base.@kwdef mutable struct Options
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
  pivot_tol_rel::Float64 = 1e-10
end

# struct sblx_results (0.2.0): pointers, two capacities
struct Results
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
  n_voltage_violations::Ptr{Int32}
  n_overloaded::Ptr{Int32}
  n_tiny_pivots::Ptr{Int32}
  viol_capacity::Int64
  viol_scenario::Ptr{Int32}
  viol_bus::Ptr{Int32}
  viol_count::Ptr{Int64}
  over_capacity::Int64
  over_scenario::Ptr{Int32}
  over_branch::Ptr{Int32}
  over_loading_pct::Ptr{Float64}
  over_count::Ptr{Int64}
  qlog_capacity::Int64
  qlog_scenario::Ptr{Int32}
  qlog_iter::Ptr{Int32}
  qlog_bus::Ptr{Int32}
  qlog_side::Ptr{Int32}
  qlog_count::Ptr{Int64}
  flows::Ptr{Float64}
end

struct SblxError <: Exception
  code::Int
  msg::String
end

last_error(h::Ptr{Cvoid}) = unsafe_string(ccall((:sblx_last_error, libsblx), Cstring, (Ptr{Cvoid},), h))
multi_last_error(m::Ptr{Cvoid}) = unsafe_string(ccall((:sblx_multi_last_error, libsblx), Cstring, (Ptr{Cvoid},), m))
check(rc::Integer, h::Ptr{Cvoid} = C_NULL) = rc == 0 ? nothing : throw(SblxError(rc, last_error(h)))
check_multi(rc::Integer, m::Ptr{Cvoid}) = rc == 0 ? nothing : throw(SblxError(rc, multi_last_error(m)))

"One engine handle (ndev == 1: sblx_create) or one engine per GPU with an NCCL communicator (sblx_create_multi)."
mutable struct Handle
  ptr::Ptr{Cvoid}
  multi::Bool
  n::Int
  nbranch::Int
  function Handle(ptr, multi, n, nbranch)
    h = new(ptr, multi, n, nbranch)
    finalizer(destroy!, h)
    return h
  end
end

function destroy!(h::Handle)
  h.ptr == C_NULL && return
  h.multi ? ccall((:sblx_destroy_multi, libsblx), Cvoid, (Ptr{Cvoid},), h.ptr) : ccall((:sblx_destroy, libsblx), Cvoid, (Ptr{Cvoid},), h.ptr)
  h.ptr = C_NULL
  return
end

"""
    create(n, colptr, rowval, nz, slack, bt, vset, s, v0, qmin, qmax, qload, from, to, y11, y12, y21, y22, rating, status, opt; n_devices = 1)

All arrays are Julia-owned and only read during the call (`GC.@preserve`).  `n_devices > 1` builds the model on GPUs
0..n_devices-1 of this process and an `ncclCommInitAll` communicator inside the library.
"""
function create(n::Int, colptr::Vector{Int64}, rowval::Vector{Int64}, nz::Vector{ComplexF64}, slack::Int, bt::Vector{UInt8},
                vset::Vector{Float64}, s::Vector{ComplexF64}, v0::Vector{ComplexF64}, qmin, qmax, qload, from::Vector{Int32},
                to::Vector{Int32}, y11, y12, y21, y22, rating, status, opt::Options; n_devices::Int = 1)
  opt.struct_size = sizeof(Options)
  href = Ref{Ptr{Cvoid}}(C_NULL)
  p(x) = x === nothing ? Ptr{Float64}(C_NULL) : Ptr{Float64}(pointer(x))
  # ccall needs literal type tuples and explicit arguments (no splatting), hence the two spelled-out calls
  GC.@preserve colptr rowval nz bt vset s v0 qmin qmax qload from to y11 y12 y21 y22 rating status opt begin
    pf = isempty(from) ? Ptr{Int32}(C_NULL) : pointer(from)
    pt = isempty(to) ? Ptr{Int32}(C_NULL) : pointer(to)
    ps = status === nothing ? Ptr{UInt8}(C_NULL) : pointer(status)
    if n_devices <= 1
      check(ccall((:sblx_create, libsblx), Cint,
        (Ref{Ptr{Cvoid}}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Int64, Ptr{UInt8}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
         Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
         Ptr{Float64}, Ptr{Float64}, Ptr{UInt8}, Ref{Options}),
        href, n, pointer(colptr), pointer(rowval), p(nz), slack, pointer(bt), pointer(vset), p(s), p(v0), p(qmin), p(qmax), p(qload),
        length(from), pf, pt, p(y11), p(y12), p(y21), p(y22), p(rating), ps, opt))
    else
      check(ccall((:sblx_create_multi, libsblx), Cint,
        (Ref{Ptr{Cvoid}}, Int32, Ptr{Int32}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Int64, Ptr{UInt8}, Ptr{Float64}, Ptr{Float64},
         Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Int64, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
         Ptr{Float64}, Ptr{Float64}, Ptr{UInt8}, Ref{Options}),
        href, Int32(n_devices), Ptr{Int32}(C_NULL), n, pointer(colptr), pointer(rowval), p(nz), slack, pointer(bt), pointer(vset), p(s), p(v0),
        p(qmin), p(qmax), p(qload), length(from), pf, pt, p(y11), p(y12), p(y21), p(y22), p(rating), ps, opt))
    end
  end
  return Handle(href[], n_devices > 1, n, length(from))
end

function set_locked_pv!(h::Handle, buses::Vector{Int32})
  GC.@preserve buses begin
    if h.multi
      check_multi(ccall((:sblx_multi_set_locked_pv, libsblx), Cint, (Ptr{Cvoid}, Int64, Ptr{Int32}), h.ptr, length(buses), buses), h.ptr)
    else
      check(ccall((:sblx_set_locked_pv, libsblx), Cint, (Ptr{Cvoid}, Int64, Ptr{Int32}), h.ptr, length(buses), buses), h.ptr)
    end
  end
end

"handle of device 0 (for sblx_solve_one)"
first_handle(h::Handle) = h.multi ? ccall((:sblx_multi_handle, libsblx), Ptr{Cvoid}, (Ptr{Cvoid}, Int32), h.ptr, Int32(0)) : h.ptr

function solve_one(h::Handle)
  Vout = zeros(ComplexF64, h.n); conv = Ref{UInt8}(0); it = Ref{Int32}(0); res = Ref{Float64}(0.0); st = Ref{Int32}(0)
  h0 = first_handle(h)
  GC.@preserve Vout check(ccall((:sblx_solve_one, libsblx), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ref{UInt8}, Ref{Int32}, Ref{Float64}, Ref{Int32}),
                                h0, C_NULL, Ptr{Float64}(pointer(Vout)), conv, it, res, st), h0)
  return Vout, conv[] != 0, Int(it[]), res[], Int(st[])
end

struct BatchResults
  converged::Vector{UInt8}; status::Vector{Int32}; iterations::Vector{Int32}; n_switch_events::Vector{Int32}
  island_count::Vector{Int32}; final_mismatch::Vector{Float64}; vmin_pu::Vector{Float64}; vmax_pu::Vector{Float64}
  max_loading_pct::Vector{Float64}; n_voltage_violations::Vector{Int32}; n_overloaded::Vector{Int32}; n_tiny_pivots::Vector{Int32}
  violations::Vector{Vector{Int}}      # per scenario: bus ids outside the voltage band (device list, contingency.jl:160-162)
  overloads::Vector{Vector{Int}}       # per scenario: branch ids with loading > 100 % (contingency.jl:175)
  V::Union{Nothing,Matrix{ComplexF64}}
end

"""
    solve_outages(h, cases; want_v = false) -> BatchResults

`cases[s]` = branch ids (1-based) out of service in scenario s.  The violation / overload lists are compacted ON THE
DEVICE; bus voltages cross PCIe only with `want_v = true`.  With a multi-GPU handle the cases are sharded over the GPUs
inside the call (one ncclAllGather of the per-case records).
"""
function solve_outages(h::Handle, cases::Vector{Vector{Int}}; want_v::Bool = false, list_capacity::Int = 0)
  B = length(cases)
  optr = Int32[0]; obr = Int32[]
  for c in cases
    append!(obr, Int32.(c)); push!(optr, Int32(length(obr)))
  end
  isempty(obr) && push!(obr, Int32(0))
  cap = list_capacity > 0 ? list_capacity : 16 * B + 65536
  z32() = zeros(Int32, B); zf() = zeros(Float64, B)
  conv = zeros(UInt8, B); st = z32(); it = z32(); nsw = z32(); isl = z32(); fm = zf(); vmin = zf(); vmax = zf(); ld = zf()
  nvv = z32(); nov = z32(); ntp = z32()
  vs = zeros(Int32, cap); vb = zeros(Int32, cap); vc = Ref{Int64}(0)
  os = zeros(Int32, cap); ob = zeros(Int32, cap); op = zeros(Float64, cap); oc = Ref{Int64}(0)
  V = want_v ? zeros(ComplexF64, h.n, B) : nothing
  GC.@preserve optr obr conv st it nsw isl fm vmin vmax ld nvv nov ntp vs vb os ob op V begin
    res = Results(pointer(conv), pointer(st), pointer(it), pointer(nsw), pointer(isl), pointer(fm), pointer(vmin), pointer(vmax), pointer(ld),
                  want_v ? Ptr{Float64}(pointer(V)) : Ptr{Float64}(C_NULL), Ptr{UInt8}(C_NULL), pointer(nvv), pointer(nov), pointer(ntp),
                  cap, pointer(vs), pointer(vb), Base.unsafe_convert(Ptr{Int64}, vc), cap, pointer(os), pointer(ob), pointer(op),
                  Base.unsafe_convert(Ptr{Int64}, oc), 0, Ptr{Int32}(C_NULL), Ptr{Int32}(C_NULL), Ptr{Int32}(C_NULL), Ptr{Int32}(C_NULL),
                  Ptr{Int64}(C_NULL), Ptr{Float64}(C_NULL))
    GC.@preserve vc oc begin
      if h.multi
        check_multi(ccall((:sblx_multi_solve_outages, libsblx), Cint, (Ptr{Cvoid}, Int64, Ptr{Int32}, Ptr{Int32}, Ref{Results}), h.ptr, B, optr, obr, res), h.ptr)
      else
        check(ccall((:sblx_solve_outages, libsblx), Cint, (Ptr{Cvoid}, Int64, Ptr{Int32}, Ptr{Int32}, Ref{Results}), h.ptr, B, optr, obr, res), h.ptr)
      end
    end
  end
  if vc[] > cap || oc[] > cap           # lists truncated: the counts are complete, repeat with room for everything
    return solve_outages(h, cases; want_v = want_v, list_capacity = Int(max(vc[], oc[])) + 16)
  end
  violations = [Int[] for _ in 1:B]; overloads = [Int[] for _ in 1:B]
  for k in 1:vc[]
    push!(violations[vs[k] + 1], Int(vb[k]))      # scenario positions are 0-based, element ids follow index_base = 1
  end
  for k in 1:oc[]
    push!(overloads[os[k] + 1], Int(ob[k]))
  end
  return BatchResults(conv, st, it, nsw, isl, fm, vmin, vmax, ld, nvv, nov, ntp, violations, overloads, V)
end

"""
    solve_injection_sweep(h, B; first = 0, seed = 20260724, sigma = 0.1, lo = 0.5, hi = 1.5) -> BatchResults

Monte-Carlo load sweep generated ON THE DEVICE (`sblx_solve_injection_sweep`): scenario s scales P and Q of every load bus
by an independent truncated-normal factor from Philox4x32-10 keyed on (seed, first + s, bus) - the recipe of
examples/powerflow/mc_probabilistic_powerflow.jl:28-41.  Single-GPU handles only; split a study over several handles /
processes with `first` (the generator is counter based: a draw does not depend on batching).
"""
function solve_injection_sweep(h::Handle, B::Int; first::Int = 0, seed::Integer = 20260724, sigma::Float64 = 0.1, lo::Float64 = 0.5,
                               hi::Float64 = 1.5)
  h.multi && error("solve_injection_sweep: use one single-GPU handle per device and split the study with `first`")
  z32() = zeros(Int32, B); zf() = zeros(Float64, B)
  conv = zeros(UInt8, B); st = z32(); it = z32(); nsw = z32(); isl = z32(); fm = zf(); vmin = zf(); vmax = zf(); ld = zf()
  nvv = z32(); nov = z32(); ntp = z32()
  GC.@preserve conv st it nsw isl fm vmin vmax ld nvv nov ntp begin
    res = Results(pointer(conv), pointer(st), pointer(it), pointer(nsw), pointer(isl), pointer(fm), pointer(vmin), pointer(vmax), pointer(ld),
                  Ptr{Float64}(C_NULL), Ptr{UInt8}(C_NULL), pointer(nvv), pointer(nov), pointer(ntp),
                  0, Ptr{Int32}(C_NULL), Ptr{Int32}(C_NULL), Ptr{Int64}(C_NULL), 0, Ptr{Int32}(C_NULL), Ptr{Int32}(C_NULL), Ptr{Float64}(C_NULL),
                  Ptr{Int64}(C_NULL), 0, Ptr{Int32}(C_NULL), Ptr{Int32}(C_NULL), Ptr{Int32}(C_NULL), Ptr{Int32}(C_NULL), Ptr{Int64}(C_NULL),
                  Ptr{Float64}(C_NULL))
    check(ccall((:sblx_solve_injection_sweep, libsblx), Cint, (Ptr{Cvoid}, Int64, Int64, UInt64, Float64, Float64, Float64, Ref{Results}),
                h.ptr, B, first, UInt64(seed), sigma, lo, hi, res), h.ptr)
  end
  return BatchResults(conv, st, it, nsw, isl, fm, vmin, vmax, ld, nvv, nov, ntp, [Int[] for _ in 1:B], [Int[] for _ in 1:B], nothing)
end

end # module LibSblx
