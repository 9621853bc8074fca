"""ctypes binding of libsblx.so (include/sblx.h).  The library is built in-tree
(`sparlectra.jl_b200/build.sh` -> `sparlectra.jl_b200/lib/libsblx.so`); there is
no fallback: if it is missing, importing callers get a loud error."""
from __future__ import annotations

import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SBLX_LIB") or os.path.join(os.path.dirname(_HERE), "lib", "libsblx.so")

c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_u8p = C.POINTER(C.c_uint8)
c_f64p = C.POINTER(C.c_double)


class Options(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("index_base", C.c_int32), ("max_iter", C.c_int32),
        ("qlimits_enabled", C.c_int32), ("qlimit_start_iter", C.c_int32), ("cooldown_iters", C.c_int32),
        ("guard_max_switches", C.c_int32), ("freeze_after_repeated_switching", C.c_int32),
        ("device", C.c_int32), ("max_slots", C.c_int32), ("ordering", C.c_int32), ("relax_small", C.c_int32),
        ("factor_mode", C.c_int32), ("reserved1", C.c_int32),
        ("tol", C.c_double), ("damp", C.c_double), ("q_hyst_pu", C.c_double), ("vm_min_pu", C.c_double),
        ("vm_max_pu", C.c_double), ("base_mva", C.c_double), ("mem_fraction", C.c_double),
        ("relax_zero_frac", C.c_double), ("panel_smem_kb", C.c_double), ("pivot_tol_rel", C.c_double),
    ]


class Results(C.Structure):
    _fields_ = [
        ("converged", c_u8p), ("status", c_i32p), ("iterations", c_i32p), ("n_switch_events", c_i32p),
        ("island_count", c_i32p), ("final_mismatch", c_f64p), ("vmin_pu", c_f64p), ("vmax_pu", c_f64p),
        ("max_loading_pct", c_f64p), ("V", c_f64p), ("bus_type_final", c_u8p),
        ("n_voltage_violations", c_i32p), ("n_overloaded", c_i32p), ("n_tiny_pivots", c_i32p),
        ("viol_capacity", C.c_int64), ("viol_scenario", c_i32p), ("viol_bus", c_i32p), ("viol_count", c_i64p),
        ("over_capacity", C.c_int64), ("over_scenario", c_i32p), ("over_branch", c_i32p), ("over_loading_pct", c_f64p),
        ("over_count", c_i64p),
        ("qlog_capacity", C.c_int64), ("qlog_scenario", c_i32p), ("qlog_iter", c_i32p), ("qlog_bus", c_i32p), ("qlog_side", c_i32p),
        ("qlog_count", c_i64p), ("flows", c_f64p),
    ]


class Info(C.Structure):
    _fields_ = [
        ("n", C.c_int64), ("m", C.c_int64), ("nnz_y", C.c_int64), ("nnz_j_blocks", C.c_int64),
        ("nnz_j_scalar", C.c_int64), ("nnz_lu_scalar", C.c_int64), ("n_panels", C.c_int64), ("n_levels", C.c_int64),
        ("u_blocks", C.c_int64), ("c_blocks", C.c_int64), ("max_slots", C.c_int64), ("bytes_per_slot", C.c_int64),
        ("block_ops", C.c_double), ("algorithmic_bytes_per_iteration", C.c_double), ("ordering", C.c_char * 16),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("total_ms", C.c_double), ("factor_ms", C.c_double), ("solve_ms", C.c_double), ("mismatch_ms", C.c_double),
        ("jacobian_ms", C.c_double), ("other_ms", C.c_double), ("h2d_ms", C.c_double), ("d2h_ms", C.c_double),
        ("host_prep_ms", C.c_double), ("launches", C.c_int64), ("factor_launches", C.c_int64), ("ticks", C.c_int64),
        ("scenario_iterations", C.c_int64), ("mismatch_evaluations", C.c_int64), ("scenarios_solved", C.c_int64),
        ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
        ("tiny_pivots", C.c_int64), ("n_ranks", C.c_int64), ("shard_lo", C.c_int64), ("shard_hi", C.c_int64),
        ("allgather_ms", C.c_double), ("refine_passes", C.c_int64),
    ]


EXPORTS = [
    "sblx_version", "sblx_default_options", "sblx_last_error", "sblx_create", "sblx_destroy", "sblx_set_locked_pv",
    "sblx_set_v0", "sblx_get_info", "sblx_get_stats", "sblx_solve_outages", "sblx_solve_injections", "sblx_solve_one",
    "sblx_stage_outages", "sblx_run_staged", "sblx_fetch_results", "sblx_device_summary", "sblx_debug_pattern",
    "sblx_debug_newton_step", "sblx_debug_host_selftest", "sblx_analyze_only", "sblx_debug_phase_cycles",
    "sblx_debug_topology", "sblx_solve_injection_sweep", "sblx_solve_injection_scale", "sblx_sweep_factors",
    "sblx_comm_unique_id", "sblx_comm_init_rank", "sblx_comm_destroy", "sblx_solve_outages_sharded", "sblx_shard_bounds",
    "sblx_create_multi", "sblx_destroy_multi", "sblx_multi_last_error", "sblx_multi_set_v0", "sblx_multi_set_locked_pv",
    "sblx_multi_solve_outages", "sblx_multi_get_stats", "sblx_multi_handle", "sblx_debug_set_pf_mode",
    "sblx_mpc_read", "sblx_mpc_parse", "sblx_mpc_dims", "sblx_mpc_tables", "sblx_mpc_free", "sblx_mpc_last_error",
]

_lib = None


def load():
    """Load libsblx.so (once). Raises if the in-tree build is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} not found - run sparlectra.jl_b200/build.sh (or __graft_entry__.build()); "
                           "the CUDA engine has no fallback")
    lib = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    lib.sblx_version.restype = C.c_int
    lib.sblx_default_options.argtypes = [C.POINTER(Options)]
    lib.sblx_default_options.restype = None
    lib.sblx_last_error.argtypes = [vp]
    lib.sblx_last_error.restype = C.c_char_p
    lib.sblx_create.argtypes = [C.POINTER(vp), C.c_int64, c_i64p, c_i64p, c_f64p, C.c_int64, c_u8p, c_f64p, c_f64p,
                                c_f64p, c_f64p, c_f64p, c_f64p, C.c_int64, c_i32p, c_i32p, c_f64p, c_f64p, c_f64p,
                                c_f64p, c_f64p, c_u8p, C.POINTER(Options)]
    lib.sblx_destroy.argtypes = [vp]
    lib.sblx_destroy.restype = None
    lib.sblx_set_locked_pv.argtypes = [vp, C.c_int64, c_i32p]
    lib.sblx_set_v0.argtypes = [vp, c_f64p]
    lib.sblx_get_info.argtypes = [vp, C.POINTER(Info)]
    lib.sblx_get_stats.argtypes = [vp, C.POINTER(Stats)]
    lib.sblx_solve_outages.argtypes = [vp, C.c_int64, c_i32p, c_i32p, C.POINTER(Results)]
    lib.sblx_solve_injections.argtypes = [vp, C.c_int64, c_f64p, c_f64p, C.POINTER(Results)]
    lib.sblx_solve_one.argtypes = [vp, c_f64p, c_f64p, c_u8p, c_i32p, c_f64p, c_i32p]
    lib.sblx_stage_outages.argtypes = [vp, C.c_int64, c_i32p, c_i32p, C.c_int, C.c_int]
    lib.sblx_run_staged.argtypes = [vp]
    lib.sblx_fetch_results.argtypes = [vp, C.POINTER(Results)]
    lib.sblx_device_summary.argtypes = [vp, C.POINTER(vp), c_i64p]
    lib.sblx_debug_pattern.argtypes = [vp, c_i64p, c_i32p]
    lib.sblx_debug_newton_step.argtypes = [vp, C.c_int64, c_i32p, c_u8p, c_f64p, c_f64p, c_f64p, c_f64p, c_f64p]
    lib.sblx_debug_host_selftest.argtypes = [vp, C.c_uint64, c_f64p]
    lib.sblx_analyze_only.argtypes = [C.c_int64, c_i64p, c_i64p, C.c_int64, C.POINTER(Options), C.POINTER(Info), c_f64p]
    lib.sblx_debug_phase_cycles.argtypes = [c_f64p]
    lib.sblx_debug_topology.argtypes = [C.c_int64, C.c_int64, c_u8p, C.c_int64, c_i32p, c_i32p, c_u8p, C.c_int32, C.c_int64, c_i32p,
                                        c_i32p, c_i32p, c_i32p, c_i32p]
    lib.sblx_solve_injection_sweep.argtypes = [vp, C.c_int64, C.c_int64, C.c_uint64, C.c_double, C.c_double, C.c_double, C.POINTER(Results)]
    lib.sblx_solve_injection_scale.argtypes = [vp, C.c_int64, c_f64p, C.POINTER(Results)]
    lib.sblx_sweep_factors.argtypes = [vp, C.c_int64, C.c_int64, C.c_uint64, C.c_double, C.c_double, C.c_double, c_f64p]
    lib.sblx_comm_unique_id.argtypes = [c_u8p]
    lib.sblx_comm_init_rank.argtypes = [vp, C.c_int32, C.c_int32, c_u8p]
    lib.sblx_comm_destroy.argtypes = [vp]
    lib.sblx_solve_outages_sharded.argtypes = [vp, C.c_int64, c_i32p, c_i32p, C.POINTER(Results)]
    lib.sblx_shard_bounds.argtypes = [vp, C.c_int64, c_i32p, c_i32p, C.c_int32, c_i64p]
    lib.sblx_create_multi.argtypes = [C.POINTER(vp), C.c_int32, c_i32p] + lib.sblx_create.argtypes[1:]
    lib.sblx_destroy_multi.argtypes = [vp]
    lib.sblx_destroy_multi.restype = None
    lib.sblx_multi_last_error.argtypes = [vp]
    lib.sblx_multi_last_error.restype = C.c_char_p
    lib.sblx_multi_set_v0.argtypes = [vp, c_f64p]
    lib.sblx_multi_set_locked_pv.argtypes = [vp, C.c_int64, c_i32p]
    lib.sblx_multi_solve_outages.argtypes = [vp, C.c_int64, c_i32p, c_i32p, C.POINTER(Results)]
    lib.sblx_multi_get_stats.argtypes = [vp, C.c_int32, C.POINTER(Stats)]
    lib.sblx_multi_handle.argtypes = [vp, C.c_int32]
    lib.sblx_multi_handle.restype = vp
    lib.sblx_debug_set_pf_mode.argtypes = [vp, C.c_int32]
    lib.sblx_mpc_read.argtypes = [C.c_char_p, C.POINTER(vp)]
    lib.sblx_mpc_parse.argtypes = [C.c_char_p, C.POINTER(vp)]
    lib.sblx_mpc_dims.argtypes = [vp, c_f64p, c_i64p, c_i64p, c_i64p]
    lib.sblx_mpc_tables.argtypes = [vp, c_f64p, c_f64p, c_f64p]
    lib.sblx_mpc_free.argtypes = [vp]
    lib.sblx_mpc_free.restype = None
    lib.sblx_mpc_last_error.argtypes = []
    lib.sblx_mpc_last_error.restype = C.c_char_p
    for name in EXPORTS:
        if getattr(lib, name).restype is C.c_int and name not in ("sblx_version",):
            getattr(lib, name).restype = C.c_int
    _lib = lib
    return lib


def ptr(a, typ):
    if a is None:
        return None
    return a.ctypes.data_as(typ)


def default_options(**kw) -> Options:
    o = Options()
    load().sblx_default_options(C.byref(o))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise TypeError(f"unknown sblx option {k}")
        setattr(o, k, v)
    return o


class SblxError(RuntimeError):
    pass


def check(rc, handle=None):
    if rc != 0:
        msg = load().sblx_last_error(handle)
        raise SblxError(f"sblx error {rc}: {msg.decode() if msg else ''}")


def as_dict(struct) -> dict:
    out = {}
    for name, _ in struct._fields_:
        v = getattr(struct, name)
        out[name] = v.decode() if isinstance(v, bytes) else v
    return out


def c128_to_f64(a) -> np.ndarray:
    """contiguous complex128 -> interleaved float64 view"""
    a = np.ascontiguousarray(a, dtype=np.complex128)
    return a.view(np.float64)
