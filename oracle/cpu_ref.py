"""ctypes loader of the C++ CPU baseline (oracle/cpu_ref/cpu_ref.cpp, built by oracle/build_cpu_ref.sh).

TEST / MEASUREMENT INFRASTRUCTURE: used by tests/test_cpu_ref.py (checked against the Python oracle), by bench.py's
`cpu_baseline` leg and by `bench.py --impl reference`.  Never imported by the product package."""
from __future__ import annotations

import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libcpuref.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} not found - run oracle/build_cpu_ref.sh (or __graft_entry__.build())")
        _lib = C.CDLL(LIB_PATH)
        _lib.cpuref_run.restype = C.c_int
    return _lib


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def run(a, cases, variant=0, nthreads=1, max_iter=30, tol=1e-8, qlimits_enabled=True, qlimit_start_iter=2, want_v=True):
    """`a`: model arrays with the fields of sblx.model.ModelArrays (Y csc, slack, bus_type, vset, s_spec, v0, qmin, qmax,
    qload, br_from, br_to, y11..y22, br_rating, br_status, baseMVA); `cases`: list of branch-id lists.
    variant 0 = A (symbolic analysis every iteration, the reference default), 1 = B (symbolic once per thread)."""
    lib = load()
    Y = a.Y.tocsc()
    Y.sort_indices()
    n, B = a.n, len(cases)
    f64 = lambda x: np.ascontiguousarray(np.asarray(x, np.complex128)).view(np.float64)
    keep = dict(cp=np.ascontiguousarray(Y.indptr, np.int64), rv=np.ascontiguousarray(Y.indices, np.int64), nz=f64(Y.data),
                bt=np.ascontiguousarray(a.bus_type, np.uint8), vset=np.ascontiguousarray(a.vset, float), s=f64(a.s_spec), v0=f64(a.v0),
                qmin=np.ascontiguousarray(a.qmin, float), qmax=np.ascontiguousarray(a.qmax, float), ql=np.ascontiguousarray(a.qload, float),
                bf=np.ascontiguousarray(a.br_from, np.int32), bt2=np.ascontiguousarray(a.br_to, np.int32), y11=f64(a.y11), y12=f64(a.y12),
                y21=f64(a.y21), y22=f64(a.y22), rt=np.ascontiguousarray(a.br_rating, float), st=np.ascontiguousarray(a.br_status, np.uint8))
    optr = np.zeros(B + 1, np.int32)
    optr[1:] = np.cumsum([len(c) for c in cases])
    obr = np.ascontiguousarray([k for c in cases for k in c] or [0], np.int32)
    out = dict(converged=np.zeros(B, np.uint8), iterations=np.zeros(B, np.int32), status=np.zeros(B, np.int32),
               n_switch_events=np.zeros(B, np.int32), vmin_pu=np.zeros(B), vmax_pu=np.zeros(B), max_loading_pct=np.zeros(B))
    if want_v:
        out["V"] = np.zeros((B, n), np.complex128)
        out["bus_type_final"] = np.zeros((B, n), np.uint8)
    sec = np.zeros(2)
    k = keep
    rc = lib.cpuref_run(C.c_int64(n), _p(k["cp"], C.c_int64), _p(k["rv"], C.c_int64), _p(k["nz"], C.c_double), C.c_int64(a.slack),
                        _p(k["bt"], C.c_uint8), _p(k["vset"], C.c_double), _p(k["s"], C.c_double), _p(k["v0"], C.c_double),
                        _p(k["qmin"], C.c_double), _p(k["qmax"], C.c_double), _p(k["ql"], C.c_double), C.c_int64(len(k["bf"])),
                        _p(k["bf"], C.c_int32), _p(k["bt2"], C.c_int32), _p(k["y11"], C.c_double), _p(k["y12"], C.c_double),
                        _p(k["y21"], C.c_double), _p(k["y22"], C.c_double), _p(k["rt"], C.c_double), _p(k["st"], C.c_uint8),
                        C.c_double(a.baseMVA), C.c_int64(B), _p(optr, C.c_int32), _p(obr, C.c_int32), C.c_int32(variant),
                        C.c_int32(nthreads), C.c_int32(max_iter), C.c_double(tol), C.c_int32(int(qlimits_enabled)),
                        C.c_int32(qlimit_start_iter), _p(out["converged"], C.c_uint8), _p(out["iterations"], C.c_int32),
                        _p(out["status"], C.c_int32), _p(out["n_switch_events"], C.c_int32),
                        _p(out["V"].view(np.float64), C.c_double) if want_v else None,
                        _p(out["bus_type_final"], C.c_uint8) if want_v else None, _p(out["vmin_pu"], C.c_double),
                        _p(out["vmax_pu"], C.c_double), _p(out["max_loading_pct"], C.c_double), _p(sec, C.c_double))
    if rc != 0:
        raise RuntimeError(f"cpuref_run failed with {rc}")
    out["seconds"], out["nnz_lu"] = float(sec[0]), int(sec[1])
    return out
