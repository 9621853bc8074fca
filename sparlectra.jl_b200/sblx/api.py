"""Host-side mirror of the reference's plugin / contingency interface, driving the
CUDA engine through the C ABI (include/sblx.h).

Names, argument meaning and error behaviour follow the reference
(`src/solver_interface.jl`, `src/contingency.jl`): `PFModel`, `PFSolution`,
`AbstractExternalSolver`/`solvePf`, `buildPfModel`, `mismatchInf`,
`runpf_external`, `ContingencyCase`, `ContingencyResult`, `generateN1Branches`,
`runContingenciesBatched` (= `runContingencies!` with the batched backend).
Failures of a scenario are REPORTED in the result, never thrown
(src/contingency.jl:72-73); argument errors do throw (src/contingency.jl:270).
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import Optional
import numpy as np
import scipy.sparse as sp

from . import lib as L
from . import model as M

STATUS_TEXT = {
    0: None,
    1: "power flow did not converge (status 1)",
    2: "power flow did not converge (status 1)",
    3: "power flow did not converge (status 1)",
    4: "power flow did not converge (status 1)",
    5: "power flow did not converge (status 1)",
    6: "power flow did not converge (status 1)",
    7: "islanded without reference",
    8: "unknown branch",
    9: "islanded (island-wise solve not supported by the batched backend)",
}
REASON = {0: "none", 1: "nr_mismatch_not_converged", 2: "nr_nonfinite", 3: "singular_newton_step",
          4: "remaining_pv_q_limit_violations", 5: "active_pv_voltage_residual", 6: "max_switching_exceeded",
          7: "islanded_without_reference", 8: "unknown_branch", 9: "islanded_unsupported"}


# ---------------------------------------------------------------------------
# PFModel / PFSolution (src/solver_interface.jl:51-82)
# ---------------------------------------------------------------------------
@dataclass
class PFModel:
    Ybus: sp.csc_matrix
    baseMVA: float
    busIdx_net: np.ndarray
    busType: np.ndarray          # uint8 codes (model.BUS_*)
    slack_idx: int               # 0-based here
    Vset: np.ndarray
    Sspec: np.ndarray
    V0: np.ndarray
    qmin_pu: np.ndarray = field(default_factory=lambda: np.zeros(0))
    qmax_pu: np.ndarray = field(default_factory=lambda: np.zeros(0))
    arrays: Optional[M.ModelArrays] = None


@dataclass
class PFSolution:
    V: np.ndarray
    converged: bool
    iters: int
    residual_inf: float
    meta: object = None


class AbstractExternalSolver:
    pass


def buildPfModel(grid, flatstart=None, include_limits=True) -> PFModel:
    """buildPfModel (src/solver_interface.jl:193-295) on grid tables."""
    a = M.build_arrays(grid, flatstart=flatstart)
    return PFModel(Ybus=a.Y, baseMVA=a.baseMVA, busIdx_net=np.arange(a.n), busType=a.bus_type, slack_idx=a.slack,
                   Vset=a.vset, Sspec=a.s_spec, V0=a.v0,
                   qmin_pu=a.qmin if include_limits else np.zeros(0), qmax_pu=a.qmax if include_limits else np.zeros(0),
                   arrays=a)


def mismatchInf(model: PFModel, V) -> float:
    """mismatchInf (src/solver_interface.jl:310-313): inf-norm of the rectangular PQ/PV residual."""
    Sc = V * np.conj(model.Ybus @ V)
    ns = np.arange(len(V)) != model.slack_idx
    dp = np.abs(Sc.real - model.Sspec.real)[ns]
    second = np.where(model.busType == M.BUS_PV, np.abs(np.abs(V) - model.Vset), np.abs(Sc.imag - model.Sspec.imag))[ns]
    return float(max(dp.max(initial=0.0), second.max(initial=0.0)))


# ---------------------------------------------------------------------------
# Engine: one handle = one grid on one GPU
# ---------------------------------------------------------------------------
class Engine:
    def __init__(self, arrays: M.ModelArrays, index_base: int = 0, **options):
        """`index_base` = what the C ABI is driven with (0, or 1 like the Julia glue: CSC colptr / rowval, slack, branch
        terminals, outage and lock ids are shifted accordingly); the methods of this class always take 0-based ids."""
        self.lib = L.load()
        self.a = arrays
        self.n = arrays.n
        self.base = int(index_base)
        self.opt = L.default_options(index_base=self.base, base_mva=arrays.baseMVA, **options)
        k = self._pack_model(arrays, self.base)
        self._keep = k
        h = C.c_void_p()
        rc = self.lib.sblx_create(C.byref(h), self.n, *self._model_args(k, arrays, self.base), C.byref(self.opt))
        L.check(rc, None)
        self.h = h

    @staticmethod
    def _pack_model(arrays, base):
        Y = arrays.Y.tocsc()
        Y.sort_indices()
        return dict(
            colptr=np.ascontiguousarray(Y.indptr + base, np.int64), rowval=np.ascontiguousarray(Y.indices + base, np.int64),
            nz=L.c128_to_f64(Y.data), bt=np.ascontiguousarray(arrays.bus_type, np.uint8),
            vset=np.ascontiguousarray(arrays.vset, float), s=L.c128_to_f64(arrays.s_spec), v0=L.c128_to_f64(arrays.v0),
            qmin=np.ascontiguousarray(arrays.qmin, float), qmax=np.ascontiguousarray(arrays.qmax, float),
            ql=np.ascontiguousarray(arrays.qload, float), bf=np.ascontiguousarray(np.asarray(arrays.br_from) + base, np.int32),
            bt2=np.ascontiguousarray(np.asarray(arrays.br_to) + base, np.int32), y11=L.c128_to_f64(arrays.y11), y12=L.c128_to_f64(arrays.y12),
            y21=L.c128_to_f64(arrays.y21), y22=L.c128_to_f64(arrays.y22),
            rt=np.ascontiguousarray(arrays.br_rating, float), st=np.ascontiguousarray(arrays.br_status, np.uint8))

    @staticmethod
    def _model_args(k, arrays, base):
        """argument tail shared by sblx_create and sblx_create_multi (from colptr to br_status)"""
        return (L.ptr(k["colptr"], L.c_i64p), L.ptr(k["rowval"], L.c_i64p), L.ptr(k["nz"], L.c_f64p), arrays.slack + base,
                L.ptr(k["bt"], L.c_u8p), L.ptr(k["vset"], L.c_f64p), L.ptr(k["s"], L.c_f64p), L.ptr(k["v0"], L.c_f64p),
                L.ptr(k["qmin"], L.c_f64p), L.ptr(k["qmax"], L.c_f64p), L.ptr(k["ql"], L.c_f64p), len(arrays.br_from),
                L.ptr(k["bf"], L.c_i32p), L.ptr(k["bt2"], L.c_i32p), L.ptr(k["y11"], L.c_f64p), L.ptr(k["y12"], L.c_f64p),
                L.ptr(k["y21"], L.c_f64p), L.ptr(k["y22"], L.c_f64p), L.ptr(k["rt"], L.c_f64p), L.ptr(k["st"], L.c_u8p))

    def close(self):
        if getattr(self, "h", None):
            self.lib.sblx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- info -------------------------------------------------------------
    def info(self) -> dict:
        i = L.Info()
        L.check(self.lib.sblx_get_info(self.h, C.byref(i)), self.h)
        return L.as_dict(i)

    def stats(self) -> dict:
        s = L.Stats()
        L.check(self.lib.sblx_get_stats(self.h, C.byref(s)), self.h)
        return L.as_dict(s)

    def set_v0(self, v0):
        v = L.c128_to_f64(v0)
        L.check(self.lib.sblx_set_v0(self.h, L.ptr(v, L.c_f64p)), self.h)

    def set_locked_pv(self, buses):
        b = np.ascontiguousarray(np.asarray(buses, np.int64) + self.base, np.int32)
        L.check(self.lib.sblx_set_locked_pv(self.h, len(b), L.ptr(b, L.c_i32p)), self.h)

    # -- solves -------------------------------------------------------------
    @staticmethod
    def _alloc(B, n, want_v, want_types, want_lists=False, want_flows=False, nbranch=0, list_capacity=0):
        r = dict(converged=np.zeros(B, np.uint8), status=np.zeros(B, np.int32), iterations=np.zeros(B, np.int32),
                 n_switch_events=np.zeros(B, np.int32), island_count=np.zeros(B, np.int32),
                 final_mismatch=np.zeros(B), vmin_pu=np.zeros(B), vmax_pu=np.zeros(B), max_loading_pct=np.zeros(B),
                 n_voltage_violations=np.zeros(B, np.int32), n_overloaded=np.zeros(B, np.int32),
                 n_tiny_pivots=np.zeros(B, np.int32))
        if want_v:
            r["V"] = np.zeros((B, n), np.complex128)
        if want_types:
            r["bus_type_final"] = np.zeros((B, n), np.uint8)
        if want_lists:
            cap = int(list_capacity) if list_capacity else 16 * B + 65536
            r["_viol_scenario"] = np.zeros(cap, np.int32)
            r["_viol_bus"] = np.zeros(cap, np.int32)
            r["_viol_count"] = np.zeros(1, np.int64)
            r["_over_scenario"] = np.zeros(cap, np.int32)
            r["_over_branch"] = np.zeros(cap, np.int32)
            r["_over_pct"] = np.zeros(cap)
            r["_over_count"] = np.zeros(1, np.int64)
            r["_qlog"] = [np.zeros(cap, np.int32) for _ in range(4)]
            r["_qlog_count"] = np.zeros(1, np.int64)
        if want_flows:
            r["flows"] = np.zeros((B, nbranch, 2), np.complex128)
        return r

    @staticmethod
    def _results_struct(r):
        s = L.Results()
        s.converged = L.ptr(r["converged"], L.c_u8p)
        s.status = L.ptr(r["status"], L.c_i32p)
        s.iterations = L.ptr(r["iterations"], L.c_i32p)
        s.n_switch_events = L.ptr(r["n_switch_events"], L.c_i32p)
        s.island_count = L.ptr(r["island_count"], L.c_i32p)
        s.final_mismatch = L.ptr(r["final_mismatch"], L.c_f64p)
        s.vmin_pu = L.ptr(r["vmin_pu"], L.c_f64p)
        s.vmax_pu = L.ptr(r["vmax_pu"], L.c_f64p)
        s.max_loading_pct = L.ptr(r["max_loading_pct"], L.c_f64p)
        s.V = L.ptr(r["V"].view(np.float64), L.c_f64p) if "V" in r else None
        s.bus_type_final = L.ptr(r["bus_type_final"], L.c_u8p) if "bus_type_final" in r else None
        s.n_voltage_violations = L.ptr(r["n_voltage_violations"], L.c_i32p)
        s.n_overloaded = L.ptr(r["n_overloaded"], L.c_i32p)
        s.n_tiny_pivots = L.ptr(r["n_tiny_pivots"], L.c_i32p)
        if "_viol_scenario" in r:
            s.viol_capacity = len(r["_viol_scenario"])
            s.viol_scenario = L.ptr(r["_viol_scenario"], L.c_i32p)
            s.viol_bus = L.ptr(r["_viol_bus"], L.c_i32p)
            s.viol_count = L.ptr(r["_viol_count"], L.c_i64p)
            s.over_capacity = len(r["_over_scenario"])
            s.over_scenario = L.ptr(r["_over_scenario"], L.c_i32p)
            s.over_branch = L.ptr(r["_over_branch"], L.c_i32p)
            s.over_loading_pct = L.ptr(r["_over_pct"], L.c_f64p)
            s.over_count = L.ptr(r["_over_count"], L.c_i64p)
            s.qlog_capacity = len(r["_qlog"][0])
            s.qlog_scenario, s.qlog_iter, s.qlog_bus, s.qlog_side = (L.ptr(x, L.c_i32p) for x in r["_qlog"])
            s.qlog_count = L.ptr(r["_qlog_count"], L.c_i64p)
        s.flows = L.ptr(r["flows"].view(np.float64), L.c_f64p) if "flows" in r else None
        return s

    def _finish_lists(self, r):
        """trim the compacted (scenario, element) lists to their counts and shift ids back to 0-based"""
        if "_viol_scenario" not in r:
            return r
        nv, no, nq = int(r["_viol_count"][0]), int(r["_over_count"][0]), int(r["_qlog_count"][0])
        r["lists_complete"] = nv <= len(r["_viol_scenario"]) and no <= len(r["_over_scenario"]) and nq <= len(r["_qlog"][0])
        q = r.pop("_qlog")
        nq = min(nq, len(q[0]))
        r["qlimit_events"] = np.stack([q[0][:nq], q[1][:nq], q[2][:nq] - self.base, q[3][:nq]], axis=1)   # scenario, iteration, bus, side
        r["qlimit_event_total"] = int(r.pop("_qlog_count")[0])
        nv, no = min(nv, len(r["_viol_scenario"])), min(no, len(r["_over_scenario"]))
        r["violations"] = np.stack([r.pop("_viol_scenario")[:nv], r.pop("_viol_bus")[:nv] - self.base], axis=1)
        r["overloads"] = np.stack([r.pop("_over_scenario")[:no], r.pop("_over_branch")[:no] - self.base], axis=1)
        r["overload_pct"] = r.pop("_over_pct")[:no]
        r["violation_total"], r["overload_total"] = int(r.pop("_viol_count")[0]), int(r.pop("_over_count")[0])
        return r

    @staticmethod
    def pack_outages(cases):
        """list of branch-id lists -> (ptr int32[B+1], br int32[])"""
        ptr = np.zeros(len(cases) + 1, np.int32)
        for i, c in enumerate(cases):
            ptr[i + 1] = ptr[i] + len(c)
        br = np.array([b for c in cases for b in c], np.int32) if ptr[-1] > 0 else np.zeros(1, np.int32)
        return ptr, br

    def solve_outages(self, cases, want_v=True, want_types=True, want_lists=False, want_flows=False, list_capacity=0,
                      sharded=False) -> dict:
        """`want_lists`: compacted (scenario, bus) voltage-violation and (scenario, branch) overload lists computed on the
        device (r["violations"], r["overloads"], r["overload_pct"]); `want_flows`: per-branch complex flows [B, nbranch, 2];
        `sharded`: collective call over the communicator set with `comm_init` (every rank passes the same batch)."""
        optr, obr = self.pack_outages(cases) if not isinstance(cases, tuple) else cases
        B = len(optr) - 1
        if self.base:
            obr = np.ascontiguousarray(obr + self.base, np.int32)
        r = self._alloc(B, self.n, want_v, want_types, want_lists, want_flows, len(self.a.br_from), list_capacity)
        s = self._results_struct(r)
        fn = self.lib.sblx_solve_outages_sharded if sharded else self.lib.sblx_solve_outages
        L.check(fn(self.h, B, L.ptr(optr, L.c_i32p), L.ptr(obr, L.c_i32p), C.byref(s)), self.h)
        return self._finish_lists(r)

    # -- multi-GPU (one process per GPU): NCCL communicator inside the library ----------------
    def comm_unique_id(self) -> bytes:
        buf = np.zeros(128, np.uint8)
        L.check(self.lib.sblx_comm_unique_id(L.ptr(buf, L.c_u8p)), None)
        return buf.tobytes()

    def comm_init(self, nranks: int, rank: int, uid: bytes):
        buf = np.frombuffer(uid, np.uint8).copy()
        L.check(self.lib.sblx_comm_init_rank(self.h, nranks, rank, L.ptr(buf, L.c_u8p)), self.h)

    def shard_bounds(self, B: int, nranks: int):
        b = np.zeros(nranks + 1, np.int64)
        L.check(self.lib.sblx_shard_bounds(self.h, B, None, None, nranks, L.ptr(b, L.c_i64p)), self.h)
        return b

    def stage_outages(self, optr, obr, want_v=False, want_types=False):
        if self.base:
            obr = np.ascontiguousarray(obr + self.base, np.int32)
        L.check(self.lib.sblx_stage_outages(self.h, len(optr) - 1, L.ptr(optr, L.c_i32p), L.ptr(obr, L.c_i32p),
                                            int(want_v), int(want_types)), self.h)
        self._staged = (len(optr) - 1, want_v, want_types)

    def run_staged(self):
        L.check(self.lib.sblx_run_staged(self.h), self.h)

    def fetch(self) -> dict:
        B, wv, wt = self._staged
        r = self._alloc(B, self.n, wv, wt)
        s = self._results_struct(r)
        L.check(self.lib.sblx_fetch_results(self.h, C.byref(s)), self.h)
        return r

    def device_summary(self):
        p = C.c_void_p()
        nb = C.c_int64()
        L.check(self.lib.sblx_device_summary(self.h, C.byref(p), C.byref(nb)), self.h)
        return p.value, nb.value

    def solve_injections(self, S, qload=None, want_v=True, want_types=True) -> dict:
        S = np.ascontiguousarray(S, np.complex128)
        B = S.shape[0]
        r = self._alloc(B, self.n, want_v, want_types)
        s = self._results_struct(r)
        ql = None if qload is None else np.ascontiguousarray(qload, float)
        L.check(self.lib.sblx_solve_injections(self.h, B, L.ptr(S.view(np.float64), L.c_f64p), L.ptr(ql, L.c_f64p), C.byref(s)), self.h)
        return r

    def sweep(self, n_scen, first=0, seed=20260724, sigma=0.1, lo=0.5, hi=1.5, want_v=False, want_types=False, want_lists=False):
        """Device-side injection sweep (BASELINE config 4): per scenario an independent truncated-normal factor
        f = clip(1 + sigma z, lo, hi) on P and Q of every load bus, drawn ON THE GPU from Philox4x32-10 keyed on
        (seed, first + scenario, bus) - the reference's Monte-Carlo recipe (examples/powerflow/mc_probabilistic_powerflow.jl:28-41);
        only the parameters cross PCIe."""
        r = self._alloc(n_scen, self.n, want_v, want_types, want_lists)
        s = self._results_struct(r)
        L.check(self.lib.sblx_solve_injection_sweep(self.h, n_scen, first, seed, sigma, lo, hi, C.byref(s)), self.h)
        return self._finish_lists(r)

    def sweep_factors(self, n_scen, first=0, seed=20260724, sigma=0.1, lo=0.5, hi=1.5):
        """the factors `sweep` applies, [n_scen, n] (1.0 on non-load buses), evaluated by the same device code"""
        out = np.zeros((n_scen, self.n))
        L.check(self.lib.sblx_sweep_factors(self.h, n_scen, first, seed, sigma, lo, hi, L.ptr(out, L.c_f64p)), self.h)
        return out

    def scale_sweep(self, scale, want_v=False, want_types=False):
        """one load-scaling factor per scenario (uniform scaling of all load buses)"""
        sc = np.ascontiguousarray(scale, float)
        r = self._alloc(len(sc), self.n, want_v, want_types)
        s = self._results_struct(r)
        L.check(self.lib.sblx_solve_injection_scale(self.h, len(sc), L.ptr(sc, L.c_f64p), C.byref(s)), self.h)
        return r

    def injection_sweep(self, s_spec0, qload0, n_scen, seed=20260724, sigma=0.1, lo=0.5, hi=1.5, chunk=4096):
        """Load / injection sweep of BASELINE config 4: per scenario an independent truncated-normal factor
        f ~ N(1, sigma) clipped to [lo, hi] on P and Q of every load bus (the reference's Monte-Carlo recipe,
        examples/powerflow/mc_probabilistic_powerflow.jl:28-41), solved in chunks so that the per-scenario
        injection matrices never exceed chunk x n.  Returns the per-scenario summary arrays (no voltages)."""
        rng = np.random.Generator(np.random.PCG64(seed))
        s0 = np.asarray(s_spec0, np.complex128)
        q0 = np.asarray(qload0, float)
        isload = s0.real < 0
        keys = ("converged", "status", "iterations", "n_switch_events", "final_mismatch", "vmin_pu", "vmax_pu", "max_loading_pct")
        out = {k: [] for k in keys}
        def draw(nb):
            f = np.clip(rng.normal(1.0, sigma, size=(nb, self.n)), lo, hi)
            return np.where(isload, s0 * f, s0), np.where(isload, q0 * f, q0)

        from concurrent.futures import ThreadPoolExecutor
        sizes = [min(chunk, n_scen - b0) for b0 in range(0, n_scen, chunk)]
        with ThreadPoolExecutor(1) as pool:            # the next chunk is drawn while the GPU solves this one
            nxt = pool.submit(draw, sizes[0])
            for i, nb in enumerate(sizes):
                S, ql = nxt.result()
                if i + 1 < len(sizes):
                    nxt = pool.submit(draw, sizes[i + 1])
                r = self.solve_injections(S, ql, want_v=False, want_types=False)
                for k in keys:
                    out[k].append(r[k])
        return {k: np.concatenate(v) for k, v in out.items()}

    def solve_one(self, v_start=None):
        vout = np.zeros(self.n, np.complex128)
        c = C.c_uint8()
        it = C.c_int32()
        res = C.c_double()
        st = C.c_int32()
        vs = None if v_start is None else L.c128_to_f64(v_start)
        L.check(self.lib.sblx_solve_one(self.h, L.ptr(vs, L.c_f64p), L.ptr(vout.view(np.float64), L.c_f64p), C.byref(c),
                                        C.byref(it), C.byref(res), C.byref(st)), self.h)
        return vout, bool(c.value), int(it.value), float(res.value), int(st.value)

    # -- parity hooks -------------------------------------------------------
    def pattern(self):
        i = self.info()
        rp = np.zeros(i["m"] + 1, np.int64)
        col = np.zeros(i["nnz_j_blocks"], np.int32)
        L.check(self.lib.sblx_debug_pattern(self.h, L.ptr(rp, L.c_i64p), L.ptr(col, L.c_i32p)), self.h)
        return rp, col

    def newton_step(self, V, outages=(), bus_type=None):
        i = self.info()
        m, nb = i["m"], i["nnz_j_blocks"]
        F = np.zeros(2 * m)
        J = np.zeros(4 * nb)
        dx = np.zeros(2 * m)
        mm = np.zeros(1)
        ob = np.ascontiguousarray(list(outages) or [0], np.int32)
        bt = None if bus_type is None else np.ascontiguousarray(bus_type, np.uint8)
        v = L.c128_to_f64(V)
        L.check(self.lib.sblx_debug_newton_step(self.h, len(outages), L.ptr(ob, L.c_i32p), L.ptr(bt, L.c_u8p),
                                                L.ptr(v, L.c_f64p), L.ptr(F, L.c_f64p), L.ptr(J, L.c_f64p),
                                                L.ptr(dx, L.c_f64p), L.ptr(mm, L.c_f64p)), self.h)
        return F, J.reshape(nb, 2, 2), dx, float(mm[0])

    def host_selftest(self, seed=1) -> float:
        r = C.c_double()
        L.check(self.lib.sblx_debug_host_selftest(self.h, seed, C.byref(r)), self.h)
        return r.value


def topology_only(a, cases):
    """Host-only topology verdicts of sblx_solve_outages for the outage sets `cases` (no CUDA):
    -> (island_count, status, n_isolated); status -1 one system, -2 island-wise, 7 islanded without reference."""
    lib = L.load()
    ptr = np.zeros(len(cases) + 1, np.int32)
    ptr[1:] = np.cumsum([len(c) for c in cases])
    br = np.ascontiguousarray([k for c in cases for k in c], np.int32)
    if br.size == 0:
        br = np.zeros(1, np.int32)
    isl = np.zeros(len(cases), np.int32)
    st = np.zeros(len(cases), np.int32)
    niso = np.zeros(len(cases), np.int32)
    bt = np.ascontiguousarray(a.bus_type, np.uint8)
    f = np.ascontiguousarray(a.br_from, np.int32)
    t = np.ascontiguousarray(a.br_to, np.int32)
    bs = np.ascontiguousarray(a.br_status, np.uint8)
    L.check(lib.sblx_debug_topology(a.n, a.slack, L.ptr(bt, L.c_u8p), len(f), L.ptr(f, L.c_i32p), L.ptr(t, L.c_i32p), L.ptr(bs, L.c_u8p), 0,
                                    len(cases), L.ptr(ptr, L.c_i32p), L.ptr(br, L.c_i32p), L.ptr(isl, L.c_i32p), L.ptr(st, L.c_i32p),
                                    L.ptr(niso, L.c_i32p)), None)
    return isl, st, niso


def analyze_only(Y: sp.csc_matrix, slack: int, **options):
    """Symbolic analysis without CUDA (CPU tests)."""
    lib = L.load()
    Y = sp.csc_matrix(Y)
    Y.sort_indices()
    opt = L.default_options(index_base=0, **options)
    cp = np.ascontiguousarray(Y.indptr, np.int64)
    rv = np.ascontiguousarray(Y.indices, np.int64)
    info = L.Info()
    res = C.c_double()
    L.check(lib.sblx_analyze_only(Y.shape[0], L.ptr(cp, L.c_i64p), L.ptr(rv, L.c_i64p), slack, C.byref(opt), C.byref(info),
                                  C.byref(res)), None)
    return L.as_dict(info), res.value


# ---------------------------------------------------------------------------
# B200BatchSolver: AbstractExternalSolver backend (src/solver_interface.jl:134-145;
# template: ext/SparlectraAnalyticLoadFlowExt.jl:44-136)
# ---------------------------------------------------------------------------
class B200BatchSolver(AbstractExternalSolver):
    def __init__(self, **options):
        self.options = options


def solvePf(solver: AbstractExternalSolver, model: PFModel, tol: float = 1e-8, **kwargs) -> PFSolution:
    """solvePf(solver, model; tol, kwargs...) -> PFSolution (src/solver_interface.jl:143-145)."""
    if not isinstance(solver, B200BatchSolver):
        raise NotImplementedError("solvePf(::AbstractExternalSolver, ::PFModel) is not implemented for this solver.")
    a = model.arrays
    if a is None:
        raise ValueError("solvePf: model was not built by buildPfModel")
    opts = dict(solver.options)
    opts.update(kwargs)
    eng = Engine(a, tol=tol, **opts)
    try:
        V, conv, it, res, st = eng.solve_one(model.V0)
    finally:
        eng.close()
    return PFSolution(V=V, converged=conv, iters=it, residual_inf=res, meta=dict(status=st, reason=REASON.get(st)))


def runpf_external(grid, solver, tol=1e-8, flatstart=None, include_limits=False, **solver_kwargs):
    """runpf_external! (src/solver_interface.jl:402-442): build model, solve, re-check the
    canonical mismatch, return (iters, status, sol)."""
    model = buildPfModel(grid, flatstart=flatstart, include_limits=include_limits)
    sol = solvePf(solver, model, tol=tol, qlimits_enabled=int(include_limits), **solver_kwargs)
    res = mismatchInf(model, sol.V) if (sol.residual_inf <= 0.0 or not math.isfinite(sol.residual_inf)) else sol.residual_inf
    sol2 = PFSolution(V=sol.V, converged=sol.converged and res <= tol, iters=sol.iters, residual_inf=res, meta=sol.meta)
    return sol2.iters, (0 if sol2.converged else 1), sol2


# ---------------------------------------------------------------------------
# contingency API (src/contingency.jl)
# ---------------------------------------------------------------------------
@dataclass
class ContingencyCase:
    name: str
    kind: str = "branch"
    element: str = ""

    def __post_init__(self):
        if self.kind != "branch":
            raise ValueError(f"ContingencyCase: only kind = :branch is supported in this version, got :{self.kind}.")
        if not self.element:
            self.element = self.name


@dataclass
class ContingencyResult:
    name: str
    converged: bool
    iterations: int
    max_vm_pu: float
    min_vm_pu: float
    max_branch_loading_pct: float
    overloaded_branches: list
    voltage_violations: list
    island_count: int
    error: Optional[str]


def generateN1Branches(grid, include_transformers: bool = True):
    """generateN1Branches (src/contingency.jl:99-119): one case per in-service branch;
    parallel circuits sharing a name are disambiguated as "<name>#<branchIdx>" (1-based)."""
    count = {}
    for nm in grid.br_name:
        count[nm] = count.get(nm, 0) + 1
    cases = []
    for k in range(grid.nbranch):
        if grid.br_status[k] != 1:
            continue
        if not include_transformers and grid.br_ratio[k] != 0.0:
            continue
        nm = grid.br_name[k]
        el = f"{nm}#{k + 1}" if count[nm] > 1 else nm
        cases.append(ContingencyCase(el, "branch", el))
    return cases


def _resolve_branch(grid, element: str):
    """_resolve_contingency_branch (src/contingency.jl:124-135)."""
    pos = element.rfind("#")
    if pos >= 0:
        try:
            idx = int(element[pos + 1:])
        except ValueError:
            return None
        name = element[:pos]
        if 1 <= idx <= grid.nbranch and grid.br_name[idx - 1] == name:
            return idx - 1
        return None
    for k, nm in enumerate(grid.br_name):
        if nm == element:
            return k
    return None


PER_SCENARIO_KEYS = ("converged", "status", "iterations", "n_switch_events", "island_count", "final_mismatch", "vmin_pu", "vmax_pu",
                     "max_loading_pct", "n_voltage_violations", "n_overloaded", "n_tiny_pivots", "V", "bus_type_final", "flows")


def _solve_with_lists(eng, lists, return_raw):
    """one batched call; voltage-violation / overload lists come compacted from the device, so the bus voltages only
    cross PCIe when the caller asks for the raw arrays"""
    raw = eng.solve_outages(lists, want_v=return_raw, want_types=return_raw, want_lists=True)
    if not raw["lists_complete"]:
        raw = eng.solve_outages(lists, want_v=return_raw, want_types=return_raw, want_lists=True,
                                list_capacity=max(raw["violation_total"], raw["overload_total"], raw["qlimit_event_total"]) + 16)
    return raw


def _group(pairs, B):
    """(scenario, element) rows sorted by scenario -> list of element arrays per scenario"""
    out = [[] for _ in range(B)]
    for sc, el in pairs:
        out[int(sc)].append(int(el))
    return out


def runContingenciesBatched(grid, cases, vm_min_pu=0.9, vm_max_pu=1.1, maxIte=30, tol=1e-8, backend=None,
                            return_raw=False, retry_flat_start=False, n_devices=1, **kwargs):
    """runContingencies! (src/contingency.jl:257-319) with the batched B200 backend: the
    base case is solved first (warm-start template; on failure the cases start flat), then
    ALL cases go to the GPU(s) in one call.  Results come back in input order.  `retry_flat_start`: the cases that did
    not converge get one second batch from a flat start, iterations summed (contingency.jl:195-207).
    `n_devices` > 1: the cases are sharded over that many GPUs of this process inside the library (contiguous chunks
    like contingency.jl:304, one ncclAllGather of the per-case records)."""
    if not (vm_min_pu < vm_max_pu):
        raise ValueError("runContingencies!: vm_min_pu must be below vm_max_pu.")
    if len(cases) == 0:
        return []
    opts = dict(backend.options) if isinstance(backend, B200BatchSolver) else {}
    opts.update(kwargs)
    opts.update(max_iter=maxIte, tol=tol, vm_min_pu=vm_min_pu, vm_max_pu=vm_max_pu,
                cooldown_iters=grid.cooldown_iters, q_hyst_pu=grid.q_hyst_pu)
    mk = (lambda arr: MultiEngine(arr, n_devices, **opts)) if n_devices > 1 else (lambda arr: Engine(arr, **opts))
    a = M.build_arrays(grid)
    eng = mk(a)
    try:
        Vb, conv, _, _, _ = eng.solve_one(None)
        if conv:
            # template nodes carry the solved vm / va (rectangular_result_updates.jl:30-47)
            vm, va = np.abs(Vb), np.degrees(np.angle(Vb))
            a2 = M.build_arrays(grid, vm=vm, va_deg=va, flatstart=False)
        else:
            a2 = M.build_arrays(grid, flatstart=True)           # contingency.jl:283-287
            eng.close()
            eng = mk(a2)                                          # Vset may differ for a flat start
        eng.set_v0(a2.v0)
        first_by_name = {}
        for k, nm in enumerate(grid.br_name):
            first_by_name.setdefault(nm, k)               # _resolve_contingency_branch takes the first match by name

        def resolve(el):
            return _resolve_branch(grid, el) if "#" in el else first_by_name.get(el)
        idx = [resolve(c.element) for c in cases]
        lists = [[k] if k is not None else [-1] for k in idx]
        raw = _solve_with_lists(eng, lists, return_raw)
        viol, over = _group(raw["violations"], len(cases)), _group(raw["overloads"], len(cases))
        if retry_flat_start:
            again = [i for i in range(len(cases)) if idx[i] is not None and not raw["converged"][i] and 1 <= int(raw["status"][i]) <= 6]
            if again:
                eng.close()
                vmt, vat = (np.abs(Vb), np.degrees(np.angle(Vb))) if conv else (None, None)
                a3 = M.build_arrays(grid, vm=vmt, va_deg=vat, flatstart=True)     # cnet.flatstart = true on the template copy
                eng = mk(a3)
                raw2 = _solve_with_lists(eng, [lists[i] for i in again], return_raw)
                v2, o2 = _group(raw2["violations"], len(again)), _group(raw2["overloads"], len(again))
                for j, i in enumerate(again):
                    it1 = int(raw["iterations"][i])
                    for key in PER_SCENARIO_KEYS:
                        if key in raw and key in raw2:
                            raw[key][i] = raw2[key][j]
                    raw["iterations"][i] += it1
                    viol[i], over[i] = v2[j], o2[j]
    finally:
        eng.close()
    out = []
    for i, c in enumerate(cases):
        st = int(raw["status"][i])
        if idx[i] is None:
            out.append(ContingencyResult(c.name, False, 0, math.nan, math.nan, math.nan, [], [], 0,
                                         f"unknown branch {c.element} (not found in net.branchVec)"))
            continue
        if not raw["converged"][i]:
            out.append(ContingencyResult(c.name, False, int(raw["iterations"][i]), math.nan, math.nan, math.nan, [], [],
                                         int(raw["island_count"][i]), STATUS_TEXT.get(st, f"status {st}")))
            continue
        out.append(ContingencyResult(c.name, True, int(raw["iterations"][i]), float(raw["vmax_pu"][i]),
                                     float(raw["vmin_pu"][i]), float(raw["max_loading_pct"][i]),
                                     [grid.br_name[b] for b in over[i]], [grid.bus_name[b] for b in viol[i]],
                                     int(raw["island_count"][i]), None))
    return (out, raw) if return_raw else out


class MultiEngine:
    """One process, several GPUs (sblx_create_multi): the same model on `n_devices` B200s, batches sharded inside the
    library with one ncclAllGather of the per-scenario records.  Same solve interface as `Engine`."""

    def __init__(self, arrays: M.ModelArrays, n_devices: int, devices=None, index_base: int = 0, **options):
        self.lib = L.load()
        self.a, self.n, self.base, self.nd = arrays, arrays.n, int(index_base), int(n_devices)
        self.opt = L.default_options(index_base=self.base, base_mva=arrays.baseMVA, **options)
        k = Engine._pack_model(arrays, self.base)
        self._keep = k
        dv = None if devices is None else np.ascontiguousarray(devices, np.int32)
        h = C.c_void_p()
        rc = self.lib.sblx_create_multi(C.byref(h), self.nd, L.ptr(dv, L.c_i32p), self.n, *Engine._model_args(k, arrays, self.base), C.byref(self.opt))
        L.check(rc, None)
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.sblx_destroy_multi(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            msg = self.lib.sblx_multi_last_error(self.h)
            raise L.SblxError(f"sblx error {rc}: {msg.decode() if msg else ''}")

    def set_v0(self, v0):
        self._check(self.lib.sblx_multi_set_v0(self.h, L.ptr(L.c128_to_f64(v0), L.c_f64p)))

    def set_locked_pv(self, buses):
        b = np.ascontiguousarray(np.asarray(buses, np.int64) + self.base, np.int32)
        self._check(self.lib.sblx_multi_set_locked_pv(self.h, len(b), L.ptr(b, L.c_i32p)))

    def stats(self, dev_index=0) -> dict:
        s = L.Stats()
        self._check(self.lib.sblx_multi_get_stats(self.h, dev_index, C.byref(s)))
        return L.as_dict(s)

    def solve_one(self, v_start=None):
        h0 = self.lib.sblx_multi_handle(self.h, 0)
        vout = np.zeros(self.n, np.complex128)
        c, it, res, st = C.c_uint8(), C.c_int32(), C.c_double(), C.c_int32()
        vs = None if v_start is None else L.c128_to_f64(v_start)
        L.check(self.lib.sblx_solve_one(h0, L.ptr(vs, L.c_f64p), L.ptr(vout.view(np.float64), L.c_f64p), C.byref(c),
                                        C.byref(it), C.byref(res), C.byref(st)), h0)
        return vout, bool(c.value), int(it.value), float(res.value), int(st.value)

    def solve_outages(self, cases, want_v=True, want_types=True, want_lists=False, want_flows=False, list_capacity=0) -> dict:
        optr, obr = Engine.pack_outages(cases) if not isinstance(cases, tuple) else cases
        B = len(optr) - 1
        if self.base:
            obr = np.ascontiguousarray(obr + self.base, np.int32)
        r = Engine._alloc(B, self.n, want_v, want_types, want_lists, want_flows, len(self.a.br_from), list_capacity)
        s = Engine._results_struct(r)
        self._check(self.lib.sblx_multi_solve_outages(self.h, B, L.ptr(optr, L.c_i32p), L.ptr(obr, L.c_i32p), C.byref(s)))
        return Engine._finish_lists(self, r)


# ---------------------------------------------------------------------------
# reporting helpers of the contingency API (host only; src/contingency.jl:321-380)
# ---------------------------------------------------------------------------
def _fit_column(s: str, width: int) -> str:
    return s if len(s) <= width else s[: width - 1] + "~"


def printContingencyResults(results, io=None, max_rows: int = 50):
    """printContingencyResults (src/contingency.jl:329-363): fixed-width table of the cases."""
    import sys
    io = io or sys.stdout
    w = io.write
    w(f"N-1 contingency results ({len(results)} case(s), {sum(1 for r in results if r.converged)} converged)\n")
    w("-" * 116 + "\n")
    w("%-28s %-9s %5s %9s %9s %12s %8s %8s %7s  %s\n" % ("case", "converged", "iter", "Vmin[pu]", "Vmax[pu]", "loading[%]", "overld", "V-viol", "islands", "error"))
    w("-" * 116 + "\n")
    fmt = lambda x, f: "-" if (x is None or math.isnan(x)) else f % x
    for r in results[:max_rows]:
        w("%-28s %-9s %5d %9s %9s %12s %8d %8d %7d  %s\n" % (
            _fit_column(r.name, 28), "yes" if r.converged else "NO", r.iterations, fmt(r.min_vm_pu, "%.4f"), fmt(r.max_vm_pu, "%.4f"),
            fmt(r.max_branch_loading_pct, "%.1f"), len(r.overloaded_branches), len(r.voltage_violations), r.island_count, r.error or ""))
    if len(results) > max_rows:
        w(f"... {len(results) - max_rows} more row(s) not shown (max_rows = {max_rows})\n")
    w("-" * 116 + "\n")


def writeContingencyResultsCSV(path: str, results) -> str:
    """writeContingencyResultsCSV (src/contingency.jl:365-380): semicolon-separated, one row per case, input order."""
    jl = lambda v: ("true" if v else "false") if isinstance(v, (bool, np.bool_)) else ("NaN" if isinstance(v, float) and math.isnan(v) else str(v))
    with open(path, "w") as f:
        f.write("name;converged;iterations;min_vm_pu;max_vm_pu;max_branch_loading_pct;overloaded_branches;voltage_violations;island_count;error\n")
        for r in results:
            f.write(";".join([r.name, jl(bool(r.converged)), jl(int(r.iterations)), jl(float(r.min_vm_pu)), jl(float(r.max_vm_pu)),
                              jl(float(r.max_branch_loading_pct)), ",".join(r.overloaded_branches), ",".join(r.voltage_violations),
                              jl(int(r.island_count)), r.error or ""]) + "\n")
    return path
