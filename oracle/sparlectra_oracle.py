"""CPU ORACLE - TEST INFRASTRUCTURE ONLY.

A float64 restatement (numpy/scipy) of the reference's rectangular
Newton-Raphson contingency path, function by function, each citing the
reference file:line it follows (paths relative to /root/reference).  It is the
checker for the CUDA engine; nothing in the product path may import it (only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline/reference
legs do).

Pinning status: the reference is 100 % Julia and Julia is not installed in this
image, so the reference itself cannot be executed here.  The oracle is pinned
against every known-answer fixture the reference's own tests hold for this path
(tests/test_oracle_reference_fixtures.py lists them with file:line) and against
the published IEEE-14 solution; the linear solve (UMFPACK in the reference, an
un-vendored Julia stdlib dependency: SparseArrays/SuiteSparse_jll pinned only
by `julia = "1.12"` in Project.toml:43) is restated with SuperLU
(`scipy.sparse.linalg.splu`, partial pivoting) - the reference's tests pin
solution quality (residuals), not UMFPACK's bit pattern, so bit-level parity
with UMFPACK pivot order is UNPINNED.

Input is a `Grid` table object (sparlectra.jl_b200/sblx/grid.py, neutral data).
Indices are 0-based here; the reference is 1-based.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

SLACK, PV, PQ, ISO = 3, 2, 1, 0   # bus type codes used by the oracle (and the C ABI)

# status / rejection reasons (rectangular_network_solver.jl:702,783,902,951,1034;
# rectangular_final_status.jl:74)
R_NONE = 0
R_NOT_CONVERGED = 1
R_NONFINITE = 2
R_SINGULAR = 3
R_REMAINING_PV_Q = 4
R_ACTIVE_PV_VRES = 5
R_MAX_SWITCHING = 6
R_ISLANDED_NO_REF = 7
R_UNKNOWN_BRANCH = 8
REASON_NAMES = {
    0: "none", 1: "nr_mismatch_not_converged", 2: "nr_nonfinite", 3: "singular_newton_step",
    4: "remaining_pv_q_limit_violations", 5: "active_pv_voltage_residual", 6: "max_switching_exceeded",
    7: "islanded_without_reference", 8: "unknown_branch",
}


# ---------------------------------------------------------------------------
# a10: per-scenario setup helpers
# ---------------------------------------------------------------------------
def branch_ratio(ratio: float, shift_deg: float) -> complex:
    """calcBranchRatio (src/branch.jl:446-456) + calcComplexRatio (src/equicircuit.jl:126-128)."""
    rat = 1.0 if ratio == 0.0 else ratio
    sh = 0.0 if ratio == 0.0 else shift_deg
    if rat != 1.0 or sh != 0.0:
        a = math.radians(sh)
        return rat * complex(math.cos(a), math.sin(a))
    return 1.0 + 0.0j


def branch_stamp(r, x, b, g, ratio, shift_deg):
    """calcAdmittance (src/branch.jl:297-310)."""
    ys = 1.0 / complex(r, x)
    ysh = complex(g, b)
    t = branch_ratio(ratio, shift_deg)
    y11 = (ys + 0.5 * ysh) / (abs(t) ** 2)
    y12 = -1.0 * ys / t.conjugate()
    y21 = -1.0 * ys / t
    y22 = ys + 0.5 * ysh
    return y11, y12, y21, y22


def in_service_branches(grid, removed=()):
    rem = set(int(k) for k in removed)
    return [k for k in range(grid.nbranch) if grid.br_status[k] == 1 and k not in rem]


def mark_isolated(grid, removed=()):
    """markIsolatedBuses! (src/network.jl:1881-1930): a bus with no closed branch."""
    conn = np.zeros(grid.n, bool)
    for k in in_service_branches(grid, removed):
        conn[grid.br_from[k]] = True
        conn[grid.br_to[k]] = True
    return [i for i in range(grid.n) if not conn[i]]


def bus_types_from_prosumers(grid, iso=()):
    """refreshBusTypesFromProsumers! (src/network.jl:497-525); Isolated is kept."""
    n = grid.n
    has_slack = np.zeros(n, bool)
    has_reg = np.zeros(n, bool)
    for k in range(len(grid.ps_bus)):
        b = grid.ps_bus[k]
        if grid.ps_ref[k]:
            has_slack[b] = True
        if grid.ps_is_gen[k] and not math.isnan(grid.ps_vm[k]):   # isRegulating: has a vm setpoint
            has_reg[b] = True
    t = np.where(has_slack, SLACK, np.where(has_reg, PV, PQ)).astype(np.int64)
    for i in iso:
        t[i] = ISO
    return t


def ac_islands(grid, removed=(), iso=()):
    """_active_ac_island_components (ac_islands.jl:50-93): components over
    in-service branches among non-isolated buses, ordered by smallest member."""
    n = grid.n
    parent = list(range(n))

    def find(i):
        while parent[i] != i:
            parent[i] = parent[parent[i]]
            i = parent[i]
        return i

    isoset = set(iso)
    for k in in_service_branches(grid, removed):
        f, t = int(grid.br_from[k]), int(grid.br_to[k])
        if f in isoset or t in isoset:
            continue
        ra, rb = find(f), find(t)
        if ra != rb:
            if ra < rb:
                parent[rb] = ra
            else:
                parent[ra] = rb
    groups = {}
    for b in range(n):
        if b in isoset:
            continue
        groups.setdefault(find(b), []).append(b)
    return sorted(groups.values(), key=min)


def create_ybus(grid, removed=()):
    """createYBUS (src/equicircuit.jl:583-645) re-embedded to n x n with zero
    rows/cols for isolated buses (rectangular_network_solver.jl:170-197)."""
    n = grid.n
    rows, cols, vals = [], [], []
    for k in in_service_branches(grid, removed):
        f, t = int(grid.br_from[k]), int(grid.br_to[k])
        y11, y12, y21, y22 = branch_stamp(grid.br_r[k], grid.br_x[k], grid.br_b[k], grid.br_g[k],
                                          grid.br_ratio[k], grid.br_shift_deg[k])
        rows += [f, t, f, t]
        cols += [f, t, t, f]
        vals += [y11, y22, y12, y21]
    iso = set(mark_isolated(grid, removed))
    for k in range(len(grid.sh_bus)):
        b = int(grid.sh_bus[k])
        if b in iso:
            continue
        rows.append(b)
        cols.append(b)
        vals.append(grid.sh_y[k])
    Y = sp.coo_matrix((np.array(vals, np.complex128), (rows, cols)), shape=(n, n)).tocsc()
    Y.sum_duplicates()
    return Y


def initial_vrect(grid, vm, va_deg, types, flatstart=False):
    """initialVrect (src/network.jl:2309-2348). `grid` is only used for its size;
    vm/va_deg/types may be island-local sub-arrays."""
    n = len(vm)
    slack = int(np.flatnonzero(np.asarray(types) == SLACK)[0])
    V0 = np.empty(n, np.complex128)
    for k in range(n):
        v = 1.0 if (math.isnan(vm[k]) or vm[k] <= 0.0) else float(vm[k])
        if flatstart:
            if k == slack:
                a = math.radians(va_deg[k])
                V0[k] = complex(v * math.cos(a), v * math.sin(a))
            elif types[k] == PV:
                V0[k] = complex(v, 0.0)
            else:
                V0[k] = 1.0 + 0.0j
        else:
            a = math.radians(va_deg[k])
            V0[k] = complex(v * math.cos(a), v * math.sin(a))
    return V0, slack


def complex_svec(grid):
    """buildComplexSVec (src/network.jl:2359-2402), :Y shunts only."""
    n = grid.n
    pg = np.zeros(n); qg = np.zeros(n); pl = np.zeros(n); ql = np.zeros(n)
    for k in range(len(grid.ps_bus)):
        b = grid.ps_bus[k]
        p = 0.0 if math.isnan(grid.ps_p[k]) else grid.ps_p[k]
        q = 0.0 if math.isnan(grid.ps_q[k]) else grid.ps_q[k]
        if grid.ps_is_gen[k]:
            pg[b] += p; qg[b] += q
        else:
            pl[b] += p; ql[b] += q
    return ((pg - pl) / grid.baseMVA) + 1j * ((qg - ql) / grid.baseMVA)


def qload_pu(grid):
    """build_qload_pu (src/solver_core.jl:302-313)."""
    q = np.zeros(grid.n)
    for k in range(len(grid.ps_bus)):
        if grid.ps_is_gen[k]:
            continue
        q[grid.ps_bus[k]] += (0.0 if math.isnan(grid.ps_q[k]) else grid.ps_q[k]) / grid.baseMVA
    return q


def q_limits_pu(grid):
    """_buildQLimits! (src/network.jl:1932-1981) + _sanitize_q_limits! (src/limits.jl:211-225).
    Buses without generators keep the (+Inf, -Inf) sentinels => non-finite => no limit."""
    n = grid.n
    qmin = np.full(n, np.inf)
    qmax = np.full(n, -np.inf)
    for k in range(len(grid.ps_bus)):
        if not grid.ps_is_gen[k]:
            continue
        b = grid.ps_bus[k]
        lo = -np.inf if math.isnan(grid.ps_qmin[k]) else grid.ps_qmin[k] / grid.baseMVA
        hi = np.inf if math.isnan(grid.ps_qmax[k]) else grid.ps_qmax[k] / grid.baseMVA
        qmin[b] = (qmin[b] + lo) if np.isfinite(qmin[b]) else lo
        qmax[b] = (qmax[b] + hi) if np.isfinite(qmax[b]) else hi
    return qmin, qmax


def voltage_setpoints(grid, vm, flatstart=False):
    """_bus_voltage_setpoints_from_prosumers (rectangular_voltage_setpoints.jl:51-147):
    first regulating generator's vm_pu at the bus, else the node's vm, else 1.0;
    the slack prosumer's setpoint only counts for a flat start."""
    n = grid.n
    vset = np.array([1.0 if math.isnan(vm[k]) else float(vm[k]) for k in range(n)])
    seen = set()
    for k in range(len(grid.ps_bus)):
        b = int(grid.ps_bus[k])
        ok = (flatstart and grid.ps_ref[k]) or (grid.ps_is_gen[k] and not math.isnan(grid.ps_vm[k]))
        if not ok or math.isnan(grid.ps_vm[k]) or b in seen:
            continue
        seen.add(b)
        vset[b] = float(grid.ps_vm[k])
    return vset


# ---------------------------------------------------------------------------
# a1/a2: injections and mismatch
# ---------------------------------------------------------------------------
def calc_injections(Y, V):
    """src/solver_core.jl:35-47."""
    I = Y @ V
    return V * np.conj(I)


def mismatch_rectangular(Y, V, S, types, vset, slack):
    """rectangular_core_equations.jl:79-148: F = [dP_i, dQ_i | dV_i] per non-slack bus."""
    n = len(V)
    Sc = calc_injections(Y, V)
    F = np.zeros(2 * (n - 1))
    row = 0
    for i in range(n):
        if i == slack:
            continue
        if types[i] == PQ:
            F[row] = Sc[i].real - S[i].real
            F[row + 1] = Sc[i].imag - S[i].imag
        elif types[i] == PV:
            F[row] = Sc[i].real - S[i].real
            F[row + 1] = abs(V[i]) - vset[i]
        else:
            raise ValueError(f"mismatch_rectangular: unsupported bus type at bus {i}")
        row += 2
    return F


def mismatch_rectangular_fast(Y, V, S, types, vset, slack):
    """Vectorised twin of `mismatch_rectangular` (same arithmetic per entry)."""
    Sc = V * np.conj(Y @ V)
    ns = np.array([i for i in range(len(V)) if i != slack])
    F = np.empty(2 * len(ns))
    F[0::2] = Sc.real[ns] - S.real[ns]
    second = np.where(types[ns] == PV, np.abs(V[ns]) - vset[ns], Sc.imag[ns] - S.imag[ns])
    F[1::2] = second
    return F


def max_abs(F):
    """maximum(abs.(F)) (rectangular_network_solver.jl:764) - NaN propagating."""
    a = np.abs(F)
    if np.isnan(a).any():
        return float("nan")
    return float(a.max()) if len(a) else 0.0


# ---------------------------------------------------------------------------
# a3: Jacobian (rectangular_jacobian_builders.jl:212-429)
# ---------------------------------------------------------------------------
def jacobian_triplets(Y, V, types, slack, structural_pattern=False, vm_eps=1e-9):
    """Emit-order restatement of _assemble_rectangular_jacobian! (loops, CSC walk).
    Rows interleaved per non-slack bus, columns [Vr(non-slack); Vi(non-slack)]."""
    n = len(V)
    Y = sp.csc_matrix(Y)
    I = Y @ V
    pos = -np.ones(n, np.int64)
    rb = -np.ones(n, np.int64)
    k = 0
    for i in range(n):
        if i == slack:
            continue
        pos[i] = k
        rb[i] = 2 * k
        k += 1
    R, C, X = [], [], []

    def emit(r, c, v):
        R.append(r); C.append(c); X.append(v)

    for j in range(n):
        if pos[j] < 0:
            continue
        cr, ci = pos[j], (n - 1) + pos[j]
        for p in range(Y.indptr[j], Y.indptr[j + 1]):
            i = Y.indices[p]
            if i == slack:
                continue
            yij = Y.data[p]
            a = V[i] * np.conj(yij)
            dvr = a
            dvi = 1j * (-a)
            if i == j:
                dvr = dvr + np.conj(I[i])
                dvi = dvi + 1j * np.conj(I[i])
            if structural_pattern or abs(dvr.real) > 0.0:
                emit(rb[i], cr, dvr.real)
            if structural_pattern or abs(dvi.real) > 0.0:
                emit(rb[i], ci, dvi.real)
            if types[i] == PQ:
                if structural_pattern or abs(dvr.imag) > 0.0:
                    emit(rb[i] + 1, cr, dvr.imag)
                if structural_pattern or abs(dvi.imag) > 0.0:
                    emit(rb[i] + 1, ci, dvi.imag)
            elif types[i] != PV:
                raise ValueError("unsupported bus type")
    for i in range(n):
        if i == slack or types[i] != PV:
            continue
        vm = abs(V[i])
        if vm == 0.0 and not structural_pattern:
            continue
        d = vm if vm > 0.0 else vm_eps
        a, b = V[i].real / d, V[i].imag / d
        if structural_pattern or abs(a) > 0.0:
            emit(rb[i] + 1, pos[i], a)
        if structural_pattern or abs(b) > 0.0:
            emit(rb[i] + 1, (n - 1) + pos[i], b)
    return np.array(R, np.int64), np.array(C, np.int64), np.array(X, float)


def jacobian_sparse(Y, V, types, slack, structural_pattern=False):
    """build_rectangular_jacobian_pq_pv_sparse (rectangular_jacobian_builders.jl:95-206)."""
    n = len(V)
    R, C, X = jacobian_triplets(Y, V, types, slack, structural_pattern)
    m = 2 * (n - 1)
    return sp.coo_matrix((X, (R, C)), shape=(m, m)).tocsc()


def jacobian_sparse_fast(Y, V, types, slack):
    """Vectorised twin of `jacobian_sparse` (structural pattern; numerically the
    same matrix - explicit zeros do not change a direct solve, note 8 of SURVEY 8a)."""
    n = len(V)
    Yc = sp.coo_matrix(Y)
    I = Y @ V
    pos = -np.ones(n, np.int64)
    ns = np.array([i for i in range(n) if i != slack])
    pos[ns] = np.arange(n - 1)
    i, j, y = Yc.row, Yc.col, Yc.data
    keep = (pos[i] >= 0) & (pos[j] >= 0)
    i, j, y = i[keep], j[keep], y[keep]
    a = V[i] * np.conj(y)
    diag = i == j
    dvr = a + np.where(diag, np.conj(I[i]), 0.0)
    dvi = 1j * (-a + np.where(diag, np.conj(I[i]), 0.0))
    rP = 2 * pos[i]
    cr, ci = pos[j], (n - 1) + pos[j]
    ispq = types[i] == PQ
    R = [rP, rP, rP[ispq] + 1, rP[ispq] + 1]
    C = [cr, ci, cr[ispq], ci[ispq]]
    X = [dvr.real, dvi.real, dvr.imag[ispq], dvi.imag[ispq]]
    pv = np.array([b for b in ns if types[b] == PV], np.int64)
    if len(pv):
        vm = np.abs(V[pv])
        d = np.where(vm > 0.0, vm, 1e-9)
        R += [2 * pos[pv] + 1, 2 * pos[pv] + 1]
        C += [pos[pv], (n - 1) + pos[pv]]
        X += [V[pv].real / d, V[pv].imag / d]
    m = 2 * (n - 1)
    return sp.coo_matrix((np.concatenate(X), (np.concatenate(R), np.concatenate(C))), shape=(m, m)).tocsc()


def jacobian_dense_wirtinger(Y, V, types, slack):
    """Independent second path: dense Jacobian from the Wirtinger blocks of
    build_complex_jacobian (rectangular_core_equations.jl:43-60):
    dS = J11 dV + J12 dV*, with dV = dVr + j dVi."""
    n = len(V)
    Yd = np.asarray(sp.csc_matrix(Y).todense())
    I = Yd @ V
    J11 = np.diag(np.conj(I))
    J12 = np.diag(V) @ np.conj(Yd)
    dS_dVr = J11 + J12
    dS_dVi = 1j * (J11 - J12)
    ns = [i for i in range(n) if i != slack]
    m = 2 * (n - 1)
    J = np.zeros((m, m))
    for a, i in enumerate(ns):
        for b, j in enumerate(ns):
            J[2 * a, b] = dS_dVr[i, j].real
            J[2 * a, (n - 1) + b] = dS_dVi[i, j].real
            if types[i] == PQ:
                J[2 * a + 1, b] = dS_dVr[i, j].imag
                J[2 * a + 1, (n - 1) + b] = dS_dVi[i, j].imag
        if types[i] == PV:
            vm = abs(V[i])
            J[2 * a + 1, a] = V[i].real / vm
            J[2 * a + 1, (n - 1) + a] = V[i].imag / vm
    return J


# ---------------------------------------------------------------------------
# a4: linear solve
# ---------------------------------------------------------------------------
class SingularStep(Exception):
    pass


def solve_linear(J, b):
    """solve_sparse_system (src/solver_core.jl:60-82): sparse LU; on a singular
    matrix the sparse-QR fallback (here: dense least squares, min-norm, which
    coincides with the basic QR solution on the zero rows/cols isolated buses
    produce); if that fails too -> SingularException."""
    J = sp.csc_matrix(J)
    try:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("error")
            lu = spla.splu(J)
            x = lu.solve(b)
        if not np.all(np.isfinite(x)):
            raise RuntimeError("non-finite LU solution")
        return x
    except Exception:
        if J.shape[0] > 4000:
            raise SingularStep()
        x, *_ = np.linalg.lstsq(J.toarray(), b, rcond=None)
        if not np.all(np.isfinite(x)):
            raise SingularStep()
        return x


# ---------------------------------------------------------------------------
# a6/a7: Q-limit active set
# ---------------------------------------------------------------------------
@dataclass
class QLimitState:
    log: list = field(default_factory=list)       # (iter, bus, side)  net.qLimitLog
    events: dict = field(default_factory=dict)    # bus -> side          net.qLimitEvents

    def count(self, bus):                         # qlimit_switch_count (src/limits.jl:86-88)
        return sum(1 for e in self.log if e[1] == bus)

    def last_iter(self, bus):                     # lastQLimitIter (src/limits.jl:99-106)
        for e in reversed(self.log):
            if e[1] == bus:
                return e[0]
        return None


def active_set_q_limits(st: QLimitState, it, nb, get_qreq, is_pv, make_pq, make_pv, qmin, qmax, pv_orig,
                        allow_reenable, q_hyst_pu, cooldown_iters, lock=(), guard_max_switches=10,
                        freeze=True):
    """active_set_q_limits! (src/limits.jl:398-568), default violation mode (:delayed_switch)."""
    changed = reenabled = False
    lockset = set(lock)
    for bus in range(nb):
        if not is_pv(bus) or bus in lockset:
            continue
        qreq = get_qreq(bus)
        has_hi = np.isfinite(qmax[bus])
        has_lo = np.isfinite(qmin[bus])
        if not (has_hi or has_lo):
            continue
        side, clamp = None, 0.0
        if has_hi and qreq > qmax[bus] + q_hyst_pu:
            side, clamp = "max", qmax[bus]
        elif has_lo and qreq < qmin[bus] - q_hyst_pu:
            side, clamp = "min", qmin[bus]
        if side is None:
            continue
        if freeze and st.count(bus) >= guard_max_switches:
            continue
        make_pq(bus, clamp, side)
        st.log.append((it, bus, side))
        st.events[bus] = side
        changed = True
    if allow_reenable:
        for bus in range(nb):
            if not pv_orig[bus] or bus not in st.events or is_pv(bus):
                continue
            if not (np.isfinite(qmin[bus]) or np.isfinite(qmax[bus])):
                continue
            if st.count(bus) > 1:
                continue
            qreq = get_qreq(bus)
            lo = qmin[bus] + q_hyst_pu if np.isfinite(qmin[bus]) else -np.inf
            hi = qmax[bus] - q_hyst_pu if np.isfinite(qmax[bus]) else np.inf
            ready = (qreq > lo) and (qreq < hi)
            if ready and cooldown_iters > 0:
                li = st.last_iter(bus)
                if li is not None and (it - li) < cooldown_iters:
                    ready = False
            if ready:
                make_pv(bus)
                del st.events[bus]
                reenabled = True
    return changed, reenabled


# ---------------------------------------------------------------------------
# a8: the NR loop + acceptance (runpf_rectangular!)
# ---------------------------------------------------------------------------
@dataclass
class NRResult:
    V: np.ndarray
    converged: bool
    iters: int
    reason: int
    types: np.ndarray
    history: list
    n_events: int
    numerical_converged: bool
    S: np.ndarray
    log: list = field(default_factory=list)      # net.qLimitLog: (iteration, bus, side) in push order


def run_nr(Y, V0, S0, types0, vset, slack, qmin, qmax, qload, *, maxiter=30, tol=1e-8, damp=1.0,
           qlimits_enabled=True, qlimit_start_iter=2, q_hyst_pu=0.0, cooldown_iters=0, lock=(),
           guard_max_switches=10, freeze=True, fast=True):
    """runpf_rectangular! NR loop (rectangular_network_solver.jl:746-922) and final
    acceptance (:931-1094).  `types0` must already map Isolated -> PQ with S = 0
    (:508-512).  V0[slack] is expected to carry |Vset[slack]| (:523)."""
    n = len(V0)
    Y = sp.csr_matrix(Y)
    V = V0.astype(np.complex128).copy()
    S = S0.astype(np.complex128).copy()
    types = np.array(types0).copy()
    pv_orig = types == PV
    allow_reenable = (cooldown_iters > 0) or (q_hyst_pu > 0.0)
    st = QLimitState()
    history = []
    converged = False
    reason = R_NOT_CONVERGED
    iters = 0
    mism = mismatch_rectangular_fast if fast else mismatch_rectangular
    jac = (lambda: jacobian_sparse_fast(Y, V, types, slack)) if fast else (lambda: jacobian_sparse(Y, V, types, slack))
    ns = np.array([i for i in range(n) if i != slack])
    for it in range(1, maxiter + 1):
        iters = it
        F = mism(Y, V, S, types, vset, slack)
        max_mis = max_abs(F)
        history.append(max_mis)
        if not math.isfinite(max_mis):
            converged = False
            reason = R_NONFINITE
            break
        changed = reenabled = False
        conv_now = max_mis <= tol
        if qlimits_enabled:
            # _handle_rectangular_qlimit_iteration! (rectangular_qlimit_iteration.jl:25-210)
            Scalc = calc_injections(Y, V)
            conv_now = history[-1] <= tol
            ready = (it >= qlimit_start_iter) or conv_now
            if ready:
                def make_pq(bus, clamp, side):
                    types[bus] = PQ
                    S[bus] = complex(S[bus].real, clamp - qload[bus])

                def make_pv(bus):
                    types[bus] = PV

                changed, reenabled = active_set_q_limits(
                    st, it, n,
                    get_qreq=lambda b: 0.0 if types[b] == SLACK else Scalc[b].imag + qload[b],
                    is_pv=lambda b: types[b] == PV, make_pq=make_pq, make_pv=make_pv,
                    qmin=qmin, qmax=qmax, pv_orig=pv_orig, allow_reenable=allow_reenable,
                    q_hyst_pu=q_hyst_pu, cooldown_iters=cooldown_iters, lock=lock,
                    guard_max_switches=guard_max_switches, freeze=freeze)
                if changed or reenabled:
                    F = mism(Y, V, S, types, vset, slack)
                    max_mis = max_abs(F)
                    history[-1] = max_mis
        if conv_now and not (changed or reenabled):
            converged = True
            reason = R_NONE
            break
        # complex_newton_step_rectangular (rectangular_newton_step.jl:474-604), fixed damping
        F0 = mism(Y, V, S, types, vset, slack)
        J = jac()
        try:
            dx = solve_linear(J, -F0)
        except SingularStep:
            converged = False
            reason = R_SINGULAR
            break
        Vn = V.copy()
        Vn[ns] = V[ns] + (damp * dx[: n - 1] + 1j * (damp * dx[n - 1:]))   # _apply_rectangular_delta! (:38-55)
        Vn[slack] = V0[slack]                                                # :919
        V = Vn
    numerical = converged
    # final acceptance (:939-1094)
    pv_res = 0.0
    for k in range(n):
        if types[k] == PV and k not in st.events:
            pv_res = max(pv_res, abs(abs(V[k]) - vset[k]))        # rectangular_core_equations.jl:150-161
    if numerical and pv_res > tol:
        converged = False
        reason = R_ACTIVE_PV_VRES
    if numerical and qlimits_enabled:
        Sbus = calc_injections(Y, V)
        viol = 0
        for k in range(n):
            if types[k] != PV or not (np.isfinite(qmin[k]) and np.isfinite(qmax[k])):
                continue
            qc = Sbus[k].imag + qload[k]                            # rectangular_network_solver.jl:64-95
            if qc < qmin[k] - tol or qc > qmax[k] + tol:
                viol += 1
        if viol > 0:
            converged = False
            reason = R_REMAINING_PV_Q                                # rectangular_final_status.jl:70-74
    if numerical and freeze:
        counts = {}
        for e in st.log:
            counts[e[1]] = counts.get(e[1], 0) + 1
        osc = sum(1 for c in counts.values() if c >= max(guard_max_switches, 1))
        ok = numerical and pv_res <= tol and (reason != R_REMAINING_PV_Q) and not (osc > 0)
        if osc > 0 and not ok:
            converged = False
            reason = R_MAX_SWITCHING                                 # :1027-1036
    return NRResult(V=V, converged=converged, iters=iters, reason=reason, types=types, history=history,
                    n_events=len(st.log), numerical_converged=numerical, S=S, log=list(st.log))


# ---------------------------------------------------------------------------
# scenario level: runpf! with islands (rectangular_network_solver.jl:1783-2444)
# ---------------------------------------------------------------------------
@dataclass
class PFResult:
    converged: bool
    iters: int
    reason: int
    V: np.ndarray
    types: np.ndarray          # final solver bus types (ISO kept for isolated buses)
    island_count: int
    n_events: int
    final_mismatch: float
    iso: list
    log: list = field(default_factory=list)      # Q-limit events (iteration, bus, side), single-solve path only


def runpf(grid, vm, va_deg, removed=(), *, maxiter=30, tol=1e-8, damp=1.0, qlimits_enabled=True,
          qlimit_start_iter=2, lock=(), s_override=None, qload_override=None, flatstart=None, fast=True) -> PFResult:
    """runpf! on the net with `removed` branches out (contingency.jl:189-194):
    isolated buses, bus types from prosumers, island detection, single solve or
    per-island sub-solves with MATPOWER-like reference promotion."""
    flat = grid.flatstart if flatstart is None else flatstart
    n = grid.n
    iso = mark_isolated(grid, removed)
    vm = np.array(vm, float).copy()
    va = np.array(va_deg, float).copy()
    for i in iso:                      # setVmVa!(vm=0, va=0) in markIsolatedBuses!
        vm[i] = 0.0
        va[i] = 0.0
    types_net = bus_types_from_prosumers(grid, iso)
    comps = ac_islands(grid, removed, iso)
    Y = create_ybus(grid, removed)
    S = complex_svec(grid) if s_override is None else np.array(s_override, np.complex128).copy()
    ql = qload_pu(grid) if qload_override is None else np.array(qload_override, float)
    qmin, qmax = q_limits_pu(grid)
    vset = voltage_setpoints(grid, vm, flat)
    ins = in_service_branches(grid, removed)
    nbr_in = lambda comp: sum(1 for k in ins if grid.br_from[k] in comp and grid.br_to[k] in comp)
    multi = len(comps) > 1 and any(nbr_in(set(c)) > 0 for c in comps)
    kw = dict(maxiter=maxiter, tol=tol, damp=damp, qlimits_enabled=qlimits_enabled,
              qlimit_start_iter=qlimit_start_iter, q_hyst_pu=grid.q_hyst_pu, cooldown_iters=grid.cooldown_iters,
              fast=fast)
    Vout = np.ones(n, np.complex128)
    types_out = types_net.copy()
    if not multi:
        if not np.any(types_net == SLACK):
            return PFResult(False, 0, R_ISLANDED_NO_REF, Vout, types_out, len(comps), 0, float("nan"), iso)
        V0, slack = initial_vrect(grid, vm, va, types_net, flat)
        t = types_net.copy()
        Sx = S.copy()
        for i in range(n):
            if t[i] == ISO:
                t[i] = PQ
                Sx[i] = 0.0
        V0[slack] = _mag_keep_angle(V0[slack], vset[slack])
        r = run_nr(Y, V0, Sx, t, vset, slack, qmin, qmax, ql, lock=lock, **kw)
        tt = r.types.copy()
        for i in iso:
            tt[i] = ISO
        return PFResult(r.converged, r.iters, r.reason, r.V, tt, len(comps), r.n_events,
                        r.history[-1] if r.history else float("nan"), iso, log=r.log)
    # multi-island path (:1919-2243): validate references, solve each island with branches
    refs = []
    for comp in comps:
        sl = [b for b in comp if types_net[b] == SLACK]
        pv = [b for b in comp if types_net[b] == PV]
        refs.append(min(sl) if sl else (min(pv) if pv else -1))
    if any(r < 0 for r in refs):
        return PFResult(False, 0, R_ISLANDED_NO_REF, Vout, types_out, len(comps), 0, float("nan"), iso)
    total_it = 0
    ok = True
    reason = R_NONE
    nev = 0
    fm = 0.0
    Ycsr = sp.csr_matrix(Y)
    Vout = np.array([complex(1.0, 0.0) if (math.isnan(vm[k]) or vm[k] <= 0.0) else
                     vm[k] * complex(math.cos(math.radians(va[k])), math.sin(math.radians(va[k]))) for k in range(n)])
    for comp, ref in zip(comps, refs):
        idx = np.array(comp)
        if len(idx) == 1 and nbr_in(set(comp)) == 0:
            continue
        Ysub = Ycsr[idx][:, idx]
        t = types_net[idx].copy()
        loc = {b: k for k, b in enumerate(comp)}
        t[loc[ref]] = SLACK
        sl = loc[ref]
        V0, _ = initial_vrect(grid, vm[idx], va[idx], t, flat)
        V0[sl] = _mag_keep_angle(V0[sl], vset[idx][sl])
        # lock_pv_to_pq_buses is handed unchanged to the island sub-net (rectangular_network_solver.jl:1991): its
        # indices address the sub-net's own numbering; ids beyond the island's size are ignored (limits.jl:434-438)
        r = run_nr(Ysub, V0, S[idx], t, vset[idx], sl, qmin[idx], qmax[idx], ql[idx],
                   lock=tuple(b for b in lock if 0 <= b < len(idx)), **kw)
        total_it += r.iters
        nev += r.n_events
        fm = max(fm, r.history[-1] if r.history else 0.0)
        Vout[idx] = r.V
        types_out[idx] = r.types
        if not r.converged:
            # an island failure throws out of runpf! (rectangular_network_solver.jl:2086-2095, serial path throws at
            # the first failing island), so _run_one_contingency reports iterations = 0 (contingency.jl:192-194,213-222)
            return PFResult(False, 0, r.reason, Vout, types_out, len(comps), nev, fm, iso)
    return PFResult(ok, total_it, R_NONE, Vout, types_out, len(comps), nev, fm, iso)


def _mag_keep_angle(v, vm):
    """_apply_voltage_magnitude_preserving_angle (rectangular_voltage_helpers.jl:24-31)."""
    if v == 0:
        return complex(vm, 0.0)
    a = math.atan2(v.imag, v.real)
    return complex(vm * math.cos(a), vm * math.sin(a))


# ---------------------------------------------------------------------------
# a12: branch flows and metrics
# ---------------------------------------------------------------------------
def branch_flows(grid, V, removed=()):
    """calcNetLosses! per-branch flows (src/losses.jl:53-85,160-170), p.u."""
    m = grid.nbranch
    sf = np.zeros(m, np.complex128)
    stt = np.zeros(m, np.complex128)
    for k in in_service_branches(grid, removed):
        f, t = int(grid.br_from[k]), int(grid.br_to[k])
        rat = grid.br_ratio[k] if grid.br_ratio[k] != 0.0 else 1.0
        ang = grid.br_shift_deg[k] if grid.br_ratio[k] != 0.0 else 0.0
        a = math.radians(ang)
        tap = rat * complex(math.cos(a), math.sin(a))
        yik = 1.0 / complex(grid.br_r[k], grid.br_x[k])
        y0 = 0.5 * complex(grid.br_g[k], grid.br_b[k])
        ui, uj = V[f] / tap, V[t]
        sf[k] = abs(ui) ** 2 * np.conj(y0 + yik) - ui * np.conj(uj) * np.conj(yik)
        ui, uj = V[t], V[f] / tap
        stt[k] = abs(ui) ** 2 * np.conj(y0 + yik) - ui * np.conj(uj) * np.conj(yik)
    return sf, stt


def contingency_metrics(grid, V, iso, removed, vm_min=0.9, vm_max=1.1):
    """_contingency_metrics (src/contingency.jl:149-178)."""
    isoset = set(iso)
    vms = [abs(V[i]) for i in range(grid.n) if i not in isoset and math.isfinite(abs(V[i]))]
    vmin = min(vms) if vms else float("nan")
    vmax = max(vms) if vms else float("nan")
    viol = [i for i in range(grid.n) if i not in isoset and (abs(V[i]) < vm_min or abs(V[i]) > vm_max)]
    sf, stt = branch_flows(grid, V, removed)
    max_loading = float("nan")
    over = []
    for k in in_service_branches(grid, removed):
        rt = grid.br_rating[k]
        if math.isnan(rt) or not math.isfinite(rt) or rt <= 0.0:
            continue
        loading = 100.0 * max(abs(sf[k]), abs(stt[k])) * grid.baseMVA / rt
        if math.isnan(max_loading) or loading > max_loading:
            max_loading = loading
        if loading > 100.0:
            over.append(k)
    return dict(vmin=vmin, vmax=vmax, violations=viol, max_loading=max_loading, overloaded=over)


# ---------------------------------------------------------------------------
# a11: runContingencies! (src/contingency.jl:257-319)
# ---------------------------------------------------------------------------
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
    error: object
    V: object = None
    reason: int = 0
    types: object = None


def generate_n1_branches(grid):
    """generateN1Branches (src/contingency.jl:99-119): one case per in-service branch."""
    return [[k] for k in range(grid.nbranch) if grid.br_status[k] == 1]


def solve_base(grid, maxiter=30, tol=1e-8, **kw):
    """Base-case template solve (src/contingency.jl:275-287). Returns (vm, va_deg, flat)."""
    r = runpf(grid, grid.vm0, grid.va0_deg, (), maxiter=maxiter, tol=tol, **kw)
    if r.converged:
        vm = np.abs(r.V)
        va = np.degrees(np.angle(r.V))      # rectangular_result_updates.jl:30-47
        return vm, va, False, r
    return grid.vm0.copy(), grid.va0_deg.copy(), True, r


def run_contingencies(grid, cases, vm_min=0.9, vm_max=1.1, maxiter=30, tol=1e-8, keep_v=True,
                      retry_flat_start=False, **kw):
    vm, va, flat, _ = solve_base(grid, maxiter, tol, **kw)
    out = []
    for ci, removed in enumerate(cases):
        bad = any((k < 0 or k >= grid.nbranch) for k in removed)
        name = "base" if not len(removed) else ("?" if bad else "+".join(grid.br_name[k] for k in removed))
        if bad:
            out.append(ContingencyResult(name, False, 0, math.nan, math.nan, math.nan, [], [], 0,
                                         "unknown branch", reason=R_UNKNOWN_BRANCH))
            continue
        r = runpf(grid, vm, va, removed, maxiter=maxiter, tol=tol, flatstart=flat, **kw)
        if (not r.converged) and r.reason != R_ISLANDED_NO_REF and retry_flat_start:
            # one bounded second attempt from a flat start, iterations summed (contingency.jl:195-207)
            r2 = runpf(grid, vm, va, removed, maxiter=maxiter, tol=tol, flatstart=True, **kw)
            r2.iters += r.iters
            r = r2
        if r.reason == R_ISLANDED_NO_REF:
            out.append(ContingencyResult(name, False, r.iters, math.nan, math.nan, math.nan, [], [], r.island_count,
                                         "islanded without reference", V=r.V if keep_v else None, reason=r.reason,
                                         types=r.types))
            continue
        if not r.converged:
            out.append(ContingencyResult(name, False, r.iters, math.nan, math.nan, math.nan, [], [], r.island_count,
                                         "power flow did not converge (status 1)", V=r.V if keep_v else None,
                                         reason=r.reason, types=r.types))
            continue
        # calcNetLosses! reads vm / va_deg back from the nodes (src/network.jl:2206-2212)
        Vr = np.abs(r.V) * np.exp(1j * np.radians(np.degrees(np.angle(r.V))))
        m = contingency_metrics(grid, Vr, r.iso, removed, vm_min, vm_max)
        out.append(ContingencyResult(name, True, r.iters, m["vmax"], m["vmin"], m["max_loading"], m["overloaded"],
                                     m["violations"], r.island_count, None, V=r.V if keep_v else None, reason=0,
                                     types=r.types))
    return out
