"""Grid tables -> solver arrays (host side of the engine, vectorised numpy).

Mirrors what `buildPfModel` (src/solver_interface.jl:193-295) and the setup block
of `runpf_rectangular!` (rectangular_network_solver.jl:466-527,609-622) derive
from a `Net`; the per-function reference lines are cited inline.  This is product
host code - it does not import anything from `oracle/`.
"""
from __future__ import annotations

from dataclasses import dataclass
import numpy as np
import scipy.sparse as sp

BUS_ISOLATED, BUS_PQ, BUS_PV, BUS_SLACK = 0, 1, 2, 3


def branch_stamps(grid):
    """calcAdmittance (src/branch.jl:297-310): Y11=(ys+ysh/2)/|t|^2, Y12=-ys/conj(t),
    Y21=-ys/t, Y22=ys+ysh/2 with t = ratio*cis(shift) (ratio 0 => line, t=1;
    src/branch.jl:446-456)."""
    ys = 1.0 / (grid.br_r + 1j * grid.br_x)
    ysh = grid.br_g + 1j * grid.br_b
    is_line = grid.br_ratio == 0.0
    rat = np.where(is_line, 1.0, grid.br_ratio)
    sh = np.where(is_line, 0.0, grid.br_shift_deg)
    t = rat * np.exp(1j * np.deg2rad(sh))
    t = np.where((rat == 1.0) & (sh == 0.0), 1.0 + 0.0j, t)
    y11 = (ys + 0.5 * ysh) / np.abs(t) ** 2
    y12 = -ys / np.conj(t)
    y21 = -ys / t
    y22 = ys + 0.5 * ysh
    return y11, y12, y21, y22


def ybus(grid):
    """createYBUS (src/equicircuit.jl:583-645): stamp sum of in-service branches
    plus :Y shunts on the diagonal (isolated buses keep zero rows)."""
    n = grid.n
    y11, y12, y21, y22 = branch_stamps(grid)
    on = grid.br_status == 1
    f, t = grid.br_from[on], grid.br_to[on]
    rows = np.concatenate([f, t, f, t])
    cols = np.concatenate([f, t, t, f])
    vals = np.concatenate([y11[on], y22[on], y12[on], y21[on]])
    deg = np.bincount(np.concatenate([f, t]), minlength=n)
    if len(grid.sh_bus):
        keep = deg[grid.sh_bus] > 0
        rows = np.concatenate([rows, grid.sh_bus[keep]])
        cols = np.concatenate([cols, grid.sh_bus[keep]])
        vals = np.concatenate([vals, grid.sh_y[keep]])
    Y = sp.coo_matrix((vals, (rows, cols)), shape=(n, n)).tocsc()
    Y.sum_duplicates()
    Y.sort_indices()
    return Y


def bus_types(grid):
    """refreshBusTypesFromProsumers! (src/network.jl:497-525)."""
    n = grid.n
    has_slack = np.zeros(n, bool)
    has_reg = np.zeros(n, bool)
    np.logical_or.at(has_slack, grid.ps_bus[grid.ps_ref], True)
    reg = grid.ps_is_gen & ~np.isnan(grid.ps_vm)
    np.logical_or.at(has_reg, grid.ps_bus[reg], True)
    return np.where(has_slack, BUS_SLACK, np.where(has_reg, BUS_PV, BUS_PQ)).astype(np.uint8)


def s_spec(grid):
    """buildComplexSVec (src/network.jl:2359-2402)."""
    n = grid.n
    p = np.nan_to_num(grid.ps_p, nan=0.0)
    q = np.nan_to_num(grid.ps_q, nan=0.0)
    g = grid.ps_is_gen
    pg = np.bincount(grid.ps_bus[g], weights=p[g], minlength=n)
    qg = np.bincount(grid.ps_bus[g], weights=q[g], minlength=n)
    pl = np.bincount(grid.ps_bus[~g], weights=p[~g], minlength=n)
    ql = np.bincount(grid.ps_bus[~g], weights=q[~g], minlength=n)
    return (pg - pl) / grid.baseMVA + 1j * ((qg - ql) / grid.baseMVA)


def qload(grid):
    """build_qload_pu (src/solver_core.jl:302-313)."""
    g = grid.ps_is_gen
    q = np.nan_to_num(grid.ps_q, nan=0.0)
    return np.bincount(grid.ps_bus[~g], weights=q[~g] / grid.baseMVA, minlength=grid.n)


def q_limits(grid):
    """_buildQLimits! (src/network.jl:1932-1981): per-bus SUM of generator limits;
    buses without a generator keep non-finite sentinels (= no limit)."""
    n = grid.n
    qmin = np.full(n, np.inf)
    qmax = np.full(n, -np.inf)
    for k in np.flatnonzero(grid.ps_is_gen):
        b = grid.ps_bus[k]
        lo = -np.inf if np.isnan(grid.ps_qmin[k]) else grid.ps_qmin[k] / grid.baseMVA
        hi = np.inf if np.isnan(grid.ps_qmax[k]) else grid.ps_qmax[k] / grid.baseMVA
        qmin[b] = qmin[b] + lo if np.isfinite(qmin[b]) else lo
        qmax[b] = qmax[b] + hi if np.isfinite(qmax[b]) else hi
    return qmin, qmax


def vset(grid, vm, flatstart=False):
    """_bus_voltage_setpoints_from_prosumers (rectangular_voltage_setpoints.jl:51-147)."""
    out = np.where(np.isnan(vm), 1.0, vm).astype(float)
    ok = (grid.ps_is_gen & ~np.isnan(grid.ps_vm)) | (flatstart & grid.ps_ref & ~np.isnan(grid.ps_vm))
    seen = set()
    for k in np.flatnonzero(ok):
        b = int(grid.ps_bus[k])
        if b not in seen:
            seen.add(b)
            out[b] = grid.ps_vm[k]
    return out


def start_vector(grid, vm, va_deg, types, flatstart=False):
    """initialVrect (src/network.jl:2309-2348) + slack magnitude fix
    (rectangular_network_solver.jl:523)."""
    v = np.where(np.isnan(vm) | (vm <= 0.0), 1.0, vm)
    V = v * np.exp(1j * np.deg2rad(va_deg))
    if flatstart:
        slack = types == BUS_SLACK
        V = np.where(slack, V, np.where(types == BUS_PV, v + 0j, 1.0 + 0j))
    return V


@dataclass
class ModelArrays:
    n: int
    slack: int                 # 0-based
    Y: sp.csc_matrix
    bus_type: np.ndarray       # uint8
    vset: np.ndarray
    s_spec: np.ndarray         # complex
    v0: np.ndarray             # complex
    qmin: np.ndarray
    qmax: np.ndarray
    qload: np.ndarray
    br_from: np.ndarray
    br_to: np.ndarray
    y11: np.ndarray
    y12: np.ndarray
    y21: np.ndarray
    y22: np.ndarray
    br_rating: np.ndarray
    br_status: np.ndarray
    baseMVA: float


def build_arrays(grid, vm=None, va_deg=None, flatstart=None) -> ModelArrays:
    vm = grid.vm0 if vm is None else np.asarray(vm, float)
    va = grid.va0_deg if va_deg is None else np.asarray(va_deg, float)
    flat = grid.flatstart if flatstart is None else flatstart
    t = bus_types(grid)
    sl = np.flatnonzero(t == BUS_SLACK)
    if len(sl) != 1:
        raise ValueError(f"grid {grid.name}: exactly one reference bus is required, found {len(sl)}")
    slack = int(sl[0])
    vs = vset(grid, vm, flat)
    V0 = start_vector(grid, vm, va, t, flat)
    a = np.angle(V0[slack]) if V0[slack] != 0 else 0.0
    V0[slack] = vs[slack] * np.exp(1j * a)
    y11, y12, y21, y22 = branch_stamps(grid)
    qmin, qmax = q_limits(grid)
    return ModelArrays(n=grid.n, slack=slack, Y=ybus(grid), bus_type=t, vset=vs, s_spec=s_spec(grid), v0=V0,
                       qmin=qmin, qmax=qmax, qload=qload(grid), br_from=grid.br_from.astype(np.int32),
                       br_to=grid.br_to.astype(np.int32), y11=y11, y12=y12, y21=y21, y22=y22,
                       br_rating=grid.br_rating.astype(float), br_status=(grid.br_status == 1).astype(np.uint8),
                       baseMVA=grid.baseMVA)
