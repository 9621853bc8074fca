"""Grid tables: the flat, array-level description of one power grid.

This is neutral input data (no solver logic): the same tables feed the CUDA
engine's host side (`sblx.model`), the CPU oracle under `oracle/` (which
restates the reference's own net->array helpers) and `bench.py`.

The table columns mirror what the reference keeps on `Net`:
  * buses      -> `Node` (src/node.jl): start voltage vm/va
  * branches   -> `Branch` (src/branch.jl:163-246): r/x/b/g in p.u., ratio
                  (0.0 marks an AC line, the MATPOWER convention kept by
                  src/branch.jl:446-456), phase shift in degrees, status,
                  rating sn_MVA
  * prosumers  -> `ProSumer` (src/prosumer.jl:334-390): generator/load flag,
                  P/Q in MW/MVAr, optional voltage setpoint (makes a generator
                  regulating: src/network.jl:1196-1198), optional Q limits,
                  reference flag (`referencePri`)
  * shunts     -> `Shunt` with model :Y (src/equicircuit.jl:633-644)

Generators (cases) in this module:
  * `tiled_grid`          - src/synthetic_grids.jl:88-171 (+ documented augmentation)
  * `warmup_case118`      - data/webui/warmup_case118.jl (rule-generated 2x59 ladder)
  * `warmup_case3`        - data/webui/warmup_case3.jl
  * `case14`              - IEEE 14-bus MATPOWER case (public data, hand-embedded;
                            the reference downloads it: src/FetchMatpowerCase.jl)
  * `test3bus`, `pv_regression_net`, `island_chain`, `diverge_ring` - the
    programmatic fixtures of test/testgrid.jl and test/test_contingency.jl
"""
from __future__ import annotations

from dataclasses import dataclass, field
import numpy as np

NAN = float("nan")


@dataclass
class Grid:
    name: str
    baseMVA: float
    bus_name: list
    vm0: np.ndarray            # start |V| per bus (p.u.)
    va0_deg: np.ndarray        # start angle per bus (deg)
    br_from: np.ndarray        # 0-based bus index
    br_to: np.ndarray
    br_r: np.ndarray
    br_x: np.ndarray
    br_b: np.ndarray           # total line charging susceptance (p.u.)
    br_g: np.ndarray
    br_ratio: np.ndarray       # 0.0 => line (ratio treated as 1, shift as 0)
    br_shift_deg: np.ndarray
    br_status: np.ndarray      # 1 in service
    br_rating: np.ndarray      # sn_MVA, NaN = unrated
    br_name: list
    ps_bus: np.ndarray         # 0-based bus index
    ps_is_gen: np.ndarray      # bool: Injection type (generator / ext. network / sync. machine)
    ps_p: np.ndarray           # MW
    ps_q: np.ndarray           # MVAr
    ps_vm: np.ndarray          # voltage setpoint p.u. or NaN
    ps_qmin: np.ndarray        # MVAr or NaN
    ps_qmax: np.ndarray        # MVAr or NaN
    ps_ref: np.ndarray         # bool: referencePri set (slack)
    sh_bus: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int64))
    sh_y: np.ndarray = field(default_factory=lambda: np.zeros(0, np.complex128))
    cooldown_iters: int = 0
    q_hyst_pu: float = 0.0
    flatstart: bool = False

    @property
    def n(self) -> int:
        return len(self.bus_name)

    @property
    def nbranch(self) -> int:
        return len(self.br_from)

    def copy(self) -> "Grid":
        kw = {}
        for k, v in self.__dict__.items():
            kw[k] = v.copy() if isinstance(v, np.ndarray) else (list(v) if isinstance(v, list) else v)
        return Grid(**kw)


class GridBuilder:
    """Tiny incremental builder with the call shapes of the reference's
    addBus!/addACLine!/addPIModelACLine!/addPIModelTrafo!/addProsumer!."""

    def __init__(self, name: str, baseMVA: float = 100.0, cooldown_iters: int = 0, q_hyst_pu: float = 0.0):
        self.name, self.baseMVA = name, baseMVA
        self.cooldown_iters, self.q_hyst_pu = cooldown_iters, q_hyst_pu
        self.bus, self.vm, self.va = [], [], []
        self.br = []
        self.ps = []
        self.sh = []

    def idx(self, name):
        return self.bus.index(name)

    def add_bus(self, name, vm_pu=1.0, va_deg=0.0):
        self.bus.append(name)
        self.vm.append(vm_pu)
        self.va.append(va_deg)

    def add_pi_line(self, f, t, r_pu, x_pu, b_pu=0.0, g_pu=0.0, status=1, rating=NAN, ratio=0.0, shift_deg=0.0, name=None):
        fi, ti = (self.idx(f) if isinstance(f, str) else f), (self.idx(t) if isinstance(t, str) else t)
        assert fi != ti
        kind = "ACL" if ratio == 0.0 else "2WT"
        nm = name or f"B_{kind}_{fi + 1}_{ti + 1}"
        self.br.append((fi, ti, r_pu, x_pu, b_pu, g_pu, ratio, shift_deg, status, rating, nm))

    def add_ac_line(self, f, t, length, r, x, c_nf_per_km=0.0, tan_delta=0.0, rating=NAN):
        # src/lines.jl:61-92 + src/network.jl:818-834 + src/lines.jl:142-148:
        # with length > 1 the stored r/x are used as they are (paramsBasedOnLength),
        # with length <= 1 they are multiplied by the length; the shunt comes from
        # 2*pi*50*C*(tan_delta + j).
        based_on_length = length > 1.0
        c_nf = c_nf_per_km * length if based_on_length else c_nf_per_km
        ysh = 2.0 * np.pi * 50.0 * c_nf * 1e-9 * complex(tan_delta, 1.0)
        g, b = ysh.real, ysh.imag
        if not based_on_length:
            r, x, b, g = r * length, x * length, b * length, g * length
        self.add_pi_line(f, t, r, x, b, g, rating=rating)

    def add_prosumer(self, bus, kind, p=NAN, q=NAN, vm_pu=NAN, va_deg=None, qMin=NAN, qMax=NAN, reference=False):
        """kind: one of the reference's type strings (src/prosumer.jl:555-562)."""
        k = kind.upper()
        is_gen = k in ("GENERATOR", "EXTERNALNETWORKINJECTION", "SYNCHRONOUSMACHINE", "STATICVARCOMPENSATOR")
        bi = self.idx(bus) if isinstance(bus, str) else bus
        # src/network.jl:1183-1192: the node keeps its own start voltage whenever it
        # differs from the prosumer setpoint by more than 1e-6 (addBus! defaults it to
        # 1.0), so the setpoint lives on the prosumer only; equal values just refresh
        # the angle.
        if not np.isnan(vm_pu) and abs(self.vm[bi] - vm_pu) <= 1e-6:
            self.vm[bi] = vm_pu
            if va_deg is not None:
                self.va[bi] = va_deg
        self.ps.append((bi, is_gen, p, q, vm_pu, qMin, qMax, reference))

    def add_shunt(self, bus, p_mw, q_mvar):
        # MATPOWER Gs/Bs (MW/MVAr at 1 p.u.) -> y = (Gs + jBs)/baseMVA
        bi = self.idx(bus) if isinstance(bus, str) else bus
        self.sh.append((bi, complex(p_mw, q_mvar) / self.baseMVA))

    def build(self) -> Grid:
        br = self.br
        ps = self.ps
        f = lambda i, dt: np.array([b[i] for b in br], dtype=dt)
        g = lambda i, dt: np.array([p[i] for p in ps], dtype=dt)
        return Grid(
            name=self.name, baseMVA=self.baseMVA, bus_name=list(self.bus),
            vm0=np.array(self.vm, float), va0_deg=np.array(self.va, float),
            br_from=f(0, np.int64), br_to=f(1, np.int64), br_r=f(2, float), br_x=f(3, float), br_b=f(4, float),
            br_g=f(5, float), br_ratio=f(6, float), br_shift_deg=f(7, float), br_status=f(8, np.int64),
            br_rating=f(9, float), br_name=[b[10] for b in br],
            ps_bus=g(0, np.int64), ps_is_gen=g(1, bool), ps_p=g(2, float), ps_q=g(3, float), ps_vm=g(4, float),
            ps_qmin=g(5, float), ps_qmax=g(6, float), ps_ref=g(7, bool),
            sh_bus=np.array([s[0] for s in self.sh], np.int64), sh_y=np.array([s[1] for s in self.sh], np.complex128),
            cooldown_iters=self.cooldown_iters, q_hyst_pu=self.q_hyst_pu,
        )


# ---------------------------------------------------------------------------
# MATPOWER-style table import (bus/gen/branch matrices) - the column meaning of
# src/createnet_powermat.jl:316-520: PD/QD -> ENERGYCONSUMER, GS/BS -> shunt,
# gen rows -> regulating generators with VG setpoint and QMIN/QMAX, bus type 3
# -> reference, TAP == 0 -> line, RATE_A == 0 -> unrated.
# ---------------------------------------------------------------------------
def from_matpower_tables(name, baseMVA, bus, gen, branch) -> Grid:
    bus = np.asarray(bus, float)
    gen = np.asarray(gen, float)
    branch = np.asarray(branch, float)
    ids = [int(b) for b in bus[:, 0]]
    pos = {b: i for i, b in enumerate(ids)}
    gb = GridBuilder(name, baseMVA)
    for row in bus:
        gb.add_bus(f"Bus{int(row[0])}", vm_pu=row[7], va_deg=row[8])
    for row in branch:
        rate = row[5] if row[5] > 0 else NAN
        gb.add_pi_line(pos[int(row[0])], pos[int(row[1])], row[2], row[3], row[4], 0.0, status=int(row[10]),
                       rating=rate, ratio=row[8], shift_deg=row[9])
    for row in bus:
        i = pos[int(row[0])]
        if row[2] != 0.0 or row[3] != 0.0:
            gb.add_prosumer(i, "ENERGYCONSUMER", p=row[2], q=row[3])
        if row[4] != 0.0 or row[5] != 0.0:
            gb.add_shunt(i, row[4], row[5])
    for row in gen:
        if row[7] <= 0:
            continue
        i = pos[int(row[0])]
        is_ref = int(bus[i, 1]) == 3
        is_pv = int(bus[i, 1]) == 2
        if is_ref:
            gb.ps.append((i, True, row[1], row[2], row[5], row[4], row[3], True))
            gb.vm[i] = row[5]
        elif is_pv:
            gb.ps.append((i, True, row[1], row[2], row[5], row[4], row[3], False))
            gb.vm[i] = row[5]
        else:  # generator at a PQ bus: fixed injection, not regulating
            gb.ps.append((i, True, row[1], row[2], NAN, row[4], row[3], False))
    return gb.build()


def case14() -> Grid:
    """IEEE 14-bus test case in MATPOWER form (public data; `case14.m`)."""
    bus = [
        [1, 3, 0, 0, 0, 0, 1, 1.06, 0, 0, 1, 1.06, 0.94],
        [2, 2, 21.7, 12.7, 0, 0, 1, 1.045, -4.98, 0, 1, 1.06, 0.94],
        [3, 2, 94.2, 19, 0, 0, 1, 1.01, -12.72, 0, 1, 1.06, 0.94],
        [4, 1, 47.8, -3.9, 0, 0, 1, 1.019, -10.33, 0, 1, 1.06, 0.94],
        [5, 1, 7.6, 1.6, 0, 0, 1, 1.02, -8.78, 0, 1, 1.06, 0.94],
        [6, 2, 11.2, 7.5, 0, 0, 1, 1.07, -14.22, 0, 1, 1.06, 0.94],
        [7, 1, 0, 0, 0, 0, 1, 1.062, -13.37, 0, 1, 1.06, 0.94],
        [8, 2, 0, 0, 0, 0, 1, 1.09, -13.36, 0, 1, 1.06, 0.94],
        [9, 1, 29.5, 16.6, 0, 19, 1, 1.056, -14.94, 0, 1, 1.06, 0.94],
        [10, 1, 9, 5.8, 0, 0, 1, 1.051, -15.1, 0, 1, 1.06, 0.94],
        [11, 1, 3.5, 1.8, 0, 0, 1, 1.057, -14.79, 0, 1, 1.06, 0.94],
        [12, 1, 6.1, 1.6, 0, 0, 1, 1.055, -15.07, 0, 1, 1.06, 0.94],
        [13, 1, 13.5, 5.8, 0, 0, 1, 1.05, -15.16, 0, 1, 1.06, 0.94],
        [14, 1, 14.9, 5, 0, 0, 1, 1.036, -16.04, 0, 1, 1.06, 0.94],
    ]
    gen = [
        [1, 232.4, -16.9, 10, 0, 1.06, 100, 1, 332.4, 0],
        [2, 40, 42.4, 50, -40, 1.045, 100, 1, 140, 0],
        [3, 0, 23.4, 40, 0, 1.01, 100, 1, 100, 0],
        [6, 0, 12.2, 24, -6, 1.07, 100, 1, 100, 0],
        [8, 0, 17.4, 24, -6, 1.09, 100, 1, 100, 0],
    ]
    br = [
        [1, 2, 0.01938, 0.05917, 0.0528, 0, 0, 0, 0, 0, 1],
        [1, 5, 0.05403, 0.22304, 0.0492, 0, 0, 0, 0, 0, 1],
        [2, 3, 0.04699, 0.19797, 0.0438, 0, 0, 0, 0, 0, 1],
        [2, 4, 0.05811, 0.17632, 0.034, 0, 0, 0, 0, 0, 1],
        [2, 5, 0.05695, 0.17388, 0.0346, 0, 0, 0, 0, 0, 1],
        [3, 4, 0.06701, 0.17103, 0.0128, 0, 0, 0, 0, 0, 1],
        [4, 5, 0.01335, 0.04211, 0, 0, 0, 0, 0, 0, 1],
        [4, 7, 0, 0.20912, 0, 0, 0, 0, 0.978, 0, 1],
        [4, 9, 0, 0.55618, 0, 0, 0, 0, 0.969, 0, 1],
        [5, 6, 0, 0.25202, 0, 0, 0, 0, 0.932, 0, 1],
        [6, 11, 0.09498, 0.1989, 0, 0, 0, 0, 0, 0, 1],
        [6, 12, 0.12291, 0.25581, 0, 0, 0, 0, 0, 0, 1],
        [6, 13, 0.06615, 0.13027, 0, 0, 0, 0, 0, 0, 1],
        [7, 8, 0, 0.17615, 0, 0, 0, 0, 0, 0, 1],
        [7, 9, 0, 0.11001, 0, 0, 0, 0, 0, 0, 1],
        [9, 10, 0.03181, 0.0845, 0, 0, 0, 0, 0, 0, 1],
        [9, 14, 0.12711, 0.27038, 0, 0, 0, 0, 0, 0, 1],
        [10, 11, 0.08205, 0.19207, 0, 0, 0, 0, 0, 0, 1],
        [12, 13, 0.22092, 0.19988, 0, 0, 0, 0, 0, 0, 1],
        [13, 14, 0.17093, 0.34802, 0, 0, 0, 0, 0, 0, 1],
    ]
    return from_matpower_tables("case14", 100.0, bus, gen, br)


def warmup_case3() -> Grid:
    """data/webui/warmup_case3.jl: slack 1.02, PV 1.01 (25 MW), 60/20 load."""
    bus = [
        [1, 3, 0, 0, 0, 0, 1, 1.02, 0, 110, 1, 1.1, 0.9],
        [2, 2, 0, 0, 0, 0, 1, 1.01, 0, 110, 1, 1.1, 0.9],
        [3, 1, 60.0, 20.0, 0, 0, 1, 1.0, 0, 110, 1, 1.1, 0.9],
    ]
    gen = [[1, 50, 0, 100, -100, 1.02, 100, 1, 150, 0], [2, 25, 0, 80, -80, 1.01, 100, 1, 100, 0]]
    br = [
        [1, 2, 0.02, 0.08, 0.02, 999, 999, 999, 0, 0, 1],
        [1, 3, 0.03, 0.12, 0.02, 999, 999, 999, 0, 0, 1],
        [2, 3, 0.02, 0.10, 0.02, 999, 999, 999, 0, 0, 1],
    ]
    return from_matpower_tables("sparlectra_webui_warmup_case3", 100.0, bus, gen, br)


def warmup_case118() -> Grid:
    """data/webui/warmup_case118.jl regenerated from its rule (header lines 15-20):
    two 59-bus rails (1..59, 60..118; r=0.008 x=0.035 b=0.010), a rung between
    k and k+59 for every odd k (r=0.010 x=0.045 b=0.008), 5 MW / 1.5 MVAr on every
    PQ bus, slack at bus 1 (Vg 1.02, +-300 MVAr), PV support at 11,21,31,41,51,115
    (80 MW, Vg 1.00, +-150 MVAr); every branch rated 999 MVA."""
    pv = {11, 21, 31, 41, 51, 115}
    bus, gen, br = [], [], []
    for i in range(1, 119):
        typ = 3 if i == 1 else (2 if i in pv else 1)
        pd, qd = (0.0, 0.0) if typ != 1 else (5.0, 1.5)
        bus.append([i, typ, pd, qd, 0, 0, 1, 1.0, 0.0, 110.0, 1, 1.10, 0.90])
    gen.append([1, 100.0, 0.0, 300.0, -300.0, 1.02, 100.0, 1, 400.0, 0.0])
    for i in sorted(pv):
        gen.append([i, 80.0, 0.0, 150.0, -150.0, 1.00, 100.0, 1, 200.0, 0.0])
    for a in (1, 60):
        for k in range(a, a + 58):
            br.append([k, k + 1, 0.008, 0.035, 0.010, 999.0, 999.0, 999.0, 0, 0, 1])
    for k in range(1, 60, 2):
        br.append([k, k + 59, 0.010, 0.045, 0.008, 999.0, 999.0, 999.0, 0, 0, 1])
    return from_matpower_tables("sparlectra_webui_warmup_case118", 100.0, bus, gen, br)


# ---------------------------------------------------------------------------
# src/synthetic_grids.jl
# ---------------------------------------------------------------------------
def choose_tiled_grid_dims(max_buses: int, aspect_ratio: float = 1.0):
    """src/synthetic_grids.jl:26-51 - largest rows*cols <= max_buses, ties by
    aspect-ratio score then fewer rows.  (Closed-form scan over rows only; the
    inner cols loop of the reference is monotone in area for a fixed row.)"""
    assert max_buses >= 4 and aspect_ratio > 0
    best = (2, 2, 4, abs(1.0 - aspect_ratio))
    for rows in range(2, max_buses + 1):
        cmax = max_buses // rows
        if cmax < 2:
            continue
        for cols in range(2, cmax + 1) if cmax * rows >= best[2] else ():
            area = rows * cols
            score = abs(cols / rows - aspect_ratio)
            if area > best[2] or (area == best[2] and score < best[3]) or (area == best[2] and score == best[3] and rows < best[0]):
                best = (rows, cols, area, score)
    return best[0], best[1]


def tiled_grid(max_buses: int = None, rows: int = None, cols: int = None, *, r=0.01, x=0.05, g=0.0, b=0.0,
               base_mva=100.0, load_mw_per_right_corner=50.0, load_mvar_per_right_corner=15.0,
               generation_balance=0.995, vm_slack=1.0, vm_flat=1.0,
               augment: bool = False, seed: int = 20260724, pv_every: int = 7, pv_vm: float = 1.02,
               pv_q_mvar: float = 30.0, load_mw=(0.5, 2.0), load_q_ratio=0.3, rating_mva: float = NAN) -> Grid:
    """build_synthetic_tiled_grid_net (src/synthetic_grids.jl:88-171): rows x cols
    buses named B_rrr_ccc, horizontal, vertical and one upper-left -> lower-right
    diagonal branch per tile (in that emit order), slack at (1,1), a
    non-regulating generator at (rows,1), loads at (1,cols) and (rows,cols).

    `augment=True` adds the deterministic study augmentation SURVEY.md 8(d)
    config 3/4 describes (the stock grid has no PV bus, no distributed load and
    no ratings): a load P~U(load_mw) MW, Q = load_q_ratio*P on every non-corner
    bus (numpy PCG64, `seed`), a regulating generator with +-pv_q_mvar limits on
    every `pv_every`-th bus of every `pv_every`-th row covering the local load,
    and an optional uniform branch rating."""
    if rows is None:
        rows, cols = choose_tiled_grid_dims(max_buses, 1.0)
    gb = GridBuilder(f"synthetic_tiled_grid_{rows}x{cols}", base_mva)
    nm = lambda rr, cc: "B_%03d_%03d" % (rr, cc)
    for rr in range(1, rows + 1):
        for cc in range(1, cols + 1):
            gb.add_bus(nm(rr, cc), vm_pu=(vm_slack if (rr == 1 and cc == 1) else vm_flat))
    id_ = lambda rr, cc: (rr - 1) * cols + (cc - 1)
    for rr in range(1, rows + 1):
        for cc in range(1, cols):
            gb.add_pi_line(id_(rr, cc), id_(rr, cc + 1), r, x, b, g, rating=rating_mva)
    for rr in range(1, rows):
        for cc in range(1, cols + 1):
            gb.add_pi_line(id_(rr, cc), id_(rr + 1, cc), r, x, b, g, rating=rating_mva)
    for rr in range(1, rows):
        for cc in range(1, cols):
            gb.add_pi_line(id_(rr, cc), id_(rr + 1, cc + 1), r, x, b, g, rating=rating_mva)
    sched_load_mw = 2.0 * load_mw_per_right_corner
    sched_load_mvar = 2.0 * load_mvar_per_right_corner
    gb.add_prosumer(id_(1, 1), "EXTERNALNETWORKINJECTION", vm_pu=vm_slack, va_deg=0.0, reference=True)
    gb.add_prosumer(id_(rows, 1), "GENERATOR", p=generation_balance * sched_load_mw, q=generation_balance * sched_load_mvar)
    gb.add_prosumer(id_(1, cols), "ENERGYCONSUMER", p=load_mw_per_right_corner, q=load_mvar_per_right_corner)
    gb.add_prosumer(id_(rows, cols), "ENERGYCONSUMER", p=load_mw_per_right_corner, q=load_mvar_per_right_corner)
    if augment:
        rng = np.random.Generator(np.random.PCG64(seed))
        corners = {id_(1, 1), id_(rows, 1), id_(1, cols), id_(rows, cols)}
        pl = rng.uniform(load_mw[0], load_mw[1], size=rows * cols)
        total = 0.0
        for i in range(rows * cols):
            if i in corners:
                continue
            gb.add_prosumer(i, "ENERGYCONSUMER", p=float(pl[i]), q=float(load_q_ratio * pl[i]))
            total += float(pl[i])
        pv_buses = [id_(rr, cc) for rr in range(pv_every // 2 + 1, rows + 1, pv_every)
                    for cc in range(pv_every // 2 + 1, cols + 1, pv_every) if id_(rr, cc) not in corners]
        share = generation_balance * total / max(len(pv_buses), 1)
        for i in pv_buses:
            gb.add_prosumer(i, "SYNCHRONOUSMACHINE", p=share, q=0.0, vm_pu=pv_vm, qMin=-pv_q_mvar, qMax=pv_q_mvar)
    return gb.build()


# ---------------------------------------------------------------------------
# Programmatic fixtures of the reference tests
# ---------------------------------------------------------------------------
def test3bus(cooldown=0, hyst_pu=0.0, qlim_min=NAN, qlim_max=NAN) -> Grid:
    """createTest3BusNet (test/testgrid.jl:1076-1114)."""
    gb = GridBuilder("test3bus", 100.0, cooldown, hyst_pu)
    for b in ("ASTADT", "STATION1", "VERBUND"):
        gb.add_bus(b)
    for f, t in (("ASTADT", "STATION1"), ("ASTADT", "VERBUND"), ("VERBUND", "STATION1")):
        gb.add_ac_line(f, t, length=25.0, r=0.0, x=0.4, c_nf_per_km=9.55, tan_delta=0.0)
    gb.add_prosumer("VERBUND", "EXTERNALNETWORKINJECTION", vm_pu=1.018182, va_deg=0.0, reference=True)
    gb.add_prosumer("STATION1", "SYNCHRONOUSMACHINE", p=70.0, q=33.2, vm_pu=1.027273, qMax=qlim_max, qMin=qlim_min)
    gb.add_prosumer("ASTADT", "ENERGYCONSUMER", p=100.0, q=30.0)
    return gb.build()


def pv_regression_net(vset=1.05) -> Grid:
    """_create_pv_voltage_regression_net (test/test_pv_voltage_residuals.jl:20-39)."""
    gb = GridBuilder("pv_voltage_residual_regression", 100.0)
    for b in ("Slack", "Load", "PV"):
        gb.add_bus(b)
    gb.add_ac_line("Slack", "Load", length=10.0, r=0.02, x=0.20)
    gb.add_ac_line("Load", "PV", length=10.0, r=0.02, x=0.20)
    gb.add_ac_line("Slack", "PV", length=12.0, r=0.02, x=0.25)
    gb.add_prosumer("Slack", "EXTERNALNETWORKINJECTION", vm_pu=1.0, va_deg=0.0, reference=True)
    gb.add_prosumer("Load", "ENERGYCONSUMER", p=40.0, q=12.0)
    gb.add_prosumer("PV", "SYNCHRONOUSMACHINE", p=25.0, vm_pu=vset, va_deg=0.0, qMin=-500.0, qMax=500.0)
    g = gb.build()
    g.vm0[2] = vset  # setPVBusVset!
    return g


def island_chain(with_pv: bool) -> Grid:
    """test/test_contingency.jl:97-139 - chain A-B-C-D; cutting B-C leaves {C,D}
    load-only (no reference) or with a PV generator at C (promotable)."""
    gb = GridBuilder("n1_island_pv" if with_pv else "n1_island", 100.0)
    for b in "ABCD":
        gb.add_bus(b)
    gb.add_prosumer("A", "EXTERNALNETWORKINJECTION", vm_pu=1.0, va_deg=0.0, reference=True)
    gb.add_prosumer("B", "ENERGYCONSUMER", p=10.0, q=3.0)
    if with_pv:
        gb.add_prosumer("C", "GENERATOR", p=5.0, vm_pu=1.0, qMin=-10.0, qMax=10.0)
    else:
        gb.add_prosumer("C", "ENERGYCONSUMER", p=5.0, q=1.0)
    gb.add_prosumer("D", "ENERGYCONSUMER", p=4.0, q=1.0)
    for f, t in ("AB", "BC", "CD"):
        gb.add_pi_line(f, t, 0.01, 0.08, 0.0)
    return gb.build()


def diverge_ring() -> Grid:
    """test/test_contingency.jl:141-160 - 350 MW on a 3-bus ring."""
    gb = GridBuilder("n1_diverge", 100.0)
    for b in "ABC":
        gb.add_bus(b)
    gb.add_prosumer("A", "EXTERNALNETWORKINJECTION", vm_pu=1.0, va_deg=0.0, reference=True)
    gb.add_prosumer("C", "ENERGYCONSUMER", p=350.0, q=100.0)
    for f, t in ("AB", "BC", "AC"):
        gb.add_pi_line(f, t, 0.01, 0.08, 0.0)
    return gb.build()


def twin_lines() -> Grid:
    """test/test_contingency.jl:47-62 - two parallel circuits A-B."""
    gb = GridBuilder("n1_twin", 100.0)
    for b in "AB":
        gb.add_bus(b)
    gb.add_prosumer("A", "EXTERNALNETWORKINJECTION", vm_pu=1.0, va_deg=0.0, reference=True)
    gb.add_prosumer("B", "ENERGYCONSUMER", p=20.0, q=5.0)
    gb.add_pi_line("A", "B", 0.01, 0.08, 0.0)
    gb.add_pi_line("A", "B", 0.02, 0.16, 0.0)
    return gb.build()


def random_meshed_grid(n=300, seed=1, extra_edge_frac=0.35, trafo_frac=0.12, shifter_frac=0.04, pv_frac=0.12,
                       load_mw=(2.0, 12.0), pv_q_mvar=60.0, rating_mva=250.0, gen_share=0.9) -> Grid:
    """Irregular transmission-like test grid (NOT a reference generator): a random spanning tree plus
    `extra_edge_frac * n` chords (average degree about 2.7 like the PEGASE cases), a share of the branches being
    transformers with off-nominal ratio and a few phase shifters (Y12 != Y21, src/branch.jl:297-310), line
    charging, bus shunts, loads on 70 % of the buses and regulating generators with finite Q limits.  Exercises
    the minimum-degree / nested-dissection symbolic code and the asymmetric stamps on a non-mesh topology."""
    rng = np.random.Generator(np.random.PCG64(seed))
    gb = GridBuilder(f"rand{n}_{seed}", 100.0)
    for i in range(n):
        gb.add_bus(f"N{i + 1}", vm_pu=1.0, va_deg=0.0)
    edges = set()
    for i in range(1, n):                       # random recursive tree with short-range attachment
        j = int(rng.integers(max(0, i - 12), i))
        edges.add((j, i))
    target = n - 1 + int(extra_edge_frac * n)
    while len(edges) < target:
        i = int(rng.integers(0, n))
        j = int(np.clip(i + rng.integers(-25, 26), 0, n - 1))
        if i != j and (min(i, j), max(i, j)) not in edges:
            edges.add((min(i, j), max(i, j)))
    for (i, j) in sorted(edges):
        x = float(rng.uniform(0.01, 0.08))
        r = x * float(rng.uniform(0.1, 0.4))
        u = rng.random()
        if u < shifter_frac:
            gb.add_pi_line(i, j, 0.1 * r, x, 0.0, 0.0, rating=rating_mva, ratio=float(rng.uniform(0.97, 1.03)),
                           shift_deg=float(rng.uniform(-3.0, 3.0)))
        elif u < shifter_frac + trafo_frac:
            gb.add_pi_line(i, j, 0.1 * r, x, 0.0, 0.0, rating=rating_mva, ratio=float(rng.uniform(0.94, 1.06)))
        else:
            gb.add_pi_line(i, j, r, x, float(rng.uniform(0.0, 0.04)), 0.0, rating=rating_mva)
    loads = rng.random(n) < 0.7
    total = 0.0
    for i in range(1, n):
        if loads[i]:
            p = float(rng.uniform(*load_mw))
            gb.add_prosumer(i, "ENERGYCONSUMER", p=p, q=0.3 * p)
            total += p
        if rng.random() < 0.05:
            gb.add_shunt(i, 0.0, float(rng.uniform(2.0, 8.0)))
    pv = [i for i in range(1, n) if rng.random() < pv_frac]
    share = gen_share * total / max(1, len(pv))   # large n: use ~1.01 (the slack sits at one end of a long chain)
    for i in pv:
        gb.add_prosumer(i, "GENERATOR", p=share, q=0.0, vm_pu=float(rng.choice([1.0, 1.01, 1.02])), qMin=-pv_q_mvar,
                        qMax=pv_q_mvar)
    gb.add_prosumer(0, "EXTERNALNETWORKINJECTION", p=0.0, q=0.0, vm_pu=1.02, va_deg=0.0, reference=True)
    return gb.build()
