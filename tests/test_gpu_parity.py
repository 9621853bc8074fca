"""GPU parity tests: the CUDA engine (through the C ABI) against the CPU oracle on the same
inputs.  Bar (north_star): converged flag, final active set and iteration count identical;
bus voltages within 1e-8 p.u. (float64 path; tolerance written here)."""
import math

import numpy as np
import pytest
import scipy.sparse as sp

import sparlectra_oracle as O
from sblx import api, grid as G, model as M

pytestmark = pytest.mark.gpu
VTOL = 1e-8


def _engine(g, vm=None, va=None, **kw):
    a = M.build_arrays(g, vm=vm, va_deg=va)
    return api.Engine(a, **kw), a


def _oracle_arrays(g, removed=()):
    iso = O.mark_isolated(g, removed)
    types = O.bus_types_from_prosumers(g, iso)
    Y = O.create_ybus(g, removed)
    return Y, types, iso


def _jac_from_blocks(eng, J, n, slack):
    rp, col = eng.pattern()
    ns = [i for i in range(n) if i != slack]
    pos = {b: k for k, b in enumerate(ns)}
    m = n - 1
    R, C, X = [], [], []
    for a in range(m):
        for k in range(rp[a], rp[a + 1]):
            b = pos[int(col[k])]
            blk = J[k]
            R += [2 * a, 2 * a, 2 * a + 1, 2 * a + 1]
            C += [b, (n - 1) + b, b, (n - 1) + b]
            X += [blk[0, 0], blk[0, 1], blk[1, 0], blk[1, 1]]
    return sp.coo_matrix((X, (R, C)), shape=(2 * m, 2 * m)).tocsc()


@pytest.mark.parametrize("case", ["case14", "warmup118", "tiled12", "test3bus"])
def test_newton_step_kernels_match_oracle(case):
    """K1 (mismatch), K2 (Jacobian fill), K3+K4 (LU + substitution) on one state."""
    g = {"case14": G.case14, "warmup118": G.warmup_case118, "test3bus": lambda: G.test3bus(qlim_min=-15.0, qlim_max=15.0),
         "tiled12": lambda: G.tiled_grid(rows=12, cols=12, augment=True, pv_every=4)}[case]()
    eng, a = _engine(g)
    rng = np.random.default_rng(7)
    n = g.n
    V = a.v0 * (1.0 + 0.02 * rng.standard_normal(n)) * np.exp(1j * 0.05 * rng.standard_normal(n))
    V[a.slack] = a.v0[a.slack]
    for removed in ([], [1], [0, 3]) if g.nbranch > 3 else ([], [1]):
        F, J, dx, mm = eng.newton_step(V, removed)
        Y, types, iso = _oracle_arrays(g, removed)
        assert not iso
        S = O.complex_svec(g)
        vset = O.voltage_setpoints(g, g.vm0)
        Fo = O.mismatch_rectangular(Y, V, S, types, vset, a.slack)
        assert np.max(np.abs(F - Fo)) <= 1e-12 * max(1.0, np.max(np.abs(Fo)))
        assert abs(mm - np.max(np.abs(Fo))) <= 1e-12 * max(1.0, mm)
        Jo = O.jacobian_sparse(Y, V, types, a.slack, structural_pattern=True)
        Jg = _jac_from_blocks(eng, J, n, a.slack)
        assert abs(Jg - Jo).max() <= 1e-11 * max(1.0, abs(Jo).max())
        dxo = O.solve_linear(Jo, -Fo)
        # dx layout of the hook: (dVr, dVi) per non-slack bus
        dvr, dvi = dx[0::2], dx[1::2]
        ref = np.concatenate([dvr, dvi])
        assert np.max(np.abs(ref - dxo)) <= 1e-9 * max(1.0, np.max(np.abs(dxo)))
        # residual of the GPU solve against the oracle Jacobian
        assert np.max(np.abs(Jo @ ref + Fo)) <= 1e-10 * max(1.0, np.max(np.abs(Fo)))
    eng.close()


def _compare_batch(g, cases, raw, oracle_results, label, allow_unsupported=True):
    nflip = 0
    for i, (c, o) in enumerate(zip(cases, oracle_results)):
        st = int(raw["status"][i])
        assert st != 9, (label, i, c)      # island-wise solves are supported
        assert bool(raw["converged"][i]) == o.converged, (label, i, c, st, o.reason)
        assert int(raw["island_count"][i]) == o.island_count, (label, i, c)
        if o.reason == O.R_ISLANDED_NO_REF:
            assert st == 7
            continue
        assert int(raw["iterations"][i]) == o.iterations, (label, i, c, raw["iterations"][i], o.iterations)
        if o.converged:
            assert st == 0
            dv = np.max(np.abs(raw["V"][i] - o.V))
            assert dv <= VTOL, (label, i, c, dv)
            tg = raw["bus_type_final"][i]
            assert np.array_equal(tg, o.types.astype(np.uint8)), (label, i, c)
            assert abs(raw["vmin_pu"][i] - o.min_vm_pu) <= 1e-9
            assert abs(raw["vmax_pu"][i] - o.max_vm_pu) <= 1e-9
            if math.isnan(o.max_branch_loading_pct):
                assert math.isnan(raw["max_loading_pct"][i])
            else:
                assert abs(raw["max_loading_pct"][i] - o.max_branch_loading_pct) <= 1e-6
        else:
            assert st == o.reason, (label, i, c, st, o.reason)
    return nflip


def _run_both(g, cases, **opts):
    vm, va, flat, base = O.solve_base(g)
    assert base.converged
    a = M.build_arrays(g, vm=vm, va_deg=va)
    eng = api.Engine(a, cooldown_iters=g.cooldown_iters, q_hyst_pu=g.q_hyst_pu, **opts)
    raw = eng.solve_outages(cases)
    stats = eng.stats()
    eng.close()
    oracle = O.run_contingencies(g, cases)
    return raw, oracle, stats


def test_case14_base_solution_matches_published_and_oracle():
    g = G.case14()
    eng, a = _engine(g)
    V, conv, it, res, st = eng.solve_one(None)
    eng.close()
    r = O.runpf(g, g.vm0, g.va0_deg, ())
    assert conv and st == 0 and it == r.iters
    assert np.max(np.abs(V - r.V)) <= VTOL
    # published MATPOWER solution of case14 (3 decimals)
    vm_pub = [1.060, 1.045, 1.010, 1.018, 1.020, 1.070, 1.062, 1.090, 1.056, 1.051, 1.057, 1.055, 1.050, 1.036]
    assert np.max(np.abs(np.abs(V) - vm_pub)) < 6e-4


def test_case14_full_n1_parity():
    g = G.case14()
    cases = O.generate_n1_branches(g)
    raw, oracle, _ = _run_both(g, cases)
    assert len(cases) == 20
    _compare_batch(g, cases, raw, oracle, "case14")
    assert int(raw["converged"].sum()) >= 19          # test/test_contingency.jl:82


def test_warmup118_full_n1_parity():
    g = G.warmup_case118()
    cases = O.generate_n1_branches(g)
    raw, oracle, _ = _run_both(g, cases)
    _compare_batch(g, cases, raw, oracle, "warmup118")


def test_tiled_grid_n1_n2_parity_with_qlimits():
    g = G.tiled_grid(rows=14, cols=14, augment=True, pv_every=4, pv_q_mvar=24.0, rating_mva=60.0)
    n1 = O.generate_n1_branches(g)
    rng = np.random.default_rng(11)
    n2 = [sorted(rng.choice(g.nbranch, size=2, replace=False).tolist()) for _ in range(120)]
    # corner bus (1, cols) has degree 2: taking both branches isolates it
    corner = 13
    inc = [k for k in range(g.nbranch) if g.br_from[k] == corner or g.br_to[k] == corner]
    cases = n1 + n2 + [inc]
    raw, oracle, _ = _run_both(g, cases)
    _compare_batch(g, cases, raw, oracle, "tiled14")
    assert raw["n_switch_events"].sum() > 0, "augmentation must exercise PV->PQ switching"
    assert raw["bus_type_final"][-1][corner] == 0       # isolated bus reported as such


@pytest.mark.parametrize("qlim,expect_pq", [(15.0, True), (40.0, False)])
def test_three_bus_qlimit_fixture(qlim, expect_pq):
    """test/testgrid.jl:1142-1184 + test_solver_interface.jl:1228-1240: with +-15 MVAr STATION1 ends PQ
    with exactly one :max event; with a limit above 33.2 MVAr it stays PV."""
    g = G.test3bus(qlim_min=-qlim, qlim_max=qlim)
    eng, a = _engine(g)
    raw = eng.solve_outages([[]])
    eng.close()
    r = O.runpf(g, g.vm0, g.va0_deg, ())
    assert raw["converged"][0] == 1 and r.converged
    assert int(raw["iterations"][0]) == r.iters
    assert (raw["bus_type_final"][0][1] == M.BUS_PQ) == expect_pq
    assert int(raw["n_switch_events"][0]) == (1 if expect_pq else 0)
    assert np.max(np.abs(raw["V"][0] - r.V)) <= VTOL


def test_three_bus_locked_pv_is_rejected():
    """test_solver_interface.jl:310-322: +-1 MVAr with lock_pv_to_pq_buses=[2] must return erg == 1."""
    g = G.test3bus(qlim_min=-1.0, qlim_max=1.0)
    eng, a = _engine(g)
    eng.set_locked_pv([1])
    raw = eng.solve_outages([[]])
    eng.close()
    r = O.runpf(g, g.vm0, g.va0_deg, (), lock=(1,))
    assert not r.converged and raw["converged"][0] == 0
    assert int(raw["status"][0]) == r.reason == O.R_REMAINING_PV_Q


def test_islanding_and_divergence_are_reported_not_thrown():
    """test/test_contingency.jl:97-160."""
    g = G.island_chain(with_pv=False)
    res = api.runContingenciesBatched(g, [api.ContingencyCase("B_ACL_2_3")])
    assert not res[0].converged and res[0].error == "islanded without reference" and res[0].island_count >= 2
    g = G.diverge_ring()
    res = api.runContingenciesBatched(g, [api.ContingencyCase("B_ACL_1_3"), api.ContingencyCase("NO_SUCH_BRANCH")])
    assert not res[0].converged and res[0].error is not None
    assert not res[1].converged and "unknown branch" in res[1].error
    o = O.run_contingencies(g, [[2]])
    assert not o[0].converged


def test_island_with_pv_is_promoted_and_solved_island_wise():
    """test/test_contingency.jl:118-139: cutting B-C leaves {C, D} with a PV generator at C; the reference promotes it to
    the island's angle reference and solves both islands (island_count == 2, iterations summed)."""
    g = G.island_chain(with_pv=True)
    res, raw = api.runContingenciesBatched(g, [api.ContingencyCase("B_ACL_2_3")], return_raw=True)
    o = O.run_contingencies(g, [[1]])[0]
    assert o.converged and o.island_count == 2
    assert res[0].converged and res[0].island_count == 2 and res[0].iterations == o.iterations
    assert np.max(np.abs(raw["V"][0] - o.V)) <= VTOL
    assert np.array_equal(raw["bus_type_final"][0], o.types.astype(np.uint8))
    assert abs(res[0].min_vm_pu - o.min_vm_pu) <= 1e-9 and abs(res[0].max_vm_pu - o.max_vm_pu) <= 1e-9


def test_ladder_split_into_two_referenced_islands():
    """warm-up 118 ladder: taking both rails out between buses 15/16 and 74/75 splits it into two islands, the right one
    without the slack: its lowest PV bus (21) becomes the reference (ac_islands.jl:228-234)."""
    g = G.warmup_case118()
    ka = [k for k in range(g.nbranch) if (g.br_from[k], g.br_to[k]) == (14, 15)][0]
    kb = [k for k in range(g.nbranch) if (g.br_from[k], g.br_to[k]) == (73, 74)][0]
    cases = [[ka, kb], [ka], [kb], [3, 70]]
    raw, oracle, _ = _run_both(g, cases)
    assert oracle[0].island_count == 2 and oracle[0].converged
    assert raw["bus_type_final"][0][20] == 3          # promoted reference reported as Slack, like the island sub-net
    _compare_batch(g, cases, raw, oracle, "ladder-split")


def test_twin_circuits_get_their_own_case():
    g = G.twin_lines()
    g.br_name = ["B_ACL_1_2", "B_ACL_1_2"]
    cases = api.generateN1Branches(g)
    assert len(cases) == 2 and cases[0].element != cases[1].element and "#" in cases[0].element
    res = api.runContingenciesBatched(g, cases)
    assert all(r.converged for r in res)
    assert res[0].min_vm_pu != res[1].min_vm_pu


def test_split_kernel_path_matches_oracle():
    """The experimental split factorisation (k_pivot / k_subst / k_schur, factor_mode=1) must give the same answers."""
    g = G.tiled_grid(rows=12, cols=12, augment=True, pv_every=4, pv_q_mvar=24.0, rating_mva=60.0)
    cases = O.generate_n1_branches(g)[:80]
    raw, oracle, _ = _run_both(g, cases, factor_mode=1)
    _compare_batch(g, cases, raw, oracle, "split")


def test_injection_sweep_parity():
    g = G.tiled_grid(rows=10, cols=10, augment=True, pv_every=3, pv_q_mvar=20.0)
    vm, va, flat, base = O.solve_base(g)
    a = M.build_arrays(g, vm=vm, va_deg=va)
    rng = np.random.default_rng(5)
    B = 40
    S0 = a.s_spec
    ql0 = a.qload
    f = np.clip(rng.normal(1.0, 0.1, size=(B, g.n)), 0.5, 1.5)
    isload = (S0.real < 0)
    S = np.where(isload, S0 * f, S0)
    ql = np.where(isload, ql0 * f, ql0)
    eng = api.Engine(a)
    raw = eng.solve_injections(S, ql)
    eng.close()
    for s in range(B):
        r = O.runpf(g, vm, va, (), s_override=S[s], qload_override=ql[s])
        assert bool(raw["converged"][s]) == r.converged
        assert int(raw["iterations"][s]) == r.iters
        if r.converged:
            assert np.max(np.abs(raw["V"][s] - r.V)) <= VTOL
            assert np.array_equal(raw["bus_type_final"][s], r.types.astype(np.uint8))


def test_solvepf_plugin_roundtrip():
    g = G.pv_regression_net(1.05)
    model = api.buildPfModel(g)
    sol = api.solvePf(api.B200BatchSolver(), model, tol=1e-9, max_iter=40)
    assert sol.converged and api.mismatchInf(model, sol.V) <= 1e-9
    assert abs(abs(sol.V[2]) - 1.05) < 1e-7            # test_pv_voltage_residuals.jl:100-104
    it, status, sol2 = api.runpf_external(g, api.B200BatchSolver(), tol=1e-9)
    assert status == 0


def test_medium_grid_properties_54x54():
    """Size-independent properties at config-3 size: every converged scenario's returned V must
    satisfy the canonical mismatch (oracle function, cheap) and the order of results is the input order."""
    g = G.tiled_grid(rows=54, cols=54, augment=True)
    a0 = M.build_arrays(g)
    eng = api.Engine(a0)
    Vb, conv, it, res, st = eng.solve_one(None)
    assert conv
    a = M.build_arrays(g, vm=np.abs(Vb), va_deg=np.degrees(np.angle(Vb)))
    eng.set_v0(a.v0)
    cases = [[k] for k in range(0, g.nbranch, 17)]
    raw = eng.solve_outages(cases)
    eng.close()
    assert raw["converged"].mean() > 0.95
    S = O.complex_svec(g)
    vset = O.voltage_setpoints(g, np.abs(Vb))
    ql = O.qload_pu(g)
    for i in range(0, len(cases), 25):
        if not raw["converged"][i]:
            continue
        Y = O.create_ybus(g, cases[i])
        t = raw["bus_type_final"][i].astype(np.int64)
        Sx = S.copy()
        # buses switched to PQ carry the clamped Q
        qmin, qmax = O.q_limits_pu(g)
        Sc = O.calc_injections(Y, raw["V"][i])
        sw = (a.bus_type == M.BUS_PV) & (t == M.BUS_PQ)
        Sx[sw] = Sx[sw].real + 1j * Sc[sw].imag
        F = O.mismatch_rectangular_fast(Y, raw["V"][i], Sx, t, vset, a.slack)
        assert np.max(np.abs(F)) <= 1e-8
        qg = Sc.imag + ql
        assert np.all((qg[sw] >= qmax[sw] - 1e-6) | (qg[sw] <= qmin[sw] + 1e-6))


def test_full_size_10k_grid_parity_and_properties():
    """BASELINE config 4/5 grid (100x100 tiled, 10 000 buses, 29 601 branches): a spread of N-1 and N-2 outages.
    Three scenarios are compared against the oracle in full (iterations, active set, voltages <= 1e-8); every
    sampled scenario must satisfy the canonical mismatch evaluated by the oracle's function at the returned V."""
    g = G.tiled_grid(rows=100, cols=100, augment=True, rating_mva=100.0)
    a0 = M.build_arrays(g)
    eng = api.Engine(a0)
    Vb, conv, itb, resb, stb = eng.solve_one(None)
    ob = O.runpf(g, g.vm0, g.va0_deg, ())
    assert conv and ob.converged and itb == ob.iters and np.max(np.abs(Vb - ob.V)) <= VTOL
    vm, va = np.abs(Vb), np.degrees(np.angle(Vb))
    a = M.build_arrays(g, vm=vm, va_deg=va)
    eng.set_v0(a.v0)
    rng = np.random.default_rng(3)
    n1 = [[int(k)] for k in rng.choice(g.nbranch, size=200, replace=False)]
    n2 = [sorted(int(x) for x in rng.choice(g.nbranch, size=2, replace=False)) for _ in range(56)]
    cases = n1 + n2
    raw = eng.solve_outages(cases)
    info = eng.info()
    eng.close()
    assert info["nnz_lu_scalar"] > 2_000_000 and raw["converged"].mean() > 0.98
    for i in (0, 101, 230):
        o = O.runpf(g, vm, va, cases[i])
        assert bool(raw["converged"][i]) == o.converged and int(raw["iterations"][i]) == o.iters
        assert np.max(np.abs(raw["V"][i] - o.V)) <= VTOL
        assert np.array_equal(raw["bus_type_final"][i], o.types.astype(np.uint8))
    S = O.complex_svec(g)
    vset = O.voltage_setpoints(g, vm)
    for i in range(0, len(cases), 16):
        if not raw["converged"][i]:
            continue
        Y = O.create_ybus(g, cases[i])
        t = raw["bus_type_final"][i].astype(np.int64)
        Sc = O.calc_injections(Y, raw["V"][i])
        Sx = S.copy()
        sw = (a.bus_type == M.BUS_PV) & (t == M.BUS_PQ)
        Sx[sw] = Sx[sw].real + 1j * Sc[sw].imag
        t2 = np.where(t == 0, 1, t)
        F = O.mismatch_rectangular_fast(Y, raw["V"][i], Sx, t2, vset, a.slack)
        iso = (t == 0)
        if iso.any():       # isolated buses carry neutral rows
            ns = np.array([b for b in range(g.n) if b != a.slack])
            keep = np.repeat(~iso[ns], 2)
            F = F[keep]
        assert np.max(np.abs(F)) <= 1e-8


def test_full_size_10k_injection_sweep_parity():
    """BASELINE config 4 shape: 100x100 tiled grid, load/injection sweep with independent truncated-normal factors
    per load bus (f ~ N(1, 0.1) clipped to [0.5, 1.5], examples/powerflow/mc_probabilistic_powerflow.jl:28-41),
    warm start from the base solution, Q-limit switching active.  Two scenarios are compared with the oracle in full;
    every sampled scenario must satisfy the canonical mismatch at the returned V and keep switched buses at a limit."""
    g = G.tiled_grid(rows=100, cols=100, augment=True, rating_mva=100.0)
    a0 = M.build_arrays(g)
    eng = api.Engine(a0)
    Vb, conv, *_ = eng.solve_one(None)
    assert conv
    vm, va = np.abs(Vb), np.degrees(np.angle(Vb))
    a = M.build_arrays(g, vm=vm, va_deg=va)
    eng.set_v0(a.v0)
    rng = np.random.default_rng(20260724)
    B = 224          # 35.8 MB of injections: more than one 32 MB chunk of the pinned H2D staging
    f = np.clip(rng.normal(1.0, 0.1, size=(B, g.n)), 0.5, 1.5)
    isload = a.s_spec.real < 0
    S = np.where(isload, a.s_spec * f, a.s_spec)
    ql = np.where(isload, a.qload * f, a.qload)
    raw = eng.solve_injections(S, ql)
    eng.close()
    assert raw["converged"].all()
    assert raw["n_switch_events"].sum() > 0                   # the sweep does exercise PV -> PQ
    for s in (0, 57, 223):
        o = O.runpf(g, vm, va, (), s_override=S[s], qload_override=ql[s])
        assert o.converged and int(raw["iterations"][s]) == o.iters
        assert np.max(np.abs(raw["V"][s] - o.V)) <= VTOL
        assert np.array_equal(raw["bus_type_final"][s], o.types.astype(np.uint8))
    Y = O.create_ybus(g, ())
    vset = O.voltage_setpoints(g, vm)
    qmin, qmax = O.q_limits_pu(g)
    for s in range(0, B, 12):
        t = raw["bus_type_final"][s].astype(np.int64)
        Sc = O.calc_injections(Y, raw["V"][s])
        sw = (a.bus_type == M.BUS_PV) & (t == M.BUS_PQ)
        Sx = S[s].copy()
        Sx[sw] = Sx[sw].real + 1j * Sc[sw].imag
        F = O.mismatch_rectangular_fast(Y, raw["V"][s], Sx, t, vset, a.slack)
        assert np.max(np.abs(F)) <= 1e-8
        qg = Sc.imag + ql[s]
        assert np.all((qg[sw] >= qmax[sw] - 1e-6) | (qg[sw] <= qmin[sw] + 1e-6))


@pytest.mark.parametrize("seed", [1, 2])
def test_irregular_grid_with_transformers_and_phase_shifters_n1_parity(seed):
    """Non-mesh topology (random tree + chords, average degree 2.7) with off-nominal transformers, phase shifters
    (Y12 != Y21), line charging and bus shunts: full N-1 (radial branches split the grid -> island handling) plus
    random N-2, every scenario compared with the oracle."""
    g = G.random_meshed_grid(300, seed)
    n1 = O.generate_n1_branches(g)
    rng = np.random.default_rng(100 + seed)
    n2 = [sorted(rng.choice(g.nbranch, size=2, replace=False).tolist()) for _ in range(60)]
    cases = n1 + n2
    raw, oracle, _ = _run_both(g, cases)
    _compare_batch(g, cases, raw, oracle, f"rand300_{seed}")
    assert raw["n_switch_events"].sum() > 0
    assert (raw["island_count"] > 1).any()              # radial spurs: some outages island part of the grid


def test_randomised_grids_and_symbolic_options_parity():
    """Random grid shapes x random symbolic options (ordering, amalgamation, panel budget: they select different
    kernel classes and index-list paths), a handful of N-1 / N-2 scenarios each, every one against the oracle
    (the long version is tools/stress.py)."""
    rng = np.random.default_rng(20261020)
    for r in range(8):
        if r % 2 == 0:
            g = G.tiled_grid(rows=int(rng.integers(4, 30)), cols=int(rng.integers(4, 30)), augment=True,
                             pv_every=int(rng.integers(3, 8)), pv_q_mvar=float(rng.uniform(15, 40)))
        else:
            g = G.random_meshed_grid(int(rng.integers(40, 700)), int(rng.integers(1, 10**6)),
                                     extra_edge_frac=float(rng.uniform(0.1, 0.8)), gen_share=float(rng.uniform(0.9, 1.0)))
        opts = dict(ordering=int(rng.integers(0, 3)), relax_small=int(rng.integers(2, 40)),
                    relax_zero_frac=float(rng.uniform(0.0, 0.4)), panel_smem_kb=float(rng.choice([24, 48, 100, 200])))
        vm, va, flat, base = O.solve_base(g)
        eng = api.Engine(M.build_arrays(g, vm=vm, va_deg=va, flatstart=flat), **opts)
        cases = [[int(k)] for k in rng.choice(g.nbranch, size=min(10, g.nbranch), replace=False)] + \
                [sorted(int(x) for x in rng.choice(g.nbranch, size=2, replace=False)) for _ in range(3)]
        raw = eng.solve_outages(cases)
        eng.close()
        for i, c in enumerate(cases):
            o = O.runpf(g, vm, va, c)
            tag = (g.name, opts, c)
            assert bool(raw["converged"][i]) == o.converged and int(raw["iterations"][i]) == o.iters, tag
            if o.converged:
                assert np.max(np.abs(raw["V"][i] - o.V)) <= VTOL, tag
                assert np.array_equal(raw["bus_type_final"][i], o.types.astype(np.uint8)), tag


def test_retry_flat_start_matches_reference_semantics():
    """runContingencies!(...; retry_flat_start = true) (contingency.jl:195-207): a non-converged case gets exactly one
    second attempt from a flat start and reports the summed iterations; converged cases are untouched."""
    for g in (G.case14(), G.diverge_ring()):
        cases = api.generateN1Branches(g)
        plain = api.runContingenciesBatched(g, cases)
        retry = api.runContingenciesBatched(g, cases, retry_flat_start=True)
        oc = O.generate_n1_branches(g)
        want = O.run_contingencies(g, oc, retry_flat_start=True)
        assert any(not r.converged for r in plain)
        for p, r, w in zip(plain, retry, want):
            assert r.converged == w.converged and r.iterations == w.iterations and r.error == w.error
            if p.converged:
                assert r.iterations == p.iterations and r.max_vm_pu == p.max_vm_pu
            elif p.error.startswith("power flow did not converge"):
                assert r.iterations > p.iterations


@pytest.mark.parametrize("variant", ["damped", "no_qlimits", "late_qlimits", "hysteresis_reenable", "loose_tol_short"])
def test_runpf_keyword_variants_parity(variant):
    """The runpf! keywords the contingency path forwards (rectangular_network_solver.jl:1783-1883): damping, Q-limit
    switch, first eligible iteration, hysteresis / cooldown re-enable (limits.jl:527-561), tolerance and iteration cap."""
    g = G.tiled_grid(rows=12, cols=12, augment=True, pv_every=4, pv_q_mvar=22.0, rating_mva=60.0)
    okw, ekw = {}, {}
    if variant == "damped":
        okw, ekw = dict(damp=0.7), dict(damp=0.7)
    elif variant == "no_qlimits":
        okw, ekw = dict(qlimits_enabled=False), dict(qlimits_enabled=0)
    elif variant == "late_qlimits":
        okw, ekw = dict(qlimit_start_iter=4), dict(qlimit_start_iter=4)
    elif variant == "hysteresis_reenable":
        g.q_hyst_pu, g.cooldown_iters = 0.02, 2
    elif variant == "loose_tol_short":
        okw, ekw = dict(tol=1e-5, maxiter=6), dict(tol=1e-5, max_iter=6)
    vm, va, flat, base = O.solve_base(g, **okw)
    a = M.build_arrays(g, vm=vm, va_deg=va, flatstart=flat)
    eng = api.Engine(a, cooldown_iters=g.cooldown_iters, q_hyst_pu=g.q_hyst_pu, **ekw)
    cases = [[k] for k in range(0, g.nbranch, 5)] + [[3, 77], [10, 11]]
    raw = eng.solve_outages(cases)
    eng.close()
    nsw = 0
    for i, c in enumerate(cases):
        o = O.runpf(g, vm, va, c, flatstart=flat, **okw)
        assert bool(raw["converged"][i]) == o.converged and int(raw["iterations"][i]) == o.iters, (variant, c, raw["iterations"][i], o.iters)
        if o.converged:
            assert np.max(np.abs(raw["V"][i] - o.V)) <= VTOL, (variant, c)
            assert np.array_equal(raw["bus_type_final"][i], o.types.astype(np.uint8)), (variant, c)
        nsw += int(raw["n_switch_events"][i])
    assert (nsw == 0) == (variant == "no_qlimits")


def test_injection_sweep_helper_is_chunk_independent():
    """Engine.injection_sweep draws the config-4 factors in chunks while the GPU solves the previous chunk; the
    result must not depend on the chunk size and must equal a direct solve with the same factors."""
    g = G.tiled_grid(rows=9, cols=9, augment=True, pv_every=3, pv_q_mvar=20.0)
    vm, va, flat, base = O.solve_base(g)
    a = M.build_arrays(g, vm=vm, va_deg=va)
    eng = api.Engine(a)
    r1 = eng.injection_sweep(a.s_spec, a.qload, 70, seed=3, chunk=1000)
    r2 = eng.injection_sweep(a.s_spec, a.qload, 70, seed=3, chunk=70)
    rng = np.random.Generator(np.random.PCG64(3))
    f = np.clip(rng.normal(1.0, 0.1, size=(70, g.n)), 0.5, 1.5)
    isload = a.s_spec.real < 0
    r3 = eng.solve_injections(np.where(isload, a.s_spec * f, a.s_spec), np.where(isload, a.qload * f, a.qload))
    eng.close()
    for k in ("converged", "iterations", "n_switch_events", "vmin_pu", "vmax_pu"):
        assert np.array_equal(r1[k], r2[k]) and np.array_equal(r1[k], r3[k]), k
    o = O.runpf(g, vm, va, (), s_override=np.where(isload, a.s_spec * f[5], a.s_spec), qload_override=np.where(isload, a.qload * f[5], a.qload))
    assert o.converged == bool(r1["converged"][5]) and o.iters == int(r1["iterations"][5])


@pytest.mark.parametrize("name,maker", [("case14_n1_oracle.json", lambda: G.case14()),
                                        ("test3bus_q15_n1_oracle.json", lambda: G.test3bus(qlim_min=-15.0, qlim_max=15.0)),
                                        ("island_chain_pv_n1_oracle.json", lambda: G.island_chain(True))])
def test_engine_matches_committed_golden_vectors(name, maker):
    """The engine against the committed vectors of tests/golden (no oracle code on this test's path)."""
    import json, os
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name)))
    g = maker()
    eng = api.Engine(M.build_arrays(g))
    Vb, conv, itb, *_ = eng.solve_one(None)
    assert conv and itb == gold["base_iterations"] and np.max(np.abs(np.abs(Vb) - gold["base_vm"])) <= VTOL
    a = M.build_arrays(g, vm=np.abs(Vb), va_deg=np.degrees(np.angle(Vb)))
    eng.set_v0(a.v0)
    raw = eng.solve_outages([c["removed"] for c in gold["cases"]])
    eng.close()
    for i, c in enumerate(gold["cases"]):
        assert bool(raw["converged"][i]) == c["converged"] and int(raw["island_count"][i]) == c["island_count"]
        if c["error"] != "islanded without reference":
            assert int(raw["iterations"][i]) == c["iterations"]
        if c["converged"]:
            assert np.max(np.abs(np.abs(raw["V"][i]) - c["vm"])) <= VTOL
            assert raw["bus_type_final"][i].tolist() == c["types"]
            assert abs(raw["vmin_pu"][i] - c["vmin"]) <= 1e-9 and abs(raw["vmax_pu"][i] - c["vmax"]) <= 1e-9


def test_malformed_outage_sets_are_reported_per_scenario():
    """More than 8 branches in one scenario, or an id outside the model: status 8 for that scenario, the rest of
    the batch is solved (failures are reported, never thrown: contingency.jl:72-73)."""
    g = G.tiled_grid(rows=6, cols=6, augment=True)
    eng = api.Engine(M.build_arrays(g))
    raw = eng.solve_outages([[0], list(range(9)), [g.nbranch + 5], [], [1, 2]])
    eng.close()
    assert raw["status"].tolist()[1] == 8 and raw["status"].tolist()[2] == 8
    assert not raw["converged"][1] and not raw["converged"][2]
    assert raw["converged"][0] and raw["converged"][3] and raw["converged"][4]


def test_island_wise_solve_with_an_isolated_bus():
    """N-2 that isolates a leaf bus AND splits the rest into two referenced islands (found by tools/stress.py):
    the isolated bus keeps 1 angle 0 / type isolated in the recombined island-wise result."""
    g = G.random_meshed_grid(545, 808968, extra_edge_frac=0.24766309440693596, gen_share=0.9570001040188474)
    cases = [[383, 643], [383], [643]]
    raw, oracle, _ = _run_both(g, cases)
    _compare_batch(g, cases, raw, oracle, "rand545")
    assert raw["island_count"][0] == 2 and raw["bus_type_final"][0][306] == 0 and raw["V"][0][306] == 1.0


@pytest.mark.parametrize("args,kw,case", [
    ((50, 199813), dict(extra_edge_frac=0.2290426133360295, pv_frac=0.20315213184860367, gen_share=0.9051247332382303), [43, 44, 46]),
    ((65, 250666), dict(extra_edge_frac=0.12247042342733547, pv_frac=0.28330260349118397, gen_share=0.9351149305359199), [7, 15, 21, 65]),
])
def test_outage_set_that_isolates_a_cut_bus(args, kw, case):
    """N-3 / N-4 sets that strip a bus of ALL its branches where that bus was the only link between its neighbours
    (found by tools/stress_islands.py): the grid splits although every other removed branch still has connected
    endpoints - the host topology analysis must not take its connected-grid shortcut."""
    g = G.random_meshed_grid(*args, **kw)
    cases = [case, case[:1], case[1:]]
    raw, oracle, _ = _run_both(g, cases)
    _compare_batch(g, cases, raw, oracle, g.name)
    assert raw["island_count"][0] == 2


def test_topology_heavy_outage_sets_parity():
    """Nearly radial small grids with N-2..N-4 sets: most scenarios isolate buses and / or split the grid into several
    islands, with and without a reference (short version of tools/stress_islands.py)."""
    rng = np.random.default_rng(5)
    multi = 0
    for r in range(6):
        g = G.random_meshed_grid(int(rng.integers(20, 120)), int(rng.integers(1, 10**6)), extra_edge_frac=float(rng.uniform(0.02, 0.25)),
                                 pv_frac=float(rng.uniform(0.1, 0.4)), gen_share=float(rng.uniform(0.9, 1.0)))
        cases = [sorted(int(x) for x in rng.choice(g.nbranch, size=int(rng.integers(2, 5)), replace=False)) for _ in range(20)]
        raw, oracle, _ = _run_both(g, cases)
        _compare_batch(g, cases, raw, oracle, g.name)
        multi += int((raw["island_count"] > 1).sum())
    assert multi > 20


def test_locked_pv_list_in_island_wise_solves_addresses_island_numbering():
    """lock_pv_to_pq_buses + islanding: the reference hands the list unchanged to each island sub-net
    (rectangular_network_solver.jl:1991), so an id addresses the island's own bus numbering (ac_islands.jl:329-337)
    and ids beyond the island's size lock nothing (limits.jl:434-438).  Found by tools/stress_keywords.py."""
    g = G.random_meshed_grid(237, 637171, pv_q_mvar=30.0, pv_frac=0.2, gen_share=0.95)
    vm, va, flat, base = O.solve_base(g)
    a = M.build_arrays(g, vm=vm, va_deg=va, flatstart=flat)
    pv = np.flatnonzero(a.bus_type == M.BUS_PV)
    iso_cases = []
    for k in range(g.nbranch):                     # radial branches: their outage splits the grid
        o = O.runpf(g, vm, va, [k], flatstart=flat)
        if o.island_count == 2 and o.reason != O.R_ISLANDED_NO_REF:
            iso_cases.append([k])
        if len(iso_cases) >= 6:
            break
    assert iso_cases
    for lock in ([int(pv[0])], [int(pv[-1])], [int(pv[1]), int(pv[len(pv) // 2])]):
        eng = api.Engine(a)
        eng.set_locked_pv(lock)
        raw = eng.solve_outages(iso_cases + [[]])
        eng.close()
        for i, c in enumerate(iso_cases + [[]]):
            o = O.runpf(g, vm, va, c, flatstart=flat, lock=lock)
            assert bool(raw["converged"][i]) == o.converged, (lock, c)
            if o.converged:
                assert int(raw["iterations"][i]) == o.iters and np.max(np.abs(raw["V"][i] - o.V)) <= VTOL, (lock, c)
                assert np.array_equal(raw["bus_type_final"][i], o.types.astype(np.uint8)), (lock, c)


@pytest.mark.parametrize("shape", ["clique", "star", "chain"])
def test_dense_star_and_chain_topologies_parity(shape):
    """Shapes at the edges of the symbolic code: a complete graph (one dense supernode, panel chains of 16 pivots), a
    star (every branch outage isolates a leaf) and a long chain (deep tree; every outage splits the grid)."""
    rng = np.random.default_rng(9)
    n = {"clique": 36, "star": 60, "chain": 80}[shape]
    gb = G.GridBuilder(shape, 100.0)
    for i in range(n):
        gb.add_bus(f"N{i + 1}")
    if shape == "clique":
        edges = [(i, j) for i in range(n) for j in range(i + 1, n)]
    elif shape == "star":
        edges = [(0, j) for j in range(1, n)]
    else:
        edges = [(i, i + 1) for i in range(n - 1)]
    for (i, j) in edges:
        x = float(rng.uniform(0.01, 0.03)) if shape == "chain" else float(rng.uniform(0.02, 0.08))
        gb.add_pi_line(i, j, 0.2 * x, x, float(rng.uniform(0.0, 0.005 if shape == "chain" else 0.02)), 0.0, rating=400.0)
    for i in range(1, n):
        p = float(rng.uniform(5.0, 30.0)) if shape == "clique" else float(rng.uniform(0.3, 1.0)) if shape == "chain" else float(rng.uniform(1.0, 4.0))
        gb.add_prosumer(i, "ENERGYCONSUMER", p=p, q=0.3 * p)
        if i % 9 == 4:
            gb.add_prosumer(i, "GENERATOR", p=(8.0 if shape == "chain" else 3.0) * p, q=0.0, vm_pu=1.01, qMin=-15.0, qMax=15.0)
    gb.add_prosumer(0, "EXTERNALNETWORKINJECTION", p=0.0, q=0.0, vm_pu=1.02, va_deg=0.0, reference=True)
    g = gb.build()
    cases = O.generate_n1_branches(g)[:: max(1, len(edges) // 60)] + [[0, 1], [2, len(edges) - 1]]
    raw, oracle, _ = _run_both(g, cases)
    _compare_batch(g, cases, raw, oracle, shape)


@pytest.mark.parametrize("slots", [1, 3, 8, 37])
def test_slot_refill_with_few_resident_slots(slots):
    """Continuous batching: far more scenarios than resident slots, with short (4-5 iterations), long (Q-limit
    switching), non-convergent (30 iterations) and island-wise scenarios mixed, so slots are refilled at different
    ticks.  Results must equal the oracle's and must not depend on the number of slots."""
    g = G.tiled_grid(rows=9, cols=9, augment=True, pv_every=3, pv_q_mvar=14.0, rating_mva=60.0)
    rng = np.random.default_rng(77)
    n1 = O.generate_n1_branches(g)
    n2 = [sorted(rng.choice(g.nbranch, size=2, replace=False).tolist()) for _ in range(40)]
    n5 = [sorted(rng.choice(g.nbranch, size=5, replace=False).tolist()) for _ in range(25)]
    corner = 8
    inc = [k for k in range(g.nbranch) if g.br_from[k] == corner or g.br_to[k] == corner]
    cases = n1 + n2 + n5 + [inc]
    raw, oracle, stats = _run_both(g, cases, max_slots=slots)
    _compare_batch(g, cases, raw, oracle, f"refill{slots}")
    assert stats["ticks"] > (len(cases) // max(slots, 1)) * 3


def test_handles_are_independent_and_reusable():
    """Two live handles on different grids with interleaved calls, and one handle reused for batches of very different
    sizes and kinds (outages, injections, single solve): every call returns what a fresh handle returns."""
    g1, g2 = G.case14(), G.tiled_grid(rows=8, cols=8, augment=True, pv_every=3, pv_q_mvar=16.0)
    def fresh(g, cases):
        e = api.Engine(M.build_arrays(g))
        r = e.solve_outages(cases)
        e.close()
        return r
    c1 = [[k] for k in range(g1.nbranch)]
    c2 = [[k] for k in range(0, g2.nbranch, 3)]
    want1, want2 = fresh(g1, c1), fresh(g2, c2)
    e1, e2 = api.Engine(M.build_arrays(g1)), api.Engine(M.build_arrays(g2))
    a2 = M.build_arrays(g2)
    for rep in range(3):
        r2 = e2.solve_outages(c2)
        r1 = e1.solve_outages(c1)
        small = e2.solve_outages(c2[:2])
        inj = e2.solve_injections(np.tile(a2.s_spec, (5, 1)), None)
        one = e2.solve_one(None)
        for k in ("converged", "iterations", "status", "island_count"):
            assert np.array_equal(r1[k], want1[k]) and np.array_equal(r2[k], want2[k]), (rep, k)
            assert np.array_equal(small[k], want2[k][:2]), (rep, k)
        assert np.array_equal(r1["V"], want1["V"]) and np.array_equal(r2["V"], want2["V"])     # bit-identical run to run
        assert inj["converged"].all() and len(set(inj["iterations"].tolist())) == 1
        assert one[1] and one[2] == int(inj["iterations"][0])
    e1.close(); e2.close()


def test_concurrent_callers_on_separate_handles():
    """Two host threads, one handle each (the contract of include/sblx.h), solving at the same time: the results equal
    the serial ones.  A third thread hammering ONE of the handles is serialised by the handle's mutex."""
    import threading
    g1, g2 = G.tiled_grid(rows=10, cols=10, augment=True), G.random_meshed_grid(200, 4)
    jobs = []
    for g in (g1, g2):
        cases = [[k] for k in range(0, g.nbranch, 2)]
        e = api.Engine(M.build_arrays(g))
        jobs.append((e, cases, e.solve_outages(cases)))
    out = {}
    def work(tag, e, cases, reps):
        res = []
        for _ in range(reps):
            res.append(e.solve_outages(cases))
        out[tag] = res
    th = [threading.Thread(target=work, args=(0, jobs[0][0], jobs[0][1], 4)),
          threading.Thread(target=work, args=(1, jobs[1][0], jobs[1][1], 4)),
          threading.Thread(target=work, args=(2, jobs[1][0], jobs[1][1][:7], 6))]
    for t in th: t.start()
    for t in th: t.join()
    for tag, (e, cases, want) in ((0, jobs[0]), (1, jobs[1])):
        for r in out[tag]:
            assert np.array_equal(r["iterations"], want["iterations"]) and np.array_equal(r["V"], want["V"])
    for r in out[2]:
        assert np.array_equal(r["iterations"], jobs[1][2]["iterations"][:7]) and np.array_equal(r["V"], jobs[1][2]["V"][:7])
    for e, _, _ in jobs:
        e.close()


def test_one_based_abi_gives_the_same_results():
    """index_base = 1 is what the Julia glue drives the C ABI with (1-based CSC colptr / rowval, slack, branch
    terminals, outage and lock ids); it must give exactly the results of the 0-based path."""
    g = G.random_meshed_grid(180, 12, pv_q_mvar=30.0)
    vm, va, flat, base = O.solve_base(g)
    a = M.build_arrays(g, vm=vm, va_deg=va, flatstart=flat)
    pv = np.flatnonzero(a.bus_type == M.BUS_PV)
    cases = [[k] for k in range(0, g.nbranch, 3)] + [[0, 5], [g.nbranch - 1, 7], [], [g.nbranch + 3]]
    res = []
    for base_idx in (0, 1):
        e = api.Engine(a, index_base=base_idx)
        e.set_locked_pv([int(pv[0])])
        res.append(e.solve_outages(cases))
        e.close()
    for k in ("converged", "status", "iterations", "island_count", "n_switch_events"):
        assert np.array_equal(res[0][k], res[1][k]), k
    assert np.array_equal(res[0]["V"], res[1]["V"]) and np.array_equal(res[0]["bus_type_final"], res[1]["bus_type_final"])
    assert res[1]["status"][-1] == 8 and res[1]["converged"][-2]
    o = O.runpf(g, vm, va, cases[1], flatstart=flat, lock=[int(pv[0])])
    assert o.converged == bool(res[1]["converged"][1]) and o.iters == int(res[1]["iterations"][1])
