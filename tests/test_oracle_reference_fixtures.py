"""Pins the CPU oracle against every known-answer fixture the reference's own tests hold for the
hot path (SURVEY.md 8c list; file:line given per test) and against the published IEEE-14
solution.  CPU only."""
import math

import numpy as np
import scipy.sparse as sp

import sparlectra_oracle as O
from sblx import grid as G

PQ, PV, SL = O.PQ, O.PV, O.SLACK


def _fixture_4bus():
    # test/test_factorized_linear_solver.jl:121-139
    y12 = 1.0 / (0.01 + 0.1j); y23 = 1.0 / (0.02 + 0.15j); y34 = 1.0 / (0.015 + 0.12j); y14 = 1.0 / (0.03 + 0.2j)
    Y = sp.csc_matrix(np.array([[y12 + y14, -y12, 0, -y14], [-y12, y12 + y23, -y23, 0], [0, -y23, y23 + y34, -y34],
                                [-y14, 0, -y34, y14 + y34]], dtype=complex))
    V = np.array([1.0, 0.98 + 0.02j, 1.01 - 0.01j, 0.99 + 0.03j])
    types = np.array([PQ, PQ, PV, PQ])
    return Y, V, types


def test_jacobian_4bus_fixture_three_paths_agree_and_match_finite_differences():
    """(i) 4-bus Jacobian fixture: structural sparse build == default build == dense Wirtinger twin
    (rectangular_jacobian_builders.jl:446-552 keeps the same dense diagnostic twin) == finite differences
    of mismatch_rectangular."""
    Y, V, types = _fixture_4bus()
    slack = 0
    vset = np.abs(V)
    for t in (types, np.array([PQ, PQ, PQ, PQ])):
        Js = O.jacobian_sparse(Y, V, t, slack, structural_pattern=True).toarray()
        Jd = O.jacobian_sparse(Y, V, t, slack, structural_pattern=False).toarray()
        Jf = O.jacobian_sparse_fast(Y, V, t, slack).toarray()
        Jw = O.jacobian_dense_wirtinger(Y, V, t, slack)
        assert np.allclose(Js, Jd, atol=0, rtol=0)
        assert np.max(np.abs(Js - Jw)) < 1e-12
        assert np.max(np.abs(Js - Jf)) < 1e-12
        S = np.zeros(4, complex)
        F0 = O.mismatch_rectangular(Y, V, S, t, vset, slack)
        n = 4
        Jfd = np.zeros_like(Js)
        eps = 1e-7
        ns = [1, 2, 3]
        for k, b in enumerate(ns):
            for part, col in ((1.0, k), (1j, (n - 1) + k)):
                Vp = V.copy(); Vp[b] += eps * part
                Vm = V.copy(); Vm[b] -= eps * part
                Jfd[:, col] = (O.mismatch_rectangular(Y, Vp, S, t, vset, slack) - O.mismatch_rectangular(Y, Vm, S, t, vset, slack)) / (2 * eps)
        assert np.max(np.abs(Js - Jfd)) < 1e-6
        assert F0.shape == (6,)


def test_two_bus_mismatch_fixture():
    """(ii) test/testgrid.jl:420-433: Y=[-10i 10i;10i -10i], S=[0,-1-0.2i], flat V: F = calc - spec."""
    Y = sp.csc_matrix(np.array([[-10j, 10j], [10j, -10j]]))
    V = np.array([1.0 + 0j, 1.0 + 0j])
    S = np.array([0j, -1.0 - 0.2j])
    F = O.mismatch_rectangular(Y, V, S, np.array([SL, PQ]), np.ones(2), 0)
    assert np.allclose(F, [1.0, 0.2])
    full = O.max_abs(O.mismatch_rectangular(Y, np.array([1.0 + 0j, 1.0 - 1.0j]), S, np.array([SL, PQ]), np.ones(2), 0))
    assert full > 1.0      # the oversized step of the autodamp fixture really is worse


def test_three_bus_sparse_jacobian_fixture():
    """(iii) test/testgrid.jl:2740-2755."""
    Y = sp.csc_matrix((np.array([8j, -8j, -8j, 14j, -6j, -6j, 6j]), ([0, 0, 1, 1, 1, 2, 2], [0, 1, 0, 1, 2, 1, 2])), shape=(3, 3))
    V = np.ones(3, complex)
    J = O.jacobian_sparse(Y, V, np.array([SL, PQ, PQ]), 0)
    assert sp.issparse(J) and J.shape == (4, 4)
    assert np.max(np.abs(J.toarray() - O.jacobian_dense_wirtinger(Y, V, np.array([SL, PQ, PQ]), 0))) < 1e-13


def test_pv_row_moves_to_setpoint_after_one_step():
    """(iv) test/test_pv_voltage_residuals.jl:41-53,93-98: residual < 0 and |V_pv| ~ vset to 1e-5 after one step."""
    g = G.pv_regression_net(1.05)
    Y = O.create_ybus(g)
    S = O.complex_svec(g)
    types = np.array([SL, PQ, PV])
    V0, slack = O.initial_vrect(g, g.vm0, g.va0_deg, types, flatstart=True)
    V0[2] = 1.0 + 0j
    vset = np.array([1.0, 1.0, 1.05])
    F0 = O.mismatch_rectangular(Y, V0, S, types, vset, slack)
    J = O.jacobian_sparse(Y, V0, types, slack)
    dx = O.solve_linear(J, -F0)
    V1 = V0.copy()
    V1[1:] += dx[:2] + 1j * dx[2:]
    assert F0[3] < 0.0
    assert abs(abs(V1[2]) - 1.05) < abs(1.0 - 1.05)
    assert abs(abs(V1[2]) - 1.05) < 1e-5
    r = O.runpf(g, g.vm0, g.va0_deg, (), maxiter=40, tol=1e-9)
    assert r.converged and abs(abs(r.V[2]) - 1.05) < 1e-7 and r.types[2] == PV      # :100-104


def test_three_bus_qlimit_known_answers():
    """(v) test/testgrid.jl:1142-1184,2520-2582; test_solver_interface.jl:310-322,1228-1240."""
    g = G.test3bus(qlim_min=-15.0, qlim_max=15.0)
    r = O.runpf(g, g.vm0, g.va0_deg, ())
    assert r.converged and r.types[1] == PQ and r.n_events == 1
    # qlim above the 33.2 MVAr need: no hit
    g = G.test3bus(qlim_min=-40.0, qlim_max=40.0)
    r = O.runpf(g, g.vm0, g.va0_deg, ())
    assert r.converged and r.types[1] == PV and r.n_events == 0
    # locked PV with +-1 MVAr must be rejected (erg == 1)
    g = G.test3bus(qlim_min=-1.0, qlim_max=1.0)
    r = O.runpf(g, g.vm0, g.va0_deg, (), lock=(1,))
    assert not r.converged and r.reason == O.R_REMAINING_PV_Q
    # same limits unlocked: switches and converges at tol 1e-6 (test_factorized_linear_solver.jl:79-87)
    r = O.runpf(g, g.vm0, g.va0_deg, (), tol=1e-6, maxiter=20)
    assert r.converged and r.types[1] == PQ
    # start-iter gating: qlimit_start_iter=10, 3 iterations at tol 1e-12 -> erg 1, still PV, no event (testgrid.jl:2570-2575)
    g = G.test3bus(qlim_min=-15.0, qlim_max=15.0)
    r = O.runpf(g, g.vm0, g.va0_deg, (), maxiter=3, tol=1e-12, qlimit_start_iter=10)
    assert not r.converged and r.types[1] == PV and r.n_events == 0
    # power balance: total bus power == total branch losses to 1e-6 (testgrid.jl:1170-1177)
    g = G.test3bus(cooldown=2, hyst_pu=0.01, qlim_min=-15.0, qlim_max=15.0)
    r = O.runpf(g, g.vm0, g.va0_deg, (), maxiter=50, tol=1e-12)
    assert r.converged
    sf, st = O.branch_flows(g, r.V)
    Sbus = O.calc_injections(O.create_ybus(g), r.V)
    assert abs(Sbus.sum() - (sf + st).sum()) < 1e-6 / 100.0


def test_active_set_hysteresis_script():
    """(vi) test/testgrid.jl:2584-2679: qreq 1.05, 1.11, 0.0, 1.11, 0.0 with hyst = 0.10."""
    st = O.QLimitState()
    types = [PV]
    kw = dict(qmin=np.array([-1.0]), qmax=np.array([1.0]), pv_orig=np.array([True]), q_hyst_pu=0.10, cooldown_iters=0,
              is_pv=lambda b: types[b] == PV, make_pq=lambda b, c, s: types.__setitem__(b, PQ),
              make_pv=lambda b: types.__setitem__(b, PV))
    ch, re = O.active_set_q_limits(st, 1, 1, get_qreq=lambda b: 1.05, allow_reenable=False, **kw)
    assert not ch and types[0] == PV and not st.log                      # delayed
    ch, re = O.active_set_q_limits(st, 2, 1, get_qreq=lambda b: 1.11, allow_reenable=False, **kw)
    assert ch and types[0] == PQ and len(st.log) == 1                    # first switch
    ch, re = O.active_set_q_limits(st, 3, 1, get_qreq=lambda b: 0.0, allow_reenable=True, **kw)
    assert not ch and re and types[0] == PV and not st.events            # first re-enable
    ch, re = O.active_set_q_limits(st, 4, 1, get_qreq=lambda b: 1.11, allow_reenable=False, **kw)
    assert ch and types[0] == PQ and len(st.log) == 2                    # second switch
    ch, re = O.active_set_q_limits(st, 5, 1, get_qreq=lambda b: 0.0, allow_reenable=True, **kw)
    assert not ch and not re and types[0] == PQ and len(st.log) == 2     # second re-enable blocked


def test_islanding_chains():
    """(vii) test/test_contingency.jl:97-139."""
    g = G.island_chain(with_pv=False)
    res = O.run_contingencies(g, [[1]])
    assert not res[0].converged and res[0].error == "islanded without reference" and res[0].island_count >= 2
    g = G.island_chain(with_pv=True)
    res = O.run_contingencies(g, [[1]])
    assert res[0].converged and res[0].island_count == 2


def test_divergence_is_reported():
    """(viii) test/test_contingency.jl:141-160."""
    g = G.diverge_ring()
    res = O.run_contingencies(g, [[2], [99]])
    assert not res[0].converged and res[0].error is not None
    assert not res[1].converged and "unknown branch" in res[1].error


def test_case14_n1_counts_and_published_solution():
    """(ix) test/test_contingency.jl:31-37,82 + the published MATPOWER case14 solution."""
    g = G.case14()
    cases = O.generate_n1_branches(g)
    assert len(cases) == 20 and int((g.br_ratio == 0.0).sum()) == 17
    r = O.runpf(g, g.vm0, g.va0_deg, ())
    assert r.converged
    vm_pub = [1.060, 1.045, 1.010, 1.018, 1.020, 1.070, 1.062, 1.090, 1.056, 1.051, 1.057, 1.055, 1.050, 1.036]
    va_pub = [0.0, -4.98, -12.72, -10.31, -8.77, -14.22, -13.36, -13.36, -14.94, -15.10, -14.79, -15.08, -15.16, -16.03]
    assert np.max(np.abs(np.abs(r.V) - vm_pub)) < 6e-4
    assert np.max(np.abs(np.degrees(np.angle(r.V)) - va_pub)) < 1.1e-2
    res = O.run_contingencies(g, cases)
    assert sum(x.converged for x in res) >= 19
    # flat start reaches the same solution (runpf! flatstart path)
    r2 = O.runpf(g, np.ones(14), np.zeros(14), (), flatstart=True)
    assert r2.converged and np.max(np.abs(r2.V - r.V)) < 1e-8


def test_fast_and_loop_paths_are_the_same_algorithm():
    g = G.warmup_case118()
    a = O.runpf(g, g.vm0, g.va0_deg, [5], fast=True)
    b = O.runpf(g, g.vm0, g.va0_deg, [5], fast=False)
    assert a.converged and b.converged and a.iters == b.iters
    assert np.max(np.abs(a.V - b.V)) < 1e-12


def test_singular_network_returns_nonconvergence():
    """test/testgrid.jl:2757-2765: a load bus with no branch -> erg == 1."""
    gb = G.GridBuilder("singular", 100.0)
    gb.add_bus("Slack"); gb.add_bus("Load"); gb.add_bus("X")
    gb.add_prosumer("Slack", "EXTERNALNETWORKINJECTION", vm_pu=1.0, va_deg=0.0, reference=True)
    gb.add_prosumer("Load", "ENERGYCONSUMER", p=10.0, q=5.0)
    gb.add_pi_line("Slack", "X", 0.01, 0.1)
    g = gb.build()
    Y = O.create_ybus(g)
    types = np.array([SL, PQ, PQ])
    V0 = np.ones(3, complex)
    r = O.run_nr(Y, V0, O.complex_svec(g), types, np.ones(3), 0, *O.q_limits_pu(g), O.qload_pu(g), maxiter=5)
    assert not r.converged


def test_warmup118_rule_matches_shipped_file_when_available():
    """The regenerated ladder equals data/webui/warmup_case118.jl (only checkable where the reference tree exists)."""
    import os
    import re
    path = "/root/reference/data/webui/warmup_case118.jl"
    if not os.path.exists(path):
        import pytest
        pytest.skip("reference tree not present")
    txt = open(path).read()
    def block(name):
        m = re.search(name + r" = \[(.*?)\],", txt, re.S)
        rows = [r.strip().rstrip(";") for r in m.group(1).strip().splitlines() if r.strip()]
        return np.array([[float(x) for x in r.split()] for r in rows])
    bus, gen, br = block("bus"), block("gen"), block("branch")
    g = G.warmup_case118()
    assert g.n == len(bus) and g.nbranch == len(br)
    assert np.array_equal(g.br_from + 1, br[:, 0].astype(int)) and np.array_equal(g.br_to + 1, br[:, 1].astype(int))
    assert np.allclose(g.br_r, br[:, 2]) and np.allclose(g.br_x, br[:, 3]) and np.allclose(g.br_b, br[:, 4])
    S = O.complex_svec(g)
    pd = np.zeros(g.n); pd[bus[:, 0].astype(int) - 1] = bus[:, 2]
    pg = np.zeros(g.n); pg[gen[:, 0].astype(int) - 1] = gen[:, 1]
    assert np.allclose(S.real * 100.0, pg - pd)
    assert set(np.flatnonzero(O.bus_types_from_prosumers(g) == PV) + 1) == set(bus[bus[:, 1] == 2, 0].astype(int))
