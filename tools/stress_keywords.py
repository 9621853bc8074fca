"""Randomised parity stress of the NR / Q-limit control logic (GPU): random runpf! keywords (damping, first Q-limit
iteration, hysteresis, cooldown re-enable, tolerance, iteration cap, locked PV buses) on grids with tight Q limits.
python tools/stress_keywords.py [rounds] [seed]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sparlectra.jl_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import sparlectra_oracle as O
from sblx import api, grid as G, model as M
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 21)
only = int(sys.argv[3]) if len(sys.argv) > 3 else None      # replay one round with diagnostics
bad = total = events = nonconv = 0
for r in range(rounds):
    if rng.random() < 0.5:
        g = G.tiled_grid(rows=int(rng.integers(5, 16)), cols=int(rng.integers(5, 16)), augment=True, pv_every=int(rng.integers(2, 6)),
                         pv_q_mvar=float(rng.uniform(8, 30)))
    else:
        g = G.random_meshed_grid(int(rng.integers(40, 300)), int(rng.integers(1, 10**6)), pv_q_mvar=float(rng.uniform(15, 70)),
                                 pv_frac=float(rng.uniform(0.1, 0.3)), gen_share=float(rng.uniform(0.9, 1.0)))
    g.q_hyst_pu = float(rng.choice([0.0, 0.0, 0.01, 0.05]))
    g.cooldown_iters = int(rng.choice([0, 0, 1, 3]))
    # knife edges by construction are left out: a clamped bus sits EXACTLY on its limit, so re-enabling with a zero
    # hysteresis band, or testing the limits on the warm-start iterate itself (first Q-limit iteration 1), decides
    # on rounding noise (SURVEY 8d "edge-flip scenarios")
    if g.cooldown_iters > 0 and g.q_hyst_pu == 0.0:
        g.q_hyst_pu = 0.01
    okw = dict(damp=float(rng.choice([1.0, 1.0, 0.8, 0.6])), qlimit_start_iter=int(rng.integers(2, 5)), tol=float(rng.choice([1e-8, 1e-6, 1e-9])),
               maxiter=int(rng.choice([30, 30, 12, 8])), qlimits_enabled=bool(rng.random() < 0.85))
    ekw = dict(damp=okw["damp"], qlimit_start_iter=okw["qlimit_start_iter"], tol=okw["tol"], max_iter=okw["maxiter"],
               qlimits_enabled=int(okw["qlimits_enabled"]), cooldown_iters=g.cooldown_iters, q_hyst_pu=g.q_hyst_pu)
    a0 = M.build_arrays(g)
    pv = np.flatnonzero(a0.bus_type == M.BUS_PV)
    lock = sorted(int(x) for x in rng.choice(pv, size=min(len(pv), int(rng.integers(0, 3))), replace=False)) if len(pv) else []
    cases = [[int(k)] for k in rng.choice(g.nbranch, size=min(16, g.nbranch), replace=False)] + [[]]
    if only is not None and r != only:
        continue
    vm, va, flat, base = O.solve_base(g, **okw)
    a = M.build_arrays(g, vm=vm, va_deg=va, flatstart=flat)
    eng = api.Engine(a, **ekw)
    if lock:
        eng.set_locked_pv(lock)
    raw = eng.solve_outages(cases)
    eng.close()
    nb = 0
    for i, c in enumerate(cases):
        o = O.runpf(g, vm, va, c, flatstart=flat, lock=lock, **okw)
        total += 1
        nonconv += not o.converged
        ok = bool(raw["converged"][i]) == o.converged and (o.reason == O.R_ISLANDED_NO_REF or int(raw["iterations"][i]) == o.iters)
        if ok and o.converged:
            ok = np.max(np.abs(raw["V"][i] - o.V)) <= 1e-8 and np.array_equal(raw["bus_type_final"][i], o.types.astype(np.uint8))
        events += int(raw["n_switch_events"][i])
        if not ok:
            nb += 1
            print("   MISMATCH", c, "gpu", bool(raw["converged"][i]), int(raw["iterations"][i]), int(raw["status"][i]), "oracle", o.converged, o.iters, o.reason)
            if only is not None:
                Y = O.create_ybus(g, c)
                qmin, qmax = O.q_limits_pu(g)
                ql = O.qload_pu(g)
                for tag, V, ty in (("gpu", raw["V"][i], raw["bus_type_final"][i]), ("oracle", o.V, o.types)):
                    q = O.calc_injections(Y, V).imag + ql
                    pvb = np.flatnonzero(np.asarray(ty) == 2)
                    viol = [(int(b), float(q[b]), float(qmin[b]), float(qmax[b])) for b in pvb if q[b] < qmin[b] - okw["tol"] or q[b] > qmax[b] + okw["tol"]]
                    print("      ", tag, "PV buses", len(pvb), "violating", viol[:6], "locked types", [int(ty[b]) for b in lock], "events", int(raw["n_switch_events"][i]) if tag == "gpu" else o.n_events)
                print("       max|dV|", float(np.max(np.abs(raw["V"][i] - o.V))))
    bad += nb
    print(r, g.name, g.n, {**okw, "hyst": g.q_hyst_pu, "cooldown": g.cooldown_iters, "lock": lock}, "base", base.converged, "->", "OK" if nb == 0 else f"{nb} BAD", flush=True)
print(f"scenarios {total}, not converged {nonconv}, switch events {events}, TOTAL BAD {bad}")
sys.exit(1 if bad else 0)
