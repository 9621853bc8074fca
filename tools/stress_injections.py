"""Randomised parity stress of the injection-sweep entry (GPU): random grids, strong load / generation perturbations
(some scenarios diverge), per-scenario qload override on and off.  python tools/stress_injections.py [rounds] [seed]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sparlectra.jl_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import sparlectra_oracle as O
from sblx import api, grid as G, model as M
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 31)
bad = total = nonconv = events = 0
for r in range(rounds):
    if rng.random() < 0.5:
        g = G.tiled_grid(rows=int(rng.integers(4, 18)), cols=int(rng.integers(4, 18)), augment=True, pv_every=int(rng.integers(2, 6)),
                         pv_q_mvar=float(rng.uniform(10, 40)))
    else:
        g = G.random_meshed_grid(int(rng.integers(30, 400)), int(rng.integers(1, 10**6)), pv_q_mvar=float(rng.uniform(20, 80)),
                                 gen_share=float(rng.uniform(0.9, 1.0)))
    vm, va, flat, base = O.solve_base(g)
    a = M.build_arrays(g, vm=vm, va_deg=va, flatstart=flat)
    eng = api.Engine(a, ordering=int(rng.integers(0, 3)))
    B = 20
    sigma = float(rng.choice([0.1, 0.3, 0.6]))
    f = np.clip(rng.normal(1.0, sigma, size=(B, g.n)), 0.2, 2.5)
    S = a.s_spec * f
    with_q = bool(rng.random() < 0.6)
    ql = a.qload * f if with_q else None
    raw = eng.solve_injections(S, ql)
    eng.close()
    nb = 0
    for s in range(B):
        o = O.runpf(g, vm, va, (), flatstart=flat, s_override=S[s], qload_override=None if ql is None else ql[s])
        total += 1
        nonconv += not o.converged
        events += int(raw["n_switch_events"][s])
        ok = bool(raw["converged"][s]) == o.converged and int(raw["iterations"][s]) == o.iters
        if ok and o.converged:
            ok = np.max(np.abs(raw["V"][s] - o.V)) <= 1e-8 and np.array_equal(raw["bus_type_final"][s], o.types.astype(np.uint8))
        if not ok:
            nb += 1
            print("   MISMATCH scenario", s, "gpu", bool(raw["converged"][s]), int(raw["iterations"][s]), int(raw["status"][s]), "oracle", o.converged, o.iters, o.reason)
    bad += nb
    print(r, g.name, g.n, "sigma", sigma, "qload override", with_q, "base", base.converged, "->", "OK" if nb == 0 else f"{nb} BAD", flush=True)
print(f"scenarios {total}, not converged {nonconv}, switch events {events}, TOTAL BAD {bad}")
sys.exit(1 if bad else 0)
