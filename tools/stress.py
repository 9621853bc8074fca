"""Randomised parity stress (GPU): random grid shapes x random symbolic options, a handful of N-1 / N-2 scenarios
each against the oracle.  python tools/stress.py [rounds] [seed]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sparlectra.jl_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import sparlectra_oracle as O
from sblx import api, grid as G, model as M
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 24
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 7)
only = int(sys.argv[3]) if len(sys.argv) > 3 else None      # replay one round of a seed with diagnostics
bad = 0
for r in range(rounds):
    kind = rng.choice(["mesh", "rand"])
    if kind == "mesh":
        rows, cols = int(rng.integers(4, 40)), int(rng.integers(4, 40))
        g = G.tiled_grid(rows=rows, cols=cols, augment=True, pv_every=int(rng.integers(3, 8)), pv_q_mvar=float(rng.uniform(15, 40)))
    else:
        g = G.random_meshed_grid(int(rng.integers(40, 900)), int(rng.integers(1, 10**6)), extra_edge_frac=float(rng.uniform(0.1, 0.8)), gen_share=float(rng.uniform(0.9, 1.0)))
    opts = dict(ordering=int(rng.integers(0, 3)), relax_small=int(rng.integers(2, 40)), relax_zero_frac=float(rng.uniform(0.0, 0.4)),
                panel_smem_kb=float(rng.choice([24, 48, 100, 200])))
    ns = 14
    cases = [[int(k)] for k in rng.choice(g.nbranch, size=min(ns, g.nbranch), replace=False)] + \
            [sorted(int(x) for x in rng.choice(g.nbranch, size=2, replace=False)) for _ in range(4)]
    if only is not None and r != only:
        continue
    vm, va, flat, base = O.solve_base(g)
    a = M.build_arrays(g, vm=vm, va_deg=va, flatstart=flat)
    try:
        eng = api.Engine(a, **opts)
    except Exception as e:
        print(r, g.name, opts, "ENGINE ERROR", e); bad += 1; continue
    raw = eng.solve_outages(cases)
    info = eng.info(); eng.close()
    nb = 0
    for i, c in enumerate(cases):
        o = O.runpf(g, vm, va, c)
        ok = bool(raw["converged"][i]) == o.converged and int(raw["iterations"][i]) == o.iters
        if ok and o.converged:
            ok = np.max(np.abs(raw["V"][i] - o.V)) <= 1e-8 and np.array_equal(raw["bus_type_final"][i], o.types.astype(np.uint8))
        if not ok:
            nb += 1
            print("   MISMATCH case", c, "gpu", bool(raw["converged"][i]), int(raw["iterations"][i]), int(raw["status"][i]), "oracle", o.converged, o.iters)
            if o.converged and raw["converged"][i]:
                dv = np.abs(raw["V"][i] - o.V)
                tg, to = raw["bus_type_final"][i], o.types.astype(np.uint8)
                print("      max|dV| %.3e at bus %d; type differences at buses %s (gpu %s oracle %s); islands gpu %d oracle %d; switch events gpu %d"
                      % (dv.max(), int(dv.argmax()), np.flatnonzero(tg != to).tolist(), tg[tg != to].tolist(), to[tg != to].tolist(),
                         int(raw["island_count"][i]), o.island_count, int(raw["n_switch_events"][i])))
    bad += nb
    print(r, g.name, g.n, g.nbranch, opts, info["ordering"], "panels", info["n_panels"], "levels", info["n_levels"], "base", base.converged, "->", "OK" if nb == 0 else f"{nb} BAD", flush=True)
print("TOTAL BAD", bad)
sys.exit(1 if bad else 0)
