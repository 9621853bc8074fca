"""Randomised parity stress of the topology paths (GPU): small, nearly radial grids, N-2 / N-3 / N-4 outage sets, so
that most scenarios isolate buses and / or split the grid into several islands (with and without a reference).
python tools/stress_islands.py [rounds] [seed]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sparlectra.jl_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import sparlectra_oracle as O
from sblx import api, grid as G, model as M
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 11)
bad = total = multi = noref = 0
for r in range(rounds):
    g = G.random_meshed_grid(int(rng.integers(20, 140)), int(rng.integers(1, 10**6)), extra_edge_frac=float(rng.uniform(0.02, 0.25)),
                             pv_frac=float(rng.uniform(0.1, 0.4)), gen_share=float(rng.uniform(0.9, 1.0)))
    vm, va, flat, base = O.solve_base(g)
    eng = api.Engine(M.build_arrays(g, vm=vm, va_deg=va, flatstart=flat), ordering=int(rng.integers(0, 3)))
    cases = [sorted(int(x) for x in rng.choice(g.nbranch, size=int(rng.integers(2, 5)), replace=False)) for _ in range(24)]
    raw = eng.solve_outages(cases)
    eng.close()
    nb = 0
    for i, c in enumerate(cases):
        o = O.runpf(g, vm, va, c, flatstart=flat)
        total += 1
        multi += o.island_count > 1
        noref += o.reason == O.R_ISLANDED_NO_REF
        ok = bool(raw["converged"][i]) == o.converged and int(raw["island_count"][i]) == o.island_count
        if ok and o.reason == O.R_ISLANDED_NO_REF:
            ok = int(raw["status"][i]) == 7
        elif ok:
            ok = int(raw["iterations"][i]) == o.iters
        if ok and o.converged:
            ok = np.max(np.abs(raw["V"][i] - o.V)) <= 1e-8 and np.array_equal(raw["bus_type_final"][i], o.types.astype(np.uint8))
        if not ok:
            nb += 1
            print("   MISMATCH", g.name, c, "gpu", bool(raw["converged"][i]), int(raw["iterations"][i]), int(raw["status"][i]), int(raw["island_count"][i]),
                  "oracle", o.converged, o.iters, o.reason, o.island_count)
    bad += nb
    print(r, g.name, g.n, g.nbranch, "base", base.converged, "->", "OK" if nb == 0 else f"{nb} BAD", flush=True)
print(f"scenarios {total}, with several islands {multi}, islanded without reference {noref}, TOTAL BAD {bad}")
sys.exit(1 if bad else 0)
