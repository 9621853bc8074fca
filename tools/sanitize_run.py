"""Small batches of the hot path for compute-sanitizer (memcheck / racecheck / initcheck):
  compute-sanitizer --tool memcheck python tools/sanitize_run.py
case14 full N-1, the 118-bus ladder's N-1 with a split case, a 12x12 tiled grid with Q-limit switching, lists + flows,
and a device-side sweep - every kernel family of the engine runs at least once."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sparlectra.jl_b200"))
import numpy as np
from sblx import api, grid as G, model as M


def run(g, cases, **kw):
    eng = api.Engine(M.build_arrays(g), **kw)
    Vb, conv, *_ = eng.solve_one(None)
    a = M.build_arrays(g, vm=np.abs(Vb), va_deg=np.degrees(np.angle(Vb)))
    eng.set_v0(a.v0)
    r = eng.solve_outages(cases, want_lists=True, want_flows=True)
    s = eng.sweep(8, want_v=True)
    st = eng.stats()
    eng.close()
    return int(r["converged"].sum()), len(cases), int(s["converged"].sum()), st["launches"]


g = G.case14()
print("case14", run(g, [[k] for k in range(g.nbranch)]))
g = G.warmup_case118()
ka = [k for k in range(g.nbranch) if (g.br_from[k], g.br_to[k]) == (14, 15)][0]
kb = [k for k in range(g.nbranch) if (g.br_from[k], g.br_to[k]) == (73, 74)][0]
print("ladder118", run(g, [[k] for k in range(0, g.nbranch, 3)] + [[ka, kb], [10 ** 6]]))
g = G.tiled_grid(rows=12, cols=12, augment=True, pv_every=4, pv_q_mvar=24.0, rating_mva=4.0)
print("tiled12", run(g, [[k] for k in range(0, g.nbranch, 5)], max_slots=32))
g = G.tiled_grid(rows=30, cols=30, augment=True)
print("tiled30 (mid/large front classes)", run(g, [[k] for k in range(0, g.nbranch, 300)]))
