"""Throughput probe on other grid shapes (not the bench line): python tools/probe_grid.py random:2869:3 | tiled:54:54"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sparlectra.jl_b200"))
import numpy as np
from sblx import api, grid as G, model as M
spec = (sys.argv[1] if len(sys.argv) > 1 else "random:2869:3").split(":")
g = G.random_meshed_grid(int(spec[1]), int(spec[2]), gen_share=1.01) if spec[0] == "random" else G.tiled_grid(rows=int(spec[1]), cols=int(spec[2]), augment=True)
opts = dict(kv.split("=") for kv in sys.argv[2:])
nmax = int(opts.pop("nmax", 0))
opts = {k: (float(v) if "." in v else int(v)) for k, v in opts.items()}
eng = api.Engine(M.build_arrays(g), **opts)
Vb, conv, it, *_ = eng.solve_one(None)
a = M.build_arrays(g, vm=np.abs(Vb), va_deg=np.degrees(np.angle(Vb)))
eng.set_v0(a.v0)
cases = [[k] for k in range(g.nbranch)]
rep = max(1, 20000 // len(cases))
cases = cases * rep
if nmax:
    cases = cases[:: max(1, len(cases) // nmax)][:nmax]
optr, obr = eng.pack_outages(cases)
eng.stage_outages(optr, obr)
eng.run_staged()
t = []
for _ in range(3):
    eng.run_staged(); st = eng.stats(); t.append(st["total_ms"])
print("run totals ms:", [round(x, 1) for x in t], "ticks", st["ticks"], "launches", st["launches"])
res = eng.fetch(); info = eng.info()
ms = float(np.median(t))
print(f"{g.name} {opts}: buses {g.n} branches {g.nbranch} base conv {conv} it {it}; scenarios {len(cases)} ({rep}x N-1) converged {int(res['converged'].sum())} "
      f"islanded-split {int((res['island_count'] > 1).sum())}; {len(cases) / ms * 1e3:.0f} scen/s, {st['scenario_iterations'] / ms * 1e3:.0f} iter/s; "
      f"factor {st['factor_ms']:.0f} solve {st['solve_ms']:.0f} jac {st['jacobian_ms']:.0f} mis {st['mismatch_ms']:.0f} other {st['other_ms']:.0f} ms; "
      f"nnzLU {info['nnz_lu_scalar']} panels {info['n_panels']} levels {info['n_levels']} slots {info['max_slots']} ordering {info['ordering']}")
eng.close()
