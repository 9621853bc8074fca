"""Config-4 probe (not the bench line): injection sweep on the 100x100 grid, end to end through sblx_solve_injections
(host-generated factors, H2D of the per-scenario injections included).  python tools/probe_sweep.py [n_scen]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sparlectra.jl_b200"))
import numpy as np
from sblx import api, grid as G, model as M
n_scen = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
g = G.tiled_grid(rows=100, cols=100, augment=True, rating_mva=100.0)
eng = api.Engine(M.build_arrays(g))
Vb, conv, *_ = eng.solve_one(None)
a = M.build_arrays(g, vm=np.abs(Vb), va_deg=np.degrees(np.angle(Vb)))
eng.set_v0(a.v0)
eng.injection_sweep(a.s_spec, a.qload, 512)            # warm
t0 = time.perf_counter()
r = eng.injection_sweep(a.s_spec, a.qload, n_scen)
dt = time.perf_counter() - t0
st = eng.stats()
print(f"sweep: {n_scen} scenarios in {dt:.2f} s = {n_scen / dt:.0f} scen/s end to end (factor generation on the host included); "
      f"converged {int(r['converged'].sum())}, mean iterations {r['iterations'].mean():.2f}, scenarios with PV->PQ events {int((r['n_switch_events'] > 0).sum())}; "
      f"last chunk: device {st['total_ms']:.0f} ms, h2d {st['h2d_ms']:.0f} ms for {st['h2d_bytes'] / 1e6:.0f} MB")
eng.close()
