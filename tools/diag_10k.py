import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sparlectra.jl_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import sparlectra_oracle as O
from sblx import api, grid as G, model as M
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 100
g = G.tiled_grid(rows=rows, cols=rows, augment=True, rating_mva=100.0)
a0 = M.build_arrays(g)
eng = api.Engine(a0)
Vb, conv, itb, resb, stb = eng.solve_one(None)
print("gpu base: conv", conv, "iters", itb, "res", resb)
r = O.runpf(g, g.vm0, g.va0_deg, ())
print("oracle base: conv", r.converged, "iters", r.iters, "events", r.n_events, "dV", np.abs(r.V - Vb).max())
vm, va = np.abs(Vb), np.degrees(np.angle(Vb))
a = M.build_arrays(g, vm=vm, va_deg=va)
eng.set_v0(a.v0)
# one newton step check at the warm start with outage 5
Y, types, iso = O.create_ybus(g, [5]), O.bus_types_from_prosumers(g), []
S = O.complex_svec(g); vset = O.voltage_setpoints(g, vm)
F, J, dx, mm = eng.newton_step(a.v0, [5])
Fo = O.mismatch_rectangular_fast(Y, a.v0, S, types, vset, a.slack)
Jo = O.jacobian_sparse_fast(Y, a.v0, types, a.slack)
dxo = O.solve_linear(Jo, -Fo)
ref = np.concatenate([dx[0::2], dx[1::2]])
print("F diff", np.abs(F - Fo).max(), "max|F|", np.abs(Fo).max(), "dx diff", np.abs(ref - dxo).max(), "max|dx|", np.abs(dxo).max(),
      "gpu resid", np.abs(Jo @ ref + Fo).max(), "oracle resid", np.abs(Jo @ dxo + Fo).max())
cases = [[k] for k in range(8)]
raw = eng.solve_outages(cases)
print("gpu iters", raw["iterations"], "events", raw["n_switch_events"], "conv", raw["converged"])
for c in cases[:4]:
    t = time.time(); ro = O.runpf(g, vm, va, c)
    i = c[0]
    print("oracle", c, "iters", ro.iters, "events", ro.n_events, "conv", ro.converged,           "dV", np.abs(ro.V - raw["V"][i]).max(), "types equal", np.array_equal(ro.types.astype(np.uint8), raw["bus_type_final"][i]), "%.1fs" % (time.time() - t))
eng.close()
