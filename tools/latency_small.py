import sys, time
sys.path.insert(0, "/root/repo/sparlectra.jl_b200")
import numpy as np
from sblx import api, grid as G, model as M
for name, g in (("case14", G.case14()), ("warmup118", G.warmup_case118())):
    t0 = time.perf_counter(); eng = api.Engine(M.build_arrays(g)); t1 = time.perf_counter()
    Vb, conv, *_ = eng.solve_one(None)
    a = M.build_arrays(g, vm=np.abs(Vb), va_deg=np.degrees(np.angle(Vb))); eng.set_v0(a.v0)
    cases = [[k] for k in range(g.nbranch)]
    eng.solve_outages(cases)
    ts = []
    for _ in range(5):
        t2 = time.perf_counter(); r = eng.solve_outages(cases); ts.append(time.perf_counter() - t2)
    st = eng.stats(); info = eng.info()
    print(f"{name}: create {1e3*(t1-t0):.1f} ms; N-1 of {len(cases)} scenarios: {1e3*min(ts):.2f} ms wall ({len(cases)/min(ts):.0f} scen/s), converged {int(r['converged'].sum())}, ticks {st['ticks']}, launches {st['launches']}, levels {info['n_levels']}")
    eng.close()
