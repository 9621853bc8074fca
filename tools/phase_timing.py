"""Per-phase cycle breakdown of k_front (thread 0's view) by size class.  Needs the instrumented build:
  (cd sparlectra.jl_b200/csrc && nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo \
     -Xcompiler -fPIC,-O3 -shared -DSBLX_PHASE_TIMING engine.cu symbolic.cpp matpower.cpp -o ../lib/libsblx_timing.so)
The clock reads and atomics slow the kernels down (about 2x) and distort the chain phase: use it for the
relative size of the other phases and trust tools/ncu_phases.py for stall attribution."""
import sys, os, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sparlectra.jl_b200"))
import numpy as np
from sblx import lib as L
L.LIB_PATH = os.path.join(ROOT, "sparlectra.jl_b200", "lib", "libsblx_timing.so")
from sblx import api, grid as G, model as M
g = G.tiled_grid(rows=100, cols=100, augment=True, rating_mva=100.0)
eng = api.Engine(M.build_arrays(g))
Vb, conv, *_ = eng.solve_one(None)
a = M.build_arrays(g, vm=np.abs(Vb), va_deg=np.degrees(np.angle(Vb)))
eng.set_v0(a.v0)
nscen = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
optr, obr = eng.pack_outages([[k] for k in range(nscen)])
eng.stage_outages(optr, obr)
eng.run_staged()
lib = L.load()
buf0 = (C.c_double * 50)()
lib.sblx_debug_phase_cycles.argtypes = [C.POINTER(C.c_double)]
lib.sblx_debug_phase_cycles(buf0)
eng.run_staged()
buf1 = (C.c_double * 50)()
lib.sblx_debug_phase_cycles(buf1)
st = eng.stats()
d = (np.array(buf1[:]) - np.array(buf0[:])).reshape(5, 10)
names = ["ingest", "chain(w0)", "wait ingest", "subst", "U store", "schur(t0)", "cv", "tail"]
print("factor_ms", st["factor_ms"], "scenario_iterations", st["scenario_iterations"])
tot = d.sum()
for c, nt in enumerate([32, 64, 128, 256, 512]):
    row = d[c]
    n = max(row[9], 1)
    print("NT=%3d  CTAs %8d  avg cycles/CTA %8.0f | " % (nt, row[9], row[:8].sum() / n) + "  ".join("%s %6.0f" % (names[k], row[k] / n) for k in range(8)))
