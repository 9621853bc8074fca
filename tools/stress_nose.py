"""Static-pivoting stress near voltage collapse: uniform load scaling toward the nose of the PV curve (where the Jacobian
approaches singularity) on a tiled grid; every scale factor is solved by the engine (sblx_solve_injection_scale) and by
the oracle (SuperLU with partial pivoting) on the same injections.  Reports flips (converged flag / iteration count /
active set / |dV| > 1e-8), the tiny-pivot counters and the last converging factor.
  python tools/stress_nose.py [rows] [n_factors]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "sparlectra.jl_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import multiprocessing as mp
import numpy as np
import sparlectra_oracle as O
from sblx import api, grid as G, model as M

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 54
nf = int(sys.argv[2]) if len(sys.argv) > 2 else 48
g = G.tiled_grid(rows=rows, cols=rows, augment=True, rating_mva=100.0)
a0 = M.build_arrays(g)
eng = api.Engine(a0)
Vb, conv, *_ = eng.solve_one(None)
vm, va = np.abs(Vb), np.degrees(np.angle(Vb))
a = M.build_arrays(g, vm=vm, va_deg=va)
eng.set_v0(a.v0)
# bracket the nose with the engine itself (cheap), then sample densely just below and across it
lo, hi = 1.0, 64.0
for _ in range(14):
    mid = 0.5 * (lo + hi)
    r = eng.scale_sweep([mid])
    lo, hi = (mid, hi) if r["converged"][0] else (lo, mid)
scale = np.concatenate([np.linspace(0.9 * lo, lo, nf // 2), lo + (hi - lo) * np.linspace(0.0, 4.0, nf - nf // 2)])
raw = eng.scale_sweep(scale, want_v=True, want_types=True)
st = eng.stats()
eng.close()
isload = a.s_spec.real < 0
G_ = {}


def one(f):
    r = O.runpf(g, vm, va, (), s_override=np.where(isload, a.s_spec * f, a.s_spec), qload_override=np.where(isload, a.qload * f, a.qload))
    return r.converged, r.iters, r.V, r.types.astype(np.uint8)


with mp.get_context("fork").Pool(min(len(scale), os.cpu_count() or 2)) as pool:
    ref = pool.map(one, list(scale), chunksize=1)
flips = 0
for s, (conv, it, V, types) in enumerate(ref):
    same = bool(raw["converged"][s]) == conv and int(raw["iterations"][s]) == it
    if same and conv:
        same = np.max(np.abs(raw["V"][s] - V)) <= 1e-8 and np.array_equal(raw["bus_type_final"][s], types)
    if not same:
        flips += 1
        print(f"  flip at scale {scale[s]:.6f}: engine conv {int(raw['converged'][s])} it {int(raw['iterations'][s])} status {int(raw['status'][s])} "
              f"tiny {int(raw['n_tiny_pivots'][s])} | oracle conv {int(conv)} it {it}")
print(f"{g.name}: nose bracket [{lo:.5f}, {hi:.5f}] x base load; {len(scale)} factors, {int(raw['converged'].sum())} converge, "
      f"iterations {int(raw['iterations'].min())}..{int(raw['iterations'].max())}, flips {flips}, tiny pivots {int(raw['n_tiny_pivots'].sum())} "
      f"(stats.tiny_pivots {st['tiny_pivots']}), vmin at last converging factor {np.nanmin(raw['vmin_pu']):.4f}")
