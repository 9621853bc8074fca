"""Regenerates tests/golden/*.json from the CPU oracle (python tests/golden/make_golden.py).
The reference itself (Julia) cannot run in this image, so these vectors pin the ORACLE (which is pinned to the
reference's own known-answer fixtures in tests/test_oracle_reference_fixtures.py) against silent drift, and give the
GPU tests a committed target that does not depend on the oracle code at run time."""
import json, os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "sparlectra.jl_b200"))
import numpy as np
import sparlectra_oracle as O
from sblx import grid as G


def dump(name, g, cases):
    vm, va, flat, base = O.solve_base(g)
    res = O.run_contingencies(g, cases)
    out = {"grid": g.name, "n": g.n, "nbranch": g.nbranch, "base_iterations": base.iters,
           "base_vm": [round(float(x), 12) for x in np.abs(base.V)], "base_va_deg": [round(float(x), 10) for x in np.degrees(np.angle(base.V))],
           "cases": []}
    for c, r in zip(cases, res):
        out["cases"].append({"removed": list(map(int, c)), "converged": bool(r.converged), "iterations": int(r.iterations),
                             "island_count": int(r.island_count), "error": r.error,
                             "vmin": None if not r.converged else round(float(r.min_vm_pu), 12),
                             "vmax": None if not r.converged else round(float(r.max_vm_pu), 12),
                             "vm": None if not r.converged else [round(float(x), 12) for x in np.abs(r.V)],
                             "types": None if r.types is None else [int(t) for t in r.types]})
    with open(os.path.join(HERE, name), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print(name, len(cases), "cases")


if __name__ == "__main__":
    g = G.case14()
    dump("case14_n1_oracle.json", g, O.generate_n1_branches(g))
    g = G.test3bus(qlim_min=-15.0, qlim_max=15.0)
    dump("test3bus_q15_n1_oracle.json", g, O.generate_n1_branches(g))
    g = G.island_chain(True)
    dump("island_chain_pv_n1_oracle.json", g, O.generate_n1_branches(g))
