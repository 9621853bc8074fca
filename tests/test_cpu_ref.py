"""The reference-shaped C++ CPU baseline (oracle/cpu_ref, what bench.py times as `cpu_baseline` / `--impl reference`)
against the Python oracle: same converged flag, iteration count, final active set, |dV| <= 1e-8, same metrics - for
both variants (A: symbolic analysis every iteration like the reference default; B: symbolic reused)."""
import math
import os
import subprocess

import numpy as np
import pytest

import sparlectra_oracle as O
from sblx import grid as G, model as M

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def cpu_ref():
    subprocess.check_call(["bash", os.path.join(ROOT, "oracle", "build_cpu_ref.sh")])
    import cpu_ref as C
    return C


CASES = {
    "case14": G.case14,
    "ladder118": G.warmup_case118,
    "tiled12_qlim": lambda: G.tiled_grid(rows=12, cols=12, augment=True, pv_every=4, pv_q_mvar=24.0, rating_mva=60.0),
    "test3bus_q15": lambda: G.test3bus(qlim_min=-15.0, qlim_max=15.0),
    "diverge_ring": G.diverge_ring,
}


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("variant", [0, 1])
def test_cpu_baseline_matches_python_oracle(cpu_ref, name, variant):
    g = CASES[name]()
    vm, va, flat, base = O.solve_base(g)
    a = M.build_arrays(g, vm=vm, va_deg=va, flatstart=flat)
    cases = O.generate_n1_branches(g)
    if g.nbranch > 30:
        cases += [[1, 7], [0, g.nbranch - 1]]
    oracle = O.run_contingencies(g, cases)
    r = cpu_ref.run(a, cases, variant=variant, nthreads=3)
    checked = 0
    for i, o in enumerate(oracle):
        if o.island_count > 1 or r["status"][i] in (7, 9):
            assert not r["converged"][i] or o.converged      # islanding is outside the baseline's scope: never a false positive
            continue
        assert bool(r["converged"][i]) == o.converged, (name, i)
        assert int(r["iterations"][i]) == o.iterations, (name, i)
        checked += 1
        if o.converged:
            assert np.max(np.abs(r["V"][i] - o.V)) <= 1e-8
            assert np.array_equal(r["bus_type_final"][i], o.types.astype(np.uint8))
            assert abs(r["vmin_pu"][i] - o.min_vm_pu) <= 1e-9 and abs(r["vmax_pu"][i] - o.max_vm_pu) <= 1e-9
            if math.isnan(o.max_branch_loading_pct):
                assert math.isnan(r["max_loading_pct"][i])
            else:
                assert abs(r["max_loading_pct"][i] - o.max_branch_loading_pct) <= 1e-6
        else:
            assert int(r["status"][i]) == o.reason
    assert checked >= min(3, len(cases))


def test_cpu_baseline_threads_and_variants_agree(cpu_ref):
    g = G.tiled_grid(rows=16, cols=16, augment=True, pv_every=5)
    vm, va, flat, base = O.solve_base(g)
    a = M.build_arrays(g, vm=vm, va_deg=va)
    cases = O.generate_n1_branches(g)[:60]
    r1 = cpu_ref.run(a, cases, variant=0, nthreads=1)
    r4 = cpu_ref.run(a, cases, variant=1, nthreads=4)
    assert np.array_equal(r1["iterations"], r4["iterations"]) and np.array_equal(r1["converged"], r4["converged"])
    assert np.max(np.abs(r1["V"] - r4["V"])) <= 1e-12
    assert r1["nnz_lu"] > 0
