"""GPU tests of the round-2 boundary: device-side violation / overload lists and branch flows, composed summary
records, device-side injection sweeps (Philox), duplicate / malformed outage sets, tiny-pivot monitor, the in-library
multi-GPU path (skipped on a 1-GPU box), and wide oracle comparisons at benchmark size."""
import math
import multiprocessing as mp
import os

import numpy as np
import pytest

import sparlectra_oracle as O
from sblx import api, grid as G, model as M, shard

pytestmark = pytest.mark.gpu
VTOL = 1e-8


def _ngpu():
    import ctypes
    try:
        rt = ctypes.CDLL("libcudart.so")
    except OSError:
        import torch
        return torch.cuda.device_count()
    n = ctypes.c_int(0)
    return n.value if rt.cudaGetDeviceCount(ctypes.byref(n)) == 0 else 0


_POOL_ARGS = {}


def _pool_runpf(job):
    g, vm, va = _POOL_ARGS["g"], _POOL_ARGS["vm"], _POOL_ARGS["va"]
    removed, S, ql = job
    r = O.runpf(g, vm, va, removed, s_override=S, qload_override=ql)
    return r.converged, r.iters, r.V, r.types.astype(np.uint8), r.reason


def oracle_many(g, vm, va, jobs):
    """oracle runpf over many scenarios on the host cores (forked workers run numpy / SuperLU only)"""
    _POOL_ARGS.update(g=g, vm=vm, va=va)
    nproc = max(1, min(len(jobs), (os.cpu_count() or 2) - 1, 16))
    if nproc == 1:
        return [_pool_runpf(j) for j in jobs]
    with mp.get_context("fork").Pool(nproc) as pool:
        return pool.map(_pool_runpf, jobs, chunksize=1)


def _warm_engine(g, **opts):
    vm, va, flat, base = O.solve_base(g)
    assert base.converged
    a = M.build_arrays(g, vm=vm, va_deg=va)
    return api.Engine(a, cooldown_iters=g.cooldown_iters, q_hyst_pu=g.q_hyst_pu, **opts), a, vm, va


def _ladder_split(g):
    ka = [k for k in range(g.nbranch) if (g.br_from[k], g.br_to[k]) == (14, 15)][0]
    kb = [k for k in range(g.nbranch) if (g.br_from[k], g.br_to[k]) == (73, 74)][0]
    return [ka, kb]


def test_device_lists_and_flows_match_oracle_metrics():
    """overloaded_branches / voltage_violations (contingency.jl:149-178) and the branch flows (losses.jl:53-85) come
    from the device; checked per scenario against the oracle's _contingency_metrics / branch_flows."""
    g = G.tiled_grid(rows=12, cols=12, augment=True, pv_every=4, pv_q_mvar=24.0, rating_mva=4.0)
    cases = O.generate_n1_branches(g)[:120] + [[3, 40], [7, 8]]
    eng, a, vm, va = _warm_engine(g, vm_min_pu=0.995, vm_max_pu=1.015)
    raw = eng.solve_outages(cases, want_lists=True, want_flows=True)
    eng.close()
    assert raw["lists_complete"]
    oracle = O.run_contingencies(g, cases, vm_min=0.995, vm_max=1.015)
    viol = api._group(raw["violations"], len(cases))
    over = api._group(raw["overloads"], len(cases))
    n_over = n_viol = 0
    for i, o in enumerate(oracle):
        assert bool(raw["converged"][i]) == o.converged
        if not o.converged:
            assert viol[i] == [] and over[i] == [] and raw["n_overloaded"][i] == 0
            continue
        assert viol[i] == sorted(o.voltage_violations), i
        assert over[i] == sorted(o.overloaded_branches), i
        assert raw["n_voltage_violations"][i] == len(o.voltage_violations) and raw["n_overloaded"][i] == len(o.overloaded_branches)
        n_over += len(over[i]); n_viol += len(viol[i])
        Vr = np.abs(o.V) * np.exp(1j * np.radians(np.degrees(np.angle(o.V))))
        sf, st = O.branch_flows(g, Vr, cases[i])
        assert np.max(np.abs(raw["flows"][i][:, 0] - sf)) <= 1e-8 and np.max(np.abs(raw["flows"][i][:, 1] - st)) <= 1e-8
    assert n_over > 50 and n_viol > 50          # the fixture does exercise both lists
    # loading values of the list entries
    k = 0
    for (sc, br), pct in zip(raw["overloads"], raw["overload_pct"]):
        assert pct > 100.0
        k += 1
    assert k == raw["overload_total"]


def test_list_overflow_is_reported_and_retry_completes():
    g = G.tiled_grid(rows=8, cols=8, augment=True, pv_every=3, rating_mva=2.0)
    cases = O.generate_n1_branches(g)
    eng, a, vm, va = _warm_engine(g, vm_min_pu=0.999, vm_max_pu=1.001)
    small = eng.solve_outages(cases, want_v=False, want_types=False, want_lists=True, list_capacity=7)
    assert not small["lists_complete"] and len(small["violations"]) == 7
    full = eng.solve_outages(cases, want_v=False, want_types=False, want_lists=True,
                             list_capacity=max(small["violation_total"], small["overload_total"]))
    eng.close()
    assert full["lists_complete"]
    assert full["n_voltage_violations"].sum() == len(full["violations"]) == full["violation_total"]
    assert full["n_overloaded"].sum() == len(full["overloads"])
    assert np.array_equal(small["n_voltage_violations"], full["n_voltage_violations"])


def test_device_summary_equals_fetched_results_with_islands_and_bad_branches():
    """the device records the all-gather moves must hold the composed island-wise results and the host
    verdicts, not the raw slots of the first B device scenarios."""
    import ctypes
    g = G.warmup_case118()
    eng, a, vm, va = _warm_engine(g)
    split = _ladder_split(g)          # both rails out: two islands with a reference each
    assert api.topology_only(a, [split])[1][0] == -2
    cases = [[0], split, [10 ** 6], [2, 5], [], split + [3]]
    raw = eng.solve_outages(cases)
    ptr, nbytes = eng.device_summary()
    assert nbytes == len(cases) * shard.RECORD_BYTES
    import torch
    buf = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    rt = ctypes.CDLL("libcudart.so")
    rt.cudaMemcpy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
    assert rt.cudaMemcpy(buf.data_ptr(), ptr, nbytes, 3) == 0
    rec = shard.records_from_bytes(buf)
    ref = shard.pack_records(raw)
    eng.close()
    for name in rec.dtype.names:
        x, y = rec[name], ref[name]
        assert np.array_equal(x, y, equal_nan=True) if x.dtype.kind == "f" else np.array_equal(x, y), name
    assert raw["status"][2] == 8 and raw["island_count"][1] == 2 and raw["converged"][1]


def test_duplicate_outage_entries_are_one_removal():
    g = G.case14()
    eng, a, vm, va = _warm_engine(g)
    r1 = eng.solve_outages([[4], [2, 7]])
    r2 = eng.solve_outages([[4, 4, 4], [7, 2, 7, 2]])
    eng.close()
    assert np.array_equal(r1["iterations"], r2["iterations"]) and np.array_equal(r1["V"], r2["V"])


def test_outage_ptr_is_validated():
    g = G.case14()
    eng, a, vm, va = _warm_engine(g)
    with pytest.raises(api.L.SblxError):
        eng.solve_outages((np.array([0, 2, 1], np.int32), np.array([1, 2], np.int32)))
    with pytest.raises(api.L.SblxError):
        eng.solve_outages((np.array([1, 2], np.int32), np.array([1, 2], np.int32)))
    ok = eng.solve_outages([[1]])
    eng.close()
    assert ok["converged"][0]


def _philox4x32_10(c, k):
    """numpy restatement of Philox4x32-10 (Salmon et al. 2011) for counter rows c[:, 4], key k[2]"""
    M0, M1, W0, W1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), 0x9E3779B9, 0xBB67AE85
    c = c.astype(np.uint64)
    k0, k1 = int(k[0]), int(k[1])
    for _ in range(10):
        p0, p1 = M0 * c[:, 0], M1 * c[:, 2]
        n0 = (p1 >> np.uint64(32)) ^ c[:, 1] ^ np.uint64(k0)
        n1 = p1 & np.uint64(0xFFFFFFFF)
        n2 = (p0 >> np.uint64(32)) ^ c[:, 3] ^ np.uint64(k1)
        n3 = p0 & np.uint64(0xFFFFFFFF)
        c = np.stack([n0, n1, n2, n3], axis=1)
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c


def sweep_factors_reference(B, n, first, seed, sigma, lo, hi):
    sc, bus = np.meshgrid(np.arange(first, first + B, dtype=np.uint64), np.arange(n, dtype=np.uint64), indexing="ij")
    ctr = np.stack([sc.ravel() & np.uint64(0xFFFFFFFF), sc.ravel() >> np.uint64(32), bus.ravel(), np.zeros(B * n, np.uint64)], axis=1)
    r = _philox4x32_10(ctr, (seed & 0xFFFFFFFF, seed >> 32))
    u1 = (((r[:, 0] << np.uint64(32)) | r[:, 1]) >> np.uint64(11)).astype(np.float64) + 1.0
    u2 = (((r[:, 2] << np.uint64(32)) | r[:, 3]) >> np.uint64(11)).astype(np.float64) + 1.0
    u1, u2 = u1 / 9007199254740992.0, u2 / 9007199254740992.0
    z = np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)
    return np.clip(1.0 + sigma * z, lo, hi).reshape(B, n)


def test_device_sweep_generator_matches_host_reference_and_oracle():
    """sblx_solve_injection_sweep: factors from Philox4x32-10 keyed on (seed, scenario, bus); a numpy restatement of the
    same counter scheme agrees to rounding, the draw is batching independent, and the solved scenarios match the oracle
    run on the very same injections."""
    g = G.tiled_grid(rows=10, cols=10, augment=True, pv_every=3, pv_q_mvar=20.0)
    eng, a, vm, va = _warm_engine(g)
    B, seed = 48, 20260724
    f = eng.sweep_factors(B, first=5, seed=seed)
    fr = sweep_factors_reference(B, g.n, 5, seed, 0.1, 0.5, 1.5)
    isload = a.s_spec.real < 0
    assert np.all(f[:, ~isload] == 1.0)
    assert np.max(np.abs(f[:, isload] - fr[:, isload])) <= 1e-12
    assert abs(f[:, isload].mean() - 1.0) < 0.01 and abs(f[:, isload].std() - 0.1) < 0.01
    raw = eng.sweep(B, first=5, seed=seed, want_v=True, want_types=True)
    part = eng.sweep(10, first=5 + 20, seed=seed, want_v=True)          # scenarios 20..29 on their own
    assert np.array_equal(part["V"], raw["V"][20:30]) and np.array_equal(part["iterations"], raw["iterations"][20:30])
    eng.close()
    S = np.where(isload, a.s_spec * f, a.s_spec)
    ql = np.where(isload, a.qload * f, a.qload)
    for s in range(0, B, 3):
        r = O.runpf(g, vm, va, (), s_override=S[s], qload_override=ql[s])
        assert bool(raw["converged"][s]) == r.converged and int(raw["iterations"][s]) == r.iters
        if r.converged:
            assert np.max(np.abs(raw["V"][s] - r.V)) <= VTOL
            assert np.array_equal(raw["bus_type_final"][s], r.types.astype(np.uint8))


def test_scale_sweep_matches_oracle():
    g = G.tiled_grid(rows=8, cols=8, augment=True, pv_every=3, pv_q_mvar=20.0)
    eng, a, vm, va = _warm_engine(g)
    scale = np.array([0.6, 0.9, 1.0, 1.15, 1.3])
    raw = eng.scale_sweep(scale, want_v=True, want_types=True)
    eng.close()
    isload = a.s_spec.real < 0
    for s, f in enumerate(scale):
        r = O.runpf(g, vm, va, (), s_override=np.where(isload, a.s_spec * f, a.s_spec), qload_override=np.where(isload, a.qload * f, a.qload))
        assert bool(raw["converged"][s]) == r.converged and int(raw["iterations"][s]) == r.iters
        if r.converged:
            assert np.max(np.abs(raw["V"][s] - r.V)) <= VTOL


def test_injection_sweep_on_a_two_island_base_grid():
    """island sub-scenarios (device ids >= B) must read their PARENT scenario's injections."""
    g = G.warmup_case118()
    g.br_status = np.array(g.br_status).copy()
    g.br_status[_ladder_split(g)] = 0            # the intact grid already consists of two referenced islands
    vm, va, flat, base = O.solve_base(g)
    assert base.converged and base.island_count == 2
    a = M.build_arrays(g, vm=vm, va_deg=va)
    eng = api.Engine(a)
    rng = np.random.default_rng(11)
    B = 6
    f = np.clip(rng.normal(1.0, 0.1, size=(B, g.n)), 0.5, 1.5)
    isload = a.s_spec.real < 0
    S = np.where(isload, a.s_spec * f, a.s_spec)
    ql = np.where(isload, a.qload * f, a.qload)
    raw = eng.solve_injections(S, ql)
    eng.close()
    for s in range(B):
        r = O.runpf(g, vm, va, (), s_override=S[s], qload_override=ql[s])
        assert r.island_count == 2 and int(raw["island_count"][s]) == 2
        assert bool(raw["converged"][s]) == r.converged and int(raw["iterations"][s]) == r.iters
        assert np.max(np.abs(raw["V"][s] - r.V)) <= VTOL


def test_run_contingencies_batched_names_from_device_lists():
    """runContingenciesBatched no longer pulls V: names come from the device lists; compare with the oracle."""
    g = G.tiled_grid(rows=9, cols=9, augment=True, pv_every=3, pv_q_mvar=24.0, rating_mva=3.0)
    cases = api.generateN1Branches(g)
    res = api.runContingenciesBatched(g, cases, vm_min_pu=0.99, vm_max_pu=1.015)   # not 1.02: PV buses sit ON their 1.02 setpoint
    oracle = O.run_contingencies(g, [[k] for k in range(g.nbranch)], vm_min=0.99, vm_max=1.015)
    assert len(res) == len(oracle)
    some = 0
    for r, o in zip(res, oracle):
        assert r.converged == o.converged and r.iterations == o.iterations and r.island_count == o.island_count
        if o.converged:
            assert r.overloaded_branches == [g.br_name[k] for k in o.overloaded_branches]
            assert r.voltage_violations == [g.bus_name[b] for b in o.voltage_violations]
            some += len(r.overloaded_branches) + len(r.voltage_violations)
    assert some > 20


def test_tiny_pivot_monitor_is_quiet_on_regular_cases_and_counts_when_forced():
    g = G.case14()
    eng, a, vm, va = _warm_engine(g)
    raw = eng.solve_outages(O.generate_n1_branches(g))
    assert raw["n_tiny_pivots"].sum() == 0 and eng.stats()["tiny_pivots"] == 0
    eng.close()
    eng2 = api.Engine(a, pivot_tol_rel=1e3)          # absurd threshold: every pivot block counts
    raw2 = eng2.solve_outages([[1]])
    assert raw2["n_tiny_pivots"][0] > 0 and eng2.stats()["tiny_pivots"] == raw2["n_tiny_pivots"][0]
    assert raw2["converged"][0] and np.array_equal(raw2["V"], eng2.solve_outages([[1]])["V"])
    eng2.close()


def test_config3_54x54_oracle_comparison_16_scenarios():
    """BASELINE config 3 grid (54 x 54 = 2 916 buses, 8 533 branches): 16 N-1 scenarios compared with the oracle in
    full (converged flag, iterations, active set, |dV| <= 1e-8)."""
    g = G.tiled_grid(rows=54, cols=54, augment=True, rating_mva=100.0)
    eng, a, vm, va = _warm_engine(g)
    rng = np.random.default_rng(54)
    cases = [[int(k)] for k in rng.choice(g.nbranch, size=16, replace=False)]
    raw = eng.solve_outages(cases)
    eng.close()
    for i, (conv, it, V, types, reason) in enumerate(oracle_many(g, vm, va, [(c, None, None) for c in cases])):
        assert bool(raw["converged"][i]) == conv and int(raw["iterations"][i]) == it, (i, cases[i])
        if conv:
            assert np.max(np.abs(raw["V"][i] - V)) <= VTOL
            assert np.array_equal(raw["bus_type_final"][i], types)


def test_config5_10k_oracle_comparison_32_scenarios():
    """BASELINE config 4/5 grid (100 x 100): 24 N-1 + 8 N-2 scenarios against the oracle in full."""
    g = G.tiled_grid(rows=100, cols=100, augment=True, rating_mva=100.0)
    a0 = M.build_arrays(g)
    eng = api.Engine(a0)
    Vb, conv, *_ = eng.solve_one(None)
    assert conv
    vm, va = np.abs(Vb), np.degrees(np.angle(Vb))
    a = M.build_arrays(g, vm=vm, va_deg=va)
    eng.set_v0(a.v0)
    rng = np.random.default_rng(1005)
    cases = [[int(k)] for k in rng.choice(g.nbranch, size=24, replace=False)]
    cases += [sorted(int(x) for x in rng.choice(g.nbranch, size=2, replace=False)) for _ in range(8)]
    raw = eng.solve_outages(cases)
    eng.close()
    for i, (conv, it, V, types, reason) in enumerate(oracle_many(g, vm, va, [(c, None, None) for c in cases])):
        assert bool(raw["converged"][i]) == conv and int(raw["iterations"][i]) == it, (i, cases[i])
        if conv:
            assert np.max(np.abs(raw["V"][i] - V)) <= VTOL
            assert np.array_equal(raw["bus_type_final"][i], types)


def test_config4_10k_device_sweep_oracle_comparison_32_scenarios():
    """BASELINE config 4: device-generated sweep on the 10k grid, 32 scenarios against the oracle on the same factors."""
    g = G.tiled_grid(rows=100, cols=100, augment=True, rating_mva=100.0)
    a0 = M.build_arrays(g)
    eng = api.Engine(a0)
    Vb, conv, *_ = eng.solve_one(None)
    vm, va = np.abs(Vb), np.degrees(np.angle(Vb))
    a = M.build_arrays(g, vm=vm, va_deg=va)
    eng.set_v0(a.v0)
    B = 32
    f = eng.sweep_factors(B, first=1000)
    raw = eng.sweep(B, first=1000, want_v=True, want_types=True)
    eng.close()
    assert raw["n_switch_events"].sum() > 0
    isload = a.s_spec.real < 0
    jobs = [((), np.where(isload, a.s_spec * f[s], a.s_spec), np.where(isload, a.qload * f[s], a.qload)) for s in range(B)]
    for s, (conv, it, V, types, reason) in enumerate(oracle_many(g, vm, va, jobs)):
        assert bool(raw["converged"][s]) == conv and int(raw["iterations"][s]) == it, s
        if conv:
            assert np.max(np.abs(raw["V"][s] - V)) <= VTOL
            assert np.array_equal(raw["bus_type_final"][s], types)


# ---------------------------------------------------------------------------------------------
# multi-GPU inside the library (needs >= 2 GPUs; the 1-GPU box runs the single-rank degenerate forms)
# ---------------------------------------------------------------------------------------------
def test_sharded_call_without_a_communicator_is_the_plain_solve():
    g = G.case14()
    eng, a, vm, va = _warm_engine(g)
    cases = O.generate_n1_branches(g)
    r1 = eng.solve_outages(cases, want_lists=True)
    r2 = eng.solve_outages(cases, want_lists=True, sharded=True)
    st = eng.stats()
    eng.close()
    for k in ("converged", "status", "iterations", "V", "island_count", "n_voltage_violations"):
        assert np.array_equal(r1[k], r2[k]), k
    assert st["n_ranks"] == 1 and st["shard_lo"] == 0 and st["shard_hi"] == len(cases)


def test_shard_bounds_match_python_mirror():
    g = G.case14()
    eng, a, vm, va = _warm_engine(g)
    for B in (0, 1, 7, 20, 29601):
        for w in (1, 2, 3, 8):
            b = eng.shard_bounds(B, w)
            assert [tuple(shard.shard_range(B, r, w)) for r in range(w)] == [(int(b[r]), int(b[r + 1])) for r in range(w)]
    eng.close()


@pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs")
def test_multi_engine_two_gpus_matches_single_gpu():
    """sblx_create_multi + sblx_multi_solve_outages: one process, two GPUs, one ncclAllGather inside the call."""
    g = G.tiled_grid(rows=12, cols=12, augment=True, pv_every=4, pv_q_mvar=24.0, rating_mva=4.0)
    cases = O.generate_n1_branches(g) + [[3, 40], [10 ** 6], [7, 8]]
    eng, a, vm, va = _warm_engine(g, vm_min_pu=0.995, vm_max_pu=1.015)
    one = eng.solve_outages(cases, want_lists=True, want_flows=True, list_capacity=400_000)
    eng.close()
    me = api.MultiEngine(a, 2, vm_min_pu=0.995, vm_max_pu=1.015)
    two = me.solve_outages(cases, want_lists=True, want_flows=True, list_capacity=400_000)
    st0, st1 = me.stats(0), me.stats(1)
    me.close()
    assert one["lists_complete"] and two["lists_complete"]
    for k in ("converged", "status", "iterations", "n_switch_events", "island_count", "V", "bus_type_final", "flows",
              "n_voltage_violations", "n_overloaded", "violations", "overloads", "qlimit_events"):
        assert np.array_equal(one[k], two[k]), k
    for k in ("final_mismatch", "vmin_pu", "vmax_pu", "max_loading_pct"):
        assert np.array_equal(one[k], two[k], equal_nan=True), k
    assert st0["n_ranks"] == 2 and st0["shard_hi"] == st1["shard_lo"] and st1["shard_hi"] == len(cases)


def _rank_worker(rank, world, uid_q, out_q, cases):
    g = G.tiled_grid(rows=12, cols=12, augment=True, pv_every=4, pv_q_mvar=24.0)
    vm, va, flat, base = O.solve_base(g)
    a = M.build_arrays(g, vm=vm, va_deg=va)
    eng = api.Engine(a, device=rank)          # one process per GPU: rank r owns device r
    if rank == 0:
        uid = eng.comm_unique_id()
        for _ in range(world - 1):
            uid_q.put(uid)
    else:
        uid = uid_q.get(timeout=120)
    eng.comm_init(world, rank, uid)
    r = eng.solve_outages(cases, want_v=False, want_types=False, sharded=True)
    st = eng.stats()
    eng.close()
    out_q.put((rank, r["converged"], r["iterations"], r["vmin_pu"], st["shard_lo"], st["shard_hi"], st["n_ranks"]))


@pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs")
def test_sharded_solve_two_processes_every_rank_gets_all_records():
    """one process per GPU: sblx_comm_unique_id / sblx_comm_init_rank / sblx_solve_outages_sharded"""
    g = G.tiled_grid(rows=12, cols=12, augment=True, pv_every=4, pv_q_mvar=24.0)
    cases = O.generate_n1_branches(g)
    ctx = mp.get_context("spawn")
    uid_q, out_q = ctx.Queue(), ctx.Queue()
    ps = [ctx.Process(target=_rank_worker, args=(r, 2, uid_q, out_q, cases)) for r in range(2)]
    for p in ps:
        p.start()
    got = [out_q.get(timeout=300) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    eng, a, vm, va = _warm_engine(g)
    ref = eng.solve_outages(cases, want_v=False, want_types=False)
    eng.close()
    got.sort()
    for rank, conv, it, vmin, lo, hi, nr in got:
        assert nr == 2 and np.array_equal(conv, ref["converged"]) and np.array_equal(it, ref["iterations"])
        assert np.array_equal(vmin, ref["vmin_pu"], equal_nan=True)
    assert got[0][4] == 0 and got[0][5] == got[1][4] and got[1][5] == len(cases)


def test_qlimit_event_log_matches_oracle_push_order():
    """net.qLimitLog (src/limits.jl:75-84): one (iteration, bus, side) entry per PV->PQ switch, in the reference's push
    order; the device appends them to a compacted list, the host restores (scenario, iteration, bus) order."""
    g = G.tiled_grid(rows=10, cols=10, augment=True, pv_every=3, pv_q_mvar=8.0)
    eng, a, vm, va = _warm_engine(g)
    cases = O.generate_n1_branches(g)[:60]
    raw = eng.solve_outages(cases, want_lists=True)
    eng.close()
    ev = raw["qlimit_events"]
    assert raw["qlimit_event_total"] == len(ev) == int(raw["n_switch_events"].sum()) > 0
    per = [[] for _ in cases]
    for sc, it, bus, side in ev:
        per[sc].append((int(it), int(bus), "max" if side > 0 else "min"))
    for i, c in enumerate(cases):
        r = O.runpf(g, vm, va, c)
        assert per[i] == [(it, b, s) for it, b, s in r.log], (i, c)


@pytest.mark.parametrize("prearmed", [True, False])
def test_iterative_refinement_path_keeps_parity(prearmed):
    """The safety net of the static-pivoting LU: a step whose factorisation met a tiny pivot and whose residual is bad gets
    one pass of iterative refinement (second factor + solve on the residual).  Forced here on healthy systems (every pivot
    flagged, every flagged step treated as bad): the refined steps must leave converged flag, iteration count, active set
    and voltages unchanged.  `prearmed` = False also exercises the first offender's path (step withheld, iteration
    repeated one tick later with refinement on)."""
    g = G.tiled_grid(rows=12, cols=12, augment=True, pv_every=4, pv_q_mvar=24.0)
    vm, va, flat, base = O.solve_base(g)
    a = M.build_arrays(g, vm=vm, va_deg=va)
    cases = O.generate_n1_branches(g)[:48]
    eng = api.Engine(a, pivot_tol_rel=1e3)
    ref = eng.solve_outages(cases)
    assert eng.stats()["refine_passes"] == 0 and ref["n_tiny_pivots"].min() > 0
    eng.lib.sblx_debug_set_pf_mode(eng.h, 1024 | (2048 if prearmed else 0))
    raw = eng.solve_outages(cases)
    st = eng.stats()
    eng.close()
    assert st["refine_passes"] >= int(raw["iterations"].sum()) - 2 * len(cases)
    oracle = O.run_contingencies(g, cases)
    for i, o in enumerate(oracle):
        assert bool(raw["converged"][i]) == o.converged and int(raw["iterations"][i]) == o.iterations, i
        if o.converged:
            assert np.max(np.abs(raw["V"][i] - o.V)) <= VTOL
            assert np.array_equal(raw["bus_type_final"][i], o.types.astype(np.uint8))
    assert np.array_equal(raw["iterations"], ref["iterations"]) and np.max(np.abs(raw["V"] - ref["V"])) <= 1e-10


def test_explicit_injection_sweep_streams_in_chunks(monkeypatch):
    """sblx_solve_injections streams the per-scenario injection matrices through the device in chunks of scenarios
    (a large sweep never needs B x n resident); forced here with a tiny budget: rows, lists and statistics of the
    chunked run equal the single-batch run."""
    g = G.tiled_grid(rows=10, cols=10, augment=True, pv_every=3, pv_q_mvar=20.0, rating_mva=3.0)
    eng, a, vm, va = _warm_engine(g, vm_min_pu=0.99, vm_max_pu=1.015)
    rng = np.random.default_rng(99)
    B = 100
    f = np.clip(rng.normal(1.0, 0.15, size=(B, g.n)), 0.5, 1.5)
    isload = a.s_spec.real < 0
    S, ql = np.where(isload, a.s_spec * f, a.s_spec), np.where(isload, a.qload * f, a.qload)

    def run():
        r = eng._alloc(B, g.n, True, True, want_lists=True)
        s = eng._results_struct(r)
        api.L.check(eng.lib.sblx_solve_injections(eng.h, B, api.L.ptr(np.ascontiguousarray(S).view(np.float64), api.L.c_f64p),
                                                  api.L.ptr(np.ascontiguousarray(ql), api.L.c_f64p), api.C.byref(s)), eng.h)
        return eng._finish_lists(r), eng.stats()

    one, st1 = run()
    monkeypatch.setenv("SBLX_SWEEP_CHUNK_GB", "0.0001")          # 100 KB: 41 scenarios per chunk
    many, st2 = run()
    eng.close()
    for k in ("converged", "iterations", "status", "V", "bus_type_final", "n_voltage_violations", "n_overloaded", "violations", "overloads",
              "qlimit_events"):
        assert np.array_equal(one[k], many[k]), k
    assert len(one["overloads"]) > 0 and len(one["violations"]) > 0
    assert st2["scenarios_solved"] == st1["scenarios_solved"] == B and st2["scenario_iterations"] == st1["scenario_iterations"]
    assert st2["ticks"] > st1["ticks"]


@pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs")
def test_multi_engine_with_fewer_scenarios_than_gpus_and_repeated_calls():
    """a rank without scenarios only takes part in the gather; handles are reusable across batch sizes"""
    g = G.case14()
    eng, a, vm, va = _warm_engine(g)
    cases = O.generate_n1_branches(g)
    ref = eng.solve_outages(cases, want_lists=True)
    eng.close()
    me = api.MultiEngine(a, 2)
    for sel in ([3], [0, 1, 2], list(range(len(cases))), [5]):
        r = me.solve_outages([cases[i] for i in sel], want_lists=True)
        for k in ("converged", "status", "iterations", "V", "bus_type_final", "island_count", "n_switch_events"):
            assert np.array_equal(r[k], ref[k][sel]), (k, sel)
    me.close()
