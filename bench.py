#!/usr/bin/env python
"""bench.py - N-1 power flows solved per second on a ~10k-bus grid (BASELINE.json metric).

A "step" is one pass of the hot path over one batch of synthetic scenarios.  Workloads (--config):

  n1_10k (default)  full branch N-1 of the 100x100 tiled grid (src/synthetic_grids.jl topology, 10 000 buses /
                    29 601 branches, documented PV / load augmentation).  N > 1: the SAME 29 601 cases are split into
                    contiguous shards over the ranks ("scaling": "strong") - the reference's own fan-out
                    (src/contingency.jl:299-318) - and every rank gets the records of all cases back through ONE
                    ncclAllGather inside the library call (sblx_solve_outages_sharded).
  n1n2_10k          BASELINE config 5: per GPU 29 601 scenarios - the N-1 list extended with seeded random N-2 pairs as the
                    rank count grows ("scaling": "weak"), same sharded call.
  n1_2916_strong    BASELINE config 3: full N-1 (8 533 cases) of the 54x54 tiled grid split over the ranks (strong).
  sweep_100k        BASELINE config 4: 100 000 load / injection draws on the 10k grid with Q-limit switching, factors
                    generated ON THE DEVICE (sblx_solve_injection_sweep), split over the ranks (strong).

  value   : scenarios / s of device time (CUDA events on the engine's stream, inputs resident in HBM), max over ranks;
            N > 1: device run + the all-gather
  e2e     : scenarios / s by the wall clock through the public C-ABI call with HOST buffers (host topology analysis,
            H2D, solve, all-gather, D2H inside the timed region)
  parity  : rank 0, N = 1: the engine's rows for the CPU sample's scenarios against the CPU baseline (C++, pinned to
            the Python oracle by tests/test_cpu_ref.py) and a few scenarios against the Python oracle directly:
            converged flag, iteration count, final bus types identical, max |dV| <= 1e-8.  A flip exits non-zero.
  roofline: dominant kernel k_front (numeric multifrontal LU); roofline_fp64 beside it
  cpu_baseline / --impl reference : oracle/cpu_ref (reference-shaped multithreaded C++ baseline, variant A = symbolic
            analysis every iteration like the reference default, variant B = symbolic reused) on a bounded sample
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "sparlectra.jl_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "n1_power_flows_per_second"
UNIT = "scenarios/s"
FP64_NOMINAL_TFLOPS = 37.2      # 148 SMs x 64 DFMA / clk x 2 flop x 1.965 GHz (no measured FP64 peak in MEASURED_PEAKS.json)
CONFIGS = ("n1_10k", "n1n2_10k", "n1_2916_strong", "sweep_100k")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="n1_10k", choices=CONFIGS)
    ap.add_argument("--rows", type=int, default=0)
    ap.add_argument("--cols", type=int, default=0)
    ap.add_argument("--scenarios", type=int, default=0, help="scenarios per step (0 = the config's full list)")
    ap.add_argument("--max-slots", type=int, default=0)
    ap.add_argument("--cpu-sample", type=int, default=0, help="scenarios in the cpu_baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-api-e2e", action="store_true", help="skip the runContingenciesBatched end-to-end number")
    ap.add_argument("--ordering", type=int, default=0)
    ap.add_argument("--relax-small", type=int, default=0)
    ap.add_argument("--relax-frac", type=float, default=0.0)
    ap.add_argument("--panel-kb", type=float, default=0.0)
    ap.add_argument("--factor-mode", type=int, default=-1)
    return ap.parse_args()


def workload(args, world):
    """-> grid, scenario list (None for the sweep), scenarios per step (whole job), name, scaling"""
    from sblx import grid as G
    cfg = args.config
    rows, cols = (54, 54) if cfg == "n1_2916_strong" else (100, 100)
    rows, cols = args.rows or rows, args.cols or cols
    g = G.tiled_grid(rows=rows, cols=cols, augment=True, rating_mva=100.0)
    n1 = [[k] for k in range(g.nbranch)]
    if cfg == "sweep_100k":
        total = args.scenarios or 100_000
        return g, None, total, f"tiled_{rows}x{cols}_aug_injection_sweep_{total}_truncnormal_sigma0.1_device_philox", "strong"
    if cfg == "n1n2_10k":
        per_rank = args.scenarios or len(n1)
        total = per_rank * world
        cases = list(n1[:total])
        rng = np.random.Generator(np.random.PCG64(20260724))
        while len(cases) < total:
            a, b = rng.integers(0, g.nbranch, size=2)
            if a != b:
                cases.append(sorted((int(a), int(b))))
        return g, cases, total, f"tiled_{rows}x{cols}_aug_N-1" + ("+randomN-2" if total > len(n1) else ""), "weak"
    cases = n1[: args.scenarios] if args.scenarios else n1
    return g, cases, len(cases), f"tiled_{rows}x{cols}_aug_N-1", "strong"


class ClockSampler:
    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self._stop = threading.Event()
        self.th = None

    def start(self):
        def run():
            q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
            while not self._stop.is_set():
                try:
                    out = subprocess.run(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                         capture_output=True, text=True, timeout=5).stdout.strip()
                    if out:
                        self.rows.append([x.strip() for x in out.split(",")])
                except Exception:
                    pass
                self._stop.wait(0.2)
        self.th = threading.Thread(target=run, daemon=True)
        self.th.start()

    def stop(self):
        self._stop.set()
        if self.th:
            self.th.join(timeout=6)
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for k, nm in enumerate(names):
                    if r[3 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------
# CPU baseline: oracle/cpu_ref (C++), bounded sample
# ---------------------------------------------------------------------------------------------------------------
def build_cpu_ref():
    # rebuilt on the machine that runs it (-march=native: the .so that travelled with the snapshot may target another CPU)
    so = os.path.join(ROOT, "oracle", "_build", "libcpuref.so")
    rc = subprocess.call(["bash", os.path.join(ROOT, "oracle", "build_cpu_ref.sh")], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    if rc != 0 and not os.path.exists(so):
        raise SystemExit("bench.py: oracle/build_cpu_ref.sh failed and no prebuilt libcpuref.so is present")
    import cpu_ref
    return cpu_ref


def cpu_sample_size(g, ncores, given=0):
    """about 12 s of variant-A work: the C++ baseline needs ~0.05 ms per bus per scenario and core on the tiled grids"""
    if given:
        return given
    per_scen_s = max(2e-3, 5.5e-5 * g.n)
    return int(min(max(ncores, ncores * round(12.0 / per_scen_s)), 8192))


def sweep_cases_arrays(a, factors):
    isload = a.s_spec.real < 0
    return np.where(isload, a.s_spec * factors, a.s_spec), np.where(isload, a.qload * factors, a.qload)


def cpu_baseline_run(cpu_ref, a, cases, ncores):
    """variant A (reference default: symbolic analysis every iteration) and B (symbolic reused) on the same sample"""
    rA = cpu_ref.run(a, cases, variant=0, nthreads=ncores, want_v=True)
    rB = cpu_ref.run(a, cases, variant=1, nthreads=ncores, want_v=False)
    return rA, rB


_ORC = {}


def _orc_one(c):
    import sparlectra_oracle as O
    r = O.runpf(_ORC["g"], _ORC["vm"], _ORC["va"], c)
    return r.converged, r.iters, r.V, r.types.astype(np.uint8)


def python_oracle_rows(g, vm, va, cases):
    import multiprocessing as mp
    _ORC.update(g=g, vm=vm, va=va)
    nproc = max(1, min(len(cases), (os.cpu_count() or 2)))
    with mp.get_context("fork").Pool(nproc) as pool:
        return pool.map(_orc_one, cases, chunksize=1)


def parity_gate(eng, g, a, vm, va, cases, rA, n_python=8):
    """engine rows of the sample scenarios against the C++ baseline rows and n_python scenarios of the Python oracle"""
    raw = eng.solve_outages(cases, want_v=True, want_types=True)
    flips, max_dv, skipped = 0, 0.0, 0
    for i in range(len(cases)):
        if rA["status"][i] in (7, 9) or raw["island_count"][i] != 1:
            skipped += 1          # islanding is outside the C++ baseline's scope
            continue
        same = bool(raw["converged"][i]) == bool(rA["converged"][i]) and int(raw["iterations"][i]) == int(rA["iterations"][i])
        if same and raw["converged"][i]:
            dv = float(np.max(np.abs(raw["V"][i] - rA["V"][i])))
            max_dv = max(max_dv, dv)
            same = dv <= 1e-8 and np.array_equal(raw["bus_type_final"][i], rA["bus_type_final"][i])
        flips += 0 if same else 1
    pick = list(range(0, len(cases), max(1, len(cases) // n_python)))[:n_python]
    py_flips, py_dv = 0, 0.0
    for i, (conv, it, V, types) in zip(pick, python_oracle_rows(g, vm, va, [cases[i] for i in pick])):
        same = bool(raw["converged"][i]) == conv and int(raw["iterations"][i]) == it
        if same and conv:
            dv = float(np.max(np.abs(raw["V"][i] - V)))
            py_dv = max(py_dv, dv)
            same = dv <= 1e-8 and np.array_equal(raw["bus_type_final"][i], types)
        py_flips += 0 if same else 1
    return {"n": len(cases) - skipped, "flips": flips, "max_dv": max_dv, "skipped_islanded": skipped,
            "checker": "oracle/cpu_ref (C++ baseline, variant A) on the cpu_baseline sample",
            "python_oracle": {"n": len(pick), "flips": py_flips, "max_dv": py_dv},
            "criteria": "converged flag, iteration count, final bus types identical; max |dV| <= 1e-8 p.u."}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return reference_arm(args, rank, world)

    import torch
    import torch.distributed as dist
    from sblx import api, model as M, shard
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    g, cases, total, wname, scaling = workload(args, world)
    sweep = cases is None
    a0 = M.build_arrays(g)
    t0 = time.perf_counter()
    extra = {}
    if args.relax_small > 0:
        extra["relax_small"] = args.relax_small
    if args.relax_frac > 0:
        extra["relax_zero_frac"] = args.relax_frac
    if args.panel_kb > 0:
        extra["panel_smem_kb"] = args.panel_kb
    if args.factor_mode >= 0:
        extra["factor_mode"] = args.factor_mode
    eng = api.Engine(a0, device=local_rank, max_slots=args.max_slots, ordering=args.ordering, **extra)
    t_create = time.perf_counter() - t0
    Vb, conv, itb, resb, stb = eng.solve_one(None)
    if not conv:
        raise SystemExit("bench.py: base case did not converge")
    vm, va = np.abs(Vb), np.degrees(np.angle(Vb))
    a = M.build_arrays(g, vm=vm, va_deg=va)
    eng.set_v0(a.v0)
    if os.environ.get("SBLX_MODE_AFTER_BASE"):        # tuning probes: switch an experiment on after the base solve
        eng.lib.sblx_debug_set_pf_mode(eng.h, int(os.environ["SBLX_MODE_AFTER_BASE"]))
    lo, hi = shard.shard_range(total, rank, world)
    if world > 1 and not sweep:
        uid = [eng.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        eng.comm_init(world, rank, uid[0])      # NCCL communicator INSIDE the library: the gather is part of the call
    if not sweep:
        optr, obr = eng.pack_outages(cases)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    AGG = ("factor_ms", "solve_ms", "mismatch_ms", "jacobian_ms", "other_ms", "launches", "factor_launches", "scenario_iterations",
           "mismatch_evaluations", "ticks")
    agg = {k: 0 for k in AGG}
    ag_ms = 0.0

    def one_step(e2e_call):
        """one pass over this rank's share; returns device ms of the pass (run + gather)"""
        nonlocal ag_ms
        if sweep:
            r = eng.sweep(hi - lo, first=lo)
        elif world > 1:
            r = eng.solve_outages((optr, obr), want_v=False, want_types=False, sharded=True)
        elif e2e_call:
            r = eng.solve_outages((optr, obr), want_v=False, want_types=False)
        else:
            eng.run_staged()
            r = None
        st = eng.stats()
        for k in AGG:
            agg[k] += st[k]
        ag_ms += st["allgather_ms"]
        return st["total_ms"] + st["allgather_ms"], r, st

    # ---- device-time arm (value).  N = 1 outage configs: inputs staged once, sblx_run_staged only.
    staged_arm = (world == 1 and not sweep)
    if staged_arm:
        eng.stage_outages(optr, obr)
    for _ in range(args.warmup):
        one_step(False)
    info = eng.info()
    agg = {k: 0 for k in AGG}
    ag_ms = 0.0
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    dev_ms, wall0, last = 0.0, time.perf_counter(), None
    for _ in range(args.steps):
        ms, r, st = one_step(False)
        dev_ms += ms
        last = r if r is not None else last
    barrier()
    wall_ms = (time.perf_counter() - wall0) * 1e3
    clocks = sampler.stop()
    if staged_arm:
        last = eng.fetch()
    if sweep and world > 1:      # records of the other ranks' draws: one all-gather of the engine's device buffer
        ptr, nbytes = eng.device_summary()

        class _Raw:
            __cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}
        rec = shard.records_from_bytes(shard.gather_records(torch.as_tensor(_Raw(), device=f"cuda:{local_rank}"), total, world, dist))
        conv_all, it_sum = int(rec["converged"].sum()), int(rec["iterations"].sum())
    else:
        conv_all, it_sum = int(last["converged"].sum()), int(last["iterations"].sum())
    t = torch.tensor([dev_ms, wall_ms], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max, wall_ms_max = t.tolist()
    agg_value, ag_value = dict(agg), ag_ms

    # ---- end-to-end arm: the public C-ABI call with host buffers.  Sharded / sweep calls already ARE that call: the
    # wall clock of the timed loop above is the end-to-end number; the staged N = 1 arm gets its own loop here.
    if staged_arm:
        e2e_steps = max(1, min(args.steps, 2))
        one_step(True)   # warm
        barrier()
        w0 = time.perf_counter()
        for _ in range(e2e_steps):
            _, _, st = one_step(True)
        barrier()
        e2e_s = time.perf_counter() - w0
    else:
        e2e_steps, e2e_s = args.steps, wall_ms_max * 1e-3
    h2d, d2h = st["h2d_bytes"], st["d2h_bytes"]

    # ---- second end-to-end number: the reference-shaped call runContingenciesBatched (base solve, engine creation,
    # symbolic analysis, device-side violation / overload lists, result objects) - rank 0, N = 1, outage configs
    api_e2e = None
    if rank == 0 and world == 1 and not sweep and not args.no_api_e2e:
        cc = api.generateN1Branches(g)[: len(cases)]
        w0 = time.perf_counter()
        res = api.runContingenciesBatched(g, cc)
        dt = time.perf_counter() - w0
        api_e2e = {"value": len(cc) / dt, "unit": UNIT, "seconds": dt, "converged": sum(1 for r_ in res if r_.converged),
                   "call": "runContingenciesBatched(grid, generateN1Branches(grid)): engine creation + symbolic analysis + base solve + "
                           "one batched solve with device-side lists + ContingencyResult objects"}

    # ---- BASELINE configs 1 / 2 as latencies (rank 0, N = 1): wall clock of one sblx_solve_outages call on tiny batches
    small = None
    if rank == 0 and world == 1 and args.config == "n1_10k" and not args.scenarios:
        from sblx import grid as G
        small = {}
        for nm, gs in (("ieee14_full_n1", G.case14()), ("ladder118_full_n1", G.warmup_case118())):
            es = api.Engine(M.build_arrays(gs), device=local_rank)
            Vs, *_ = es.solve_one(None)
            es.set_v0(M.build_arrays(gs, vm=np.abs(Vs), va_deg=np.degrees(np.angle(Vs))).v0)
            cs = [[k] for k in range(gs.nbranch)]
            es.solve_outages(cs)
            ts = []
            for _ in range(5):
                w0 = time.perf_counter()
                rs = es.solve_outages(cs)
                ts.append(time.perf_counter() - w0)
            sst = es.stats()
            small[nm] = {"scenarios": len(cs), "ms": 1e3 * min(ts), "converged": int(rs["converged"].sum()), "ticks": sst["ticks"], "launches": sst["launches"]}
            es.close()

    # ---- CPU baseline + parity gate (rank 0, N = 1)
    cpu = parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not sweep:
        cpu_ref = build_cpu_ref()
        ncores = os.cpu_count() or 1
        ns = min(len(cases), cpu_sample_size(g, ncores, args.cpu_sample))
        sample = cases[:ns]
        rA, rB = cpu_baseline_run(cpu_ref, a, sample, ncores)
        cpu = {"value": ns / rA["seconds"], "unit": UNIT, "cores": ncores, "kind": "port",
               "variant_A_symbolic_every_iteration": ns / rA["seconds"], "variant_B_symbolic_reused": ns / rB["seconds"],
               "solver": "oracle/cpu_ref: own multifrontal LU (approximate minimum degree on A+A^T, supernodes, partial pivoting "
                         "inside the fronts), std::thread, contiguous chunks per thread",
               "nnz_lu": rA["nnz_lu"],
               "sample": f"first {ns} N-1 scenarios of the same workload, reference-shaped C++ baseline on {ncores} threads: "
                         f"variant A {rA['seconds']:.1f} s ({int(rA['iterations'].sum())} NR iterations), variant B {rB['seconds']:.1f} s"}
        parity = parity_gate(eng, g, a, vm, va, sample, rA)
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650 GB/s"
        nnzj, nnzlu, N = info["nnz_j_scalar"], info["nnz_lu_scalar"], 2 * info["m"]
        agg = agg_value
        units = agg["scenario_iterations"]
        front_bytes_unit = 8.0 * nnzj + 8.0 * nnzlu + 16.0 * N
        traffic, traffic_src = None, None
        for fn in ("r2_k_front_traffic.json", "r1_k_front_traffic.json"):
            try:   # ncu dram__bytes_read+write of all k_front launches of one tick / scenarios in that tick (same grid)
                tj = json.load(open(os.path.join(ROOT, "profiles", fn)))
                if g.n == 10000:
                    traffic = tj["k_front_dram_bytes_per_unit"] * units / max(1, agg["factor_launches"])
                    traffic_src = fn
                    break
            except Exception:
                pass
        own_ms = dev_ms - ag_value                       # rank 0's device time without the gather
        ach_front = units * front_bytes_unit / (agg["factor_ms"] * 1e-3) / 1e9 if agg["factor_ms"] > 0 else 0.0
        ach_step = units * info["algorithmic_bytes_per_iteration"] / (own_ms * 1e-3) / 1e9 if own_ms > 0 else 0.0
        fp64 = units * info["block_ops"] * 16.0 / (agg["factor_ms"] * 1e-3) / 1e12 if agg["factor_ms"] > 0 else 0.0
        line = {
            "metric": METRIC, "value": total * args.steps / (dev_ms_max * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wname, "name": args.config, "buses": g.n, "branches": g.nbranch,
                       "scenarios_total_per_step": total, "scenarios_rank0_per_step": hi - lo, "tol": 1e-8, "max_iter": 30, "qlimits": True,
                       "l2_policy": "per-step working set (>= 20 MB per resident scenario x slots) far exceeds the 126 MB L2",
                       "ordering": info["ordering"], "slots": info["max_slots"], "nnz_j": nnzj, "nnz_lu": nnzlu,
                       "levels": info["n_levels"], "panels": info["n_panels"],
                       "multi_gpu": None if world == 1 else ("contiguous shards + one ncclAllGather of 64-byte records inside sblx_solve_outages_sharded"
                                                             if not sweep else "contiguous draw ranges per rank (counter-based generator), records gathered once")},
            "converged": conv_all, "iterations_per_scenario": it_sum / max(1, total),
            "wall_ms_per_step": wall_ms_max / args.steps,
            "clocks": clocks,
            "e2e": {"value": total * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "call": "sblx_solve_injection_sweep" if sweep else ("sblx_solve_outages_sharded" if world > 1 else "sblx_solve_outages")},
            "gpu_launches": int(agg["launches"]),
            "roofline": {"bound": "hbm", "kernel": "k_front (numeric multifrontal LU, all launches of one pass)",
                         "achieved": ach_front, "peak": hbm, "unit": "GB/s", "frac": ach_front / hbm, "traffic": traffic,
                         "algorithmic_bytes_per_launch": units * front_bytes_unit / max(1, agg["factor_launches"]),
                         "traffic_note": f"bytes per launch: ncu dram read+write per scenario-iteration (profiles/{traffic_src}) x units / launches; "
                                         "the excess over the algorithmic bytes is the multifrontal update-matrix arena",
                         "peak_source": peak_src, "bytes_per_unit": front_bytes_unit, "units": units,
                         "avg_launch_ms": agg["factor_ms"] / max(1, agg["factor_launches"]), "launches": agg["factor_launches"],
                         "kernel_ms": agg["factor_ms"]},
            "roofline_fp64": {"bound": "fp64", "kernel": "k_front", "achieved": fp64, "peak": FP64_NOMINAL_TFLOPS, "unit": "TFLOP/s",
                              "frac": fp64 / FP64_NOMINAL_TFLOPS, "flop_per_unit": info["block_ops"] * 16.0,
                              "peak_source": "nominal: 148 SMs x 64 DFMA/clk x 2 x 1.965 GHz"},
            "roofline_step": {"achieved": ach_step, "peak": hbm, "unit": "GB/s", "frac": ach_step / hbm,
                              "bytes_per_unit": info["algorithmic_bytes_per_iteration"],
                              "scenario_iterations_per_s": units / (own_ms * 1e-3) if own_ms > 0 else 0.0},
            "kernel_ms": {k: agg[k] for k in ("factor_ms", "solve_ms", "mismatch_ms", "jacobian_ms", "other_ms")},
            "allgather_ms_per_step": ag_value / args.steps if world > 1 else None, "create_s": t_create,
        }
        if api_e2e:
            line["e2e_api"] = api_e2e
        if small:
            line["small_batch_latency"] = small
        if cpu:
            line["cpu_baseline"] = cpu
        if parity:
            line["parity"] = parity
        print(json.dumps(line))
    eng.close()
    if world > 1:
        dist.destroy_process_group()
    if parity and (parity["flips"] or parity["python_oracle"]["flips"]):
        raise SystemExit(f"bench.py: parity gate failed: {parity}")


def reference_arm(args, rank, world):
    """Times the reference's CPU implementation of the path: oracle/cpu_ref (reference-shaped multithreaded C++ baseline;
    Julia is not in the image), variant A = symbolic analysis every iteration (the reference default), with every host
    core on a bounded sample of the same workload.  Variant B (symbolic reused) is reported beside it."""
    if rank != 0:
        return
    import sparlectra_oracle as O
    from sblx import model as M
    g, cases, total, wname, scaling = workload(args, 1)
    cpu_ref = build_cpu_ref()
    vm, va, flat, base = O.solve_base(g)
    a = M.build_arrays(g, vm=vm, va_deg=va)
    ncores = os.cpu_count() or 1
    if cases is None:
        print(json.dumps({"impl": "reference", "unavailable": "the C++ baseline takes outage scenarios; the sweep config has no reference arm"}))
        return
    ns = min(len(cases), cpu_sample_size(g, ncores, args.cpu_sample))
    sample = cases[:ns]
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_ref.run(a, sample[: max(ncores, ns // 8)], variant=0, nthreads=ncores, want_v=False)
    rates, t_all = [], 0.0
    for _ in range(max(1, min(args.steps, 2))):
        r = cpu_ref.run(a, sample, variant=0, nthreads=ncores, want_v=False)
        rates.append(ns / r["seconds"])
        t_all += r["seconds"]
    rB = cpu_ref.run(a, sample, variant=1, nthreads=ncores, want_v=False)
    v = float(np.mean(rates))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": len(rates),
            "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * t_all / len(rates), "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wname, "name": args.config, "buses": g.n, "branches": g.nbranch},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": ncores, "kind": "port",
                             "variant_A_symbolic_every_iteration": v, "variant_B_symbolic_reused": ns / rB["seconds"],
                             "solver": "oracle/cpu_ref: own multifrontal LU (approximate minimum degree, supernodes, partial pivoting in the fronts)",
                             "converged": int(r["converged"].sum()),
                             "sample": f"{ns} N-1 scenarios per step, reference-shaped C++ baseline (variant A), {ncores} threads, contiguous chunks"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
